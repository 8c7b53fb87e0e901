// shim of <boost/date_time/posix_time/posix_time.hpp>: ptime / time_duration as used by source.cc:101-102,
// :1213, :1328-1330 (local clock, microsecond ticks, "YYYY-Mon-DD HH:MM:SS" and "HH:MM:SS" output)
#pragma once
#include <chrono>
#include <cstdio>
#include <ctime>
#include <ostream>
#include "../../cstdint_shim.hpp"
namespace boost {
namespace gregorian {
struct date { int y, m, d; date(int yy, int mm, int dd) : y(yy), m(mm), d(dd) {} };
}
namespace posix_time {
class time_duration {
public:
    explicit time_duration(int64_t us = 0) : us_(us) {}
    int64_t ticks() const { return us_; }
    int64_t total_seconds() const { return us_ / 1000000; }
private:
    int64_t us_;
};
class ptime {
public:
    ptime() : us_(0) {}
    explicit ptime(const gregorian::date &d)
    {
        std::tm tm = std::tm();
        tm.tm_year = d.y - 1900; tm.tm_mon = d.m - 1; tm.tm_mday = d.d;
        us_ = static_cast<int64_t>(timegm(&tm)) * 1000000;
    }
    static ptime from_us(int64_t us) { ptime p; p.us_ = us; return p; }
    int64_t us() const { return us_; }
    time_duration operator-(const ptime &o) const { return time_duration(us_ - o.us_); }
private:
    int64_t us_;   // microseconds since 1970-01-01 on the LOCAL clock
};
inline int64_t local_now_us()
{
    using namespace std::chrono;
    const int64_t utc = duration_cast<microseconds>(system_clock::now().time_since_epoch()).count();
    std::time_t t = static_cast<std::time_t>(utc / 1000000);
    std::tm tm;
    localtime_r(&t, &tm);
    return utc + static_cast<int64_t>(tm.tm_gmtoff) * 1000000;
}
struct microsec_clock { static ptime local_time() { return ptime::from_us(local_now_us()); } };
struct second_clock { static ptime local_time() { return ptime::from_us(local_now_us() / 1000000 * 1000000); } };
inline std::ostream &operator<<(std::ostream &os, const ptime &p)
{
    std::time_t t = static_cast<std::time_t>(p.us() / 1000000);
    std::tm tm;
    gmtime_r(&t, &tm);
    char buf[64];
    strftime(buf, sizeof(buf), "%Y-%b-%d %H:%M:%S", &tm);
    return os << buf;
}
inline std::ostream &operator<<(std::ostream &os, const time_duration &d)
{
    const int64_t s = d.total_seconds();
    char buf[64];
    snprintf(buf, sizeof(buf), "%02lld:%02lld:%02lld", (long long)(s / 3600), (long long)((s / 60) % 60), (long long)(s % 60));
    return os << buf;
}
}  // namespace posix_time
}  // namespace boost
