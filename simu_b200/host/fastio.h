// Buffered text I/O for the host driver: SURVEY.md 8(f) rows 3-4 (the .bim name lookup and the
// .pheno / .causals writers are the largest host costs once the kernels are fast: an 11M-variant
// run writes 11M .causals lines and parses 11M .bim lines).
//
//   TextWriter  formats exactly like the reference's `ofstream << value` at default precision
//               (source.cc:1062-1069, :1078-1105): doubles as printf("%g") with 6 significant
//               digits -- std::to_chars(general, 6) is specified to produce the same characters
//               (checked against ostringstream over 2e7 values incl. inf/nan/-0 in
//               tests/test_host_fastio.py) -- into a 1 MB buffer instead of per-field ostream calls.
//   FileText    a whole text file in memory, tokenised in place (no per-field std::string).
#pragma once

#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

namespace simu {

// The same operator<< set into a growing memory buffer: worker threads format slices of a big table
// (.causals of an 11M-variant run) in parallel, the owner writes the slices out in order.
class MemWriter {
public:
    MemWriter &operator<<(std::string_view s) { buf_.append(s.data(), s.size()); return *this; }
    MemWriter &operator<<(const std::string &s) { buf_.append(s); return *this; }
    MemWriter &operator<<(const char *s) { buf_.append(s); return *this; }
    MemWriter &operator<<(char c) { buf_.push_back(c); return *this; }
    MemWriter &operator<<(double v)       // == ostream << double at precision 6 == printf("%g")
    {
        char tmp[40];
        buf_.append(tmp, (size_t)(std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::general, 6).ptr - tmp));
        return *this;
    }
    MemWriter &operator<<(long long v)
    {
        char tmp[24];
        buf_.append(tmp, (size_t)(std::to_chars(tmp, tmp + sizeof(tmp), v).ptr - tmp));
        return *this;
    }
    MemWriter &operator<<(int v) { return *this << (long long)v; }
    const std::string &str() const { return buf_; }
    void reserve(size_t n) { buf_.reserve(n); }

private:
    std::string buf_;
};

class TextWriter {
public:
    explicit TextWriter(const std::string &path) : fp_(fopen(path.c_str(), "wb")), buf_(1 << 20) {}
    ~TextWriter() { flush(); if (fp_) fclose(fp_); }
    TextWriter(const TextWriter &) = delete;
    TextWriter &operator=(const TextWriter &) = delete;

    TextWriter &operator<<(std::string_view s)
    {
        if (s.size() > buf_.size() - n_) { flush(); if (s.size() > buf_.size()) { raw(s.data(), s.size()); return *this; } }
        memcpy(buf_.data() + n_, s.data(), s.size());
        n_ += s.size();
        return *this;
    }
    TextWriter &operator<<(const std::string &s) { return *this << std::string_view(s); }
    TextWriter &operator<<(const char *s) { return *this << std::string_view(s); }
    TextWriter &operator<<(char c) { room(1); buf_[n_++] = c; return *this; }
    TextWriter &operator<<(double v)       // == ostream << double at precision 6 == printf("%g")
    {
        room(40);
        n_ = (size_t)(std::to_chars(buf_.data() + n_, buf_.data() + buf_.size(), v, std::chars_format::general, 6).ptr - buf_.data());
        return *this;
    }
    TextWriter &operator<<(long long v)
    {
        room(24);
        n_ = (size_t)(std::to_chars(buf_.data() + n_, buf_.data() + buf_.size(), v).ptr - buf_.data());
        return *this;
    }
    TextWriter &operator<<(int v) { return *this << (long long)v; }
    void flush() { raw(buf_.data(), n_); n_ = 0; }

private:
    void room(size_t k) { if (buf_.size() - n_ < k) flush(); }
    void raw(const char *p, size_t k) { if (fp_ && k) fwrite(p, 1, k, fp_); }   // an unopenable file is silently skipped, like ofstream
    FILE *fp_;
    std::vector<char> buf_;
    size_t n_ = 0;
};

// A text file read in one piece; lines and blank-separated fields are handed out as views into
// the (mutable) buffer.  Field separators: runs of ' ', '\t', '\r' (libplinkio's delimiter set plus
// the CR that getline-based parsing leaves behind); empty lines are skipped by the callers.
class FileText {
public:
    explicit FileText(const std::string &path)
    {
        FILE *fp = fopen(path.c_str(), "rb");
        if (!fp) return;
        ok_ = true;
        if (fseek(fp, 0, SEEK_END) == 0) {        // regular file: one read of the whole thing
            const long sz = ftell(fp);
            rewind(fp);
            if (sz > 0) data_.reserve((size_t)sz);
        }
        char chunk[1 << 16];
        size_t k;
        while ((k = fread(chunk, 1, sizeof(chunk), fp)) > 0) data_.append(chunk, k);
        fclose(fp);
    }
    bool ok() const { return ok_; }
    size_t count_lines() const       // upper bound on the number of records, for reserve()
    {
        size_t n = 1;
        const char *p = data_.data(), *end = p + data_.size();
        while ((p = (const char *)memchr(p, '\n', (size_t)(end - p)))) { n++; p++; }
        return n;
    }
    // next line (without '\n') into [*b, *e); false at end of file
    bool next_line(char **b, char **e)
    {
        if (pos_ >= data_.size()) return false;
        char *base = &data_[0];
        char *p = base + pos_, *end = base + data_.size();
        char *q = (char *)memchr(p, '\n', (size_t)(end - p));
        if (!q) q = end;
        *b = p; *e = q;
        pos_ = (size_t)(q - base) + 1;
        return true;
    }
    char *begin() { return data_.empty() ? nullptr : &data_[0]; }
    size_t size() const { return data_.size(); }
    // cut the text into `parts` pieces that end on line boundaries: piece p = [cuts[p], cuts[p+1])
    std::vector<size_t> cut_at_lines(int parts) const
    {
        std::vector<size_t> cuts(1, 0);
        for (int p = 1; p < parts; p++) {
            size_t at = data_.size() * (size_t)p / (size_t)parts;
            if (at < cuts.back()) at = cuts.back();
            const void *nl = at < data_.size() ? memchr(data_.data() + at, '\n', data_.size() - at) : nullptr;
            cuts.push_back(nl ? (size_t)((const char *)nl - data_.data()) + 1 : data_.size());
        }
        cuts.push_back(data_.size());
        return cuts;
    }
    // next line of [*pos, end) (without '\n') into [*b, *e); false when the piece is exhausted
    bool next_line_in(size_t *pos, size_t end, char **b, char **e)
    {
        if (*pos >= end) return false;
        char *base = &data_[0];
        char *p = base + *pos, *stop = base + end;
        char *q = (char *)memchr(p, '\n', (size_t)(stop - p));
        if (!q) q = stop;
        *b = p; *e = q;
        *pos = (size_t)(q - base) + 1;
        return true;
    }
    // split [b, e) into at most max_fields views; returns the number of fields found (may exceed max_fields)
    static int split(char *b, char *e, std::string_view *out, int max_fields)
    {
        int n = 0;
        char *p = b;
        while (p < e) {
            while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
            char *q = p;
            while (q < e && *q != ' ' && *q != '\t' && *q != '\r') q++;
            if (q > p) { if (n < max_fields) out[n] = std::string_view(p, (size_t)(q - p)); n++; }
            p = q;
        }
        return n;
    }

private:
    std::string data_;
    size_t pos_ = 0;
    bool ok_ = false;
};

// host threads for the text stages (SIMU_B200_IO_THREADS, default min(8, cores))
inline int text_threads()
{
    if (const char *e = getenv("SIMU_B200_IO_THREADS")) return std::max(1, atoi(e));
    return (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
}

// run fn(t) for t in [0, n) on n threads (n == 1: inline)
template <typename F> inline void parallel_for_threads(int n, F fn)
{
    if (n <= 1) { fn(0); return; }
    std::vector<std::thread> pool;
    for (int t = 1; t < n; t++) pool.emplace_back(fn, t);
    fn(0);
    for (auto &th : pool) th.join();
}

// Write `count` table rows: row(out, j) prints row j into `out` (a TextWriter or a MemWriter -- the same
// operator<< set).  Beyond 200K rows the slices of 1M-row batches are formatted into memory by several
// threads and written out in order; the file gets the same bytes either way.
template <typename RowFn> inline void write_rows(TextWriter &file, size_t count, RowFn row)
{
    const int threads = count > 200000 ? text_threads() : 1;
    if (threads == 1) {
        for (size_t j = 0; j < count; j++) row(file, j);
        return;
    }
    const size_t batch = (size_t)1 << 20;
    for (size_t b0 = 0; b0 < count; b0 += batch) {
        const size_t b1 = std::min(count, b0 + batch);
        std::vector<MemWriter> parts((size_t)threads);
        parallel_for_threads(threads, [&](int t) {
            const size_t lo = b0 + (b1 - b0) * (size_t)t / (size_t)threads, hi = b0 + (b1 - b0) * (size_t)(t + 1) / (size_t)threads;
            parts[(size_t)t].reserve((hi - lo) * 72);
            for (size_t j = lo; j < hi; j++) row(parts[(size_t)t], j);
        });
        for (const MemWriter &m : parts) file << std::string_view(m.str());
    }
}

// complete-token numeric parses (strtoll / strtod semantics of the reference's parsers, which
// require the whole field to be consumed: bim_parse.c:112-175)
inline bool full_long(std::string_view s, long long *v)
{
    if (s.empty()) return false;
    if (s.size() <= 18) {                        // plain decimal digits: no strtoll, cannot overflow
        size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
        long long acc = 0;
        bool digits = i < s.size();
        for (size_t k = i; k < s.size(); k++) {
            if (s[k] < '0' || s[k] > '9') { digits = false; break; }
            acc = acc * 10 + (s[k] - '0');
        }
        if (digits) { *v = s[0] == '-' ? -acc : acc; return true; }
    }
    char tmp[64];
    std::string big;
    const char *c;
    if (s.size() < sizeof(tmp)) { memcpy(tmp, s.data(), s.size()); tmp[s.size()] = '\0'; c = tmp; }
    else { big.assign(s); c = big.c_str(); }
    char *end = nullptr;
    *v = strtoll(c, &end, 10);
    return end && *end == '\0';
}

inline bool full_double(std::string_view s, double *v)
{
    if (s.empty()) return false;
    if (s.size() == 1 && s[0] >= '0' && s[0] <= '9') { *v = s[0] - '0'; return true; }   // the usual "0" cM column
    char tmp[128];
    std::string big;
    const char *c;
    if (s.size() < sizeof(tmp)) { memcpy(tmp, s.data(), s.size()); tmp[s.size()] = '\0'; c = tmp; }
    else { big.assign(s); c = big.c_str(); }
    char *end = nullptr;
    *v = strtod(c, &end);
    return end && *end == '\0';
}

}  // namespace simu
