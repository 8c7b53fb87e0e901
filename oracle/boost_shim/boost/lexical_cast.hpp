// shim of <boost/lexical_cast.hpp>: whole-token conversions, bad_lexical_cast on failure
#pragma once
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>
#include "cstdint_shim.hpp"
namespace boost {
class bad_lexical_cast : public std::bad_cast {
public:
    const char *what() const noexcept override { return "bad lexical cast: source type value could not be interpreted as target"; }
};
namespace shim_detail {
template <class T> struct caster {
    static T from(const std::string &s)
    {
        if (s.empty() || s[0] == ' ' || s[0] == '\t' || s[0] == '\n') throw bad_lexical_cast();
        std::istringstream is(s);
        T v;
        is >> v;
        if (is.fail() || !is.eof()) { char c; if (is.fail() || is.get(c)) throw bad_lexical_cast(); }
        return v;
    }
};
template <> struct caster<double> {
    static double from(const std::string &s)
    {
        if (s.empty() || s[0] == ' ' || s[0] == '\t' || s[0] == '\n') throw bad_lexical_cast();
        if (s.size() > 1 && (s[1] == 'x' || s[1] == 'X')) throw bad_lexical_cast();
        if (s.size() > 2 && (s[0] == '-' || s[0] == '+') && (s[2] == 'x' || s[2] == 'X')) throw bad_lexical_cast();
        char *end = nullptr;
        const double v = std::strtod(s.c_str(), &end);
        if (!end || *end != '\0' || end == s.c_str()) throw bad_lexical_cast();
        return v;
    }
};
template <> struct caster<float> {
    static float from(const std::string &s) { caster<double>::from(s); return std::strtof(s.c_str(), nullptr); }
};
template <> struct caster<std::string> {
    static std::string from(const std::string &s) { return s; }
};
}  // namespace shim_detail
template <class Target, class Source> Target lexical_cast(const Source &src)
{
    std::ostringstream os;
    os << src;
    return shim_detail::caster<Target>::from(os.str());
}
}  // namespace boost
