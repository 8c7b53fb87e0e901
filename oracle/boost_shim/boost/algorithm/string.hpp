// shim of <boost/algorithm/string.hpp>: trim_if, split, is_any_of, token_compress_on, contains, replace_all
#pragma once
#include <string>
#include <vector>
#include "../cstdint_shim.hpp"
namespace boost {
struct is_any_of_pred {
    std::string set;
    bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline is_any_of_pred is_any_of(const std::string &s) { is_any_of_pred p; p.set = s; return p; }
enum token_compress_mode_type { token_compress_on, token_compress_off };
template <class Pred> void trim_if(std::string &s, Pred p)
{
    size_t a = 0, b = s.size();
    while (a < b && p(s[a])) a++;
    while (b > a && p(s[b - 1])) b--;
    s = s.substr(a, b - a);
}
template <class Pred>
std::vector<std::string> &split(std::vector<std::string> &out, const std::string &in, Pred p,
                                token_compress_mode_type mode = token_compress_off)
{
    out.clear();
    std::string cur;
    size_t i = 0;
    const size_t n = in.size();
    while (true) {
        while (i < n && !p(in[i])) cur.push_back(in[i++]);
        out.push_back(cur);
        cur.clear();
        if (i >= n) break;
        i++;                                   // the separator
        if (mode == token_compress_on) while (i < n && p(in[i])) i++;
    }
    return out;
}
inline bool contains(const std::string &s, const std::string &what) { return s.find(what) != std::string::npos; }
inline void replace_all(std::string &s, const std::string &from, const std::string &to)
{
    if (from.empty()) return;
    for (size_t p = 0; (p = s.find(from, p)) != std::string::npos; p += to.size()) s.replace(p, from.size(), to);
}
}  // namespace boost
