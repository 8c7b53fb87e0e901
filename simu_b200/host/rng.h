// Host random-number layer: reproduces the draw sequence of the reference, which
// uses ONE boost::mt19937 for everything (/root/reference/src/source.cc:1222) through
//   * variate_generator<mt19937&, uniform_int<>>   called as rand(n) by std::random_shuffle
//     (source.cc:737-741, :814-831, :1045-1047)
//   * variate_generator<mt19937&, normal_distribution<>>   (source.cc:894, :914, :937, :1015)
//
// Boost is an un-vendored, un-pinned dependency of the reference (CMakeLists.txt:14 lists
// 1.40-1.57) and is absent from this image, so the algorithms below are restated from the
// published Boost.Random sources of that version range:
//   * mt19937: identical to std::mt19937 (same recurrence, same 32-bit seeding); the int64
//     seed narrows to uint32 exactly as boost's arithmetic constructor does.
//   * uniform_int<>()(eng, n): generate_uniform_int(eng, 0, n-1) with the bucket rejection
//     rule (bucket = 0xFFFFFFFF/(range+1), +1 if the remainder equals range).
//   * normal_distribution<>: the cached Box-Muller of Boost <= 1.55
//     (rho = sqrt(-2 log(1-r2)); cos(2 pi r1) then sin(2 pi r1)); uniform_01 takes ONE 32-bit
//     draw per double (u * 2^-32).  Boost >= 1.56 switched to a 128-layer ziggurat whose
//     tables cannot be restated here; draws against such a build are NOT pinned.
//   * std::random_shuffle(first, last, rand): libstdc++'s loop
//     (for i in 1..n-1: j = rand(i+1); if i != j swap).
// Hot-path parity does not depend on any of this: the CPU checker and the GPU path consume the
// same x and e.
#pragma once

#include <cmath>
#include <cstdint>
#include <random>
#include <utility>
#include <vector>

namespace simu {

class Mt19937 {
public:
    explicit Mt19937(int64_t seed) : eng_(static_cast<uint32_t>(seed)) {}
    uint32_t operator()() { return static_cast<uint32_t>(eng_()); }

private:
    std::mt19937 eng_;
};

// uniform_int<>()(eng, n): integer in [0, n)
inline int uniform_int_below(Mt19937 &eng, int n)
{
    const uint32_t range = static_cast<uint32_t>(n - 1);
    if (range == 0) return 0;
    const uint32_t brange = 0xFFFFFFFFu;
    if (range == brange) return static_cast<int>(eng());
    uint32_t bucket = brange / (range + 1);
    if (brange % (range + 1) == range) ++bucket;
    for (;;) {
        const uint32_t r = eng() / bucket;
        if (r <= range) return static_cast<int>(r);
    }
}

// std::random_shuffle(first, last, rand) as libstdc++ implements it
template <typename T>
void random_shuffle(std::vector<T> &v, Mt19937 &eng)
{
    for (size_t i = 1; i < v.size(); ++i) {
        const size_t j = static_cast<size_t>(uniform_int_below(eng, static_cast<int>(i + 1)));
        if (i != j) std::swap(v[i], v[j]);
    }
}

// a fresh object per stage, like the reference's `variate_generator ... random_normal(...)`
class NormalDraws {
public:
    explicit NormalDraws(Mt19937 &eng) : eng_(eng) {}
    double operator()()
    {
        const double pi = 3.14159265358979323846;
        if (!valid_) {
            r1_ = uniform01();
            r2_ = uniform01();
            rho_ = std::sqrt(-2.0 * std::log(1.0 - r2_));
            valid_ = true;
        } else {
            valid_ = false;
        }
        return rho_ * (valid_ ? std::cos(2.0 * pi * r1_) : std::sin(2.0 * pi * r1_)) * 1.0 + 0.0;
    }

private:
    double uniform01() { return static_cast<double>(eng_()) * (1.0 / 4294967296.0); }
    Mt19937 &eng_;
    double r1_ = 0, r2_ = 0, rho_ = 0;
    bool valid_ = false;
};

}  // namespace simu
