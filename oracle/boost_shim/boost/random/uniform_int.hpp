// shim of <boost/random/uniform_int.hpp>: operator()(eng, n) -> integer in [0, n), by the bucket
// rejection rule of boost::random::detail::generate_uniform_int for a 32-bit engine.
#pragma once
#include "../cstdint_shim.hpp"
namespace boost {
template <class IntType = int> class uniform_int {
public:
    typedef IntType result_type;
    explicit uniform_int(IntType lo = 0, IntType hi = 9) : lo_(lo), hi_(hi) {}
    template <class Engine> result_type operator()(Engine &eng) { return generate(eng, lo_, hi_); }
    template <class Engine> result_type operator()(Engine &eng, IntType n) { return generate(eng, 0, n - 1); }
private:
    template <class Engine> static result_type generate(Engine &eng, IntType lo, IntType hi)
    {
        const uint32_t range = static_cast<uint32_t>(hi) - static_cast<uint32_t>(lo);
        const uint32_t brange = 0xFFFFFFFFu;
        if (range == 0) return lo;
        if (range == brange) return static_cast<IntType>(eng() + static_cast<uint32_t>(lo));
        uint32_t bucket = brange / (range + 1);
        if (brange % (range + 1) == range) ++bucket;
        for (;;) {
            const uint32_t r = eng() / bucket;
            if (r <= range) return static_cast<IntType>(r + static_cast<uint32_t>(lo));
        }
    }
    IntType lo_, hi_;
};
}  // namespace boost
