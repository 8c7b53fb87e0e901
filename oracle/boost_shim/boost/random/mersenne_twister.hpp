// shim of <boost/random/mersenne_twister.hpp>: boost::mt19937 == the MT19937 of the C++ standard
// (same recurrence and init_genrand seeding); an integral seed narrows to 32 bits.
#pragma once
#include <random>
#include "../cstdint_shim.hpp"
namespace boost {
class mt19937 {
public:
    typedef uint32_t result_type;
    mt19937() : e_(5489u) {}
    template <class T> explicit mt19937(const T &seed) : e_(static_cast<uint32_t>(seed)) {}
    result_type operator()() { return static_cast<result_type>(e_()); }
    static result_type min() { return 0u; }
    static result_type max() { return 0xFFFFFFFFu; }
private:
    std::mt19937 e_;
};
}  // namespace boost
