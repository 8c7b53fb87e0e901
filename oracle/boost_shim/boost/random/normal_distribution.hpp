// shim of <boost/random/normal_distribution.hpp>: the cached Box-Muller of Boost <= 1.55;
// uniform_01 takes one 32-bit draw per double.  (Boost >= 1.56 uses a ziggurat: not restated.)
#pragma once
#include <cmath>
#include "../cstdint_shim.hpp"
namespace boost {
template <class RealType = double> class normal_distribution {
public:
    typedef RealType result_type;
    explicit normal_distribution(RealType mean = 0, RealType sigma = 1)
        : mean_(mean), sigma_(sigma), r1_(0), r2_(0), rho_(0), valid_(false) {}
    template <class Engine> result_type operator()(Engine &eng)
    {
        const result_type pi = result_type(3.14159265358979323846);
        if (!valid_) {
            r1_ = u01(eng);
            r2_ = u01(eng);
            rho_ = std::sqrt(-result_type(2) * std::log(result_type(1) - r2_));
            valid_ = true;
        } else {
            valid_ = false;
        }
        return rho_ * (valid_ ? std::cos(result_type(2) * pi * r1_) : std::sin(result_type(2) * pi * r1_)) * sigma_ + mean_;
    }
private:
    template <class Engine> static result_type u01(Engine &eng) { return result_type(eng()) * (result_type(1) / result_type(4294967296.0)); }
    RealType mean_, sigma_, r1_, r2_, rho_;
    bool valid_;
};
}  // namespace boost
