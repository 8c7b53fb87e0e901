// shim of <boost/random/variate_generator.hpp>
#pragma once
#include "../cstdint_shim.hpp"
namespace boost {
template <class Engine, class Distribution> class variate_generator {   // Engine is a reference type here
public:
    typedef typename Distribution::result_type result_type;
    variate_generator(Engine e, Distribution d) : eng_(e), dist_(d) {}
    result_type operator()() { return dist_(eng_); }
    template <class T> result_type operator()(const T &value) { return dist_(eng_, value); }
private:
    Engine eng_;
    Distribution dist_;
};
}  // namespace boost
