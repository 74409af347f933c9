// Minimal stand-in for boost::icl::discrete_interval. TEST INFRASTRUCTURE ONLY:
// lets the UNMODIFIED reference src/snpCaller/call_vC.cpp compile in an image
// without boost (used at call_vC.cpp:276). Not a copy of boost; written from the
// documented semantics of the three ICL entry points the reference uses.
#pragma once
#include <functional>
#include <utility>
namespace boost { namespace icl {

struct interval_bounds {
    int bits;
    static interval_bounds closed() { interval_bounds b; b.bits = 3; return b; }
};

// std::less as a template-template argument makes namespace std an associated
// namespace, so the reference's unqualified make_pair(...) finds std::make_pair
// through ADL exactly as it does with real boost.
template <class T, template <class> class Compare = std::less>
class discrete_interval {
public:
    discrete_interval() : lo_(T()), hi_(T()) {}
    discrete_interval(T lo, T hi) : lo_(lo), hi_(hi) {}
    T lower() const { return lo_; }
    T upper() const { return hi_; }
    bool contains(T x) const { return !(x < lo_) && !(hi_ < x); }
private:
    T lo_, hi_;   // closed bounds
};

template <class IntervalT, class T>
IntervalT construct(T lo, T hi, interval_bounds) { return IntervalT(lo, hi); }

}} // namespace boost::icl
