// Minimal stand-in for boost::icl::split_interval_map. TEST INFRASTRUCTURE ONLY.
// The reference (call_vC.cpp:83,207,276-278,567) needs clear(), operator+= of an
// (interval, value) pair and point lookup operator()(key). Real ICL keeps split
// segments whose codomain is the "+="-aggregate, in insertion order, of every value
// whose interval covers the segment (GeneDef::operator== is always false, so nothing
// is ever absorbed or joined: gene.h:139-150). A point lookup therefore returns
// first_inserted += second_inserted += ... ; we compute that on demand.
#pragma once
#include <vector>
#include <utility>
#include "discrete_interval.hpp"
namespace boost { namespace icl {

template <class DomainT, class CodomainT>
class split_interval_map {
public:
    typedef discrete_interval<DomainT> interval_type;
    void clear() { items_.clear(); }
    split_interval_map& operator+=(const std::pair<interval_type, CodomainT>& kv) {
        items_.push_back(kv);
        return *this;
    }
    CodomainT operator()(const DomainT& x) const {
        bool first = true;
        CodomainT acc;
        for (size_t i = 0; i < items_.size(); ++i) {
            if (!items_[i].first.contains(x)) continue;
            if (first) { acc = items_[i].second; first = false; }
            else acc += items_[i].second;
        }
        return acc;
    }
private:
    std::vector<std::pair<interval_type, CodomainT> > items_;
};

}} // namespace boost::icl
