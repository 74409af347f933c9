// Stand-in for <boost/icl/closed_interval.hpp> (included at call_vC.cpp:21, unused).
#pragma once
#include "discrete_interval.hpp"
