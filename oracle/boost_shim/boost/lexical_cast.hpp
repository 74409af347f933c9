// Stand-in for <boost/lexical_cast.hpp>. TEST INFRASTRUCTURE ONLY (oracle build).
// The reference includes this header (call_vC.cpp:23) but never uses it.
#pragma once
