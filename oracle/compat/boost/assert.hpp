// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#include <cassert>
#ifndef BOOST_ASSERT
#define BOOST_ASSERT(x) assert(x)
#endif
