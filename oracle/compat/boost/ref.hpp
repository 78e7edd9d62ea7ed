// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_REF
#define ISL_COMPAT_BOOST_REF
#include <functional>
namespace boost {
using std::ref;
using std::cref;
using std::reference_wrapper;
}
#endif
