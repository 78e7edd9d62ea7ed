// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_TYPE_TRAITS
#define ISL_COMPAT_BOOST_TYPE_TRAITS
#include <type_traits>
namespace boost {
using std::is_same;
using std::remove_const;
using std::remove_reference;
using std::remove_pointer;
using std::remove_cv;
using std::is_fundamental;
using std::is_pointer;
using std::is_const;
using std::is_arithmetic;
using std::is_integral;
using std::is_floating_point;
using std::add_const;
using std::add_pointer;
using std::is_base_of;
using std::is_convertible;
}
#endif
