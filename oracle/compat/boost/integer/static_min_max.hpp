// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_STATIC_MIN_MAX
#define ISL_COMPAT_BOOST_STATIC_MIN_MAX
namespace boost {
template <unsigned long A, unsigned long B> struct static_unsigned_max { static const unsigned long value = (A > B) ? A : B; };
template <unsigned long A, unsigned long B> struct static_unsigned_min { static const unsigned long value = (A < B) ? A : B; };
template <long A, long B> struct static_signed_max { static const long value = (A > B) ? A : B; };
template <long A, long B> struct static_signed_min { static const long value = (A < B) ? A : B; };
}
#endif
