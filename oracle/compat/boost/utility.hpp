// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_UTILITY
#define ISL_COMPAT_BOOST_UTILITY
#include <iterator>
#include <memory>
#include <utility>
namespace boost {
class noncopyable {
protected:
    noncopyable() {}
    ~noncopyable() {}

private:
    noncopyable(const noncopyable&) = delete;
    noncopyable& operator=(const noncopyable&) = delete;
};
using std::addressof;
template <class T> T next(T x) { return ++x; }
template <class T, class D> T next(T x, D n) { std::advance(x, n); return x; }
template <class T> T prior(T x) { return --x; }
template <class T, class D> T prior(T x, D n) { std::advance(x, -n); return x; }
}
#endif
