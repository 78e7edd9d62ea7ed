// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_UNORDERED_SET
#define ISL_COMPAT_BOOST_UNORDERED_SET
#include <functional>
#include <unordered_set>
#include <utility>
namespace boost {
template <class T>
struct hash : std::hash<T> {};
template <class A, class B>
struct hash<std::pair<A, B> > {
    std::size_t operator()(const std::pair<A, B>& p) const {
        std::size_t seed = 0;
        seed ^= std::hash<A>()(p.first) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        seed ^= std::hash<B>()(p.second) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};
template <class K, class H = hash<K>, class E = std::equal_to<K> >
using unordered_set = std::unordered_set<K, H, E>;
}
#endif
