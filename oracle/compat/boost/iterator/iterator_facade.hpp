// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_ITERATOR_FACADE
#define ISL_COMPAT_BOOST_ITERATOR_FACADE
#include <cstddef>
#include <iterator>
#include <memory>
#include <type_traits>
namespace boost {
struct incrementable_traversal_tag {};
struct single_pass_traversal_tag : incrementable_traversal_tag {};
struct forward_traversal_tag : single_pass_traversal_tag {};
struct bidirectional_traversal_tag : forward_traversal_tag {};
struct random_access_traversal_tag : bidirectional_traversal_tag {};

class iterator_core_access {
public:
    template <class F> static typename F::reference dereference(const F& f) { return f.dereference(); }
    template <class F> static void increment(F& f) { f.increment(); }
    template <class F> static void decrement(F& f) { f.decrement(); }
    template <class F, class D> static void advance(F& f, D n) { f.advance(n); }
    template <class F> static bool equal(const F& a, const F& b) { return a.equal(b); }
    template <class F> static typename F::difference_type distance_from(const F& a, const F& b) { return -a.distance_to(b); }
    template <class F> static typename F::difference_type distance_to(const F& a, const F& b) { return a.distance_to(b); }
};

namespace compat_detail {
template <class Tag> struct std_category { typedef std::input_iterator_tag type; };
template <> struct std_category<forward_traversal_tag> { typedef std::forward_iterator_tag type; };
template <> struct std_category<bidirectional_traversal_tag> { typedef std::bidirectional_iterator_tag type; };
template <> struct std_category<random_access_traversal_tag> { typedef std::random_access_iterator_tag type; };
// operator-> for iterators whose reference is a value
template <class Ref>
struct arrow_proxy {
    Ref r;
    Ref* operator->() { return std::addressof(r); }
};
template <class Ref, bool IsRef = std::is_reference<Ref>::value>
struct arrow {
    typedef arrow_proxy<typename std::remove_const<Ref>::type> type;
    static type make(Ref r) { return type{r}; }
};
template <class Ref>
struct arrow<Ref, true> {
    typedef typename std::remove_reference<Ref>::type* type;
    static type make(Ref r) { return std::addressof(r); }
};
}  // namespace compat_detail

template <class Derived, class Value, class Traversal, class Reference = Value&, class Difference = std::ptrdiff_t>
class iterator_facade {
    Derived& derived() { return *static_cast<Derived*>(this); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }

public:
    typedef typename std::remove_const<Value>::type value_type;
    typedef Reference reference;
    typedef Difference difference_type;
    typedef typename compat_detail::arrow<Reference>::type pointer;
    typedef typename compat_detail::std_category<Traversal>::type iterator_category;

    reference operator*() const { return iterator_core_access::dereference(derived()); }
    pointer operator->() const { return compat_detail::arrow<Reference>::make(*derived()); }
    reference operator[](difference_type n) const { Derived t(derived()); t += n; return *t; }
    Derived& operator++() { iterator_core_access::increment(derived()); return derived(); }
    Derived operator++(int) { Derived t(derived()); ++*this; return t; }
    Derived& operator--() { iterator_core_access::decrement(derived()); return derived(); }
    Derived operator--(int) { Derived t(derived()); --*this; return t; }
    Derived& operator+=(difference_type n) { iterator_core_access::advance(derived(), n); return derived(); }
    Derived& operator-=(difference_type n) { iterator_core_access::advance(derived(), -n); return derived(); }
    Derived operator+(difference_type n) const { Derived t(derived()); t += n; return t; }
    Derived operator-(difference_type n) const { Derived t(derived()); t -= n; return t; }
};

#define ISL_FACADE_TPL template <class D, class V, class T, class R, class Diff>
#define ISL_FACADE iterator_facade<D, V, T, R, Diff>
ISL_FACADE_TPL D operator+(Diff n, const ISL_FACADE& d) { D t(static_cast<const D&>(d)); t += n; return t; }
ISL_FACADE_TPL bool operator==(const ISL_FACADE& a, const ISL_FACADE& b) {
    return iterator_core_access::equal(static_cast<const D&>(a), static_cast<const D&>(b));
}
ISL_FACADE_TPL bool operator!=(const ISL_FACADE& a, const ISL_FACADE& b) { return !(a == b); }
ISL_FACADE_TPL Diff operator-(const ISL_FACADE& a, const ISL_FACADE& b) {
    return iterator_core_access::distance_to(static_cast<const D&>(b), static_cast<const D&>(a));
}
ISL_FACADE_TPL bool operator<(const ISL_FACADE& a, const ISL_FACADE& b) { return (b - a) > 0; }
ISL_FACADE_TPL bool operator>(const ISL_FACADE& a, const ISL_FACADE& b) { return (b - a) < 0; }
ISL_FACADE_TPL bool operator<=(const ISL_FACADE& a, const ISL_FACADE& b) { return (b - a) >= 0; }
ISL_FACADE_TPL bool operator>=(const ISL_FACADE& a, const ISL_FACADE& b) { return (b - a) <= 0; }
#undef ISL_FACADE_TPL
#undef ISL_FACADE
}  // namespace boost
#endif
