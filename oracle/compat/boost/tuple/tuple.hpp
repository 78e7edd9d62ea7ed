// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_TUPLE
#define ISL_COMPAT_BOOST_TUPLE
#include <functional>
#include <tuple>
namespace boost {
namespace tuples {
struct null_type {};
template <class... T>
class tuple : public std::tuple<T...> {
    typedef std::tuple<T...> Base;

public:
    tuple() {}
    template <class... U, class = typename std::enable_if<sizeof...(U) == sizeof...(T) && (sizeof...(U) > 0)>::type>
    tuple(U&&... u) : Base(std::forward<U>(u)...) {}
    template <class... U>
    tuple(const tuple<U...>& o) : Base(static_cast<const std::tuple<U...>&>(o)) {}
    template <class... U>
    tuple(const std::tuple<U...>& o) : Base(o) {}
    template <class... U>
    tuple& operator=(const tuple<U...>& o) { Base::operator=(static_cast<const std::tuple<U...>&>(o)); return *this; }
    template <int N>
    typename std::tuple_element<N, Base>::type& get() { return std::get<N>(static_cast<Base&>(*this)); }
    template <int N>
    const typename std::tuple_element<N, Base>::type& get() const { return std::get<N>(static_cast<const Base&>(*this)); }
};
template <int N, class TUP>
struct element;
template <int N, class... T>
struct element<N, tuple<T...> > { typedef typename std::tuple_element<N, std::tuple<T...> >::type type; };
template <int N, class... T>
struct element<N, const tuple<T...> > { typedef const typename std::tuple_element<N, std::tuple<T...> >::type type; };
template <class TUP>
struct length;
template <class... T>
struct length<tuple<T...> > { static const int value = sizeof...(T); };
template <int N, class... T>
typename std::tuple_element<N, std::tuple<T...> >::type& get(tuple<T...>& t) { return t.template get<N>(); }
template <int N, class... T>
const typename std::tuple_element<N, std::tuple<T...> >::type& get(const tuple<T...>& t) { return t.template get<N>(); }
template <class T> struct unwrap { typedef T type; };
template <class T> struct unwrap<std::reference_wrapper<T> > { typedef T& type; };
template <class... T>
tuple<typename unwrap<typename std::decay<T>::type>::type...> make_tuple(T&&... t) {
    return tuple<typename unwrap<typename std::decay<T>::type>::type...>(std::forward<T>(t)...);
}
template <class... T>
tuple<T&...> tie(T&... t) { return tuple<T&...>(t...); }
}  // namespace tuples
using tuples::tuple;
using tuples::make_tuple;
using tuples::tie;
using tuples::get;
}  // namespace boost
#endif
