// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_FUNCTION
#define ISL_COMPAT_BOOST_FUNCTION
#include <functional>
#include <tuple>
namespace boost {
namespace compat_detail {
template <std::size_t I, class Tup, bool = (I < std::tuple_size<Tup>::value)>
struct arg_or_void { typedef void type; };
template <std::size_t I, class Tup>
struct arg_or_void<I, Tup, true> { typedef typename std::tuple_element<I, Tup>::type type; };
}
template <class Sig>
class function;
template <class R, class... A>
class function<R(A...)> : public std::function<R(A...)> {
    typedef std::function<R(A...)> Base;
    typedef std::tuple<A...> Args;

public:
    typedef R result_type;
    typedef typename compat_detail::arg_or_void<0, Args>::type arg1_type;
    typedef typename compat_detail::arg_or_void<1, Args>::type arg2_type;
    typedef typename compat_detail::arg_or_void<2, Args>::type arg3_type;
    typedef typename compat_detail::arg_or_void<3, Args>::type arg4_type;
    typedef typename compat_detail::arg_or_void<4, Args>::type arg5_type;
    typedef typename compat_detail::arg_or_void<5, Args>::type arg6_type;
    typedef arg1_type argument_type;
    typedef arg1_type first_argument_type;
    typedef arg2_type second_argument_type;
    static const int arity = sizeof...(A);
    function() {}
    template <class F>
    function(F f) : Base(f) {}
    template <class F>
    function& operator=(F f) { Base::operator=(f); return *this; }
    bool empty() const { return !static_cast<bool>(*this); }
    void clear() { Base::operator=(nullptr); }
};
}  // namespace boost
#endif
