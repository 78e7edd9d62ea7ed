// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_LEXICAL_CAST
#define ISL_COMPAT_BOOST_LEXICAL_CAST
#include <iomanip>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
namespace boost {
class bad_lexical_cast : public std::bad_cast {
public:
    const char* what() const noexcept override { return "bad lexical cast"; }
};
template <class T, class S>
T lexical_cast(const S& s) {
    std::stringstream ss;
    if (std::is_floating_point<S>::value) ss << std::setprecision(std::numeric_limits<double>::digits10 + 2);
    T t;
    if (!(ss << s) || !(ss >> t)) throw bad_lexical_cast();
    ss >> std::ws;
    if (!ss.eof()) throw bad_lexical_cast();
    return t;
}
template <>
inline std::string lexical_cast<std::string, std::string>(const std::string& s) { return s; }
}
#endif
