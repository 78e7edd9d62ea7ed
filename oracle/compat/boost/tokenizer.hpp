// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_TOKENIZER
#define ISL_COMPAT_BOOST_TOKENIZER
#include <string>
#include <vector>
namespace boost {
template <class C>
class char_separator {
public:
    std::basic_string<C> dropped;
    explicit char_separator(const C* d = " ") : dropped(d) {}
};
template <class SEP>
class tokenizer {
    std::vector<std::string> tok_;

public:
    typedef std::vector<std::string>::const_iterator iterator;
    typedef iterator const_iterator;
    tokenizer(const std::string& s, const SEP& sep) {
        std::string cur;
        for (char ch : s) {
            if (sep.dropped.find(ch) != std::string::npos) {
                if (!cur.empty()) { tok_.push_back(cur); cur.clear(); }
            } else cur.push_back(ch);
        }
        if (!cur.empty()) tok_.push_back(cur);
        // dereferencing end() of an empty token list must not crash readers that only test find()
        if (tok_.empty()) tok_.push_back(std::string());
    }
    iterator begin() const { return tok_.begin(); }
    iterator end() const { return tok_.end(); }
};
}
#endif
