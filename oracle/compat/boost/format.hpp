// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_FORMAT
#define ISL_COMPAT_BOOST_FORMAT
#include <cstdlib>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>
namespace boost {
// positional "%N%" directives and absolute tabs "%|Nt|" (what base/io/Format.hpp:155-205 builds); feeding an argument
// after the object was streamed starts a new argument list, like boost::format does
class format {
    std::string fmt_;
    std::vector<std::string> args_;
    mutable bool dumped_ = false;

public:
    format() {}
    explicit format(const std::string& f) : fmt_(f) {}
    format& parse(const std::string& f) { fmt_ = f; args_.clear(); return *this; }
    template <class T>
    format& operator%(const T& t) {
        if (dumped_) { args_.clear(); dumped_ = false; }
        std::ostringstream s;
        s << t;
        args_.push_back(s.str());
        return *this;
    }
    std::string str() const {
        std::string out;
        std::size_t line0 = 0;  // start of the current output line
        for (std::size_t i = 0; i < fmt_.size(); ++i) {
            const char ch = fmt_[i];
            if (ch != '%') {
                out.push_back(ch);
                if (ch == '\n') line0 = out.size();
                continue;
            }
            std::size_t j = i + 1, num = 0;
            bool digits = false;
            while (j < fmt_.size() && fmt_[j] >= '0' && fmt_[j] <= '9') { num = 10 * num + (fmt_[j] - '0'); ++j; digits = true; }
            if (digits && j < fmt_.size() && fmt_[j] == '%') {
                if (num >= 1 && num <= args_.size()) out += args_[num - 1];
                i = j;
            } else if (j < fmt_.size() && fmt_[j] == '|') {
                const std::size_t k = fmt_.find('|', j + 1);
                if (k == std::string::npos) break;
                const std::string spec = fmt_.substr(j + 1, k - j - 1);
                if (!spec.empty() && spec[spec.size() - 1] == 't') {
                    const std::size_t col = static_cast<std::size_t>(std::atol(spec.c_str()));
                    while (out.size() - line0 < col) out.push_back(' ');
                }
                i = k;
            } else out.push_back('%');
        }
        dumped_ = true;
        return out;
    }
    void clear() { args_.clear(); }
    friend std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
};
}
#endif
