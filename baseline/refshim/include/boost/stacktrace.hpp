// boost::stacktrace interface subset on top of glibc backtrace(): enough for the reference's
// crash handler (operator<< of a default-constructed stacktrace).
#ifndef REFSHIM_BOOST_STACKTRACE_HPP
#define REFSHIM_BOOST_STACKTRACE_HPP
#include <execinfo.h>
#include <cstdlib>
#include <ostream>
#include <string>
#include <vector>
namespace boost { namespace stacktrace {
class frame {
   public:
    explicit frame(const void* a = nullptr, std::string s = std::string()) : addr_(a), sym_(std::move(s)) {}
    const void* address() const { return addr_; }
    std::string name() const { return sym_; }
   private:
    const void* addr_;
    std::string sym_;
};
class stacktrace {
   public:
    stacktrace() {
        void* buf[64];
        const int n = ::backtrace(buf, 64);
        char** syms = ::backtrace_symbols(buf, n);
        for (int i = 0; i < n; ++i) frames_.emplace_back(buf[i], syms ? syms[i] : "");
        std::free(syms);
    }
    size_t size() const { return frames_.size(); }
    const frame& operator[](size_t i) const { return frames_[i]; }
    std::vector<frame>::const_iterator begin() const { return frames_.begin(); }
    std::vector<frame>::const_iterator end() const { return frames_.end(); }
   private:
    std::vector<frame> frames_;
};
inline std::ostream& operator<<(std::ostream& os, const frame& f) { return os << f.name(); }
inline std::ostream& operator<<(std::ostream& os, const stacktrace& st) {
    for (size_t i = 0; i < st.size(); ++i) os << " " << i << "# " << st[i].name() << "\n";
    return os;
}
inline std::string to_string(const stacktrace& st) {
    std::string s;
    for (size_t i = 0; i < st.size(); ++i) s += st[i].name() + "\n";
    return s;
}
}}  // namespace boost::stacktrace
#endif
