// Clean-room boost::intrusive_ptr (interface subset) for building the unmodified reference on
// an image without Boost headers; only used by baseline/_ref (bench.py --impl reference).
#ifndef REFSHIM_BOOST_INTRUSIVE_PTR_HPP
#define REFSHIM_BOOST_INTRUSIVE_PTR_HPP
#include <cstddef>
#include <functional>
#include <utility>
namespace boost {
template <class T>
class intrusive_ptr {
   public:
    typedef T element_type;
    intrusive_ptr() noexcept : p_(nullptr) {}
    intrusive_ptr(std::nullptr_t) noexcept : p_(nullptr) {}
    intrusive_ptr(T* p, bool add_ref = true) : p_(p) {
        if (p_ && add_ref) intrusive_ptr_add_ref(p_);
    }
    intrusive_ptr(const intrusive_ptr& o) : p_(o.p_) {
        if (p_) intrusive_ptr_add_ref(p_);
    }
    intrusive_ptr(intrusive_ptr&& o) noexcept : p_(o.p_) { o.p_ = nullptr; }
    ~intrusive_ptr() {
        if (p_) intrusive_ptr_release(p_);
    }
    intrusive_ptr& operator=(intrusive_ptr o) noexcept {
        std::swap(p_, o.p_);
        return *this;
    }
    void reset() { intrusive_ptr().swap(*this); }
    void reset(T* p) { intrusive_ptr(p).swap(*this); }
    T* get() const noexcept { return p_; }
    T* detach() noexcept {
        T* r = p_;
        p_ = nullptr;
        return r;
    }
    T& operator*() const noexcept { return *p_; }
    T* operator->() const noexcept { return p_; }
    explicit operator bool() const noexcept { return p_ != nullptr; }
    void swap(intrusive_ptr& o) noexcept { std::swap(p_, o.p_); }

   private:
    T* p_;
};
template <class T, class U>
bool operator==(const intrusive_ptr<T>& a, const intrusive_ptr<U>& b) noexcept { return a.get() == b.get(); }
template <class T, class U>
bool operator!=(const intrusive_ptr<T>& a, const intrusive_ptr<U>& b) noexcept { return a.get() != b.get(); }
template <class T>
bool operator==(const intrusive_ptr<T>& a, std::nullptr_t) noexcept { return a.get() == nullptr; }
template <class T>
bool operator!=(const intrusive_ptr<T>& a, std::nullptr_t) noexcept { return a.get() != nullptr; }
template <class T>
bool operator<(const intrusive_ptr<T>& a, const intrusive_ptr<T>& b) noexcept { return std::less<T*>()(a.get(), b.get()); }
}  // namespace boost
#endif
