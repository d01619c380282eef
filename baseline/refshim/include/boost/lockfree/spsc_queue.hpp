// Clean-room single-producer/single-consumer ring with the boost::lockfree::spsc_queue
// interface subset the reference uses (push / pop / capacity<N> or a run-time size).
#ifndef REFSHIM_BOOST_LOCKFREE_SPSC_QUEUE_HPP
#define REFSHIM_BOOST_LOCKFREE_SPSC_QUEUE_HPP
#include <atomic>
#include <cstddef>
#include <vector>
namespace boost { namespace lockfree {
template <size_t N>
struct capacity { static constexpr size_t value = N; };
namespace detail {
template <class... O> struct cap_of { static constexpr size_t value = 0; };
template <size_t N, class... O> struct cap_of<capacity<N>, O...> { static constexpr size_t value = N; };
}
template <class T, class... Options>
class spsc_queue {
   public:
    spsc_queue() : buf_(detail::cap_of<Options...>::value + 1), head_(0), tail_(0) {}
    explicit spsc_queue(size_t n) : buf_(n + 1), head_(0), tail_(0) {}
    bool push(const T& v) {
        const size_t t = tail_.load(std::memory_order_relaxed);
        const size_t next = (t + 1) % buf_.size();
        if (next == head_.load(std::memory_order_acquire)) return false;
        buf_[t] = v;
        tail_.store(next, std::memory_order_release);
        return true;
    }
    bool pop(T& out) {
        const size_t h = head_.load(std::memory_order_relaxed);
        if (h == tail_.load(std::memory_order_acquire)) return false;
        out = buf_[h];
        head_.store((h + 1) % buf_.size(), std::memory_order_release);
        return true;
    }
    bool empty() const { return head_.load(std::memory_order_acquire) == tail_.load(std::memory_order_acquire); }
    size_t read_available() const {
        const size_t h = head_.load(std::memory_order_acquire), t = tail_.load(std::memory_order_acquire);
        return (t + buf_.size() - h) % buf_.size();
    }
    size_t write_available() const { return buf_.size() - 1 - read_available(); }

   private:
    std::vector<T> buf_;
    std::atomic<size_t> head_, tail_;
};
}}  // namespace boost::lockfree
#endif
