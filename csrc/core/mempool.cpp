#include "mempool.h"

#include <algorithm>

#include "log.h"

namespace istore {

MemoryPool::MemoryPool(size_t pool_bytes, size_t granule, int device)
    : granule_(granule), total_blocks_(pool_bytes / granule), device_(device) {
    bitmap_.assign((total_blocks_ + 63) / 64, 0);
    // bits past the end are permanently "used" so that searches never hand them out
    const size_t tail = total_blocks_ & 63;
    if (tail) bitmap_.back() = ~uint64_t(0) << tail;
}

void MemoryPool::mark(size_t first, size_t k, bool used) {
    size_t i = first;
    const size_t end = first + k;
    while (i < end) {
        const size_t w = i >> 6, b = i & 63;
        const size_t n = std::min<size_t>(64 - b, end - i);
        const uint64_t mask = (n == 64 ? ~uint64_t(0) : ((uint64_t(1) << n) - 1)) << b;
        if (used)
            bitmap_[w] |= mask;
        else
            bitmap_[w] &= ~mask;
        i += n;
    }
}

bool MemoryPool::all_used(size_t first, size_t k) const {
    size_t i = first;
    const size_t end = first + k;
    while (i < end) {
        const size_t w = i >> 6, b = i & 63;
        const size_t n = std::min<size_t>(64 - b, end - i);
        const uint64_t mask = (n == 64 ? ~uint64_t(0) : ((uint64_t(1) << n) - 1)) << b;
        if ((bitmap_[w] & mask) != mask) return false;
        i += n;
    }
    return true;
}

// First fit from the hint, wrapping once.  Words that are completely full are skipped 64
// granules at a time; inside a word the next free granule is found with ctz.
int64_t MemoryPool::find_run(size_t k) {
    if (k == 0 || k > total_blocks_ - used_blocks_) return -1;
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0: runs inside [hint, total); pass 1: runs that start below the hint
        size_t pos = pass == 0 ? hint_ : 0;
        const size_t end = pass == 0 ? total_blocks_ : std::min(total_blocks_, hint_ + k - 1);
        while (pos + k <= end) {
            const size_t w = pos >> 6;
            const uint64_t freebits = ~bitmap_[w] & (~uint64_t(0) << (pos & 63));
            if (!freebits) {
                pos = (w + 1) << 6;
                continue;
            }
            const size_t start = (w << 6) + size_t(__builtin_ctzll(freebits));
            if (start + k > end) break;
            // measure the free run beginning at `start`
            size_t run = 0, p = start;
            while (run < k) {
                const size_t pb = p & 63;
                const uint64_t used = bitmap_[p >> 6] >> pb;
                if (used == 0) {
                    run += 64 - pb;
                    p += 64 - pb;
                } else {
                    const size_t z = size_t(__builtin_ctzll(used));
                    run += z;
                    p += z;
                    break;
                }
            }
            if (run >= k) return int64_t(start);
            pos = p + 1;  // p is a used granule: restart after it
        }
    }
    return -1;
}

int64_t MemoryPool::allocate(size_t size) {
    const size_t k = blocks_for(size);
    const int64_t first = find_run(k);
    if (first < 0) return -1;
    mark(size_t(first), k, true);
    used_blocks_ += k;
    hint_ = size_t(first) + k;
    if (hint_ >= total_blocks_) hint_ = 0;
    return first * int64_t(granule_);
}

bool MemoryPool::allocate_n(size_t size, size_t n, std::vector<uint64_t>& out) {
    const size_t k = blocks_for(size);
    if (k == 0 || n * k > free_blocks()) return false;
    const size_t mark_at = out.size();
    const size_t saved_hint = hint_;
    for (size_t i = 0; i < n; ++i) {
        const int64_t off = allocate(size);
        if (off < 0) {  // fragmentation: roll back
            for (size_t j = mark_at; j < out.size(); ++j) deallocate(out[j], size);
            out.resize(mark_at);
            hint_ = saved_hint;
            return false;
        }
        out.push_back(uint64_t(off));
    }
    return true;
}

bool MemoryPool::deallocate(uint64_t offset, size_t size) {
    const size_t k = blocks_for(size);
    if (offset % granule_ != 0) return false;
    const size_t first = offset / granule_;
    if (k == 0 || first + k > total_blocks_) return false;
    if (!all_used(first, k)) {
        LOG_ERROR("double free or corrupt free: offset=%llu size=%zu",
                  (unsigned long long)offset, size);
        return false;
    }
    mark(first, k, false);
    used_blocks_ -= k;
    if (first < hint_) hint_ = first;  // reuse low addresses first
    return true;
}

uint32_t MM::add_pool(size_t pool_bytes, size_t granule, int device) {
    pools_.push_back(std::make_unique<MemoryPool>(pool_bytes, granule, device));
    return uint32_t(pools_.size() - 1);
}

bool MM::allocate(size_t size, size_t n, int device_hint, std::vector<Allocation>& out) {
    if (n == 0) return true;
    const size_t mark_at = out.size();
    size_t remaining = n;
    // pools on the hinted device first, then the rest in creation order
    // The NVLS-replicated region (device -2) is a separate address space: it serves exactly
    // the requests that ask for it and is never used as overflow for ordinary blocks.
    constexpr int kReplica = -2;
    std::vector<size_t> order;
    order.reserve(pools_.size());
    if (device_hint == kReplica) {
        for (size_t p = 0; p < pools_.size(); ++p)
            if (pools_[p]->device() == kReplica) order.push_back(p);
    } else {
        for (size_t p = 0; p < pools_.size(); ++p)
            if (device_hint >= 0 && pools_[p]->device() == device_hint) order.push_back(p);
        for (size_t p = 0; p < pools_.size(); ++p)
            if (!(device_hint >= 0 && pools_[p]->device() == device_hint) &&
                pools_[p]->device() != kReplica)
                order.push_back(p);
    }

    std::vector<uint64_t> offs;
    for (size_t p : order) {
        if (!remaining) break;
        MemoryPool& pool = *pools_[p];
        const size_t k = (size + pool.granule() - 1) / pool.granule();
        size_t take = std::min(remaining, k ? pool.free_blocks() / k : 0);
        // fragmentation can defeat the estimate: halve until it fits
        while (take) {
            offs.clear();
            if (pool.allocate_n(size, take, offs)) break;
            take /= 2;
        }
        if (!take) continue;
        for (uint64_t o : offs) out.push_back(Allocation{uint32_t(p), o});
        remaining -= take;
    }
    if (remaining) {
        for (size_t j = mark_at; j < out.size(); ++j)
            pools_[out[j].seg]->deallocate(out[j].offset, size);
        out.resize(mark_at);
        return false;
    }
    return true;
}

bool MM::deallocate(uint32_t seg, uint64_t offset, size_t size) {
    if (seg >= pools_.size()) return false;
    return pools_[seg]->deallocate(offset, size);
}

bool MM::need_extend() const {
    for (size_t p = pools_.size(); p > 0; --p)  // last ordinary pool
        if (pools_[p - 1]->device() != -2) return pools_[p - 1]->usage() > kExtendUsageRatio;
    return false;
}

size_t MM::used_bytes() const {
    size_t n = 0;
    for (auto& p : pools_) n += p->used_blocks() * p->granule();
    return n;
}

size_t MM::total_bytes() const {
    size_t n = 0;
    for (auto& p : pools_) n += p->bytes();
    return n;
}

}  // namespace istore
