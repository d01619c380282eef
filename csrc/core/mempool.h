// Bitmap block allocator + multi-pool manager.
//
// Capability parity with the reference's MemoryPool / MM (src/mempool.h:21-84,
// src/mempool.cpp:57-192): fixed-granule pools, first-fit search for a contiguous run of
// ceil(size/granule) granules starting at a moving hint, n objects per call, double-free
// detection, a "last pool more than half full" extension trigger.
//
// Deliberate differences:
//   * the pool only does accounting over byte offsets; the memory itself (HBM segment or
//     shm) is owned by fabric::SegmentOwner, so the same allocator serves both backends;
//   * n-object allocation is all-or-nothing and the caller decides about duplicates
//     BEFORE allocating, which removes the reference's leak of deduplicated slots
//     (its bitmap is marked before the callback can reject the key) and its
//     half-applied batches on out-of-memory;
//   * the bitmap holds exactly ceil(blocks/64) words.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace istore {

class MemoryPool {
   public:
    MemoryPool(size_t pool_bytes, size_t granule, int device);

    // Offset of a contiguous run able to hold `size` bytes, or -1 when none is free.
    int64_t allocate(size_t size);
    // All-or-nothing allocation of n equally sized objects; appends offsets to `out`.
    bool allocate_n(size_t size, size_t n, std::vector<uint64_t>& out);
    // Returns false (and changes nothing) on a double free / out-of-range request.
    bool deallocate(uint64_t offset, size_t size);

    size_t granule() const { return granule_; }
    size_t total_blocks() const { return total_blocks_; }
    size_t used_blocks() const { return used_blocks_; }
    size_t free_blocks() const { return total_blocks_ - used_blocks_; }
    size_t bytes() const { return total_blocks_ * granule_; }
    int device() const { return device_; }
    double usage() const { return total_blocks_ ? double(used_blocks_) / total_blocks_ : 1.0; }

   private:
    size_t blocks_for(size_t size) const { return (size + granule_ - 1) / granule_; }
    int64_t find_run(size_t k);
    void mark(size_t first, size_t k, bool used);
    bool all_used(size_t first, size_t k) const;

    size_t granule_;
    size_t total_blocks_;
    size_t used_blocks_ = 0;
    size_t hint_ = 0;
    int device_;
    std::vector<uint64_t> bitmap_;  // bit set = granule in use
};

struct Allocation {
    uint32_t seg;
    uint64_t offset;
};

// Pool manager.  Not thread safe: owned by the reactor thread.
class MM {
   public:
    static constexpr double kExtendUsageRatio = 0.5;  // reference: BLOCK_USAGE_RATIO

    uint32_t add_pool(size_t pool_bytes, size_t granule, int device);
    // All-or-nothing allocation of n objects of `size` bytes.  Pools whose device equals
    // `device_hint` are tried first (hint < 0: creation order).
    bool allocate(size_t size, size_t n, int device_hint, std::vector<Allocation>& out);
    bool deallocate(uint32_t seg, uint64_t offset, size_t size);

    bool need_extend() const;
    size_t num_pools() const { return pools_.size(); }
    const MemoryPool& pool(size_t i) const { return *pools_[i]; }
    size_t used_bytes() const;
    size_t total_bytes() const;

   private:
    std::vector<std::unique_ptr<MemoryPool>> pools_;
};

}  // namespace istore
