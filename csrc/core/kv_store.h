// Host-side key -> block index with the reference's visibility rules.
//
// Parity (reference: src/infinistore.cpp:63-65, src/infinistore.h:30-44):
//   * a key enters the map at allocate / write-accept time, uncommitted  (C3);
//   * reads and check_exist need `committed`; get_match_last_index does not (C3);
//   * an existing key - committed or not - makes a new write a silent no-op (C4);
//   * blocks are reference counted: the pool space is returned when the last holder
//     (map entry, in-flight write, reader lease) lets go, so purge is safe while
//     transfers are in flight;
//   * get_match_last_index replays the reference's exact binary search (C7).
// Additions: per-connection ownership of uncommitted blocks so that a writer that dies
// before committing does not leave a permanently reserved key, and an
// LRU order over the committed blocks so that a full pool can evict instead of answering
// 507 until an operator purges it (the reference has no eviction).
//
// Layout.  The reference keeps `unordered_map<string, intrusive_ptr<PTR>>`: a node, a key
// string and a PTR allocation per block.  Here a block is ONE allocation (header + key
// bytes, intrusive count as in the reference's PTR) and the map is a flat open-addressing
// table of {64-bit key hash, block pointer}: allocate costs ~100 ns per key instead of ~250,
// and the commit path touches one cache line per block.
// Threading: a KVStore and every BlockPtr to its blocks are used under the server mutex only.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "hash.h"
#include "mempool.h"
#include "../wire/protocol.h"

namespace istore {

struct Block {
    uint32_t refs = 1;  // intrusive count (server mutex); the creator holds the first one
    uint32_t seg;
    MM* mm;
    uint64_t offset;
    uint32_t size;
    uint32_t gen;
    uint64_t owner;  // connection id that reserved it (0 once committed)
    uint64_t hash;   // table hash of the key
    uint32_t key_len;
    uint16_t key_off;  // key bytes start this far behind `this` (header size of the block kind)
    bool committed = false;
    bool in_lru = false;

    std::string_view key() const {
        return std::string_view(reinterpret_cast<const char*>(this) + key_off, key_len);
    }
    uint64_t addr() const { return make_addr(seg, offset); }
};

// A block of a store that evicts: the recency links live in a derived type so that the
// default configuration (no eviction, the reference's behaviour) keeps the smaller header -
// the commit path is bound by cache misses on these objects (measured: 16 more bytes per
// block cost ~5 ns per committed block, 20 % of the 4 KB write rate).
struct LruBlock : Block {
    LruBlock* lru_prev = nullptr;  // towards more recently used; linked only while committed
    LruBlock* lru_next = nullptr;  // towards less recently used
};

// Counted reference to a block; dropping the last one returns the pool space.
class BlockPtr {
   public:
    BlockPtr() = default;
    explicit BlockPtr(Block* adopt) : b_(adopt) {}  // takes over one existing count
    BlockPtr(const BlockPtr& o) : b_(o.b_) {
        if (b_) ++b_->refs;
    }
    BlockPtr(BlockPtr&& o) noexcept : b_(o.b_) { o.b_ = nullptr; }
    BlockPtr& operator=(BlockPtr o) noexcept {
        std::swap(b_, o.b_);
        return *this;
    }
    ~BlockPtr() { reset(); }
    void reset() {
        if (b_ && --b_->refs == 0) {
            b_->mm->deallocate(b_->seg, b_->offset, b_->size);
            std::free(b_);
        }
        b_ = nullptr;
    }
    Block* get() const { return b_; }
    Block* operator->() const { return b_; }
    Block& operator*() const { return *b_; }
    explicit operator bool() const { return b_ != nullptr; }
    uint32_t use_count() const { return b_ ? b_->refs : 0; }

   private:
    Block* b_ = nullptr;
};

class KVStore {
   public:
    // track_recency: keep the LRU order that evict() needs (every block is an LruBlock)
    explicit KVStore(MM* mm, bool track_recency = false);
    ~KVStore();
    KVStore(const KVStore&) = delete;
    KVStore& operator=(const KVStore&) = delete;

    // Reserve blocks for `keys`.  out[i] is the locator of key i, or the fake (0,0) block
    // when the key already exists (first writer wins).  Returns kFinish, or kOutOfMemory
    // with nothing reserved.
    int reserve(const std::vector<std::string_view>& keys, size_t size, int device_hint,
                uint64_t conn, std::vector<RemoteBlock>& out);
    // Flip the blocks at `addrs` to committed.  Unknown addresses are ignored.
    size_t commit(const uint64_t* addrs, size_t n);
    // Pull the headers of the in-flight blocks at `addrs` towards the cache (a commit for them
    // is about to arrive).  No state changes.
    void warm(const uint64_t* addrs, size_t n);
    // Locators of committed keys; kKeyNotFound if any key is missing or uncommitted.
    // `lease` receives references that keep the blocks alive until the caller drops them.
    // kInvalidReq if a stored block is smaller than `need` bytes.
    // A hit makes the block the most recently used one.
    int lookup(const std::vector<std::string_view>& keys, size_t need,
               std::vector<RemoteBlock>& out, std::vector<BlockPtr>* lease);
    bool exists_committed(std::string_view key) const;
    // Recency hint: the committed blocks among `keys` become the most recently used ones (in
    // list order).  Returns how many were refreshed; a store that does not evict returns 0.
    size_t touch(const std::vector<std::string_view>& keys);
    bool present(std::string_view key) const { return find(key) != nullptr; }
    int match_last_index(const std::vector<std::string_view>& keys) const;
    struct Victim;
    // Drop every uncommitted block reserved by `conn` (connection closed).  With `victims`
    // the blocks are handed to the caller instead of being released: a writer kernel may
    // already have claimed or published their device-index ways, which must be erased
    // BEFORE the space can be reused (same contract as evict()).
    size_t drop_uncommitted(uint64_t conn, std::vector<Victim>* victims = nullptr);
    // The same for the in-flight blocks at `addrs` that `conn` reserved (a writer reports
    // that its kernels failed: nothing of that batch may ever become visible).
    size_t drop_inflight(const uint64_t* addrs, size_t n, uint64_t conn,
                         std::vector<Victim>* victims = nullptr);
    // Empties the map.  Blocks that a reader still leases stay alive (and their space stays
    // reserved) until the lease is dropped - purge is safe while transfers are in flight.
    size_t purge();
    // Remove least-recently-used committed blocks from the map until they cover at least
    // `bytes` of pool space (rounded to granules) or none is left.  Blocks that a reader
    // still leases are skipped.  The victims are handed to the caller, which must make them
    // unreachable for device-side readers (index erase) BEFORE dropping the references -
    // dropping the last reference returns the space to the pool.
    // `replica`: take victims from the NVLS-replicated region (true) or from the ordinary
    // pools (false) - space of one kind cannot serve requests for the other.
    struct Victim {
        BlockPtr block;
        KeyHash hash;  // fingerprint of the evicted key (its device-index entry)
    };
    size_t evict(size_t bytes, bool replica, std::vector<Victim>& victims);
    uint64_t evicted() const { return evicted_; }
    size_t size() const { return count_; }
    // Visit every committed block (checkpointing).
    template <typename F>
    void for_each_committed(F&& fn) const {
        for (const Slot& s : table_)
            if (s.block && s.block->committed) fn(std::string(s.block->key()), *s.block);
    }
    size_t inflight() const { return inflight_count_; }

   private:
    // Open addressing, linear probing, backward-shift deletion.  The slot holds one count.
    struct Slot {
        uint64_t hash = 0;
        Block* block = nullptr;  // nullptr = empty
    };
    static uint64_t hash_of(std::string_view key) {
        return hash_bytes(reinterpret_cast<const uint8_t*>(key.data()), key.size(), kHashSeed1);
    }
    Block* find(std::string_view key) const { return find(key, hash_of(key)); }
    Block* find(std::string_view key, uint64_t h) const;
    void insert(Block* b);          // b->hash set; the table takes over one count
    BlockPtr remove(Block* b);      // out of the table; the table's count moves to the result
    void grow();
    // placement-construct a block (header + key bytes) in `mem`
    Block* init_block(void* mem, std::string_view key, uint64_t h, const Allocation& a,
                      size_t size, uint32_t gen, uint64_t conn) const;

    // In-flight (reserved, uncommitted) blocks are found by address in O(1): one slot per
    // allocation granule of every pool, holding the block that starts there.
    Block*& inflight_slot(uint32_t seg, uint64_t offset);
    void lru_push_front(LruBlock* b);
    void lru_unlink(LruBlock* b);

    MM* mm_;
    const bool track_lru_;
    uint32_t next_gen_ = 1;
    std::vector<Slot> table_;
    size_t count_ = 0;
    std::vector<std::vector<Block*>> inflight_;  // [segment][granule]
    size_t inflight_count_ = 0;
    LruBlock* lru_head_ = nullptr;  // most recently used
    LruBlock* lru_tail_ = nullptr;  // eviction candidate
    uint64_t evicted_ = 0;
};

}  // namespace istore
