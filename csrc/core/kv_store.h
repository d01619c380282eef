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
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "hash.h"
#include "mempool.h"
#include "../wire/protocol.h"

namespace istore {

struct Block {
    MM* mm;
    uint32_t seg;
    uint64_t offset;
    uint32_t size;
    uint32_t gen;
    bool committed = false;
    uint64_t owner = 0;  // connection id that reserved it (0 once committed)
    const std::string* key = nullptr;  // the map node's key (node addresses are stable)
    Block(MM* m, uint32_t s, uint64_t o, uint32_t sz, uint32_t g, uint64_t own)
        : mm(m), seg(s), offset(o), size(sz), gen(g), owner(own) {}
    ~Block() { mm->deallocate(seg, offset, size); }
    Block(const Block&) = delete;
    Block& operator=(const Block&) = delete;
    uint64_t addr() const { return make_addr(seg, offset); }
};
using BlockPtr = std::shared_ptr<Block>;

// A block of a store that evicts: the recency links live in a derived type so that the
// default configuration (no eviction, the reference's behaviour) keeps the smaller block -
// the commit path is bound by cache misses on these objects (measured: 16 more bytes per
// block cost ~5 ns per committed block, 20 % of the 4 KB write rate).
struct LruBlock : Block {
    using Block::Block;
    LruBlock* lru_prev = nullptr;  // towards more recently used; linked only while committed
    LruBlock* lru_next = nullptr;  // towards less recently used
    bool in_lru = false;
};

struct StrHash {
    using is_transparent = void;
    size_t operator()(std::string_view s) const {
        return size_t(hash_bytes(reinterpret_cast<const uint8_t*>(s.data()), s.size(), kHashSeed1));
    }
    size_t operator()(const std::string& s) const { return (*this)(std::string_view(s)); }
};
struct StrEq {
    using is_transparent = void;
    bool operator()(std::string_view a, std::string_view b) const { return a == b; }
};

class KVStore {
   public:
    // track_recency: keep the LRU order that evict() needs (every block is an LruBlock)
    explicit KVStore(MM* mm, bool track_recency = false) : mm_(mm), track_lru_(track_recency) {}

    // Reserve blocks for `keys`.  out[i] is the locator of key i, or the fake (0,0) block
    // when the key already exists (first writer wins).  Returns kFinish, or kOutOfMemory
    // with nothing reserved.
    int reserve(const std::vector<std::string_view>& keys, size_t size, int device_hint,
                uint64_t conn, std::vector<RemoteBlock>& out);
    // Flip the blocks at `addrs` to committed.  Unknown addresses are ignored.
    size_t commit(const uint64_t* addrs, size_t n);
    // Locators of committed keys; kKeyNotFound if any key is missing or uncommitted.
    // `lease` receives references that keep the blocks alive until the caller drops them.
    // kInvalidReq if a stored block is smaller than `need` bytes.
    // A hit makes the block the most recently used one.
    int lookup(const std::vector<std::string_view>& keys, size_t need,
               std::vector<RemoteBlock>& out, std::vector<BlockPtr>* lease);
    bool exists_committed(std::string_view key) const;
    bool present(std::string_view key) const { return map_.find(key) != map_.end(); }
    int match_last_index(const std::vector<std::string_view>& keys) const;
    // Drop every uncommitted block reserved by `conn` (connection closed).
    size_t drop_uncommitted(uint64_t conn);
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
    size_t size() const { return map_.size(); }
    // Visit every committed block (checkpointing).
    template <typename F>
    void for_each_committed(F&& fn) const {
        for (auto& kv : map_)
            if (kv.second && kv.second->committed) fn(kv.first, *kv.second);
    }
    size_t inflight() const { return inflight_count_; }

   private:
    // In-flight (reserved, uncommitted) blocks are found by address in O(1): one slot per
    // allocation granule of every pool, holding the block that starts there.
    Block*& inflight_slot(uint32_t seg, uint64_t offset);
    void lru_push_front(LruBlock* b);
    void lru_unlink(LruBlock* b);

    MM* mm_;
    uint32_t next_gen_ = 1;
    std::unordered_map<std::string, BlockPtr, StrHash, StrEq> map_;
    std::vector<std::vector<Block*>> inflight_;  // [segment][granule]
    size_t inflight_count_ = 0;
    LruBlock* lru_head_ = nullptr;  // most recently used
    LruBlock* lru_tail_ = nullptr;  // eviction candidate
    uint64_t evicted_ = 0;
    const bool track_lru_;
};

}  // namespace istore
