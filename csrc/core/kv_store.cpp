#include "kv_store.h"

#include <algorithm>
#include <new>

#include "log.h"

namespace istore {

namespace {
constexpr size_t kInitialSlots = 1024;
}

KVStore::KVStore(MM* mm, bool track_recency) : mm_(mm), track_lru_(track_recency) {
    table_.resize(kInitialSlots);
}

KVStore::~KVStore() { purge(); }

// ---------------------------------------------------------------- flat table
Block* KVStore::find(std::string_view key, uint64_t h) const {
    const size_t mask = table_.size() - 1;
    for (size_t i = size_t(h) & mask;; i = (i + 1) & mask) {
        const Slot& s = table_[i];
        if (!s.block) return nullptr;
        if (s.hash == h && s.block->key() == key) return s.block;
    }
}

void KVStore::grow() {
    std::vector<Slot> old(table_.size() * 2);
    old.swap(table_);
    const size_t mask = table_.size() - 1;
    for (const Slot& s : old) {
        if (!s.block) continue;
        size_t i = size_t(s.hash) & mask;
        while (table_[i].block) i = (i + 1) & mask;
        table_[i] = s;
    }
}

void KVStore::insert(Block* b) {
    if ((count_ + 1) * 10 > table_.size() * 6) grow();  // load <= 0.6
    const size_t mask = table_.size() - 1;
    size_t i = size_t(b->hash) & mask;
    while (table_[i].block) i = (i + 1) & mask;
    table_[i] = Slot{b->hash, b};
    ++count_;
}

BlockPtr KVStore::remove(Block* b) {
    const size_t mask = table_.size() - 1;
    size_t i = size_t(b->hash) & mask;
    while (table_[i].block != b) {
        if (!table_[i].block) return BlockPtr();  // not in the table
        i = (i + 1) & mask;
    }
    // backward-shift deletion keeps every probe sequence intact without tombstones
    size_t hole = i;
    for (size_t j = (hole + 1) & mask; table_[j].block; j = (j + 1) & mask) {
        const size_t home = size_t(table_[j].hash) & mask;
        // may entry j move into the hole?  only if its home is not cyclically in (hole, j]
        const bool stays = hole <= j ? (home > hole && home <= j) : (home > hole || home <= j);
        if (stays) continue;
        table_[hole] = table_[j];
        hole = j;
    }
    table_[hole] = Slot{};
    --count_;
    return BlockPtr(b);  // adopts the count the table held
}

Block* KVStore::init_block(void* mem, std::string_view key, uint64_t h, const Allocation& a,
                           size_t size, uint32_t gen, uint64_t conn) const {
    const size_t header = track_lru_ ? sizeof(LruBlock) : sizeof(Block);
    Block* b = track_lru_ ? static_cast<Block*>(new (mem) LruBlock()) : new (mem) Block();
    b->seg = a.seg;
    b->mm = mm_;
    b->offset = a.offset;
    b->size = uint32_t(size);
    b->gen = gen;
    b->owner = conn;
    b->hash = h;
    b->key_len = uint32_t(key.size());
    b->key_off = uint16_t(header);
    std::memcpy(reinterpret_cast<char*>(mem) + header, key.data(), key.size());
    return b;
}

// ---------------------------------------------------------------- recency list
void KVStore::lru_push_front(LruBlock* b) {
    b->lru_prev = nullptr;
    b->lru_next = lru_head_;
    if (lru_head_) lru_head_->lru_prev = b;
    lru_head_ = b;
    if (!lru_tail_) lru_tail_ = b;
    b->in_lru = true;
}

void KVStore::lru_unlink(LruBlock* b) {
    if (!b->in_lru) return;
    (b->lru_prev ? b->lru_prev->lru_next : lru_head_) = b->lru_next;
    (b->lru_next ? b->lru_next->lru_prev : lru_tail_) = b->lru_prev;
    b->lru_prev = b->lru_next = nullptr;
    b->in_lru = false;
}

Block*& KVStore::inflight_slot(uint32_t seg, uint64_t offset) {
    if (inflight_.size() <= seg) inflight_.resize(seg + 1);
    std::vector<Block*>& v = inflight_[seg];
    const MemoryPool& pool = mm_->pool(seg);
    if (v.size() != pool.total_blocks()) v.assign(pool.total_blocks(), nullptr);
    return v[offset / pool.granule()];
}

// ---------------------------------------------------------------- store operations
int KVStore::reserve(const std::vector<std::string_view>& keys, size_t size, int device_hint,
                     uint64_t conn, std::vector<RemoteBlock>& out) {
    out.assign(keys.size(), RemoteBlock{0, 0, 0});
    // Decide about duplicates first, then allocate exactly what is needed: nothing leaks for
    // deduplicated keys.  Duplicates inside the batch are caught by comparing against the
    // fresh keys of the batch that share the hash (rare: checked linearly among equals only).
    struct Fresh {
        size_t idx;
        uint64_t hash;
    };
    std::vector<Fresh> fresh;
    fresh.reserve(keys.size());
    // small open-addressing set of the batch's own hashes -> position in `fresh`
    size_t cap = 16;
    while (cap < keys.size() * 2) cap <<= 1;
    std::vector<uint32_t> seen(cap, UINT32_MAX);
    for (size_t i = 0; i < keys.size(); ++i) {
        const uint64_t h = hash_of(keys[i]);
        if (find(keys[i], h)) continue;
        bool dup = false;
        size_t s = size_t(h) & (cap - 1);
        for (; seen[s] != UINT32_MAX; s = (s + 1) & (cap - 1)) {
            const Fresh& f = fresh[seen[s]];
            if (f.hash == h && keys[f.idx] == keys[i]) {
                dup = true;
                break;
            }
        }
        if (dup) continue;
        seen[s] = uint32_t(fresh.size());
        fresh.push_back(Fresh{i, h});
    }
    // Everything that can fail for lack of HOST memory happens before any state changes: the
    // table is grown, the block headers are allocated, the in-flight slot arrays are sized.
    // After that the batch cannot fail half way (all or nothing).
    try {
        while ((count_ + fresh.size()) * 10 > table_.size() * 6) grow();
    } catch (const std::bad_alloc&) {
        return kOutOfMemory;
    }
    const size_t header = track_lru_ ? sizeof(LruBlock) : sizeof(Block);
    std::vector<void*> mem(fresh.size(), nullptr);
    auto free_mem = [&] {
        for (void* m : mem) std::free(m);
    };
    for (size_t j = 0; j < fresh.size(); ++j) {
        mem[j] = std::malloc(header + keys[fresh[j].idx].size());
        if (!mem[j]) {
            free_mem();
            return kOutOfMemory;
        }
    }
    std::vector<Allocation> allocs;
    allocs.reserve(fresh.size());
    if (!mm_->allocate(size, fresh.size(), device_hint, allocs)) {
        free_mem();
        return kOutOfMemory;
    }
    try {
        for (const Allocation& a : allocs) (void)inflight_slot(a.seg, a.offset);
    } catch (const std::bad_alloc&) {
        for (const Allocation& a : allocs) mm_->deallocate(a.seg, a.offset, size);
        free_mem();
        return kOutOfMemory;
    }
    for (size_t j = 0; j < fresh.size(); ++j) {
        const size_t i = fresh[j].idx;
        uint32_t gen = next_gen_++;
        if (next_gen_ == 0) next_gen_ = 1;  // 0 means "not committed" in the device index
        Block* blk = init_block(mem[j], keys[i], fresh[j].hash, allocs[j], size, gen, conn);
        inflight_slot(allocs[j].seg, allocs[j].offset) = blk;
        ++inflight_count_;
        out[i] = RemoteBlock{allocs[j].seg + 1, gen, blk->addr()};
        insert(blk);  // the table owns the creator's count; capacity was ensured above
    }
    return kFinish;
}

size_t KVStore::commit(const uint64_t* addrs, size_t n) {
    size_t done = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t seg = addr_seg(addrs[i]);
        const uint64_t off = addr_off(addrs[i]);
        if (seg >= mm_->num_pools() || off >= mm_->pool(seg).bytes()) continue;
        Block*& slot = inflight_slot(seg, off);
        Block* b = slot;
        if (!b || b->offset != off) continue;  // unknown / already committed: ignored
        b->committed = true;
        b->owner = 0;
        if (track_lru_) lru_push_front(static_cast<LruBlock*>(b));
        slot = nullptr;
        --inflight_count_;
        ++done;
    }
    return done;
}

void KVStore::warm(const uint64_t* addrs, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const uint32_t seg = addr_seg(addrs[i]);
        const uint64_t off = addr_off(addrs[i]);
        if (seg >= mm_->num_pools() || off >= mm_->pool(seg).bytes()) continue;
        if (const Block* b = inflight_slot(seg, off)) __builtin_prefetch(b, 1);
    }
}

int KVStore::lookup(const std::vector<std::string_view>& keys, size_t need,
                    std::vector<RemoteBlock>& out, std::vector<BlockPtr>* lease) {
    out.clear();
    out.reserve(keys.size());
    const size_t lease_mark = lease ? lease->size() : 0;
    for (auto k : keys) {
        Block* b = find(k);
        int code = kFinish;
        if (!b || !b->committed)
            code = kKeyNotFound;
        else if (b->size < need)  // never let a reader run past what was written
            code = kInvalidReq;
        if (code != kFinish) {
            out.clear();
            if (lease) lease->resize(lease_mark);
            return code;
        }
        out.push_back(RemoteBlock{b->seg + 1, b->gen, b->addr()});
        if (lease) {
            ++b->refs;
            lease->push_back(BlockPtr(b));
        }
        if (track_lru_ && lru_head_ != b) {
            lru_unlink(static_cast<LruBlock*>(b));
            lru_push_front(static_cast<LruBlock*>(b));
        }
    }
    return kFinish;
}

bool KVStore::exists_committed(std::string_view key) const {
    const Block* b = find(key);
    return b && b->committed;
}

size_t KVStore::touch(const std::vector<std::string_view>& keys) {
    if (!track_lru_) return 0;
    size_t n = 0;
    for (auto k : keys) {
        Block* b = find(k);
        if (!b || !b->committed) continue;
        ++n;
        if (lru_head_ == b) continue;
        lru_unlink(static_cast<LruBlock*>(b));
        lru_push_front(static_cast<LruBlock*>(b));
    }
    return n;
}

// Exact replay of the reference's search (src/infinistore.cpp:1092-1108): presence is
// assumed prefix-monotone; on other inputs the answer is whatever this probe sequence
// yields, and callers depend on that.
int KVStore::match_last_index(const std::vector<std::string_view>& keys) const {
    int left = 0, right = int(keys.size());
    while (left < right) {
        const int mid = left + (right - left) / 2;
        if (present(keys[size_t(mid)]))
            left = mid + 1;
        else
            right = mid;
    }
    return left - 1;
}

size_t KVStore::drop_uncommitted(uint64_t conn, std::vector<Victim>* victims) {
    if (inflight_count_ == 0) return 0;
    size_t n = 0;
    for (auto& seg : inflight_) {
        for (Block*& slot : seg) {
            Block* b = slot;
            if (!b || b->owner != conn) continue;
            slot = nullptr;
            --inflight_count_;
            ++n;
            if (victims) {
                // the writer's kernel may have claimed (or even published) a way of the device
                // index for this block: the caller erases it before the space is reused
                const std::string_view key = b->key();
                const KeyHash kh = hash_key(reinterpret_cast<const uint8_t*>(key.data()), key.size());
                victims->push_back(Victim{remove(b), kh});
            } else {
                remove(b);  // the returned reference dies here: space back to the pool
            }
        }
    }
    return n;
}

size_t KVStore::drop_inflight(const uint64_t* addrs, size_t n, uint64_t conn,
                              std::vector<Victim>* victims) {
    size_t done = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t seg = addr_seg(addrs[i]);
        const uint64_t off = addr_off(addrs[i]);
        if (seg >= mm_->num_pools() || off >= mm_->pool(seg).bytes()) continue;
        Block*& slot = inflight_slot(seg, off);
        Block* b = slot;
        if (!b || b->offset != off || b->owner != conn) continue;
        slot = nullptr;
        --inflight_count_;
        ++done;
        if (victims) {
            const std::string_view key = b->key();
            const KeyHash kh = hash_key(reinterpret_cast<const uint8_t*>(key.data()), key.size());
            victims->push_back(Victim{remove(b), kh});
        } else {
            remove(b);
        }
    }
    return done;
}

size_t KVStore::evict(size_t bytes, bool replica, std::vector<Victim>& victims) {
    size_t freed = 0;
    LruBlock* b = lru_tail_;
    while (b && freed < bytes) {
        LruBlock* more_recent = b->lru_prev;
        const bool in_replica = mm_->pool(b->seg).device() == kReplicaDevice;
        if (in_replica == replica && b->refs == 1) {  // only the table holds it: nobody reads
            const size_t g = mm_->pool(b->seg).granule();
            freed += (size_t(b->size) + g - 1) / g * g;
            lru_unlink(b);
            const std::string_view key = b->key();
            const KeyHash kh = hash_key(reinterpret_cast<const uint8_t*>(key.data()), key.size());
            victims.push_back(Victim{remove(b), kh});
            ++evicted_;
        }
        b = more_recent;
    }
    return freed;
}

size_t KVStore::purge() {
    const size_t n = count_;
    for (auto& seg : inflight_) std::fill(seg.begin(), seg.end(), nullptr);
    inflight_count_ = 0;
    lru_head_ = lru_tail_ = nullptr;
    for (Slot& s : table_) {
        if (!s.block) continue;
        s.block->in_lru = false;  // a leased block outlives the list it was linked in
        BlockPtr drop(s.block);   // the table's count
        s = Slot{};
    }
    count_ = 0;
    return n;
}

}  // namespace istore
