#include "kv_store.h"

#include <algorithm>

#include "log.h"

namespace istore {

Block*& KVStore::inflight_slot(uint32_t seg, uint64_t offset) {
    if (inflight_.size() <= seg) inflight_.resize(seg + 1);
    std::vector<Block*>& v = inflight_[seg];
    const MemoryPool& pool = mm_->pool(seg);
    if (v.size() != pool.total_blocks()) v.assign(pool.total_blocks(), nullptr);
    return v[offset / pool.granule()];
}

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

int KVStore::reserve(const std::vector<std::string_view>& keys, size_t size, int device_hint,
                     uint64_t conn, std::vector<RemoteBlock>& out) {
    out.assign(keys.size(), RemoteBlock{0, 0, 0});
    // Decide about duplicates first, then allocate exactly what is needed: nothing leaks for
    // deduplicated keys.  A placeholder entry per fresh key also catches duplicates inside
    // the batch (the second occurrence finds the placeholder).
    using Iter = decltype(map_)::iterator;
    std::vector<std::pair<size_t, Iter>> fresh;
    fresh.reserve(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) {
        if (map_.find(keys[i]) != map_.end()) continue;
        fresh.emplace_back(i, map_.emplace(std::string(keys[i]), nullptr).first);
    }
    std::vector<Allocation> allocs;
    allocs.reserve(fresh.size());
    if (!mm_->allocate(size, fresh.size(), device_hint, allocs)) {
        for (auto& f : fresh) map_.erase(f.second);
        return kOutOfMemory;
    }
    for (size_t j = 0; j < fresh.size(); ++j) {
        const size_t i = fresh[j].first;
        uint32_t gen = next_gen_++;
        if (next_gen_ == 0) next_gen_ = 1;  // 0 means "not committed" in the device index
        BlockPtr blk =
            track_lru_ ? std::static_pointer_cast<Block>(std::make_shared<LruBlock>(
                             mm_, allocs[j].seg, allocs[j].offset, uint32_t(size), gen, conn))
                       : std::make_shared<Block>(mm_, allocs[j].seg, allocs[j].offset,
                                                 uint32_t(size), gen, conn);
        blk->key = &fresh[j].second->first;
        inflight_slot(allocs[j].seg, allocs[j].offset) = blk.get();
        ++inflight_count_;
        out[i] = RemoteBlock{allocs[j].seg + 1, gen, blk->addr()};
        fresh[j].second->second = std::move(blk);
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

int KVStore::lookup(const std::vector<std::string_view>& keys, size_t need,
                    std::vector<RemoteBlock>& out, std::vector<BlockPtr>* lease) {
    out.clear();
    out.reserve(keys.size());
    for (auto k : keys) {
        auto it = map_.find(k);
        if (it == map_.end() || !it->second->committed) {
            out.clear();
            return kKeyNotFound;
        }
        Block& b = *it->second;
        if (b.size < need) {  // never let a reader run past what was written
            out.clear();
            return kInvalidReq;
        }
        out.push_back(RemoteBlock{b.seg + 1, b.gen, b.addr()});
        if (lease) lease->push_back(it->second);
        if (track_lru_ && lru_head_ != &b) {
            lru_unlink(static_cast<LruBlock*>(&b));
            lru_push_front(static_cast<LruBlock*>(&b));
        }
    }
    return kFinish;
}

bool KVStore::exists_committed(std::string_view key) const {
    auto it = map_.find(key);
    return it != map_.end() && it->second->committed;
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

size_t KVStore::drop_uncommitted(uint64_t conn) {
    if (inflight_count_ == 0) return 0;
    size_t n = 0;
    for (auto& seg : inflight_) {
        for (Block*& slot : seg) {
            Block* b = slot;
            if (!b || b->owner != conn) continue;
            slot = nullptr;
            --inflight_count_;
            ++n;
            const std::string key = *b->key;  // copy: erasing frees the node that owns it
            map_.erase(key);
        }
    }
    return n;
}

size_t KVStore::evict(size_t bytes, bool replica, std::vector<Victim>& victims) {
    size_t freed = 0;
    LruBlock* b = lru_tail_;
    while (b && freed < bytes) {
        LruBlock* more_recent = b->lru_prev;
        auto it = map_.find(*b->key);
        const bool in_replica = mm_->pool(b->seg).device() == kReplicaDevice;
        if (it != map_.end() && in_replica == replica &&
            it->second.use_count() == 1) {  // nobody is reading it
            const size_t g = mm_->pool(b->seg).granule();
            freed += (size_t(b->size) + g - 1) / g * g;
            lru_unlink(b);
            const KeyHash kh =
                hash_key(reinterpret_cast<const uint8_t*>(b->key->data()), b->key->size());
            victims.push_back(Victim{std::move(it->second), kh});
            map_.erase(it);  // frees the node that owns *b->key; victims keeps the block alive
            b->key = nullptr;
            ++evicted_;
        }
        b = more_recent;
    }
    return freed;
}

size_t KVStore::purge() {
    const size_t n = map_.size();
    lru_head_ = lru_tail_ = nullptr;
    for (auto& seg : inflight_) std::fill(seg.begin(), seg.end(), nullptr);
    inflight_count_ = 0;
    map_.clear();
    return n;
}

}  // namespace istore
