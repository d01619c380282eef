#include "kv_store.h"

#include "log.h"

namespace istore {

int KVStore::reserve(const std::vector<std::string_view>& keys, size_t size, int device_hint,
                     uint64_t conn, std::vector<RemoteBlock>& out) {
    out.assign(keys.size(), RemoteBlock{0, 0, 0});
    // Decide about duplicates first (also duplicates inside the batch), then allocate
    // exactly what is needed: nothing leaks for deduplicated keys.
    std::vector<size_t> fresh;
    fresh.reserve(keys.size());
    {
        std::unordered_map<std::string_view, size_t, StrHash, StrEq> seen;
        seen.reserve(keys.size());
        for (size_t i = 0; i < keys.size(); ++i) {
            if (map_.find(keys[i]) != map_.end()) continue;
            if (!seen.emplace(keys[i], i).second) continue;
            fresh.push_back(i);
        }
    }
    std::vector<Allocation> allocs;
    allocs.reserve(fresh.size());
    if (!mm_->allocate(size, fresh.size(), device_hint, allocs)) return kOutOfMemory;
    for (size_t j = 0; j < fresh.size(); ++j) {
        const size_t i = fresh[j];
        uint32_t gen = next_gen_++;
        if (next_gen_ == 0) next_gen_ = 1;  // 0 means "not committed" in the device index
        auto blk = std::make_shared<Block>(mm_, allocs[j].seg, allocs[j].offset, uint32_t(size),
                                           gen, conn);
        auto it = map_.emplace(std::string(keys[i]), blk).first;
        inflight_[blk->addr()] = {blk, &it->first};
        out[i] = RemoteBlock{allocs[j].seg + 1, gen, blk->addr()};
    }
    return kFinish;
}

size_t KVStore::commit(const uint64_t* addrs, size_t n) {
    size_t done = 0;
    for (size_t i = 0; i < n; ++i) {
        auto it = inflight_.find(addrs[i]);
        if (it == inflight_.end()) continue;
        it->second.first->committed = true;
        it->second.first->owner = 0;
        inflight_.erase(it);
        ++done;
    }
    return done;
}

int KVStore::lookup(const std::vector<std::string_view>& keys, size_t need,
                    std::vector<RemoteBlock>& out, std::vector<BlockPtr>* lease) const {
    out.clear();
    out.reserve(keys.size());
    for (auto k : keys) {
        auto it = map_.find(k);
        if (it == map_.end() || !it->second->committed) {
            out.clear();
            return kKeyNotFound;
        }
        const Block& b = *it->second;
        if (b.size < need) {  // never let a reader run past what was written
            out.clear();
            return kInvalidReq;
        }
        out.push_back(RemoteBlock{b.seg + 1, b.gen, b.addr()});
        if (lease) lease->push_back(it->second);
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
    size_t n = 0;
    for (auto it = inflight_.begin(); it != inflight_.end();) {
        if (it->second.first->owner == conn) {
            const std::string key = *it->second.second;
            it = inflight_.erase(it);
            map_.erase(key);
            ++n;
        } else {
            ++it;
        }
    }
    return n;
}

size_t KVStore::purge() {
    const size_t n = map_.size();
    inflight_.clear();
    map_.clear();
    return n;
}

}  // namespace istore
