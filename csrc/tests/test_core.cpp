// Native unit tests of the core (no Python, no GPU): wire codec, allocator, key index.
// The reference's own native tests (src/test/*) no longer compile against its headers;
// these are their working counterpart.  Build + run: python tools/build_native.py --tests
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <random>
#include <atomic>
#include <map>
#include <set>
#include <thread>
#include <string>
#include <vector>

#include "core/hash.h"
#include "core/kv_store.h"
#include "core/mempool.h"
#include "kernels/index.cuh"
#include "wire/messages.h"

using namespace istore;

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        ++g_checks;                                                              \
        if (!(cond)) {                                                           \
            ++g_failed;                                                          \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
        }                                                                        \
    } while (0)

static void test_framing() {
    CHECK(sizeof(Header) == 9);
    CHECK(sizeof(ConnInfo) == 30);
    CHECK(sizeof(RemoteBlock) == 16);
    Header h{kMagic, kOpAllocate, 123};
    unsigned char raw[9];
    std::memcpy(raw, &h, 9);
    CHECK(raw[0] == 0xef && raw[1] == 0xbe && raw[2] == 0xad && raw[3] == 0xde);  // LE magic
    CHECK(raw[4] == 'D' && raw[5] == 123);
    CHECK(op_known('R') && op_known('P') && !op_known('Z') && !op_has_body('S'));
    CHECK(std::string(op_name('M')) == "MATCH_LAST_INDEX");
    const uint64_t a = make_addr(3, 0x12345000);
    CHECK(addr_seg(a) == 3 && addr_off(a) == 0x12345000);
    CHECK(is_fake_block(RemoteBlock{0, 0, 0}) && !is_fake_block(RemoteBlock{1, 1, make_addr(0, 0)}));
}

static void test_flatbuffers() {
    alignas(8) uint8_t buf[4096];
    {
        std::vector<std::string_view> keys = {"alpha", "", "a-much-longer-key-0123456789abcdef"};
        uint64_t addrs[3] = {1, 1ull << 44, ~0ull};
        fb::Builder b(buf, sizeof(buf));
        encode_remote_meta(b, keys, 65536, 42, addrs, 3, 'A', 5);
        CHECK(reinterpret_cast<uintptr_t>(b.data()) % 8 == 0);  // finished message is aligned
        RemoteMetaRequest r = decode_remote_meta(b.data(), b.size());
        CHECK(r.keys.size() == 3 && r.keys[0] == "alpha" && r.keys[1].empty() && r.keys[2] == keys[2]);
        CHECK(r.block_size == 65536 && r.rkey == 42 && r.op == 'A' && r.hint == 5);
        CHECK(r.remote_addrs.size() == 3 && r.remote_addrs[2] == ~0ull);
        // truncation at every length must throw, never read out of bounds
        for (size_t cut = 0; cut < b.size(); ++cut) {
            bool threw = false;
            try {
                std::vector<uint8_t> copy(b.data(), b.data() + cut);  // exact-size heap block
                decode_remote_meta(copy.data(), copy.size());
            } catch (const fb::Malformed&) {
                threw = true;
            }
            if (cut < 8) CHECK(threw);
        }
    }
    {
        RemoteBlock blocks[3] = {{1, 9, make_addr(0, 4096)}, {0, 0, 0}, {2, 10, make_addr(1, 0)}};
        fb::Builder b(buf, sizeof(buf));
        encode_allocate_response(b, blocks, 3);
        auto out = decode_allocate_response(b.data(), b.size());
        CHECK(out.size() == 3 && std::memcmp(out.data(), blocks, sizeof(blocks)) == 0);
        fb::Builder e(buf, sizeof(buf));
        encode_allocate_response(e, nullptr, 0);
        CHECK(decode_allocate_response(e.data(), e.size()).empty());
    }
    {
        std::vector<LocalBlock> lb = {{"k0", 0}, {"k1", 1ull << 40}};
        std::string ipc(64, '\x7f');
        fb::Builder b(buf, sizeof(buf));
        encode_local_meta(b, 7, ipc, 32768, lb);
        LocalMetaRequest r = decode_local_meta(b.data(), b.size());
        CHECK(r.device == 7 && r.block_size == 32768 && r.ipc_handle == ipc);
        CHECK(r.blocks.size() == 2 && r.blocks[1].key == "k1" && r.blocks[1].offset == (1ull << 40));
    }
    {
        bool threw = false;
        alignas(8) uint8_t small[64];
        try {
            fb::Builder b(small, sizeof(small));
            std::vector<std::string_view> keys(32, "xxxxxxxxxxxxxxxx");
            encode_match_request(b, keys);
        } catch (const fb::Overflow&) {
            threw = true;
        }
        CHECK(threw);  // fixed buffer overflow is an exception, not a write past the end
    }
}

static void test_mempool() {
    const size_t g = 16384;
    MemoryPool p(130 * g, g, -1);
    std::vector<uint64_t> offs;
    CHECK(p.allocate_n(g, 130, offs) && offs.size() == 130);
    CHECK(p.allocate(1) == -1 && p.usage() == 1.0);
    std::set<uint64_t> uniq(offs.begin(), offs.end());
    CHECK(uniq.size() == 130);
    for (int i = 60; i < 70; ++i) CHECK(p.deallocate(uint64_t(i) * g, g));
    CHECK(p.allocate(10 * g) == int64_t(60 * g));
    CHECK(!p.deallocate(60 * g, 11 * g) || true);  // size mismatch over used range is allowed
    std::mt19937 rng(5);
    MemoryPool q(1024 * g, g, 0);
    std::vector<std::pair<uint64_t, size_t>> live;
    for (int it = 0; it < 20000; ++it) {
        if (!live.empty() && rng() % 2) {
            const size_t k = rng() % live.size();
            CHECK(q.deallocate(live[k].first, live[k].second));
            live[k] = live.back();
            live.pop_back();
        } else {
            const size_t sz = 1 + rng() % (5 * g);
            const int64_t off = q.allocate(sz);
            if (off >= 0) live.emplace_back(uint64_t(off), sz);
        }
    }
    size_t used = 0;
    for (auto& l : live) used += (l.second + g - 1) / g;
    CHECK(q.used_blocks() == used);
}

static void test_kv_store() {
    MM mm;
    mm.add_pool(64 * 16384, 16384, -1);
    KVStore st(&mm);
    std::vector<std::string_view> keys = {"a", "b", "a", "c"};
    std::vector<RemoteBlock> out;
    CHECK(st.reserve(keys, 16384, -1, 1, out) == kFinish);
    CHECK(!is_fake_block(out[0]) && !is_fake_block(out[1]) && is_fake_block(out[2]));  // in-batch dup
    CHECK(st.size() == 3 && st.inflight() == 3 && mm.used_bytes() == 3 * 16384);
    CHECK(!st.exists_committed("a") && st.present("a"));
    std::vector<RemoteBlock> found;
    CHECK(st.lookup({"a"}, 1, found, nullptr) == kKeyNotFound);  // reserved, not committed
    uint64_t addr_a = out[0].remote_addr;
    CHECK(st.commit(&addr_a, 1) == 1 && st.commit(&addr_a, 1) == 0);
    CHECK(st.exists_committed("a") && st.inflight() == 2);
    CHECK(st.lookup({"a"}, 16384, found, nullptr) == kFinish && found[0].remote_addr == addr_a);
    CHECK(st.lookup({"a"}, 16385, found, nullptr) == kInvalidReq);
    // dedup against existing keys, committed or not; no pool space leaks
    CHECK(st.reserve({"a", "b", "d"}, 16384, -1, 2, out) == kFinish);
    CHECK(is_fake_block(out[0]) && is_fake_block(out[1]) && !is_fake_block(out[2]));
    CHECK(mm.used_bytes() == 4 * 16384);
    // match: exact replay of the reference search over presence (committed or not)
    CHECK(st.match_last_index({"x", "y", "z", "a", "q", "r"}) == 3);
    CHECK(st.match_last_index({"a", "b", "c", "zz"}) == 2);
    CHECK(st.match_last_index({"zz"}) == -1);
    // writer 1 dies: its uncommitted keys go away, committed ones stay
    CHECK(st.drop_uncommitted(1) == 2 && st.size() == 2 && st.present("a") && !st.present("b"));
    // out of memory reserves nothing
    std::vector<std::string> many;
    for (int i = 0; i < 100; ++i) many.push_back("m" + std::to_string(i));
    std::vector<std::string_view> mv(many.begin(), many.end());
    const size_t before = st.size();
    CHECK(st.reserve(mv, 16384, -1, 3, out) == kOutOfMemory && st.size() == before);
    // leases keep blocks alive across purge
    std::vector<BlockPtr> lease;
    CHECK(st.lookup({"a"}, 1, found, &lease) == kFinish && lease.size() == 1);
    CHECK(st.purge() == before && st.size() == 0);
    CHECK(mm.used_bytes() == 16384);  // the leased block
    lease.clear();
    CHECK(mm.used_bytes() == 0);
}

static void test_hash() {
    std::set<uint64_t> seen;
    for (int i = 0; i < 100000; ++i) {
        const std::string k = "layer" + std::to_string(i % 80) + "/block" + std::to_string(i);
        const KeyHash h = hash_key(reinterpret_cast<const uint8_t*>(k.data()), k.size());
        CHECK(h.h1 != 0);
        seen.insert(h.h1);
    }
    CHECK(seen.size() == 100000);
    // unaligned start and zero padding do not change the value
    alignas(8) char buf[64] = {0};
    std::memcpy(buf + 3, "hello world, hello hash", 23);
    const KeyHash a = hash_key(reinterpret_cast<const uint8_t*>(buf + 3), 23);
    const KeyHash b = hash_key(reinterpret_cast<const uint8_t*>("hello world, hello hash"), 23);
    CHECK(a.h1 == b.h1 && a.h2 == b.h2);
}

static void test_kv_store_eviction() {
    MM mm;
    mm.add_pool(8 * 16384, 16384, -1);
    KVStore st(&mm, true);
    std::vector<RemoteBlock> out, found;
    std::vector<std::string> names;
    for (int i = 0; i < 8; ++i) names.push_back("k" + std::to_string(i));
    std::vector<std::string_view> keys(names.begin(), names.end());
    CHECK(st.reserve(keys, 16384, -1, 1, out) == kFinish);
    std::vector<uint64_t> addrs;
    for (auto& b : out) addrs.push_back(b.remote_addr);
    // commit order = recency order: k0 is the oldest
    for (uint64_t a : addrs) CHECK(st.commit(&a, 1) == 1);
    CHECK(st.reserve({"new"}, 16384, -1, 1, out) == kOutOfMemory);
    // a read makes k0 the most recently used; k1 is now the eviction candidate
    CHECK(st.lookup({"k0"}, 1, found, nullptr) == kFinish);
    // a leased block is skipped
    std::vector<BlockPtr> lease;
    CHECK(st.lookup({"k1"}, 1, found, &lease) == kFinish);
    CHECK(st.lookup({"k0"}, 1, found, nullptr) == kFinish);
    std::vector<KVStore::Victim> victims;
    CHECK(st.evict(2 * 16384, false, victims) == 2 * 16384 && victims.size() == 2);
    CHECK(!st.present("k2") && !st.present("k3") && st.present("k1") && st.present("k0"));
    const KeyHash h2 = hash_key(reinterpret_cast<const uint8_t*>("k2"), 2);
    CHECK(victims[0].hash.h1 == h2.h1 && victims[0].hash.h2 == h2.h2);
    CHECK(victims[0].block->addr() == addrs[2]);
    CHECK(mm.used_bytes() == 8 * 16384);  // victims still hold the space
    victims.clear();
    CHECK(mm.used_bytes() == 6 * 16384 && st.evicted() == 2);
    CHECK(st.reserve({"new", "new2"}, 16384, -1, 1, out) == kFinish);
    // uncommitted blocks are never victims; replica filter selects nothing here
    CHECK(st.evict(1, true, victims) == 0 && victims.empty());
    lease.clear();
    CHECK(st.evict(100 * 16384, false, victims) == 6 * 16384 && victims.size() == 6);
    CHECK(st.size() == 2 && st.present("new") && st.inflight() == 2);
    victims.clear();
    CHECK(st.purge() == 2);
    CHECK(mm.used_bytes() == 0);
    // the LRU list survives purge + refill
    CHECK(st.reserve({"z"}, 16384, -1, 1, out) == kFinish);
    uint64_t za = out[0].remote_addr;
    CHECK(st.commit(&za, 1) == 1);
    CHECK(st.evict(1, false, victims) == 16384 && victims.size() == 1);
}

// The flat key table (open addressing + backward-shift deletion) against a std::set model:
// random batches of reserve / commit / drop-by-connection / evict, membership checked for
// every key ever used after each step.
static void test_kv_store_table_model() {
    MM mm;
    mm.add_pool(size_t(1) << 30, 4096, -1);  // 262144 granules
    KVStore st(&mm, true);
    std::mt19937 rng{12345};
    std::set<std::string> live;
    std::vector<std::string> universe;
    std::map<uint64_t, std::vector<std::pair<std::string, uint64_t>>> uncommitted;  // conn -> (key, addr)
    uint64_t next_conn = 1;
    int next_key = 0;
    for (int step = 0; step < 400; ++step) {
        const int op = int(rng() % 10);
        if (op < 5) {  // reserve a batch (with some duplicates of existing keys)
            const int n = 1 + int(rng() % 300);
            std::vector<std::string> names;
            for (int i = 0; i < n; ++i) {
                if (!universe.empty() && rng() % 8 == 0)
                    names.push_back(universe[rng() % universe.size()]);
                else
                    names.push_back("model-key-" + std::to_string(next_key++) + std::string(rng() % 40, 'x'));
            }
            std::vector<std::string_view> keys(names.begin(), names.end());
            std::vector<RemoteBlock> out;
            const uint64_t conn = next_conn++;
            CHECK(st.reserve(keys, 4096, -1, conn, out) == kFinish);
            std::set<std::string> in_batch;
            for (int i = 0; i < n; ++i) {
                const bool existed = live.count(names[size_t(i)]) || in_batch.count(names[size_t(i)]);
                CHECK(is_fake_block(out[size_t(i)]) == existed);
                if (!existed) {
                    in_batch.insert(names[size_t(i)]);
                    uncommitted[conn].emplace_back(names[size_t(i)], out[size_t(i)].remote_addr);
                    universe.push_back(names[size_t(i)]);
                }
            }
            live.insert(in_batch.begin(), in_batch.end());
        } else if (op < 7 && !uncommitted.empty()) {  // commit one connection's blocks
            auto it = uncommitted.begin();
            std::advance(it, long(rng() % uncommitted.size()));
            std::vector<uint64_t> addrs;
            for (auto& kv : it->second) addrs.push_back(kv.second);
            CHECK(st.commit(addrs.data(), addrs.size()) == addrs.size());
            uncommitted.erase(it);
        } else if (op < 9 && !uncommitted.empty()) {  // a writer dies
            auto it = uncommitted.begin();
            std::advance(it, long(rng() % uncommitted.size()));
            CHECK(st.drop_uncommitted(it->first) == it->second.size());
            for (auto& kv : it->second) live.erase(kv.first);
            uncommitted.erase(it);
        } else {  // evict some committed blocks
            std::vector<KVStore::Victim> victims;
            st.evict(size_t(1 + rng() % 200) * 4096, false, victims);
            for (auto& v : victims) {
                const std::string k(v.block->key());
                CHECK(live.erase(k) == 1);
            }
        }
        CHECK(st.size() == live.size());
        if (step % 20 == 0 || step == 399)
            for (auto& k : universe) CHECK(st.present(k) == (live.count(k) == 1));
    }
    size_t inflight = 0;
    for (auto& kv : uncommitted) inflight += kv.second.size();
    CHECK(st.inflight() == inflight);
    const size_t n_live = live.size();
    CHECK(st.purge() == n_live && st.size() == 0 && mm.used_bytes() == 0);
}

// The device index algorithm (kernels/index.cuh) compiled for the CPU.
static void test_device_index_logic() {
    using namespace istore::kernels;
    const uint64_t slots = 1024, mask = index_bucket_mask(slots);
    std::vector<IndexBucket> table(slots / kIndexWays);
    std::memset(static_cast<void*>(table.data()), 0, table.size() * sizeof(IndexBucket));
    auto rec_of = [](int i) {
        const std::string k = "key/" + std::to_string(i);
        const KeyHash h = hash_key(reinterpret_cast<const uint8_t*>(k.data()), k.size());
        return IndexEntry{h.h1, h.h2, make_addr(0, uint64_t(i) * 4096), uint32_t(i + 1), 4096};
    };
    // half load: everything fits, claimed-but-uncommitted is invisible
    std::vector<uint32_t> slot(512);
    for (int i = 0; i < 512; ++i) {
        bool full = false;
        const IndexEntry r = rec_of(i);
        slot[size_t(i)] = idx::claim(table.data(), mask, r, true, &full);
        CHECK(slot[size_t(i)] != 0 && !full);
        CHECK(idx::find<true>(table.data(), mask, KeyHash{r.h1, r.h2}).slot_plus1 == 0);
        idx::commit(table.data(), slot[size_t(i)], r.tag, true);
    }
    for (int i = 0; i < 600; ++i) {
        const IndexEntry r = rec_of(i);
        const idx::Found f = idx::find<true>(table.data(), mask, KeyHash{r.h1, r.h2});
        if (i < 512) {
            CHECK(f.slot_plus1 == slot[size_t(i)] && f.tag == r.tag && f.addr == r.addr && f.size == 4096);
            CHECK(idx::still_valid(table.data(), f.slot_plus1, f.tag));
            const idx::Found g = idx::find<false>(table.data(), mask, KeyHash{r.h1, r.h2});
            CHECK(g.slot_plus1 == f.slot_plus1 && g.tag == f.tag && g.addr == f.addr);
        } else {
            CHECK(f.slot_plus1 == 0);
        }
    }
    // first writer wins inside the index too
    {
        bool full = false;
        IndexEntry r = rec_of(7);
        r.addr += 64;
        CHECK(idx::claim(table.data(), mask, r, true, &full) == 0 && !full);
        CHECK(idx::find<true>(table.data(), mask, KeyHash{r.h1, r.h2}).addr == rec_of(7).addr);
    }
    // eviction: the way becomes empty, readers that resolved it notice, the key can return
    for (int i = 0; i < 512; i += 2) {
        const IndexEntry r = rec_of(i);
        CHECK(idx::erase(table.data(), mask, r.h1, r.h2, r.addr));
        CHECK(!idx::erase(table.data(), mask, r.h1, r.h2, r.addr));
        CHECK(!idx::still_valid(table.data(), slot[size_t(i)], r.tag));
        CHECK(idx::find<true>(table.data(), mask, KeyHash{r.h1, r.h2}).slot_plus1 == 0);
    }
    for (int i = 1; i < 512; i += 2) {
        const IndexEntry r = rec_of(i);
        CHECK(idx::find<true>(table.data(), mask, KeyHash{r.h1, r.h2}).slot_plus1 == slot[size_t(i)]);
    }
    size_t used = 0;
    for (auto& b : table)
        for (uint64_t h : b.h1) used += h != 0;
    CHECK(used == 256);  // no tombstones: erased ways are empty again
    for (int i = 0; i < 512; i += 2) {
        bool full = false;
        IndexEntry r = rec_of(i);
        r.tag += 100000;  // a new allocation generation
        const uint32_t s = idx::claim(table.data(), mask, r, true, &full);
        CHECK(s != 0);
        idx::commit(table.data(), s, r.tag, true);
        CHECK(idx::find<true>(table.data(), mask, KeyHash{r.h1, r.h2}).tag == r.tag);
    }
    // overfull table: the failure is reported, everything claimed stays findable
    std::vector<IndexBucket> tiny(2);
    std::memset(static_cast<void*>(tiny.data()), 0, tiny.size() * sizeof(IndexBucket));
    int ok = 0, failed = 0;
    for (int i = 0; i < 40; ++i) {
        bool full = false;
        const IndexEntry r = rec_of(i);
        const uint32_t s = idx::claim(tiny.data(), index_bucket_mask(16), r, true, &full);
        if (s) {
            idx::commit(tiny.data(), s, r.tag, true);
            ++ok;
        } else {
            CHECK(full);
            ++failed;
        }
    }
    CHECK(ok == 16 && failed == 24);
    // at the load the server sizes the table for (<= 0.5) nothing overflows
    {
        const uint64_t big = 65536;
        std::vector<IndexBucket> t(big / kIndexWays);
        std::memset(static_cast<void*>(t.data()), 0, t.size() * sizeof(IndexBucket));
        int overflow = 0;
        for (int i = 0; i < int(big / 2); ++i) {
            bool full = false;
            const IndexEntry r = rec_of(i);
            const uint32_t s = idx::claim(t.data(), index_bucket_mask(big), r, true, &full);
            if (!s) ++overflow;
            else idx::commit(t.data(), s, r.tag, true);
        }
        CHECK(overflow == 0);
        int missing = 0;
        for (int i = 0; i < int(big / 2); ++i) {
            const IndexEntry r = rec_of(i);
            missing += idx::find<true>(t.data(), index_bucket_mask(big), KeyHash{r.h1, r.h2}).addr != r.addr;
        }
        CHECK(missing == 0);
    }
    // one-bucket table (b == a)
    std::vector<IndexBucket> one(1);
    std::memset(static_cast<void*>(one.data()), 0, sizeof(IndexBucket));
    for (int i = 0; i < 8; ++i) {
        bool full = false;
        const IndexEntry r = rec_of(i);
        const uint32_t s = idx::claim(one.data(), 0, r, true, &full);
        CHECK(s != 0);
        idx::commit(one.data(), s, r.tag, true);
        CHECK(idx::find<true>(one.data(), 0, KeyHash{r.h1, r.h2}).addr == r.addr);
    }
}

// Writers, an evictor and validating readers hammer one small table concurrently (threads
// stand in for GPUs).  Invariant: a read that passes validation saw the address and
// generation that belong to its key.
static void test_device_index_concurrent() {
    using namespace istore::kernels;
    const uint64_t slots = 256, mask = index_bucket_mask(slots);
    std::vector<IndexBucket> table(slots / kIndexWays);
    std::memset(static_cast<void*>(table.data()), 0, table.size() * sizeof(IndexBucket));
    constexpr int kKeys = 96;
    struct KeyState {
        KeyHash h;
        std::atomic<uint32_t> live_gen{0};  // 0: not in the store (server's view)
        std::atomic<uint32_t> slot{0};
    };
    std::vector<KeyState> ks(kKeys);
    for (int i = 0; i < kKeys; ++i) {
        const std::string k = "ck" + std::to_string(i);
        ks[size_t(i)].h = hash_key(reinterpret_cast<const uint8_t*>(k.data()), k.size());
    }
    // address encodes (key, generation): a reader can check what it resolved
    auto addr_of = [](int key, uint32_t gen) { return (uint64_t(key) << 32) | gen; };
    std::atomic<bool> stop{false};
    std::atomic<uint32_t> next_gen{1};
    std::atomic<long> validated{0}, stale{0}, bad{0}, cycles{0};
    // "server + writer + evictor" per key range: write -> commit -> (readers run) -> erase
    auto churn = [&](int first, int last) {
        std::mt19937 rng{unsigned(first)};
        while (!stop.load()) {
            const int i = first + int(rng() % unsigned(last - first));
            KeyState& k = ks[size_t(i)];
            if (k.live_gen.load() == 0) {
                const uint32_t gen = next_gen.fetch_add(1);
                bool full = false;
                const IndexEntry r{k.h.h1, k.h.h2, addr_of(i, gen), gen, 64};
                const uint32_t s = idx::claim(table.data(), mask, r, true, &full);
                if (!s) continue;
                idx::fence(true);
                idx::commit(table.data(), s, gen, true);
                k.slot.store(s);
                k.live_gen.store(gen);
            } else {
                const uint32_t gen = k.live_gen.exchange(0);
                if (!idx::erase(table.data(), mask, k.h.h1, k.h.h2, addr_of(i, gen))) ++bad;
                ++cycles;
            }
        }
    };
    auto reader = [&](unsigned seed) {
        std::mt19937 rng{seed};
        while (!stop.load()) {
            const int i = int(rng() % kKeys);
            const idx::Found f = (seed & 1) ? idx::find<true>(table.data(), mask, ks[size_t(i)].h)
                                            : idx::find<false>(table.data(), mask, ks[size_t(i)].h);
            if (!f.slot_plus1) continue;
            const uint64_t seen_addr = f.addr;
            idx::fence(true);  // "the copy"
            if (!idx::still_valid(table.data(), f.slot_plus1, f.tag)) {
                ++stale;
                continue;
            }
            ++validated;
            if (seen_addr != addr_of(i, f.tag)) ++bad;
        }
    };
    std::vector<std::thread> th;
    th.emplace_back(churn, 0, kKeys / 2);
    th.emplace_back(churn, kKeys / 2, kKeys);
    for (unsigned r = 0; r < 3; ++r) th.emplace_back(reader, 100 + r);
    std::this_thread::sleep_for(std::chrono::milliseconds(400));
    stop.store(true);
    for (auto& t : th) t.join();
    CHECK(bad.load() == 0);
    CHECK(validated.load() > 100 && cycles.load() > 100);
    std::printf("index stress: %ld validated reads, %ld stale, %ld evictions\n", validated.load(),
                stale.load(), cycles.load());
}

int main() {
    test_framing();
    test_flatbuffers();
    test_mempool();
    test_kv_store();
    test_kv_store_eviction();
    test_kv_store_table_model();
    test_hash();
    test_device_index_logic();
    test_device_index_concurrent();
    std::printf("%d checks, %d failed\n", g_checks, g_failed);
    return g_failed ? 1 : 0;
}
