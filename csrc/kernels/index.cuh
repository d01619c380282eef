// The key index algorithm: claim / find / erase / validate on the bucketed table
// (kernels.h: IndexBucket).  Written once for both sides: the sm_100a kernels run it with
// system-scope PTX memory operations on (usually peer-mapped) HBM, the native unit tests
// run the same code on the CPU with __atomic builtins, threads standing in for GPUs.
//
// Replaces the reference's std::unordered_map<string> probes on the server CPU
// (src/infinistore.cpp:63-65, 1077-1108) for every client that reads through the device
// path; the server's host map stays the authority for allocation and first-writer-wins.
//
// Concurrency contract
//   writer : claim (CAS h1 0 -> fingerprint) ... data stores ... fence ... tag := gen
//   reader : find  (fingerprints, tag with acquire, then the fields) ... copy ...
//            validate (tag unchanged)   [validate only when the server evicts]
//   evictor: tag := 0, fence, h1 := 0   (server, before the block's space is reused)
// A reader that loses a race against evictor + new writer of the same way reads fields or
// bytes of another block; the changed tag exposes it and the read is reported as a miss.
#pragma once

#include <cstdint>

#include "../core/hash.h"
#include "kernels.h"

#if defined(__CUDACC__)
#include "common.cuh"
#endif

namespace istore::kernels::idx {

// ---------------------------------------------------------------- memory primitives
IS_HD uint64_t ld_u64(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
    return dev::ld_relaxed_sys_u64(p);
#else
    return __atomic_load_n(p, __ATOMIC_RELAXED);
#endif
}
IS_HD uint32_t ld_u32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#else
    return __atomic_load_n(p, __ATOMIC_RELAXED);
#endif
}
IS_HD uint32_t ld_acquire_u32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return dev::ld_acquire_sys(p);
#else
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#endif
}
// the eight fingerprints of a bucket: four 16-byte loads, one 64-byte line
IS_HD void ld_fingerprints(const IndexBucket* b, uint64_t (&h)[kIndexWays]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < int(kIndexWays); i += 2)
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];"
                     : "=l"(h[i]), "=l"(h[i + 1])
                     : "l"(&b->h1[i])
                     : "memory");
#else
    for (uint32_t i = 0; i < kIndexWays; ++i) h[i] = __atomic_load_n(&b->h1[i], __ATOMIC_RELAXED);
#endif
}
IS_HD uint64_t cas_u64(uint64_t* p, uint64_t cmp, uint64_t val, bool sys) {
#if defined(__CUDA_ARCH__)
    // Atomics on one address are serialised at the L2 that owns it whatever the scope
    // qualifier, so a table in this GPU's own HBM may be claimed at gpu scope.
    return sys ? dev::cas_relaxed_sys_u64(p, cmp, val) : dev::cas_relaxed_gpu_u64(p, cmp, val);
#else
    (void)sys;
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
#endif
}
IS_HD void st_u64(uint64_t* p, uint64_t v) {
#if defined(__CUDA_ARCH__)
    dev::st_relaxed_sys_u64(p, v);
#else
    __atomic_store_n(p, v, __ATOMIC_RELAXED);
#endif
}
IS_HD void st_u32(uint32_t* p, uint32_t v, bool sys) {
#if defined(__CUDA_ARCH__)
    if (sys)
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    (void)sys;
    __atomic_store_n(p, v, __ATOMIC_RELAXED);
#endif
}
IS_HD void fence(bool sys) {
#if defined(__CUDA_ARCH__)
    if (sys)
        dev::fence_sys();
    else
        dev::fence_gpu();
#else
    (void)sys;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
#endif
}

// ---------------------------------------------------------------- placement
IS_HD uint64_t bucket_a(uint64_t h1, uint64_t mask) { return h1 & mask; }
IS_HD uint64_t bucket_b(uint64_t h1, uint64_t h2, uint64_t mask) {
    uint64_t b = ((h1 >> 32) ^ h2) & mask;
    if (b == (h1 & mask)) b = (b + 1) & mask;  // equal only for a one-bucket table
    return b;
}
IS_HD uint32_t first_way(uint64_t h2) { return uint32_t(h2 >> 61); }
IS_HD uint32_t slot_id(uint64_t bucket, uint32_t way) { return uint32_t(bucket * kIndexWays + way); }
IS_HD IndexWay* way_of(IndexBucket* table, uint32_t slot) {
    return &table[slot / kIndexWays].way[slot % kIndexWays];
}
IS_HD const IndexWay* way_of(const IndexBucket* table, uint32_t slot) {
    return &table[slot / kIndexWays].way[slot % kIndexWays];
}

// ---------------------------------------------------------------- writer
// Reserve a way for `rec`.  Returns slot + 1, or 0 when nothing is to be committed: the key
// is already published (first writer wins) or both buckets are full (*full set).
// Fast path: one CAS on the key's preferred way of bucket A - ONE fabric round trip for an
// insertion into a lightly loaded table; the fields follow as posted stores, tag stays 0.
// Slow path (that way is taken): read both buckets' fingerprints and take a free way of the
// emptier bucket.  Two-choice placement keeps "both buckets full" out of reach at the load
// the server sizes the table for (<= 0.5): none in 4e4 insertions at load 0.5, 3e-4 at 0.75.
IS_HD uint32_t take_way(IndexBucket* bk, uint64_t bi, uint32_t w, const IndexEntry& rec,
                        bool sys) {
    st_u64(&bk->way[w].h2, rec.h2);  // posted; the tag stays 0
    st_u64(&bk->way[w].addr, rec.addr);
    st_u32(&bk->way[w].size, rec.size, sys);
    return slot_id(bi, w) + 1;
}
IS_HD bool published_as(const IndexBucket* bk, uint32_t w, const IndexEntry& rec) {
    // same fingerprint: the same key (already published: first writer wins) or a 64-bit
    // collision with another key; the authoritative copy is the server's map either way
    return ld_acquire_u32(&bk->way[w].tag) != 0 && ld_u64(&bk->way[w].h2) == rec.h2;
}
#if defined(__CUDACC__)
#define IS_HD_NOINLINE __host__ __device__ __noinline__
#else
#define IS_HD_NOINLINE inline
#endif
// out of line: its 16 fingerprint registers must not inflate the copy kernels that inline
// the one-CAS fast path
IS_HD_NOINLINE uint32_t claim_slow(IndexBucket* table, uint64_t mask, const IndexEntry& rec,
                                   bool sys, bool* full);

IS_HD uint32_t claim(IndexBucket* table, uint64_t mask, const IndexEntry& rec, bool sys,
                     bool* full) {
    const uint64_t a = bucket_a(rec.h1, mask), b = bucket_b(rec.h1, rec.h2, mask);
    const uint32_t w0 = first_way(rec.h2);
    IndexBucket* ba = table + a;
    IndexBucket* bb = table + b;
    const uint64_t cur = cas_u64(&ba->h1[w0], 0, rec.h1, sys);
    if (cur == 0) return take_way(ba, a, w0, rec, sys);
    if (cur == rec.h1 && published_as(ba, w0, rec)) return 0;
    return claim_slow(table, mask, rec, sys, full);
}

IS_HD_NOINLINE uint32_t claim_slow(IndexBucket* table, uint64_t mask, const IndexEntry& rec,
                                   bool sys, bool* full) {
    const uint64_t a = bucket_a(rec.h1, mask), b = bucket_b(rec.h1, rec.h2, mask);
    IndexBucket* ba = table + a;
    IndexBucket* bb = table + b;
    for (int attempt = 0; attempt < 4; ++attempt) {  // retried only when a CAS loses a race
        uint64_t ha[kIndexWays], hb[kIndexWays];
        ld_fingerprints(ba, ha);
        ld_fingerprints(bb, hb);
        uint32_t free_a = 0, free_b = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t w = 0; w < kIndexWays; ++w) {
            free_a += ha[w] == 0;
            free_b += hb[w] == 0 && b != a;
            if (ha[w] == rec.h1 && published_as(ba, w, rec)) return 0;
            if (b != a && hb[w] == rec.h1 && published_as(bb, w, rec)) return 0;
        }
        if (free_a == 0 && free_b == 0) break;
        const bool use_b = free_b > free_a;
        IndexBucket* bk = use_b ? bb : ba;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t w = 0; w < kIndexWays; ++w) {
            if ((use_b ? hb[w] : ha[w]) != 0) continue;
            if (cas_u64(&bk->h1[w], 0, rec.h1, sys) == 0)
                return take_way(bk, use_b ? b : a, w, rec, sys);
        }
    }
    *full = true;
    return 0;
}

// The caller has fenced the block's data stores (and the claim's field stores): fence +
// relaxed store is a release pattern, the tag needs no second MEMBAR.
IS_HD void commit(IndexBucket* table, uint32_t slot_plus1, uint32_t tag, bool sys) {
    if (slot_plus1) st_u32(&way_of(table, slot_plus1 - 1)->tag, tag, sys);
}

// ---------------------------------------------------------------- reader
struct Found {
    uint32_t slot_plus1;  // 0 = not found (absent, or reserved by a writer still copying)
    uint32_t tag;
    uint64_t addr;
    uint32_t size;
};

// Round trip 1: the 16 fingerprints of both buckets and, speculatively, the tag of the
// preferred way of bucket A (where the key sits unless that way was taken when it was
// written).  Round trip 2: the fields.  A miss costs one round trip.
IS_HD Found find(const IndexBucket* table, uint64_t mask, const KeyHash& kh) {
    const uint64_t a = bucket_a(kh.h1, mask), b = bucket_b(kh.h1, kh.h2, mask);
    const uint32_t w0 = first_way(kh.h2);
    uint64_t ha[kIndexWays], hb[kIndexWays];
    ld_fingerprints(table + a, ha);
    ld_fingerprints(table + b, hb);
    const uint32_t tag0 = ld_acquire_u32(&table[a].way[w0].tag);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int which = 0; which < 2; ++which) {
        const IndexBucket* bk = table + (which ? b : a);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t w = 0; w < kIndexWays; ++w) {
            if ((which ? hb[w] : ha[w]) != kh.h1) continue;
            const IndexWay* wy = &bk->way[w];
            const uint32_t tag = (which == 0 && w == w0) ? tag0 : ld_acquire_u32(&wy->tag);
            if (tag == 0) continue;  // claimed, not committed
            if (ld_u64(&wy->h2) != kh.h2) continue;
            return Found{slot_id(which ? b : a, w) + 1, tag, ld_u64(&wy->addr), ld_u32(&wy->size)};
        }
        if (b == a) break;
    }
    return Found{0, 0, 0, 0};
}

// After the copy: is the entry the reader resolved still the one in the table?
IS_HD bool still_valid(const IndexBucket* table, uint32_t slot_plus1, uint32_t tag) {
    return ld_acquire_u32(&way_of(table, slot_plus1 - 1)->tag) == tag;
}

// ---------------------------------------------------------------- evictor
// Empty the way that holds (h1, h2, addr).  tag := 0 first, so a reader sees "not
// committed" before the way can be claimed by another key.  false: no such entry.
IS_HD bool erase(IndexBucket* table, uint64_t mask, uint64_t h1, uint64_t h2, uint64_t addr) {
    const uint64_t a = bucket_a(h1, mask), b = bucket_b(h1, h2, mask);
    for (int which = 0; which < 2; ++which) {
        if (which && b == a) break;
        IndexBucket* bk = table + (which ? b : a);
        for (uint32_t w = 0; w < kIndexWays; ++w) {
            if (ld_u64(&bk->h1[w]) != h1) continue;
            if (ld_u64(&bk->way[w].h2) != h2 || ld_u64(&bk->way[w].addr) != addr) continue;
            st_u32(&bk->way[w].tag, 0u, true);
            fence(true);
            st_u64(&bk->h1[w], 0);
            return true;
        }
    }
    return false;
}

}  // namespace istore::kernels::idx
