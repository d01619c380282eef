// In-band commit: how a writer kernel makes a block visible in the HBM-resident index.
//
// Replaces the reference's COMMIT control message (client SEND after the RDMA writes,
// src/libinfinistore.cpp:362-395; server flips `committed`, src/infinistore.cpp:255-271)
// with a release-ordered publication from the kernel that moved the data: a reader on any
// GPU that observes the entry's tag (ld.acquire.sys) is guaranteed to observe the block.
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace istore::kernels {

using namespace dev;

struct Publish {
    const IndexEntry* recs;
    IndexEntry* table;
    uint64_t mask;
    uint32_t* done;
    uint32_t* status;
};

// Insert `rec` into the open-addressing table.  The slot is claimed with a 64-bit CAS on
// h1 (works on peer memory over NVLink); the tag is written last with release.sys, which
// is what makes the block visible to readers on any GPU.
__device__ inline void publish_entry(const Publish& pub, const IndexEntry& rec) {
    uint64_t slot = rec.h1 & pub.mask;
    for (uint64_t probe = 0; probe <= pub.mask; ++probe) {
        IndexEntry* e = pub.table + slot;
        uint64_t cur = ld_relaxed_sys_u64(&e->h1);
        if (cur == 0) cur = cas_relaxed_sys_u64(&e->h1, 0, rec.h1);
        if (cur == 0) {  // slot is ours
            st_relaxed_sys_u64(&e->h2, rec.h2);
            st_relaxed_sys_u64(&e->addr, rec.addr);
            e->size = rec.size;
            st_release_sys(&e->tag, rec.tag);
            return;
        }
        if (cur == rec.h1) {
            // Same key already present (first writer wins) or a 64-bit collision with a
            // different key; either way the authoritative copy is the server's map.
            const uint32_t tag = ld_acquire_sys(&e->tag);
            if (tag != 0 && e->h2 == rec.h2) return;
        }
        slot = (slot + 1) & pub.mask;
    }
    if (pub.status) atomicAdd(pub.status + kStatPublishFail, 1u);
}

// Called by every thread of the CTA after its last data store.  `first`, `count`, `stride`
// enumerate the work items this CTA handled; an item belongs to block item / cpb.
//
// Ordering: bar.sync makes every thread's data stores visible to warp 0 at CTA scope; each
// lane of warp 0 then issues ONE fence.acq_rel.sys, which is cumulative, so the stores of
// the whole CTA are performed system-wide before that lane's counter increment.  (One
// fence per lane of one warp instead of one per thread: MEMBAR.SYS is the expensive part
// of the epilogue and 256 of them per CTA serialise in the memory system.)
__device__ inline void publish_done_blocks(const Publish& pub, uint32_t first, uint32_t count,
                                           uint32_t stride, uint32_t cpb) {
    __syncthreads();
    if (threadIdx.x >= 32 || threadIdx.x >= count) return;
    fence_sys();
    for (uint32_t k = threadIdx.x; k < count; k += 32) {
        const uint32_t block = (first + k * stride) / cpb;
        const uint32_t arrived = cpb == 1 ? 1 : atomicAdd(pub.done + block, 1u) + 1;
        if (arrived == cpb) {  // last chunk of this block, chip-wide
            if (cpb != 1) {
                fence_sys();          // acquire side: other CTAs' stores happen-before us
                pub.done[block] = 0;  // leave the counter area zeroed for the next launch
            }
            publish_entry(pub, pub.recs[block]);
        }
    }
}

}  // namespace istore::kernels
