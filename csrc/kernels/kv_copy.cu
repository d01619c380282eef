// kv_copy: the batched page mover behind write_cache (client GPU -> pool HBM) and
// read_cache (pool HBM -> client GPU).
//
// One launch moves a whole batch of pages (a layer's worth) described by a descriptor
// array; either side of every descriptor may be a peer-mapped address, so the bytes cross
// NVLink 5 / NVSwitch from inside the kernel.  This replaces, for one batch,
//   * N cudaMemcpyAsync calls on a fresh stream + event (reference local path,
//     src/infinistore.cpp:623-624,747-748),
//   * N RDMA_WRITE work requests chained 32 at a time (src/libinfinistore.cpp:905-970,
//     src/infinistore.cpp:456-530), and
//   * the COMMIT message round (src/libinfinistore.cpp:362-395 -> src/infinistore.cpp:
//     255-271): the kernel publishes each block in the HBM-resident index itself, with
//     release semantics at system scope, once the block's bytes have landed.
//
// Two data paths, selected per launch (measured, not guessed — see profiles/):
//   kCopyLdSt : every thread streams 128-bit (or 256-bit) vectors, 4 in flight per thread;
//   kCopyTma  : one elected thread per CTA drives an SMEM ring with 1-D bulk async copies
//               (cp.async.bulk, mbarrier completion in, bulk-group completion out).  A few
//               CTAs of one warp each keep megabytes in flight, leaving the SMs to the
//               model's own kernels when the transfer overlaps prefill.
#include <algorithm>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "copy_span.cuh"
#include "publish.cuh"

namespace istore::kernels {

namespace {

using namespace dev;

constexpr uint32_t kLdStChunk = 32u << 10;  // work item of the ld/st path
constexpr int kTmaMaxStages = 32;
constexpr uint32_t kTmaChunk = 16u << 10;   // largest bulk copy / ring slot
constexpr uint32_t kTmaRingBytes = 128u << 10;

// ---------------------------------------------------------------- ld/st path (copy_span.cuh)
// Descriptors of small batches travel in the kernel parameters (constant bank): reading
// them from the pinned host ring costs a PCIe round trip at the start of every launch
// (+3-4 us of 18 on a 32 MB launch, profiles/r1_launch_overhead_v1.json).
constexpr int kParamDescs = 256;
template <int N>
struct DescParam {
    CopyDesc d[N];
};

// VEC = 16 / 32: vector width; VEC = 1: byte fallback for unaligned tensors.
// Threads [0, 256) copy; warp 8 is the control warp (in-band commit, publish.cuh).
template <int VEC, bool PARAM, bool MC = false>
__global__ void __launch_bounds__(kLdStThreads + 32)
    kv_copy_ldst_kernel(const CopyDesc* __restrict__ descs,
                        const __grid_constant__ DescParam<PARAM ? kParamDescs : 1> pd, uint32_t n,
                        uint32_t bytes, uint32_t chunk, uint32_t cpb, Publish pub) {
    const uint32_t total = n * cpb;
    if (threadIdx.x >= kLdStThreads) {
        if (!pub.recs) return;
        const uint32_t count = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        control_warp(pub, threadIdx.x - kLdStThreads, blockIdx.x, count, gridDim.x, cpb,
                     kLdStThreads + 32);
        return;
    }
    auto desc_at = [&](uint32_t block) -> CopyDesc {
        if constexpr (PARAM)
            return pd.d[block];
        else
            return descs[block];
    };
    uint32_t item = blockIdx.x;
    CopyDesc next = item < total ? desc_at(item / cpb) : CopyDesc{0, 0};
    for (; item < total; item += gridDim.x) {
        const CopyDesc d = next;
        const uint32_t nxt = item + gridDim.x;
        if (nxt < total) next = desc_at(nxt / cpb);  // prefetch: descriptors may sit in host memory
        if (d.src == 0) {  // key not found by the device lookup
            if (threadIdx.x == 0 && item % cpb == 0 && pub.status)
                atomicAdd(pub.status + kStatMiss, 1u);
            continue;
        }
        const uint32_t off = (item % cpb) * chunk;
        const uint32_t len = min(chunk, bytes - off);
        uint8_t* dst = reinterpret_cast<uint8_t*>(d.dst) + off;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(d.src) + off;
        if constexpr (VEC == 1) {
            for (uint32_t b = threadIdx.x; b < len; b += kLdStThreads) dst[b] = src[b];
        } else {
            copy_span<VEC, MC>(dst, src, len);
        }
    }
    if (pub.recs) ctrl_barrier_arrive(kLdStThreads + 32);
}

// ---------------------------------------------------------------- bulk-async (TMA) path
// Two warps per CTA.  Warp 0: lane 0 drives an SMEM ring with 1-D bulk async copies; all lanes
// prefetch descriptors (32 at a time, one coalesced read even when they live in mapped host
// memory).  Warp 1 is the control warp (in-band commit).
//
// Ring: `stages` slots of `stage_bytes` (chosen per launch: 16 KB slots for big blocks, one
// slot per block for small ones, up to 32 slots / 128 KB).  Global->shared copies complete on
// the slot's mbarrier (complete_tx); shared->global copies are tracked as bulk groups, and a
// slot is refilled once the store that used it has finished READING shared memory
// (wait_group.read), with kPendingStores groups allowed to lag so that the issuing thread
// never waits for the store it has just issued.
constexpr int kPendingStores = 2;

template <bool PARAM>
__global__ void __launch_bounds__(64)
    kv_copy_tma_kernel(const CopyDesc* __restrict__ descs,
                       const __grid_constant__ DescParam<PARAM ? kParamDescs : 1> pd, uint32_t n,
                       uint32_t bytes, uint32_t stage_bytes, uint32_t stages, uint32_t cpb,
                       Publish pub) {
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) uint64_t full[kTmaMaxStages];
    const uint32_t total = n * cpb;
    const uint32_t grid = gridDim.x;
    const uint32_t nitems = blockIdx.x < total ? (total - blockIdx.x + grid - 1) / grid : 0;
    if (threadIdx.x >= 32) {
        if (pub.recs) control_warp(pub, threadIdx.x - 32, blockIdx.x, nitems, grid, cpb, 64);
        return;
    }
    const uint32_t lane = threadIdx.x;

    if (lane == 0) {
        for (uint32_t s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    // item k of this CTA -> global item blockIdx.x + k * grid
    auto fetch = [&](uint32_t k0) -> CopyDesc {  // lane l gets the descriptor of item k0 + l
        const uint32_t k = k0 + lane;
        if (k >= nitems) return CopyDesc{0, 0};
        const uint32_t block = (blockIdx.x + k * grid) / cpb;
        if constexpr (PARAM)
            return pd.d[block];
        else
            return descs[block];
    };
    // The descriptors of items [w, w+32) live in `cur`, of [w+32, w+64) in `nxt`.
    CopyDesc cur = fetch(0), nxt = fetch(32);
    uint32_t window = 0;
    auto desc_of = [&](uint32_t k) -> CopyDesc {  // warp-uniform k in [window, window + 64)
        const bool in_cur = k < window + 32;
        const uint32_t l = (k - window) & 31;
        CopyDesc r;
        r.src = __shfl_sync(0xffffffffu, in_cur ? cur.src : nxt.src, l);
        r.dst = __shfl_sync(0xffffffffu, in_cur ? cur.dst : nxt.dst, l);
        return r;
    };
    uint32_t loaded = 0;  // loads issued so far
    auto pump = [&](uint32_t limit) {  // issue loads up to (excluding) item `limit`
        while (loaded < nitems && loaded < limit) {
            const CopyDesc d = desc_of(loaded);
            const uint32_t s = loaded % stages;
            const uint32_t off = ((blockIdx.x + loaded * grid) % cpb) * stage_bytes;
            const uint32_t len = min(stage_bytes, bytes - off);
            if (lane == 0 && d.src != 0) {
                mbar_expect_tx(&full[s], len);
                bulk_g2s(ring + size_t(s) * stage_bytes,
                         reinterpret_cast<const uint8_t*>(d.src) + off, len, &full[s]);
            }
            ++loaded;
        }
    };
    pump(stages);  // prologue: fill the ring (stages <= 32: inside the descriptor window)

    for (uint32_t k = 0; k < nitems; ++k) {
        if (k >= window + 32) {  // slide the descriptor window
            window += 32;
            cur = nxt;
            nxt = fetch(window + 32);
        }
        const CopyDesc d = desc_of(k);
        const uint32_t s = k % stages;
        const uint32_t item = blockIdx.x + k * grid;
        const uint32_t off = (item % cpb) * stage_bytes;
        const uint32_t len = min(stage_bytes, bytes - off);
        if (lane == 0) {
            if (d.src != 0) {
                mbar_wait(&full[s], (k / stages) & 1);
                bulk_s2g(reinterpret_cast<uint8_t*>(d.dst) + off, ring + size_t(s) * stage_bytes, len);
            } else if (off == 0 && pub.status) {
                atomicAdd(pub.status + kStatMiss, 1u);
            }
            bulk_commit();  // one group per item keeps the wait_group arithmetic simple
            // stores up to item k - kPendingStores have drained their slots
            bulk_wait_read<kPendingStores>();
        }
        // slot of store j is reused by load j + stages; loads < k + 32 stay in the window
        if (k >= kPendingStores) pump(k - kPendingStores + 1 + stages);
    }
    if (lane == 0) {
        bulk_wait<0>();        // every bulk store of this CTA has completed its writes
        fence_proxy_async();   // order async-proxy writes before the generic-proxy commit
    }
    __syncwarp();
    if (pub.recs) ctrl_barrier_arrive(64);
}

std::mutex g_attr_mu;
bool g_tma_attr_set[64] = {false};

}  // namespace

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

cudaError_t launch_kv_copy(const CopyLaunch& a, cudaStream_t stream) {
    if (a.n == 0 || a.bytes == 0) return cudaSuccess;
    Publish pub{a.recs, a.table, a.table_mask, a.done, a.status, a.n, a.trace, !a.all_local, a.debug};
    if (!a.table || !a.done) pub.recs = nullptr;
    const int sms = sm_count();

    int variant = a.variant;
    // Alignment: bulk copies need 16-byte aligned addresses and sizes (the caller ORs every
    // local address into align_or; pool blocks are granule aligned).
    const bool aligned16 = (a.bytes % 16) == 0 && (a.align_or & 15) == 0;
    const bool aligned32 = (a.bytes % 32) == 0 && (a.align_or & 31) == 0;
    if (a.multicast) {
        if (!aligned16) return cudaErrorInvalidValue;  // multimem.st moves 16-byte vectors
        variant = kCopyLdSt;
    }
    if (variant == kCopyAuto) variant = aligned32 ? kCopyLdSt256 : kCopyLdSt;
    if (!aligned16 && (variant == kCopyTma || variant == kCopyLdSt256)) variant = kCopyLdSt;
    if (variant == kCopyLdSt256 && !aligned32) variant = kCopyLdSt;

    // small batches: descriptors ride in the kernel parameters
    const bool param = a.descs_host != nullptr && a.n <= uint32_t(kParamDescs) && aligned16;
    DescParam<kParamDescs> pd;
    if (param) std::memcpy(pd.d, a.descs_host, size_t(a.n) * sizeof(CopyDesc));
    const DescParam<1> none{};

    if (variant == kCopyTma) {
        // slot = 16 KB for big blocks, the block itself (rounded up to 16 B) for small ones
        const uint32_t stage_bytes = std::min(kTmaChunk, (a.bytes + 15u) & ~15u);
        const uint32_t stages = std::min<uint32_t>(kTmaMaxStages, kTmaRingBytes / stage_bytes);
        const uint32_t cpb = (a.bytes + stage_bytes - 1) / stage_bytes;
        const uint64_t total = uint64_t(a.n) * cpb;
        int ctas = a.max_ctas > 0 ? a.max_ctas : 2 * sms;
        ctas = int(std::min<uint64_t>(uint64_t(ctas), total));
        const size_t smem = size_t(stages) * stage_bytes;
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(g_attr_mu);
            if (dev >= 0 && dev < 64 && !g_tma_attr_set[dev]) {
                cudaError_t e = cudaFuncSetAttribute(kv_copy_tma_kernel<false>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     int(kTmaRingBytes));
                if (e == cudaSuccess)
                    e = cudaFuncSetAttribute(kv_copy_tma_kernel<true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(kTmaRingBytes));
                if (e != cudaSuccess) return e;
                g_tma_attr_set[dev] = true;
            }
        }
        if (param)
            kv_copy_tma_kernel<true><<<ctas, 64, smem, stream>>>(a.descs, pd, a.n, a.bytes, stage_bytes,
                                                                 stages, cpb, pub);
        else
            kv_copy_tma_kernel<false><<<ctas, 64, smem, stream>>>(a.descs, none, a.n, a.bytes,
                                                                  stage_bytes, stages, cpb, pub);
        return cudaGetLastError();
    }

    // Work item = one 32 KB chunk, or - when there are at least as many blocks as SMs - one
    // whole block: then a block is moved by a single CTA, its commit needs no cross-CTA
    // counter (claim, one fence, one store) and a reader resolves every key exactly once.
    uint32_t chunk = std::min(a.bytes, kLdStChunk);
    if (a.n >= uint32_t(sms) && a.bytes <= (1u << 20)) chunk = a.bytes;
    const uint32_t cpb = (a.bytes + chunk - 1) / chunk;
    const uint64_t total = uint64_t(a.n) * cpb;
    constexpr int T = kLdStThreads + 32;
    // One wave: every CTA is resident and loops over its items, so the commit epilogue
    // (fence + release) runs once per CTA instead of once per item and per wave.
    static int resident16 = 0, resident32 = 0;
    if (!resident16) {
        int b16 = 0, b32 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b16, kv_copy_ldst_kernel<16, false>, T, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b32, kv_copy_ldst_kernel<32, false>, T, 0);
        resident16 = std::max(b16, 1);
        resident32 = std::max(b32, 1);
    }
    const int per_sm = (variant == kCopyLdSt256) ? resident32 : resident16;
    int ctas = a.max_ctas > 0 ? a.max_ctas : per_sm * sms;
    ctas = int(std::min<uint64_t>(uint64_t(ctas), total));
    if (a.multicast && param)
        kv_copy_ldst_kernel<16, true, true><<<ctas, T, 0, stream>>>(a.descs, pd, a.n, a.bytes, chunk, cpb, pub);
    else if (a.multicast)
        kv_copy_ldst_kernel<16, false, true><<<ctas, T, 0, stream>>>(a.descs, none, a.n, a.bytes, chunk, cpb, pub);
    else if (!aligned16)
        kv_copy_ldst_kernel<1, false><<<ctas, T, 0, stream>>>(a.descs, none, a.n, a.bytes, chunk, cpb, pub);
    else if (variant == kCopyLdSt256 && param)
        kv_copy_ldst_kernel<32, true><<<ctas, T, 0, stream>>>(a.descs, pd, a.n, a.bytes, chunk, cpb, pub);
    else if (variant == kCopyLdSt256)
        kv_copy_ldst_kernel<32, false><<<ctas, T, 0, stream>>>(a.descs, none, a.n, a.bytes, chunk, cpb, pub);
    else if (param)
        kv_copy_ldst_kernel<16, true><<<ctas, T, 0, stream>>>(a.descs, pd, a.n, a.bytes, chunk, cpb, pub);
    else
        kv_copy_ldst_kernel<16, false><<<ctas, T, 0, stream>>>(a.descs, none, a.n, a.bytes, chunk, cpb, pub);
    return cudaGetLastError();
}

}  // namespace istore::kernels
