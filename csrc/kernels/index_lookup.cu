// kv_index_lookup: hash + probe + compare of the HBM-resident key index on the GPU.
//
// Serves three API calls that the reference answers on the server CPU with
// std::unordered_map<string> probes (src/infinistore.cpp:1077-1108, 436-446):
//   * read_cache        : resolve every key of the batch to a pool address and emit the
//                         copy descriptors consumed by kv_copy in the same stream;
//   * get_match_last_index : presence bitmap over the key list, then the last CTA replays
//                         the reference's binary search over that bitmap, bit-exact on any
//                         (also non-monotone) input;
//   * check_exist       : the same with one key.
// One thread per key: the key bytes are hashed twice (core/hash.h, identical on host and
// device) and the key's two buckets of the table — which lives in the pool GPU's HBM and is
// usually a PEER mapping read over NVLink — are searched (index.cuh).  An entry counts only
// once its tag has been published with release semantics by the writer's kv_copy.
// Also here: the eviction kernels (erase on the server side, post-copy validation on the
// reader side).
#include "../core/hash.h"
#include "common.cuh"
#include "index.cuh"
#include "kernels.h"

namespace istore::kernels {

namespace {

using namespace dev;

// one warp per CTA: a batch of lookups spreads over many SMs instead of queueing hundreds of
// fabric loads behind one SM's load unit
constexpr int kLookupThreads = 32;
__global__ void __launch_bounds__(kLookupThreads)
    kv_index_lookup_kernel(const __grid_constant__ LookupLaunch a) {
    const uint32_t i = blockIdx.x * kLookupThreads + threadIdx.x;
    bool found = false;
    if (i < a.n) {
        const KeyHash kh = hash_key(a.key_bytes + a.key_off[i], a.key_len[i]);
        // reads resolve present keys (bucket A first), match / exist probes mostly absent ones
        const idx::Found h = a.present ? idx::find<true>(a.table, a.table_mask, kh)
                                       : idx::find<false>(a.table, a.table_mask, kh);
        found = h.slot_plus1 != 0;
        if (a.out_descs) {
            uint64_t src = 0;
            if (found && h.size >= a.need_bytes) {
                const uint32_t seg = uint32_t(h.addr >> 44) - 1;
                if (seg < a.nsegs && a.seg_base[seg])
                    src = a.seg_base[seg] + (h.addr & ((1ull << 44) - 1));
            }
            a.out_descs[i] = CopyDesc{src, a.dst_base + a.dst_off[i]};
            if (a.found_at)
                a.found_at[i] = LookupLaunch::FoundAt{src ? h.slot_plus1 : 0u, h.tag};
        }
    }
    if (!a.present) return;
    // 32 consecutive keys per warp -> one bitmap word, no atomics
    const uint32_t word = __ballot_sync(0xffffffffu, found);
    if ((threadIdx.x & 31) == 0 && i < a.n) a.present[i >> 5] = word;
    if (!a.want_match) return;

    // elect the last CTA; it sees every CTA's bitmap words
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(a.ticket, 1u);
        is_last = t == gridDim.x - 1;
        if (is_last) *a.ticket = 0;  // self-cleaning for the next launch
    }
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    __threadfence();
    const volatile uint32_t* present = a.present;
    // Exact replay of the reference's search (src/infinistore.cpp:1093-1104).
    int left = 0, right = int(a.n);
    while (left < right) {
        const int mid = left + (right - left) / 2;
        if ((present[mid >> 5] >> (mid & 31)) & 1u)
            left = mid + 1;
        else
            right = mid;
    }
    a.status[kStatMatch] = uint32_t(left - 1);
    __threadfence_system();
}

// One thread per block that was read: the tag must be unchanged (see ValidateLaunch).
__global__ void __launch_bounds__(kLookupThreads)
    kv_index_validate_kernel(const __grid_constant__ ValidateLaunch a) {
    const uint32_t i = blockIdx.x * kLookupThreads + threadIdx.x;
    if (i >= a.n) return;
    const LookupLaunch::FoundAt f = a.found_at[i];
    if (!f.slot_plus1) return;  // a miss was counted by the lookup
    if (!idx::still_valid(a.table, f.slot_plus1, f.tag)) {
        atomicAdd(a.status + kStatMiss, 1u);
        atomicAdd(a.status + kStatStale, 1u);
    }
}

// One thread per evicted block (index.cuh: tag := 0, fence, h1 := 0).
__global__ void __launch_bounds__(kLookupThreads)
    kv_index_erase_kernel(const __grid_constant__ EraseLaunch a) {
    const uint32_t i = blockIdx.x * kLookupThreads + threadIdx.x;
    if (i >= a.n) return;
    const EraseRec r = a.recs[i];
    idx::erase(a.table, a.table_mask, r.h1, r.h2, r.addr);
}

}  // namespace

cudaError_t launch_index_validate(const ValidateLaunch& a, cudaStream_t stream) {
    if (a.n == 0) return cudaSuccess;
    const unsigned grid = (a.n + kLookupThreads - 1) / kLookupThreads;
    kv_index_validate_kernel<<<grid, kLookupThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_index_erase(const EraseLaunch& a, cudaStream_t stream) {
    if (a.n == 0) return cudaSuccess;
    const unsigned grid = (a.n + kLookupThreads - 1) / kLookupThreads;
    kv_index_erase_kernel<<<grid, kLookupThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_index_lookup(const LookupLaunch& a, cudaStream_t stream) {
    if (a.n == 0) return cudaSuccess;
    const unsigned grid = (a.n + kLookupThreads - 1) / kLookupThreads;
    kv_index_lookup_kernel<<<grid, kLookupThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace istore::kernels
