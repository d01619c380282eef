// kv_bcast_nvls: one writer -> every reader replication of KV blocks through an NVLS
// multicast mapping.
//
// The reference serves a shared prefix with N independent unicast reads, one RDMA_WRITE
// chain per client (src/infinistore.cpp:424-533).  On an NVSwitch box the writer instead
// stores each 16-byte vector ONCE to a multicast address (multimem.st) and the switch
// replicates it into the replica region of every bound GPU: writer egress stays at one
// copy while N copies are delivered.  Readers then consume their local replica at HBM
// speed.  Descriptors: src = local page, dst = address inside the multicast mapping.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace istore::kernels {

namespace {

using namespace dev;

constexpr int kThreads = 256;
constexpr uint32_t kChunk = 32u << 10;

__device__ __forceinline__ void multimem_st_u4(void* mc, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc),
                 "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
                 "f"(__uint_as_float(v.w))
                 : "memory");
}

__global__ void __launch_bounds__(kThreads)
    kv_bcast_nvls_kernel(const CopyDesc* __restrict__ descs, uint32_t n, uint32_t bytes,
                         uint32_t chunk, uint32_t cpb) {
    constexpr int U = 4;
    const uint32_t total = n * cpb;
    for (uint32_t item = blockIdx.x; item < total; item += gridDim.x) {
        const CopyDesc d = descs[item / cpb];
        const uint32_t off = (item % cpb) * chunk;
        const uint32_t len = min(chunk, bytes - off);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(d.src) + off;
        uint8_t* dst = reinterpret_cast<uint8_t*>(d.dst) + off;
        const uint32_t nvec = len / 16;
        uint32_t i = threadIdx.x;
        for (; i + (U - 1) * kThreads < nvec; i += U * kThreads) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_stream_v4(src + size_t(i + u * kThreads) * 16);
#pragma unroll
            for (int u = 0; u < U; ++u) multimem_st_u4(dst + size_t(i + u * kThreads) * 16, v[u]);
        }
        for (; i < nvec; i += kThreads) multimem_st_u4(dst + size_t(i) * 16, ld_stream_v4(src + size_t(i) * 16));
    }
    // make the replicated stores visible to every reader before the kernel's completion
    // is signalled (readers synchronise on a flag / stream event afterwards)
    fence_sys();
}

}  // namespace

cudaError_t launch_kv_bcast_nvls(const BcastLaunch& a, cudaStream_t stream) {
    if (a.n == 0 || a.bytes == 0) return cudaSuccess;
    if (a.bytes % 16) return cudaErrorInvalidValue;
    const uint32_t chunk = std::min(a.bytes, kChunk);
    const uint32_t cpb = (a.bytes + chunk - 1) / chunk;
    const uint64_t total = uint64_t(a.n) * cpb;
    int ctas = a.max_ctas > 0 ? a.max_ctas : 4 * sm_count();
    ctas = int(std::min<uint64_t>(uint64_t(ctas), total));
    kv_bcast_nvls_kernel<<<ctas, kThreads, 0, stream>>>(a.descs, a.n, a.bytes, chunk, cpb);
    return cudaGetLastError();
}

}  // namespace istore::kernels
