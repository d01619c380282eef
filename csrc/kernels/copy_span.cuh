// The vector copy loop shared by the page movers: 256 threads stream one contiguous span
// with 128-bit (LDG.E.NA.128 / STG.E.128) or 256-bit (LDG.E.ENL2.256 / STG.E.ENL2.256)
// accesses, four vectors in flight per thread.
#pragma once

#include "common.cuh"

namespace istore::kernels {

using namespace dev;

constexpr int kLdStThreads = 256;

// NVLS: one 16-byte store to a multicast address lands in every replica  (multimem.st)
__device__ __forceinline__ void st_multicast_v4(void* mc, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc),
                 "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
                 "f"(__uint_as_float(v.w))
                 : "memory");
}

// MC = true: `dst` is a multicast address (VEC must be 16).
template <int VEC, bool MC = false>
__device__ __forceinline__ void copy_span(uint8_t* dst, const uint8_t* src, uint32_t len) {
    constexpr int U = 4;
    const uint32_t nvec = len / VEC;
    const uint32_t tid = threadIdx.x;
    uint32_t i = tid;
    for (; i + (U - 1) * kLdStThreads < nvec; i += U * kLdStThreads) {
        if constexpr (VEC == 16) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_stream_v4(src + size_t(i + u * kLdStThreads) * 16);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if constexpr (MC)
                    st_multicast_v4(dst + size_t(i + u * kLdStThreads) * 16, v[u]);
                else
                    st_v4(dst + size_t(i + u * kLdStThreads) * 16, v[u]);
            }
        } else {
            u32x8 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_v8(src + size_t(i + u * kLdStThreads) * 32);
#pragma unroll
            for (int u = 0; u < U; ++u) st_v8(dst + size_t(i + u * kLdStThreads) * 32, v[u]);
        }
    }
    for (; i < nvec; i += kLdStThreads) {
        if constexpr (VEC == 16 && MC)
            st_multicast_v4(dst + size_t(i) * 16, ld_stream_v4(src + size_t(i) * 16));
        else if constexpr (VEC == 16)
            st_v4(dst + size_t(i) * 16, ld_stream_v4(src + size_t(i) * 16));
        else
            st_v8(dst + size_t(i) * 32, ld_v8(src + size_t(i) * 32));
    }
    for (uint32_t b = nvec * VEC + tid; b < len; b += kLdStThreads) dst[b] = src[b];
}

}  // namespace istore::kernels
