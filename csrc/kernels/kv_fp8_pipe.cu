// fp8 KV path on the TMA pipeline: quantise / dequantise between two bulk-async copies.
//
//   write : bf16 page --cp.async.bulk--> SMEM tile (64 rows x 128)   [loader warp]
//           4 compute warps: per-row amax, e4m3 cast, scale           (shared -> shared)
//           SMEM payload + scales --cp.async.bulk--> pool block       [storer warp]
//           then the in-band commit                                   [control warp]
//   read  : payload + scales --cp.async.bulk--> SMEM, dequantise to bf16 in SMEM,
//           --cp.async.bulk--> destination page.
//   fused read : the (otherwise idle) third warp is the RESOLVER of resolve.cuh - it hashes the
//           keys, probes the HBM index and feeds {pool block, destination} to loader and storer
//           through the shared-memory queue, then re-checks the entries after the copy:
//           read_cache_fp8 through the device index is ONE launch instead of lookup +
//           read + validate.
// Compared with kv_fp8.cu (every thread loads 16 bytes, stores 8, one scalar store per row
// for the scale): the fabric sees whole 8 KB / 256 B bulk requests instead of 8-byte stores,
// the scales of a tile travel as one vector, and the global loads no longer sit in the
// dependency chain of the conversion.  Half the NVLink bytes of a bf16 write and no separate
// cast kernel - the reference moves opaque bytes only (infinistore/lib.py:377-379).
//
// Pool block layout (unchanged): [elems x e4m3][elems/128 x fp32 scale].
// Needs elems % 512 == 0 (scale vectors of a tile are 16-byte multiples) and 16-byte aligned
// pages; launch_kv_{write,read}_fp8 fall back to kv_fp8.cu otherwise.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <mutex>
#include <type_traits>

#include "balance.h"
#include "common.cuh"
#include "kernels.h"
#include "publish.cuh"
#include "resolve.cuh"

namespace istore::kernels {

namespace {

using namespace dev;

constexpr uint32_t kRow = 128;                  // elements sharing a scale
constexpr uint32_t kTileRows = 64;              // rows per tile
constexpr uint32_t kTileElems = kRow * kTileRows;
constexpr uint32_t kTileBf16 = kTileElems * 2;  // 16 KB
constexpr uint32_t kTileQ = kTileElems;         // 8 KB
constexpr uint32_t kTileScale = kTileRows * 4;  // 256 B
constexpr int kInStages = 4, kOutStages = 3, kLag = 2;
constexpr float kE4m3Max = 448.f;

__device__ __forceinline__ uint16_t cvt_e4m3x2(float hi, float lo) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_f16x2_e4m3x2(uint16_t v) {
    uint32_t r;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(v));
    return r;
}
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(float lo, float hi) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&b);
}

struct Smem {
    // [kInStages][in tile] then [kOutStages][out tile]; sizes depend on the direction
    uint64_t full_in[kInStages], empty_in[kInStages], full_out[kOutStages], empty_out[kOutStages];
};

// Work schedule: item k of a CTA = chunk (item % cpb) of page (item / cpb); a chunk is a run
// of whole tiles (the last tile of a page may be short: rows % 4 == 0 is guaranteed).
struct Fp8Shape {
    uint32_t n, elems, chunk_elems, cpb;
};

// NCW compute warps (4 or 8) besides the loader, storer and control / resolver warps.
// FUSED (reads only): descriptors come from the resolver warp's queue instead of `descs`.
template <bool WRITE, int NCW, bool FUSED>
__global__ void __launch_bounds__((3 + NCW) * 32)
    kv_fp8_pipe_kernel(const CopyDesc* __restrict__ descs, const Fp8Shape sh, Publish pub,
                       uint32_t* status, const __grid_constant__ ResolveArgs ra) {
    static_assert(!(WRITE && FUSED), "only reads resolve keys");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ Smem bars;
    __shared__ __align__(16) ResolveQueue rq;
    constexpr uint32_t kInTile = WRITE ? kTileBf16 : (kTileQ + kTileScale);
    constexpr uint32_t kOutTile = WRITE ? (kTileQ + kTileScale) : kTileBf16;
    uint8_t* in_ring = smem;
    uint8_t* out_ring = smem + kInStages * kInTile;
    const uint32_t total = sh.n * sh.cpb;
    const uint32_t grid = gridDim.x;
    const uint32_t nitems = blockIdx.x < total ? (total - blockIdx.x + grid - 1) / grid : 0;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kInStages; ++s) {
            mbar_init(&bars.full_in[s], 1);
            mbar_init(&bars.empty_in[s], NCW);
        }
        for (int t = 0; t < kOutStages; ++t) {
            mbar_init(&bars.full_out[t], NCW);
            mbar_init(&bars.empty_out[t], 1);
        }
        if (FUSED) resolve_queue_init(rq);
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == 2) {  // control warp: in-band commit (writes) / resolver (fused reads)
        if (WRITE && pub.recs) control_warp(pub, lane, blockIdx.x, nitems, grid, sh.cpb, 64);
        if (FUSED)
            resolver_warp(ra, rq, sh.elems + sh.elems / kRow * 4, blockIdx.x, grid,
                          sh.cpb, nitems, lane);
        return;
    }
    // Every role walks the same sequence of tiles.  `role`: 0 loader, 1 storer (hands queue
    // halves back to the resolver), 2 compute (never looks at a descriptor when FUSED: a
    // missing block's tile is converted from whatever the slot holds and not stored).
    auto for_each_tile = [&](auto role, auto&& fn) {
        [[maybe_unused]] constexpr int kRole = decltype(role)::value;
        uint32_t q = 0;  // running tile number of this CTA
        for (uint32_t k = 0; k < nitems; ++k) {
            const uint32_t item = blockIdx.x + k * grid;
            CopyDesc d;
            if constexpr (!FUSED)
                d = descs[item / sh.cpb];
            else if constexpr (kRole == 2)
                d = CopyDesc{1, 0};
            else  // one lane per role makes this call: not warp wide
                d = resolved_desc<kRole == 1, false>(rq, k, 0);
            const uint32_t e0 = (item % sh.cpb) * sh.chunk_elems;
            const uint32_t e1 = min(sh.elems, e0 + sh.chunk_elems);
            for (uint32_t e = e0; e < e1; e += kTileElems, ++q)
                fn(q, d, e, min(kTileElems, e1 - e) / kRow, e == 0);
        }
    };
    if (warp == 0) {  // ---- loader
        if (lane != 0) return;
        for_each_tile(std::integral_constant<int, 0>{},
                      [&](uint32_t q, const CopyDesc& d, uint32_t e, uint32_t rows, bool) {
            const uint32_t s = q % kInStages, use = q / kInStages;
            if (use) mbar_wait(&bars.empty_in[s], (use - 1) & 1);
            uint8_t* tile = in_ring + s * kInTile;
            if (!d.src) {
                mbar_arrive(&bars.full_in[s]);
                return;
            }
            const uint8_t* src = reinterpret_cast<const uint8_t*>(d.src);
            if constexpr (WRITE) {
                mbar_expect_tx(&bars.full_in[s], rows * kRow * 2);
                bulk_g2s(tile, src + size_t(e) * 2, rows * kRow * 2, &bars.full_in[s]);
            } else {
                mbar_expect_tx(&bars.full_in[s], rows * kRow + rows * 4);
                bulk_g2s(tile, src + e, rows * kRow, &bars.full_in[s]);
                bulk_g2s(tile + kTileQ, src + sh.elems + size_t(e / kRow) * 4, rows * 4,
                         &bars.full_in[s]);
            }
        });
        return;
    }
    if (warp == 1) {  // ---- storer
        if (lane == 0) {
            uint32_t released = 0;
            for_each_tile(std::integral_constant<int, 1>{},
                          [&](uint32_t q, const CopyDesc& d, uint32_t e, uint32_t rows, bool first) {
                const uint32_t t = q % kOutStages;
                if (!d.src && first && status) atomicAdd(status + kStatMiss, 1u);
                mbar_wait(&bars.full_out[t], (q / kOutStages) & 1);
                const uint8_t* tile = out_ring + t * kOutTile;
                if (d.src) {
                    uint8_t* dst = reinterpret_cast<uint8_t*>(d.dst);
                    if constexpr (WRITE) {
                        bulk_s2g(dst + e, tile, rows * kRow);
                        bulk_s2g(dst + sh.elems + size_t(e / kRow) * 4, tile + kTileQ, rows * 4);
                    } else {
                        bulk_s2g(dst + size_t(e) * 2, tile, rows * kRow * 2);
                    }
                }
                bulk_commit();
                bulk_wait_read<kLag>();
                if (q >= uint32_t(kLag)) {
                    mbar_arrive(&bars.empty_out[released % kOutStages]);
                    ++released;
                }
            });
            bulk_wait<0>();
            fence_proxy_async();
            if (FUSED) resolved_done(rq, nitems, 0);
        }
        __syncwarp();
        if (WRITE && pub.recs) ctrl_barrier_arrive(64);
        return;
    }
    // ---- compute warps: half-warp per row, two rows per step
    const uint32_t cw = warp - 3;
    const uint32_t half = lane >> 4, hl = lane & 15;
    for_each_tile(std::integral_constant<int, 2>{},
                  [&](uint32_t q, const CopyDesc& d, uint32_t, uint32_t rows, bool) {
        const uint32_t s = q % kInStages, t = q % kOutStages;
        mbar_wait(&bars.full_in[s], (q / kInStages) & 1);
        if (q >= uint32_t(kOutStages)) mbar_wait(&bars.empty_out[t], (q / kOutStages - 1) & 1);
        const uint8_t* in = in_ring + s * kInTile;
        uint8_t* out = out_ring + t * kOutTile;
        if (d.src) {
            for (uint32_t r = cw * 2 + half; r < rows; r += NCW * 2) {
                if constexpr (WRITE) {
                    const uint4 v = *reinterpret_cast<const uint4*>(in + (size_t(r) * kRow + hl * 8) * 2);
                    float2 f[4] = {bf16x2_to_f2(v.x), bf16x2_to_f2(v.y), bf16x2_to_f2(v.z),
                                   bf16x2_to_f2(v.w)};
                    float amax = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) amax = fmaxf(amax, fmaxf(fabsf(f[j].x), fabsf(f[j].y)));
                    // the 16 lanes of this half own the row; rows r and r^1 of the two halves
                    // are both live or the loop bound differs only in the last step
                    const uint32_t mask = half ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(mask, amax, o));
                    const float scale = amax > 0.f ? amax * (1.f / kE4m3Max) : 1.f;
                    const float inv = 1.f / scale;
                    uint2 o2;
                    o2.x = uint32_t(cvt_e4m3x2(f[0].y * inv, f[0].x * inv)) |
                           (uint32_t(cvt_e4m3x2(f[1].y * inv, f[1].x * inv)) << 16);
                    o2.y = uint32_t(cvt_e4m3x2(f[2].y * inv, f[2].x * inv)) |
                           (uint32_t(cvt_e4m3x2(f[3].y * inv, f[3].x * inv)) << 16);
                    *reinterpret_cast<uint2*>(out + size_t(r) * kRow + hl * 8) = o2;
                    if (hl == 0) *reinterpret_cast<float*>(out + kTileQ + r * 4) = scale;
                } else {
                    const uint2 v = *reinterpret_cast<const uint2*>(in + size_t(r) * kRow + hl * 8);
                    const float sc = *reinterpret_cast<const float*>(in + kTileQ + r * 4);
                    const uint32_t w[2] = {v.x, v.y};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t h01 = cvt_f16x2_e4m3x2(uint16_t(w[j] & 0xffffu));
                        const uint32_t h23 = cvt_f16x2_e4m3x2(uint16_t(w[j] >> 16));
                        const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
                        const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
                        o[2 * j] = f2_to_bf16x2(f01.x * sc, f01.y * sc);
                        o[2 * j + 1] = f2_to_bf16x2(f23.x * sc, f23.y * sc);
                    }
                    *reinterpret_cast<uint4*>(out + (size_t(r) * kRow + hl * 8) * 2) =
                        make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        // this warp's reads of the input tile and writes of the output tile are done: make the
        // generic-proxy writes visible to the bulk store, then signal both rings
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&bars.empty_in[s]);
            mbar_arrive(&bars.full_out[t]);
        }
    });
}

std::mutex g_mu;
bool g_attr[64] = {false};

constexpr size_t kSmemWrite = kInStages * kTileBf16 + kOutStages * (kTileQ + kTileScale);
constexpr size_t kSmemRead = kInStages * (kTileQ + kTileScale) + kOutStages * kTileBf16;

cudaError_t ensure_attrs() {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    if (dev < 0 || dev >= 64 || g_attr[dev]) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    auto set = [&](auto* fn, size_t bytes) {
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    };
    set(kv_fp8_pipe_kernel<true, 4, false>, kSmemWrite);
    set(kv_fp8_pipe_kernel<true, 8, false>, kSmemWrite);
    set(kv_fp8_pipe_kernel<false, 4, false>, kSmemRead);
    set(kv_fp8_pipe_kernel<false, 8, false>, kSmemRead);
    set(kv_fp8_pipe_kernel<false, 8, true>, kSmemRead);
    if (e != cudaSuccess) return e;
    g_attr[dev] = true;
    return cudaSuccess;
}

Fp8Shape shape_of(const Fp8Launch& a, int ctas) {
    // whole pages per CTA when that keeps the grid evenly busy (single-CTA commit), else
    // chunks of at least two tiles (balance.h); never more than 512 K elements per item
    const uint32_t tiles = (a.elems + kTileElems - 1) / kTileElems;
    const ChunkPlan plan = plan_chunks(a.n, tiles, 2, 64, uint32_t(ctas));
    const uint32_t chunk = std::min(a.elems, plan.chunk * kTileElems);
    return Fp8Shape{a.n, a.elems, chunk, (a.elems + chunk - 1) / chunk};
}

}  // namespace

bool fp8_pipe_supported(const Fp8Launch& a) {
    return a.group == kRow && a.elems % 512 == 0 && a.elems > 0 && a.aligned16;
}

cudaError_t launch_kv_fp8_pipe(const Fp8Launch& a, bool write, cudaStream_t stream) {
    if (a.n == 0) return cudaSuccess;
    if (!fp8_pipe_supported(a)) return cudaErrorInvalidValue;
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    Publish pub{a.recs, a.table, a.table_mask, a.done, a.status, a.n, nullptr, !a.all_local, 0,
                a.shards};
    if (!write || !a.table || !a.done) pub.recs = nullptr;
    const int sms = sm_count();
    int ctas = a.max_ctas > 0 ? std::min(a.max_ctas, 2 * sms) : 2 * sms;  // ~90 KB smem: 2 per SM
    const Fp8Shape sh = shape_of(a, ctas);
    ctas = int(std::min<uint64_t>(uint64_t(ctas), uint64_t(sh.n) * sh.cpb));
    const bool wide = a.variant != 2;  // 8 compute warps unless variant 2 asks for 4 (A/B)
    const ResolveArgs none{};
    if (write && wide)
        kv_fp8_pipe_kernel<true, 8, false><<<ctas, (3 + 8) * 32, kSmemWrite, stream>>>(a.descs, sh, pub, a.status, none);
    else if (write)
        kv_fp8_pipe_kernel<true, 4, false><<<ctas, (3 + 4) * 32, kSmemWrite, stream>>>(a.descs, sh, pub, a.status, none);
    else if (wide)
        kv_fp8_pipe_kernel<false, 8, false><<<ctas, (3 + 8) * 32, kSmemRead, stream>>>(a.descs, sh, pub, a.status, none);
    else
        kv_fp8_pipe_kernel<false, 4, false><<<ctas, (3 + 4) * 32, kSmemRead, stream>>>(a.descs, sh, pub, a.status, none);
    return cudaGetLastError();
}

// read_cache_fp8 through the device index in one launch: keys -> (resolver warp) -> pool
// blocks -> dequantised pages.  Same contract as launch_kv_read_fused (misses and entries that
// changed under the copy are counted in status[kStatMiss] / [kStatStale]).
bool fp8_read_fused_supported(const ReadFusedLaunch& a, uint32_t elems) {
    return elems > 0 && elems % 512 == 0 && (a.align_or & 15) == 0;
}

cudaError_t launch_kv_fp8_read_fused(const ReadFusedLaunch& a, uint32_t elems, cudaStream_t stream) {
    if (a.n == 0) return cudaSuccess;
    if (!fp8_read_fused_supported(a, elems)) return cudaErrorInvalidValue;
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    ResolveArgs r{};
    r.key_bytes = a.key_bytes;
    r.key_off = a.key_off;
    r.key_len = a.key_len;
    r.dst_off = a.dst_off;
    r.dst_base = a.dst_base;
    r.table = a.table;
    r.table_mask = a.table_mask;
    r.shards = a.shards;
    r.nsegs = a.nsegs;
    for (uint32_t s = 0; s < a.nsegs && s < uint32_t(ReadFusedLaunch::kMaxSegs); ++s)
        r.seg_base[s] = a.seg_base[s];
    r.status = a.status;
    Fp8Launch shape_in;
    shape_in.n = a.n;
    shape_in.elems = elems;
    const int sms = sm_count();
    int ctas = a.max_ctas > 0 ? std::min(a.max_ctas, 2 * sms) : 2 * sms;
    const Fp8Shape sh = shape_of(shape_in, ctas);
    ctas = int(std::min<uint64_t>(uint64_t(ctas), uint64_t(sh.n) * sh.cpb));
    Publish pub{};
    pub.recs = nullptr;
    kv_fp8_pipe_kernel<false, 8, true><<<ctas, (3 + 8) * 32, kSmemRead, stream>>>(nullptr, sh, pub, a.status, r);
    return cudaGetLastError();
}

}  // namespace istore::kernels
