// Private to the client's translation units (client.cpp: control plane, client_data.cpp: data
// plane, client_doorbell.cpp: latency mode): per-device state of a connection and small helpers.
#pragma once

#include <cuda_runtime_api.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <vector>

#include "../fabric/segment.h"
#include "../kernels/kernels.h"
#include "client.h"

namespace istore {

inline constexpr size_t kBlobPayload = ~size_t(0);
inline constexpr size_t kRingBytes = 16u << 20;     // pinned, mapped staging ring per device
inline constexpr size_t kScratchBytes = 8u << 20;   // device scratch per device
inline constexpr size_t kZeroBytes = 2u << 20;      // self-cleaning zeroed counters per device
inline constexpr size_t kMaxBatch = 65536;          // blocks per kernel launch

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (dev >= 0 && dev != prev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) & ~(a - 1); }

inline uint64_t now_ns() {
    return uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(
                        std::chrono::steady_clock::now().time_since_epoch())
                        .count());
}

// A request line of the doorbell ring (kernels/kv_doorbell.cu): payload and checksum first,
// the word with the sequence number last - the worker accepts a line only when both match.
inline void doorbell_write_line(kernels::DoorbellReq* slot, uint64_t seq, uint32_t op,
                                const uint64_t (&q)[6]) {
    uint64_t line[8];
    line[0] = (seq << 2) | op;
    for (int i = 0; i < 6; ++i) line[1 + i] = q[i];
    uint64_t x = kernels::kDoorbellMagic;
    for (int i = 0; i < 7; ++i) x ^= line[i];
    line[7] = x;
    auto* dst = reinterpret_cast<volatile uint64_t*>(slot->q);
    for (int i = 1; i < 8; ++i) dst[i] = line[i];
    std::atomic_thread_fence(std::memory_order_release);
    dst[0] = line[0];
    std::atomic_thread_fence(std::memory_order_seq_cst);  // out of the store buffer now
}

// Per-device data-plane state.
struct Connection::DevCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    std::vector<std::shared_ptr<fabric::Mapping>> maps;  // by segment id
    std::vector<uint8_t*> seg_ptr;  // maps[i]->dev_ptr, cached for the per-block hot loop
    std::vector<uint8_t> seg_remote;  // 1 when the segment is not in this device's own HBM
    std::vector<uint8_t*> seg_mc;     // NVLS replica segments: multicast base (writes)
    uint8_t* ring_h = nullptr;  // pinned + mapped: descriptors, publish records, key bytes
    uint8_t* ring_d = nullptr;
    size_t ring_head = 0;
    uint8_t* scratch = nullptr;  // device memory: descriptors built by the lookup kernel
    size_t scratch_head = 0;
    uint8_t* zeros = nullptr;    // device memory kept zero between launches (counters/tickets)
    size_t zeros_head = 0;
    uint32_t* status_h = nullptr;
    uint32_t* status_d = nullptr;
    std::vector<cudaStream_t> busy;  // streams with launches since the last wait_all()
    bool dirty = false;

    // Doorbell worker (ClientConfig::doorbell): request ring + control block in pinned,
    // device-mapped host memory; `posted` / `collected` are request numbers.
    struct Doorbell {
        kernels::DoorbellReq* ring_h = nullptr;
        kernels::DoorbellReq* ring_d = nullptr;
        kernels::DoorbellCtl* ctl_h = nullptr;
        kernels::DoorbellCtl* ctl_d = nullptr;
        cudaStream_t stream = nullptr;
        uint64_t posted = 0;     // last request written to the ring
        uint64_t collected = 0;  // statuses of requests <= collected have been taken
        uint32_t epoch = 0;      // launch counter
        bool running = false;    // a launch of `epoch` has not been seen to exit
        uint64_t next_serve = 1;  // first request the next launch serves
        size_t nsegs = 0;         // pool segments known when the running worker was launched
        uint32_t misses = 0, stale = 0, publish_failures = 0;  // since the last drain
    };
    std::unique_ptr<Doorbell> db;

    // Launch streams.  Back-to-back page-mover kernels of one connection are independent of
    // each other, but in a single stream the fixed head (launch, descriptor fetch) and tail
    // (store acks, fence, commit) of every kernel are exposed: +5..30 us on a 45 us NVLink
    // launch (profiles/r1_launch_overhead_*.json).  Round-robin over a few internal streams
    // lets the tail of one kernel overlap the body of the next.  Ordering: every launch
    // waits for the caller's stream (the pages are ready); reads / lookups additionally wait
    // for earlier writes of this connection; completion is established by sync().
    static constexpr int kMaxStreams = 8;
    cudaStream_t pool[kMaxStreams] = {nullptr};
    cudaEvent_t pool_ev[kMaxStreams] = {nullptr};
    uint64_t last_write[kMaxStreams] = {0};
    uint64_t joined[kMaxStreams] = {0};
    uint64_t write_epoch = 0;
    cudaEvent_t user_ev = nullptr;
    int nstreams = 0;
    int rr = 0;

    // Stream for the next launch.  nstreams == 0: the caller's stream itself (in-stream
    // semantics, CUDA-graph capturable).
    cudaStream_t pick(cudaStream_t user, bool is_write, int want_streams) {
        if (want_streams <= 0) return user ? user : stream;
        if (nstreams < want_streams) {
            for (int i = nstreams; i < want_streams && i < kMaxStreams; ++i) {
                cudaStreamCreateWithFlags(&pool[i], cudaStreamNonBlocking);
                cudaEventCreateWithFlags(&pool_ev[i], cudaEventDisableTiming);
            }
            nstreams = std::min(want_streams, int(kMaxStreams));
            if (!user_ev) cudaEventCreateWithFlags(&user_ev, cudaEventDisableTiming);
        }
        const int i = rr++ % nstreams;
        cudaStream_t s = pool[i];
        if (user) {  // run after whatever produced the pages
            cudaEventRecord(user_ev, user);
            cudaStreamWaitEvent(s, user_ev, 0);
        }
        if (is_write) {
            last_write[i] = ++write_epoch;
        } else if (joined[i] < write_epoch) {  // reads see this connection's earlier writes
            for (int w = 0; w < nstreams; ++w) {
                if (w == i || last_write[w] <= joined[i]) continue;
                cudaEventRecord(pool_ev[w], pool[w]);
                cudaStreamWaitEvent(s, pool_ev[w], 0);
            }
            joined[i] = write_epoch;
        }
        return s;
    }

    ~DevCtx() {
        DeviceGuard g(device);
        wait_all();
        if (db) {
            // the worker leaves on a STOP request (or by itself after its idle timeout)
            if (db->running && db->ring_h) {
                const uint64_t none[6] = {0, 0, 0, 0, 0, 0};
                const uint64_t seq = ++db->posted;
                doorbell_write_line(&db->ring_h[seq % kernels::kDoorbellMaxSlots], seq,
                                    kernels::kDoorbellStop, none);
            }
            if (db->stream) {
                cudaStreamSynchronize(db->stream);
                cudaStreamDestroy(db->stream);
            }
            if (db->ring_h) cudaFreeHost(db->ring_h);
            if (db->ctl_h) cudaFreeHost(db->ctl_h);
        }
        for (int i = 0; i < nstreams; ++i) {
            cudaStreamSynchronize(pool[i]);
            cudaStreamDestroy(pool[i]);
            cudaEventDestroy(pool_ev[i]);
        }
        if (user_ev) cudaEventDestroy(user_ev);
        if (stream) {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
        maps.clear();
        if (ring_h) cudaFreeHost(ring_h);
        if (status_h) cudaFreeHost(status_h);
        if (scratch) cudaFree(scratch);
        if (zeros) cudaFree(zeros);
    }

    // Completion of everything launched so far.  A blocking cudaStreamSynchronize sleeps on an
    // interrupt (+5..10 us for a transfer that itself takes 10 us); short transfers are
    // therefore polled with cudaStreamQuery for a bounded time first.
    void wait_all() {
        DeviceGuard g(device);
        for (cudaStream_t s : busy) {
            bool done = false;
            const uint64_t t0 = now_ns();
            for (int spin = 0; spin < 4096; ++spin) {
                const cudaError_t q = cudaStreamQuery(s);
                if (q != cudaErrorNotReady) {  // finished, or failed: let synchronize report it
                    done = q == cudaSuccess;
                    break;
                }
                if ((spin & 15) == 15 && now_ns() - t0 > 60000) break;  // 60 us: not a short one
            }
            if (!done) cudaStreamSynchronize(s);
        }
        busy.clear();
        dirty = false;
    }
    cudaStream_t last = nullptr;  // stream of the most recent launch
    void mark(cudaStream_t s) {
        dirty = true;
        last = s;
        for (cudaStream_t b : busy)
            if (b == s) return;
        busy.push_back(s);
    }
    // Bump allocators.  When a region wraps, everything launched from it must be done.
    size_t ring_alloc(size_t bytes) {
        bytes = align_up(bytes, 64);
        if (ring_head + bytes > kRingBytes) {
            wait_all();
            ring_head = 0;
        }
        const size_t at = ring_head;
        ring_head += bytes;
        return at;
    }
    size_t scratch_alloc(size_t bytes) {
        bytes = align_up(bytes, 256);
        if (scratch_head + bytes > kScratchBytes) {
            wait_all();
            scratch_head = 0;
        }
        const size_t at = scratch_head;
        scratch_head += bytes;
        return at;
    }
    size_t zeros_alloc(size_t bytes) {
        bytes = align_up(bytes, 256);
        if (zeros_head + bytes > kZeroBytes) {
            wait_all();
            zeros_head = 0;
        }
        const size_t at = zeros_head;
        zeros_head += bytes;
        return at;
    }
};

}  // namespace istore
