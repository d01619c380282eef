// Latency mode of the client: the host side of the doorbell worker (kernels/kv_doorbell.cu).
#include "client.h"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstring>

#include "../core/log.h"
#include "../core/trace.h"
#include "../kernels/kernels.h"
#include "../wire/messages.h"
#include "client_dev.h"

namespace istore {

// ---------------------------------------------------------------- doorbell worker
// Latency mode (ClientConfig::doorbell).  All of this runs under mu_.

namespace {
inline uint64_t db_done(const kernels::DoorbellCtl* c) {
    return *reinterpret_cast<const volatile uint64_t*>(&c->done_seq);
}
// Has the launch `epoch` said good-bye?  *next = the first request it did not serve.
inline bool db_exited(const kernels::DoorbellCtl* c, uint32_t epoch, uint64_t* next) {
    const uint64_t st = *reinterpret_cast<const volatile uint64_t*>(&c->state);
    if (uint32_t(st >> 44) != (epoch & 0xfffffu) || (st & 3) != kernels::kDoorbellExited) return false;
    *next = (st >> 2) & ((1ull << 42) - 1);
    return true;
}
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}
}  // namespace

// May this operation go through the worker?  The worker is ordered behind no stream, so the
// caller's data must be ready and this connection's earlier launches complete.
bool Connection::doorbell_ready(DevCtx* ctx, uint64_t user_stream, size_t bytes) {
    if (!cfg_.doorbell || streams_ <= 0 || bytes == 0 || bytes > kernels::kDoorbellMaxBytes)
        return false;
    if (!server_hbm_ || segs_.empty()) return false;
    if (cudaStreamQuery(reinterpret_cast<cudaStream_t>(user_stream)) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    for (cudaStream_t s : ctx->busy) {
        if (cudaStreamQuery(s) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
    }
    return true;
}

// (Re)launch the worker with the current view of the pool and the index; it serves
// requests from `db.posted`'s successor of what has been served so far.
int Connection::doorbell_start(DevCtx* ctx) {
    DevCtx::Doorbell& db = *ctx->db;
    kernels::DoorbellLaunch L;
    L.ring = db.ring_d;
    L.ctl = db.ctl_d;
    L.slots = uint32_t(kernels::kDoorbellMaxSlots);
    L.idle_ns = uint64_t(std::max(cfg_.doorbell_idle_us, 10)) * 1000;
    L.nsegs = uint32_t(std::min<size_t>(segs_.size(), kernels::DoorbellLaunch::kMaxSegs));
    for (uint32_t s = 0; s < L.nsegs; ++s) {
        // HBM segments and the local replica of an NVLS region; a host-shm tier is not
        // device-addressable (a block there reads as a miss and the caller falls back)
        uint8_t* base = nullptr;
        if (segs_[s].kind != kSegHostShm) base = seg_dev_ptr(ctx, s);
        L.seg_base[s] = reinterpret_cast<uint64_t>(base);
    }
    if (segs_[0].index_slots && L.seg_base[0]) {
        L.table = reinterpret_cast<kernels::IndexBucket*>(L.seg_base[0] + segs_[0].index_off);
        L.table_mask = kernels::index_bucket_mask(segs_[0].index_slots);
        L.shards = index_shards(ctx, nullptr);
    }
    L.epoch = ++db.epoch;
    L.first_seq = db.next_serve;
    const cudaError_t e = kernels::launch_kv_doorbell(L, db.stream);
    if (e != cudaSuccess) {
        fail(std::string("doorbell worker failed to launch: ") + cudaGetErrorString(e));
        return -1;
    }
    db.nsegs = segs_.size();
    db.running = true;
    stats_.doorbell_launches++;
    stats_.kernel_launches++;
    return 0;
}

// Post one request; q = {local address, pool address, h1, h2, block address, gen | bytes << 32}.
int Connection::doorbell_post(DevCtx* ctx, uint32_t op, const uint64_t (&q)[6]) {
    if (!ctx->db) {
        auto db = std::make_unique<DevCtx::Doorbell>();
        void* dp = nullptr;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const size_t ring_bytes = sizeof(kernels::DoorbellReq) * kernels::kDoorbellMaxSlots;
        if (cudaHostAlloc(reinterpret_cast<void**>(&db->ring_h), ring_bytes,
                          cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
            cudaHostGetDevicePointer(&dp, db->ring_h, 0) != cudaSuccess) {
            fail(std::string("doorbell ring: ") + cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
        db->ring_d = static_cast<kernels::DoorbellReq*>(dp);
        if (cudaHostAlloc(reinterpret_cast<void**>(&db->ctl_h), sizeof(kernels::DoorbellCtl),
                          cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
            cudaHostGetDevicePointer(&dp, db->ctl_h, 0) != cudaSuccess ||
            cudaStreamCreateWithPriority(&db->stream, cudaStreamNonBlocking, hi) != cudaSuccess) {
            fail(std::string("doorbell control block: ") + cudaGetErrorString(cudaGetLastError()));
            if (db->ring_h) cudaFreeHost(db->ring_h);
            if (db->ctl_h) cudaFreeHost(db->ctl_h);
            return -1;
        }
        db->ctl_d = static_cast<kernels::DoorbellCtl*>(dp);
        std::memset(db->ring_h, 0, ring_bytes);
        std::memset(db->ctl_h, 0, sizeof(kernels::DoorbellCtl));
        ctx->db = std::move(db);
    }
    DevCtx::Doorbell& db = *ctx->db;
    // a status word is per ring slot: take the finished ones before a slot comes round again
    if (db.posted + 2 - db.collected >= uint64_t(kernels::kDoorbellMaxSlots) && doorbell_wait(ctx) != 0)
        return -1;
    uint64_t next = 0;
    if (db.running && db_exited(db.ctl_h, db.epoch, &next)) {  // idled out
        db.running = false;
        db.next_serve = next;
    }
    // the pool grew (new segments) since the worker was launched: it resolves reads with the
    // view it was launched with, so it is replaced
    if (db.running && db.nsegs != segs_.size()) doorbell_stop(ctx);
    const uint64_t seq = ++db.posted;
    doorbell_write_line(&db.ring_h[seq % kernels::kDoorbellMaxSlots], seq, op, q);
    if (!db.running && doorbell_start(ctx) != 0) return -1;
    ctx->dirty = true;
    stats_.doorbell_ops++;
    return 0;
}

void Connection::doorbell_collect(DevCtx* ctx) {
    DevCtx::Doorbell& db = *ctx->db;
    const uint64_t done = std::min(db_done(db.ctl_h), db.posted);
    for (uint64_t s = db.collected + 1; s <= done; ++s) {
        const uint32_t st = *reinterpret_cast<const volatile uint32_t*>(
            &db.ctl_h->status[s % kernels::kDoorbellMaxSlots]);
        if (st == kernels::kDoorbellMiss || st == kernels::kDoorbellStale) ++db.misses;
        if (st == kernels::kDoorbellStale) ++db.stale;
        if (st == kernels::kDoorbellIndexFull) ++db.publish_failures;
    }
    if (done > db.collected) db.collected = done;
}

// Every posted request has completed (0), or the worker is gone for good (-1).
int Connection::doorbell_wait(DevCtx* ctx) {
    DevCtx::Doorbell& db = *ctx->db;
    const uint64_t deadline = now_ns() + uint64_t(std::max(cfg_.timeout_ms, 1)) * 1000000ull;
    for (uint32_t spin = 0; db_done(db.ctl_h) < db.posted; ++spin) {
        uint64_t next = 0;
        if (db.running && db_exited(db.ctl_h, db.epoch, &next)) {
            // it left (idle timeout) without having seen the last request(s): again
            db.running = false;
            db.next_serve = next;
        }
        if (!db.running) {
            if (db_done(db.ctl_h) >= db.posted) break;
            if (doorbell_start(ctx) != 0) return -1;
        }
        cpu_relax();
        if ((spin & 4095) == 4095) {
            const cudaError_t q = cudaStreamQuery(db.stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) {
                fail(std::string("doorbell worker died: ") + cudaGetErrorString(q));
                db.running = false;
                return -1;
            }
            if (now_ns() > deadline) {
                fail("doorbell worker did not answer within the timeout");
                return -1;
            }
        }
    }
    doorbell_collect(ctx);
    return 0;
}

// Ordinary launches are ordered behind what the worker still has to do (a read kernel must
// see the blocks a doorbell write is publishing).
int Connection::doorbell_quiesce(DevCtx* ctx) {
    if (!ctx->db || ctx->db->posted <= ctx->db->collected) return 0;
    return doorbell_wait(ctx);
}

// Ask the worker to leave and wait until it has (requests posted before are served first).
void Connection::doorbell_stop(DevCtx* ctx) {
    if (!ctx->db) return;
    DevCtx::Doorbell& db = *ctx->db;
    uint64_t next = 0;
    if (db.running && db_exited(db.ctl_h, db.epoch, &next)) {
        db.running = false;
        db.next_serve = next;
    }
    if (!db.running) return;
    doorbell_collect(ctx);
    const uint64_t none[6] = {0, 0, 0, 0, 0, 0};
    const uint64_t stop_seq = ++db.posted;
    doorbell_write_line(&db.ring_h[stop_seq % kernels::kDoorbellMaxSlots], stop_seq,
                        kernels::kDoorbellStop, none);
    cudaStreamSynchronize(db.stream);  // STOP, or the idle timeout: it ends either way
    db.running = false;
    if (db_exited(db.ctl_h, db.epoch, &next)) {
        db.next_serve = next;
        if (next <= stop_seq) {
            // it idled out before it saw the STOP: the line is void (a later launch must not
            // find it), its number is reused
            *reinterpret_cast<volatile uint64_t*>(&db.ring_h[stop_seq % kernels::kDoorbellMaxSlots].q[0]) = 0;
            std::atomic_thread_fence(std::memory_order_seq_cst);
            db.posted = stop_seq - 1;
        }
    } else {
        db.next_serve = db.posted + 1;
    }
    doorbell_collect(ctx);
}

}  // namespace istore
