#include "client.h"

#include <arpa/inet.h>
#include <cuda_runtime_api.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/uio.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstring>

#include "../core/log.h"
#include "../core/trace.h"
#include "../kernels/kernels.h"
#include "../wire/messages.h"

namespace istore {

namespace {

constexpr size_t kBlobPayload = ~size_t(0);
constexpr size_t kRingBytes = 16u << 20;     // pinned, mapped staging ring per device
constexpr size_t kScratchBytes = 8u << 20;   // device scratch per device
constexpr size_t kZeroBytes = 2u << 20;      // self-cleaning zeroed counters per device
constexpr size_t kMaxBatch = 65536;          // blocks per kernel launch

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

bool send_all(int fd, const iovec* iov_in, int iovcnt) {
    iovec iov[4];
    for (int i = 0; i < iovcnt; ++i) iov[i] = iov_in[i];
    int first = 0;
    while (first < iovcnt) {
        msghdr mh{};
        mh.msg_iov = iov + first;
        mh.msg_iovlen = size_t(iovcnt - first);
        ssize_t n = sendmsg(fd, &mh, MSG_NOSIGNAL);
        if (n < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        while (n > 0 && first < iovcnt) {
            if (size_t(n) >= iov[first].iov_len) {
                n -= ssize_t(iov[first].iov_len);
                ++first;
            } else {
                iov[first].iov_base = static_cast<uint8_t*>(iov[first].iov_base) + n;
                iov[first].iov_len -= size_t(n);
                n = 0;
            }
        }
        while (first < iovcnt && iov[first].iov_len == 0) ++first;
    }
    return true;
}

bool recv_all(int fd, void* buf, size_t len) {
    uint8_t* p = static_cast<uint8_t*>(buf);
    while (len) {
        const ssize_t n = recv(fd, p, len, 0);
        if (n == 0) return false;
        if (n < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += n;
        len -= size_t(n);
    }
    return true;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) & ~(a - 1); }

inline uint64_t now_ns() {
    return uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(
                        std::chrono::steady_clock::now().time_since_epoch())
                        .count());
}

}  // namespace

void PendingHashes::grow() {
    std::vector<Entry> old;
    old.swap(slots_);
    slots_.assign(old.empty() ? 1024 : old.size() * 2, Entry{});
    count_ = 0;
    for (auto& sl : old)
        if (sl.addr) map_put(sl.addr, sl.h);
}

void PendingHashes::map_put(uint64_t addr, const KeyHash& h) {
    if ((count_ + 1) * 2 > slots_.size()) grow();
    const size_t mask = slots_.size() - 1;
    size_t i = mix(addr) & mask;
    while (slots_[i].addr && slots_[i].addr != addr) i = (i + 1) & mask;
    if (!slots_[i].addr) ++count_;
    slots_[i].addr = addr;
    slots_[i].h = h;
}

bool PendingHashes::map_take(uint64_t addr, KeyHash* out) {
    if (slots_.empty() || count_ == 0) return false;
    const size_t mask = slots_.size() - 1;
    size_t i = mix(addr) & mask;
    while (slots_[i].addr != addr) {
        if (!slots_[i].addr) return false;
        i = (i + 1) & mask;
    }
    *out = slots_[i].h;
    // backward-shift deletion keeps probe sequences intact without tombstones
    size_t j = i;
    for (;;) {
        j = (j + 1) & mask;
        if (!slots_[j].addr) break;
        const size_t home = mix(slots_[j].addr) & mask;
        if ((i <= j) ? (home <= i || home > j) : (home <= i && home > j)) {
            slots_[i] = slots_[j];
            i = j;
        }
    }
    slots_[i].addr = 0;
    --count_;
    return true;
}

void PendingHashes::spill() {
    for (size_t i = head_; i < fifo_.size(); ++i) map_put(fifo_[i].addr, fifo_[i].h);
    fifo_.clear();
    head_ = 0;
}

bool PendingHashes::take(uint64_t addr, KeyHash* out) {
    if (head_ < fifo_.size()) {
        if (fifo_[head_].addr == addr) {  // in allocation order: the common case
            *out = fifo_[head_].h;
            if (++head_ == fifo_.size()) {
                fifo_.clear();
                head_ = 0;
            }
            return true;
        }
        spill();  // out of order: from here on look the blocks up by address
    }
    return map_take(addr, out);
}

// A request line of the doorbell ring (kernels/kv_doorbell.cu): payload and checksum first,
// the word with the sequence number last - the worker accepts a line only when both match.
static void doorbell_write_line(kernels::DoorbellReq* slot, uint64_t seq, uint32_t op,
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
        uint64_t signature = 0;   // pool / index view the running worker was launched with
        size_t nsegs = 0;         // segments known at that launch
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

struct Connection::Task {
    enum Kind { kAllocate, kWaitEvent, kStop } kind = kWaitEvent;
    // allocate
    std::vector<std::string> keys;
    int block_size = 0;
    std::function<void(std::vector<RemoteBlock>)> alloc_cb;
    // wait for device work, then commit + callback
    int device = -1;
    cudaEvent_t event = nullptr;
    int status = 0;
    bool commit = false;
    std::vector<uint64_t> commits;  // addresses written by exactly this task's launches
    std::function<void(int)> done_cb;
};

Connection::Connection() {}

Connection::~Connection() { close(); }

void Connection::fail(const std::string& msg) {
    last_error_ = msg;
    LOG_ERROR("%s", msg.c_str());
}

ClientStats Connection::stats() const { return stats_; }

// ---------------------------------------------------------------- control plane

int Connection::init_connection(const ClientConfig& cfg) {
    cfg_ = cfg;
    if (fd_ >= 0) return 0;
    addrinfo hints{};
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    addrinfo* res = nullptr;
    const std::string port = std::to_string(cfg.service_port);
    if (getaddrinfo(cfg.host_addr.c_str(), port.c_str(), &hints, &res) != 0 || !res) {
        fail("cannot resolve " + cfg.host_addr);
        return -1;
    }
    int fd = -1;
    for (addrinfo* ai = res; ai; ai = ai->ai_next) {
        fd = socket(ai->ai_family, ai->ai_socktype | SOCK_CLOEXEC, ai->ai_protocol);
        if (fd < 0) continue;
        if (connect(fd, ai->ai_addr, ai->ai_addrlen) == 0) break;
        ::close(fd);
        fd = -1;
    }
    freeaddrinfo(res);
    if (fd < 0) {
        fail("cannot connect to " + cfg.host_addr + ":" + port + ": " + std::strerror(errno));
        return -1;
    }
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    timeval tv{};
    tv.tv_sec = cfg.timeout_ms / 1000;
    tv.tv_usec = (cfg.timeout_ms % 1000) * 1000;
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
    fd_ = fd;

    ConnInfo me{};
    me.qpn = uint32_t(getpid());
    me.psn = cfg.device >= 0 ? uint32_t(cfg.device) : 0xffffffffu;
    std::memcpy(me.gid, fabric::process_uuid(), 16);
    me.lid = fabric::cuda_available() ? 1 : 0;
    me.mtu = kFabricVersion;
    int32_t code = 0;
    std::vector<uint8_t> payload;
    if (transact(kOpExchange, &me, sizeof(me), &code, &payload, sizeof(ConnInfo)) != 0 ||
        code != kFinish) {
        fail("fabric exchange with the server failed");
        close();
        return -1;
    }
    ConnInfo srv{};
    std::memcpy(&srv, payload.data(), sizeof(srv));
    if (srv.mtu != kFabricVersion) {
        fail("server speaks fabric protocol v" + std::to_string(srv.mtu));
        close();
        return -1;
    }
    server_cuda_ = srv.lid & 1;
    server_hbm_ = srv.lid & 2;
    server_evicts_ = srv.lid & 4;
    index_incomplete_.store((srv.lid & 8) != 0, std::memory_order_relaxed);
    std::memcpy(server_uuid_, srv.gid, 16);
    if (!worker_.joinable()) {
        stop_ = false;
        worker_ = std::thread([this] { worker(); });
    }
    return 0;
}

int Connection::setup_rdma(const ClientConfig&) {
    if (fd_ < 0) return -1;
    // The HBM pool is the fast path: look keys up on the GPU unless told otherwise.
    device_lookup_ = false;
    return refresh_pool_map();
}

void Connection::close() {
    if (worker_.joinable()) {
        {
            std::lock_guard<std::mutex> lk(q_mu_);
            stop_ = true;
            Task t;
            t.kind = Task::kStop;
            queue_.push_back(std::move(t));
        }
        q_cv_.notify_all();
        worker_.join();
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& kv : devs_) kv.second->wait_all();
        devs_.clear();
        host_maps_.clear();
        for (auto& kv : host_regs_)
            if (kv.second.registered) cudaHostUnregister(reinterpret_cast<void*>(kv.first));
        host_regs_.clear();
    }
    if (fd_ >= 0) {
        ::close(fd_);
        fd_ = -1;
    }
}

// One request/response exchange.  fixed_payload: bytes following the code on success, or
// kBlobPayload for "u32 length + blob".
int Connection::transact(char op, const void* body, size_t len, int32_t* code,
                         std::vector<uint8_t>* payload, size_t fixed_payload,
                         const std::vector<uint8_t>* prefix) {
    std::lock_guard<std::mutex> lk(sock_mu_);
    if (fd_ < 0) return -1;
    Header h{kMagic, op, uint32_t(len)};
    // `prefix`: already framed reply-less messages (COMMIT) that go out in the same segment
    iovec iov[3];
    int niov = 0;
    if (prefix && !prefix->empty())
        iov[niov++] = iovec{const_cast<uint8_t*>(prefix->data()), prefix->size()};
    iov[niov++] = iovec{&h, sizeof(h)};
    if (len) iov[niov++] = iovec{const_cast<void*>(body), len};
    // A transaction that breaks half way (send error, reply timeout, short reply) leaves the
    // byte stream in an unknown position: a late reply would be taken for the answer to the
    // NEXT request.  The connection is closed instead; later calls fail fast.
    auto broken = [&](const std::string& why) {
        fail(why);
        ::shutdown(fd_, SHUT_RDWR);
        ::close(fd_);
        fd_ = -1;
        return -1;
    };
    if (!send_all(fd_, iov, niov))
        return broken(std::string("send ") + op_name(op) + ": " + std::strerror(errno));
    stats_.ctrl_requests++;
    if (!recv_all(fd_, code, sizeof(*code)))
        return broken(std::string("no reply to ") + op_name(op) +
                      " (timeout or connection closed): connection dropped");
    if (payload) payload->clear();
    if (*code != kFinish && *code != kTaskAccepted) return 0;  // error replies carry no payload
    if (fixed_payload == kBlobPayload) {
        uint32_t n = 0;
        if (!recv_all(fd_, &n, sizeof(n)) || n > kMaxBody + 4096)
            return broken(std::string("short reply to ") + op_name(op));
        payload->resize(n);
        if (n && !recv_all(fd_, payload->data(), n))
            return broken(std::string("short reply to ") + op_name(op));
    } else if (fixed_payload) {
        payload->resize(fixed_payload);
        if (!recv_all(fd_, payload->data(), fixed_payload))
            return broken(std::string("short reply to ") + op_name(op));
    }
    return 0;
}

int Connection::send_raw(const void* framed, size_t len) {
    std::lock_guard<std::mutex> lk(sock_mu_);
    if (fd_ < 0) return -1;
    iovec iov[1] = {{const_cast<void*>(framed), len}};
    if (!send_all(fd_, iov, 1)) return -1;
    stats_.ctrl_requests++;
    return 0;
}

int Connection::send_only(char op, const void* body, size_t len) {
    std::lock_guard<std::mutex> lk(sock_mu_);
    if (fd_ < 0) return -1;
    Header h{kMagic, op, uint32_t(len)};
    iovec iov[2] = {{&h, sizeof(h)}, {const_cast<void*>(body), len}};
    if (!send_all(fd_, iov, len ? 2 : 1)) return -1;
    stats_.ctrl_requests++;
    return 0;
}

int Connection::refresh_pool_map() {
    const uint32_t first = uint32_t(segs_.size());
    int32_t code = 0;
    std::vector<uint8_t> blob;
    if (transact(kOpPoolMap, &first, sizeof(first), &code, &blob, kBlobPayload) != 0 ||
        code != kFinish || blob.size() < 4)
        return -1;
    uint32_t count = 0;
    std::memcpy(&count, blob.data(), 4);
    if (blob.size() != 4 + size_t(count) * sizeof(SegmentInfo)) return -1;
    for (uint32_t i = 0; i < count; ++i) {
        SegmentInfo s;
        std::memcpy(&s, blob.data() + 4 + size_t(i) * sizeof(SegmentInfo), sizeof(s));
        segs_.push_back(s);
    }
    return 0;
}

// The kernels resolve at most kMaxSegs segment bases; a pool that auto-increased beyond that
// is served through the control plane (authoritative anyway).
bool Connection::device_index_usable() {
    if (segs_.empty() || !segs_[0].index_slots) return false;
    // Some writer could not publish a block in the HBM index (both of the key's buckets were
    // full): that key is reachable through the server only, so every read of this connection
    // resolves through the server from now on instead of reporting a false miss.
    if (index_incomplete_.load(std::memory_order_relaxed)) return false;
    return segs_.size() <= size_t(kernels::LookupLaunch::kMaxSegs);
}

int Connection::check_exist(const std::string& key) {
    if (device_lookup_ && server_hbm_ && device_index_usable()) {
        const int r = match_via_device_index({std::string_view(key)}, true);
        if (r >= -1) return r == 0 ? 0 : 1;
    }
    int32_t code = 0;
    std::vector<uint8_t> p;
    if (transact(kOpCheckExist, key.data(), key.size(), &code, &p, sizeof(int32_t)) != 0 ||
        code != kFinish)
        return -1;
    int32_t v;
    std::memcpy(&v, p.data(), sizeof(v));
    return v;
}

int Connection::get_match_last_index(const std::vector<std::string_view>& keys) {
    if (keys.empty()) return -1;
    if (device_lookup_ && server_hbm_ && device_index_usable()) {
        const int r = match_via_device_index(keys, false);
        if (r >= -1) return r;
    }
    const std::vector<std::string_view>& kv = keys;
    std::vector<uint8_t> buf(remote_meta_bound(kv, 0));
    fb::Builder b(buf.data(), buf.size());
    encode_match_request(b, kv);
    int32_t code = 0;
    std::vector<uint8_t> p;
    if (transact(kOpMatchLastIdx, b.data(), b.size(), &code, &p, sizeof(int32_t)) != 0 ||
        code != kFinish)
        return -2;
    int32_t v;
    std::memcpy(&v, p.data(), sizeof(v));
    return v;
}

// A device-side miss on a path that makes no SYNC round trip: ask the server whether the HBM
// index is complete - if a writer overflowed it, the key may exist all the same, and from now
// on this connection resolves its reads through the server.
void Connection::refresh_index_state() {
    if (index_incomplete_.load(std::memory_order_relaxed)) return;
    int32_t code = 0;
    std::vector<uint8_t> p;
    if (transact(kOpSync, nullptr, 0, &code, &p, sizeof(uint32_t)) == 0 && code == kFinish &&
        p.size() == sizeof(uint32_t)) {
        uint32_t remain;
        std::memcpy(&remain, p.data(), sizeof(remain));
        if (remain & kSyncIndexIncomplete) index_incomplete_.store(true, std::memory_order_relaxed);
    }
}

int Connection::sync_local() {
    NvtxRange nvtx("istore.sync");
    // one sync at a time: a staged commit list and the SYNC that applies it belong together
    std::lock_guard<std::mutex> sync_lk(sync_mu_);
    {
        // Nothing to tell the server (no commits pending, no leases held by host-mediated
        // lookups): completion of the kernels is all there is to wait for.
        bool quiet;
        {
            std::lock_guard<std::mutex> lk(mu_);
            quiet = pending_commit_.empty() && !ctrl_dirty_;
        }
        bool async_idle;
        {
            std::lock_guard<std::mutex> lk(q_mu_);
            async_idle = inflight_async_ == 0;
        }
        if (quiet && async_idle) {
            const int drained = drain_devices();
            if (drained == -kKeyNotFound) refresh_index_state();
            return drained != 0 ? drained : 0;
        }
    }
    if (cfg_.posted_commit) {
        bool leases;
        {
            std::lock_guard<std::mutex> lk(mu_);
            leases = ctrl_dirty_;
        }
        if (!leases) {
            // Posted commit: wait for the kernels, then send the commit list one-way.  The
            // blocks are already visible to device-path readers (in-band commit); the server's
            // map follows when the message arrives - no round trip on the caller's path.
            std::vector<uint64_t> mine;
            {
                std::lock_guard<std::mutex> lk(mu_);
                mine.swap(pending_commit_);
            }
            bool device_error = false;
            const int drained = drain_devices(&device_error);
            if (device_error) {
                (void)discard_blocks(mine.data(), mine.size());
                return -1;
            }
            if (!mine.empty() && send_commit(mine.data(), mine.size()) != 0) return -1;
            if (drained == -kKeyNotFound) refresh_index_state();
            return drained;
        }
    }
    // The commit list is taken BEFORE waiting for the GPU, so it names only blocks whose
    // kernels the drain below covers, and it is shipped right away as a STAGED commit: the
    // server decodes it and pulls the block headers into its cache while this thread waits
    // for the kernels; the SYNC that follows the drain applies it.  (Measured at N=1: the
    // commit of 8192 blocks was ~0.1 ms of a 0.5 ms write phase when it followed the drain.)
    std::vector<uint64_t> addrs;
    {
        std::lock_guard<std::mutex> lk(mu_);
        addrs.swap(pending_commit_);
    }
    constexpr size_t kInline = 128 * 1024;  // addresses; larger lists use the chunked path
    auto frame_of = [&](char op, int32_t block_size, const uint64_t* a, size_t n) {
        std::vector<uint8_t> buf(align_up(n * 8 + 128, 8));
        fb::Builder b(buf.data(), buf.size());
        encode_remote_meta(b, {}, block_size, block_size >= 0 ? take_publish_failures() : 0, a, n, op);
        std::vector<uint8_t> framed(sizeof(Header) + b.size());
        Header ch{kMagic, op, uint32_t(b.size())};
        std::memcpy(framed.data(), &ch, sizeof(ch));
        std::memcpy(framed.data() + sizeof(ch), b.data(), b.size());
        return framed;
    };
    bool staged = false;
    if (!addrs.empty() && addrs.size() <= kInline) {
        const std::vector<uint8_t> f = frame_of(kOpStageCommit, 0, addrs.data(), addrs.size());
        if (send_raw(f.data(), f.size()) != 0) {
            fail("commit: send failed");
            return -1;
        }
        staged = true;
    }
    bool device_error = false;
    const int drained = drain_devices(&device_error);
    if (device_error) {
        // The data of these blocks may not have landed, while the in-band commit may already
        // have published some of them in the device index: the server releases them now,
        // index entries included (reservations it cannot match die with the connection).
        if (staged) {
            const std::vector<uint8_t> f = frame_of(kOpStageCommit, -1, nullptr, 0);
            (void)send_raw(f.data(), f.size());
        } else {
            (void)discard_blocks(addrs.data(), addrs.size());
        }
        return -1;
    }
    // A read that missed (drained < 0 without a device error) does not undo the writes of
    // the same window: their kernels completed, so their commits are applied all the same.
    std::vector<uint8_t> framed;  // reply-less messages that travel with the SYNC
    if (!staged && !addrs.empty() && send_commit(addrs.data(), addrs.size()) != 0) return -1;
    if (const uint32_t pf = take_publish_failures()) {
        // index insertions that failed are known only now (after the drain): an empty COMMIT
        // carries the count, ahead of the SYNC in the same segment
        std::vector<uint8_t> buf(256);
        fb::Builder b(buf.data(), buf.size());
        encode_remote_meta(b, {}, 0, pf, nullptr, 0, kOpCommit);
        framed.resize(sizeof(Header) + b.size());
        Header ch{kMagic, kOpCommit, uint32_t(b.size())};
        std::memcpy(framed.data(), &ch, sizeof(ch));
        std::memcpy(framed.data() + sizeof(ch), b.data(), b.size());
    }
    int32_t code = 0;
    std::vector<uint8_t> p;
    if (transact(kOpSync, nullptr, 0, &code, &p, sizeof(uint32_t), &framed) != 0 ||
        code != kFinish)
        return -1;
    {
        std::lock_guard<std::mutex> lk(mu_);
        ctrl_dirty_ = false;
    }
    if (drained != 0) return drained;
    uint32_t remain;
    std::memcpy(&remain, p.data(), sizeof(remain));
    if (remain & kSyncIndexIncomplete) index_incomplete_.store(true, std::memory_order_relaxed);
    return int(remain & ~kSyncIndexIncomplete);
}

int Connection::sync_rdma() {
    {  // async operations first: their completions append to the commit list
        std::unique_lock<std::mutex> lk(q_mu_);
        if (!idle_cv_.wait_for(lk, std::chrono::milliseconds(cfg_.timeout_ms),
                               [this] { return inflight_async_ == 0; })) {
            fail("sync: timed out waiting for asynchronous operations");
            return -1;
        }
    }
    const int r = sync_local();
    return r < 0 ? r : 0;
}

int Connection::send_commit(const uint64_t* addrs, size_t count) {
    // chunk so that one message stays far below the body cap
    constexpr size_t kChunk = 256 * 1024;
    for (size_t at = 0; at < count; at += kChunk) {
        const size_t n = std::min(kChunk, count - at);
        std::vector<uint8_t> buf(align_up(n * 8 + 128, 8));
        fb::Builder b(buf.data(), buf.size());
        // rkey (unused by COMMIT in the reference) reports index insertions that failed
        encode_remote_meta(b, {}, 0, take_publish_failures(), addrs + at, n, kOpCommit);
        if (send_only(kOpCommit, b.data(), b.size()) != 0) {
            fail("commit: send failed");
            return -1;
        }
    }
    return 0;
}

int Connection::discard_blocks(const uint64_t* addrs, size_t count) {
    constexpr size_t kChunk = 256 * 1024;
    for (size_t at = 0; at < count; at += kChunk) {
        const size_t n = std::min(kChunk, count - at);
        std::vector<uint8_t> buf(align_up(n * 8 + 128, 8));
        fb::Builder b(buf.data(), buf.size());
        encode_remote_meta(b, {}, -1, 0, addrs + at, n, kOpStageCommit);
        if (send_only(kOpStageCommit, b.data(), b.size()) != 0) return -1;
    }
    return 0;
}

uint32_t Connection::take_publish_failures() {
    std::lock_guard<std::mutex> lk(mu_);
    const uint32_t n = publish_failures_;
    publish_failures_ = 0;
    return n;
}

int Connection::flush_commits() {
    std::vector<uint64_t> addrs;
    {
        std::lock_guard<std::mutex> lk(mu_);
        addrs.swap(pending_commit_);
    }
    if (addrs.empty()) return 0;
    return send_commit(addrs.data(), addrs.size());
}

// One control-plane message holds at most kMaxBody bytes: large key lists are sent in chunks.
static std::vector<std::pair<size_t, size_t>> chunk_keys(const std::vector<std::string_view>& keys) {
    constexpr size_t kBudget = 3u << 20;
    std::vector<std::pair<size_t, size_t>> out;
    size_t begin = 0, bytes = 0;
    for (size_t i = 0; i < keys.size(); ++i) {
        const size_t need = keys[i].size() + 16;
        if (bytes + need > kBudget && i > begin) {
            out.emplace_back(begin, i);
            begin = i;
            bytes = 0;
        }
        bytes += need;
    }
    if (begin < keys.size()) out.emplace_back(begin, keys.size());
    return out;
}

int Connection::allocate(const std::vector<std::string_view>& keys, int block_size,
                         std::vector<RemoteBlock>& out, int hint) {
    out.clear();
    if (keys.empty() || block_size <= 0) return -1;
    out.reserve(keys.size());
    for (auto [b0, b1] : chunk_keys(keys)) {
        const std::vector<std::string_view> kv(keys.begin() + b0, keys.begin() + b1);
        std::vector<uint8_t> buf(remote_meta_bound(kv, 0));
        fb::Builder b(buf.data(), buf.size());
        encode_remote_meta(b, kv, block_size, 0, nullptr, 0, kOpAllocate,
                           hint == kHintDefault ? cfg_.pool_hint : hint);
        int32_t code = 0;
        std::vector<uint8_t> p;
        if (transact(kOpAllocate, b.data(), b.size(), &code, &p, kBlobPayload) != 0) return -1;
        if (code != kFinish) {
            // blocks reserved by earlier chunks stay reserved-uncommitted; the server releases
            // them when this connection closes
            fail("allocate: server returned " + std::to_string(code));
            out.clear();
            return -code;
        }
        std::vector<RemoteBlock> part;
        try {
            part = decode_allocate_response(p.data(), p.size());
        } catch (const std::exception& e) {
            fail(std::string("allocate: bad reply: ") + e.what());
            out.clear();
            return -1;
        }
        if (part.size() != kv.size()) {
            out.clear();
            return -1;
        }
        out.insert(out.end(), part.begin(), part.end());
    }
    if (server_hbm_) {  // remember fingerprints: the write kernel publishes them in-band
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t i = 0; i < keys.size(); ++i) {
            if (is_fake_block(out[i])) continue;
            pending_hash_.put(out[i].remote_addr,
                              hash_key(reinterpret_cast<const uint8_t*>(keys[i].data()),
                                       keys[i].size()));
        }
    }
    return 0;
}

int Connection::touch(const std::vector<std::string_view>& keys) {
    if (keys.empty()) return 0;
    int total = 0;
    for (auto [b0, b1] : chunk_keys(keys)) {
        const std::vector<std::string_view> kv(keys.begin() + b0, keys.begin() + b1);
        std::vector<uint8_t> buf(remote_meta_bound(kv, 0));
        fb::Builder b(buf.data(), buf.size());
        encode_match_request(b, kv);
        int32_t code = 0;
        std::vector<uint8_t> p;
        if (transact(kOpTouch, b.data(), b.size(), &code, &p, sizeof(int32_t)) != 0 ||
            code != kFinish)
            return -1;
        int32_t v;
        std::memcpy(&v, p.data(), sizeof(v));
        total += v;
    }
    return total;
}

int Connection::lookup_blocks(char op, const std::vector<KeyOffset>& blocks, int block_size,
                              std::vector<RemoteBlock>& out) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        ctrl_dirty_ = true;  // the server pins looked-up blocks until our next SYNC
    }
    int32_t code = 0;
    std::vector<uint8_t> p;
    if (op == kOpLocalRead || op == kOpLocalWrite) {
        std::vector<LocalBlock> lb(blocks.size());
        for (size_t i = 0; i < blocks.size(); ++i) lb[i] = LocalBlock{blocks[i].key, blocks[i].offset};
        std::vector<uint8_t> buf(local_meta_bound(lb));
        fb::Builder b(buf.data(), buf.size());
        encode_local_meta(b, std::max(default_device_, 0), std::string_view(), block_size, lb);
        if (transact(op, b.data(), b.size(), &code, &p, kBlobPayload) != 0) return -1;
    } else {
        std::vector<std::string_view> all;
        all.reserve(blocks.size());
        for (auto& kb : blocks) all.push_back(kb.key);
        const auto chunks = chunk_keys(all);
        if (chunks.size() > 1) {  // very large batch: one request per chunk
            out.clear();
            for (auto [b0, b1] : chunks) {
                const std::vector<std::string_view> kv(all.begin() + b0, all.begin() + b1);
                std::vector<uint8_t> buf(remote_meta_bound(kv, 0));
                fb::Builder b(buf.data(), buf.size());
                encode_remote_meta(b, kv, block_size, 0, nullptr, 0, op, cfg_.pool_hint);
                if (transact(op, b.data(), b.size(), &code, &p, kBlobPayload) != 0) return -1;
                if (code != kFinish && code != kTaskAccepted) {
                    last_error_ = std::string(op_name(op)) + ": server returned " + std::to_string(code);
                    return -code;
                }
                try {
                    auto part = decode_allocate_response(p.data(), p.size());
                    out.insert(out.end(), part.begin(), part.end());
                } catch (const std::exception& e) {
                    fail(std::string("bad reply: ") + e.what());
                    return -1;
                }
            }
            return out.size() == blocks.size() ? 0 : -1;
        }
        std::vector<uint8_t> buf(remote_meta_bound(all, 0));
        fb::Builder b(buf.data(), buf.size());
        encode_remote_meta(b, all, block_size, 0, nullptr, 0, op, cfg_.pool_hint);
        if (transact(op, b.data(), b.size(), &code, &p, kBlobPayload) != 0) return -1;
    }
    if (code != kFinish && code != kTaskAccepted) {
        last_error_ = std::string(op_name(op)) + ": server returned " + std::to_string(code);
        return -code;
    }
    try {
        out = decode_allocate_response(p.data(), p.size());
    } catch (const std::exception& e) {
        fail(std::string("bad reply: ") + e.what());
        return -1;
    }
    return out.size() == blocks.size() ? 0 : -1;
}

// ---------------------------------------------------------------- data plane

Connection::DevCtx* Connection::dev_ctx(int device) {
    auto it = devs_.find(device);
    if (it != devs_.end()) return it->second.get();
    if (device < 0 || device >= fabric::cuda_device_count()) {
        fail("no such CUDA device: " + std::to_string(device));
        return nullptr;
    }
    DeviceGuard g(device);
    auto ctx = std::make_unique<DevCtx>();
    ctx->device = device;
    void* dp = nullptr;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void**>(&ctx->ring_h), kRingBytes,
                      cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaHostGetDevicePointer(&dp, ctx->ring_h, 0) != cudaSuccess) {
        fail(std::string("device context: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    ctx->ring_d = static_cast<uint8_t*>(dp);
    if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->status_h), 256,
                      cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaHostGetDevicePointer(&dp, ctx->status_h, 0) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&ctx->scratch), kScratchBytes) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&ctx->zeros), kZeroBytes) != cudaSuccess ||
        cudaMemset(ctx->zeros, 0, kZeroBytes) != cudaSuccess) {
        fail(std::string("device context: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    ctx->status_d = static_cast<uint32_t*>(dp);
    std::memset(ctx->status_h, 0, 256);
    if (default_device_ < 0) default_device_ = device;
    DevCtx* raw = ctx.get();
    devs_[device] = std::move(ctx);
    return raw;
}

std::shared_ptr<fabric::Mapping> Connection::mapping(uint32_t seg, int device) {
    if (seg >= segs_.size() && refresh_pool_map() != 0) return nullptr;
    if (seg >= segs_.size()) {
        fail("server referenced unknown pool segment " + std::to_string(seg));
        return nullptr;
    }
    if (device >= 0) {
        DevCtx* ctx = dev_ctx(device);
        if (!ctx) return nullptr;
        if (ctx->maps.size() <= seg) ctx->maps.resize(seg + 1);
        if (!ctx->maps[seg]) {
            std::string err;
            ctx->maps[seg] = fabric::map_segment(segs_[seg], device, &err);
            if (!ctx->maps[seg]) fail("cannot map pool segment: " + err);
        }
        return ctx->maps[seg];
    }
    // host-only mapping (CPU tensors against a host pool)
    if (host_maps_.size() <= seg) host_maps_.resize(seg + 1);
    if (!host_maps_[seg]) {
        std::string err;
        host_maps_[seg] = fabric::map_segment(segs_[seg], -1, &err);
        if (!host_maps_[seg]) fail("cannot map pool segment: " + err);
    }
    return host_maps_[seg];
}

int Connection::ensure_host_registered(uint64_t ptr, size_t bytes, int device, bool temporary) {
    auto it = host_regs_.upper_bound(ptr);
    if (it != host_regs_.begin()) {
        --it;
        if (ptr >= it->first && ptr + bytes <= it->first + it->second.bytes &&
            it->second.registered)
            return 0;
    }
    DeviceGuard g(device);
    const cudaError_t e = cudaHostRegister(reinterpret_cast<void*>(ptr), bytes,
                                           cudaHostRegisterMapped | cudaHostRegisterPortable);
    if (e != cudaSuccess && e != cudaErrorHostMemoryAlreadyRegistered) {
        (void)cudaGetLastError();
        return -1;
    }
    (void)cudaGetLastError();
    host_regs_[ptr] = HostReg{bytes, e == cudaSuccess, temporary};
    return 0;
}

// Implicit pins (a CPU tensor that was never register_mr'ed) live for one transfer only: the
// caller may free the tensor after sync(), and a pin that outlives its memory would map stale
// physical pages if the address is reused.
void Connection::release_temporary_host_regs() {
    for (auto it = host_regs_.begin(); it != host_regs_.end();) {
        if (it->second.temporary) {
            if (it->second.registered) cudaHostUnregister(reinterpret_cast<void*>(it->first));
            (void)cudaGetLastError();
            it = host_regs_.erase(it);
        } else {
            ++it;
        }
    }
}

int Connection::unregister_mr(uint64_t ptr) {
    if (drain_devices() != 0) { /* report via the next sync; still unpin below */ }
    std::lock_guard<std::mutex> lk(mu_);
    mrs_.erase(ptr);
    auto it = host_regs_.find(ptr);
    if (it != host_regs_.end()) {
        if (it->second.registered) cudaHostUnregister(reinterpret_cast<void*>(ptr));
        (void)cudaGetLastError();
        host_regs_.erase(it);
    }
    return 0;
}

int Connection::register_mr(uint64_t ptr, size_t size, int device) {
    std::lock_guard<std::mutex> lk(mu_);
    mrs_[ptr] = size;  // re-registering the same base replaces the old entry
    if (device < 0 && fabric::cuda_available() && server_hbm_) {
        // Pin + map host memory so that kernels can stream it over PCIe (the role
        // ibv_reg_mr plays for CPU tensors in the reference).
        const int kd = cfg_.device >= 0 ? cfg_.device : std::max(default_device_, 0);
        if (ensure_host_registered(ptr, size, kd, false) != 0)
            LOG_WARN("register_mr: could not pin host memory, falling back to staged copies");
    }
    if (device >= 0 && !dev_ctx(device)) return -1;
    return 1;
}

// Resolve the device pointer of a pool segment for `ctx` (slow path: map it first).
uint8_t* Connection::seg_dev_ptr(DevCtx* ctx, uint32_t seg) {
    if (seg < ctx->seg_ptr.size() && ctx->seg_ptr[seg]) return ctx->seg_ptr[seg];
    auto mp = mapping(seg, ctx->device);
    if (!mp || (!mp->dev_ptr && !mp->mc_ptr)) {
        fail("pool segment " + std::to_string(seg) + " is not addressable from device " +
             std::to_string(ctx->device));
        return nullptr;
    }
    if (ctx->seg_ptr.size() <= seg) {
        ctx->seg_ptr.resize(seg + 1, nullptr);
        ctx->seg_remote.resize(seg + 1, 1);
        ctx->seg_mc.resize(seg + 1, nullptr);
    }
    ctx->seg_ptr[seg] = mp->dev_ptr;
    ctx->seg_mc[seg] = mp->mc_ptr;
    // NVLink (or PCIe) on the path?  Such transfers are link-bound: a small grid saturates
    // them and leaves the SMs to whatever else runs on this GPU.
    cudaPointerAttributes attr{};
    bool local = false;
    if (cudaPointerGetAttributes(&attr, mp->dev_ptr) == cudaSuccess)
        local = attr.type == cudaMemoryTypeDevice && attr.device == ctx->device;
    (void)cudaGetLastError();
    ctx->seg_remote[seg] = local ? 0 : 1;
    return mp->dev_ptr;
}

// The index shards beyond shard 0 as seen from `ctx`'s device: the k-th HBM segment (in id
// order) that carries a table is shard k.  *all_local is cleared when a shard's table lives
// on another GPU.
kernels::IndexShards Connection::index_shards(DevCtx* ctx, bool* all_local) {
    kernels::IndexShards sh;
    uint32_t n = 0;
    for (uint32_t id = 0; id < segs_.size() && n < kernels::kMaxIndexShards; ++id) {
        if (segs_[id].kind != kSegDeviceIpc || !segs_[id].index_slots) continue;
        if (n > 0) {
            uint8_t* base = seg_dev_ptr(ctx, id);
            if (!base) break;  // cannot map it: stay with the shards found so far
            sh.table[n - 1] = reinterpret_cast<kernels::IndexBucket*>(base + segs_[id].index_off);
            sh.mask[n - 1] = kernels::index_bucket_mask(segs_[id].index_slots);
            if (all_local && ctx->seg_remote[id]) *all_local = false;
        }
        ++n;
    }
    sh.n = n;
    return sh;
}

// Move n blocks between the caller's tensor and the pool.  local_off[i] * scale is the
// byte offset of block i from base_ptr.
int Connection::move_blocks(bool write, const uint64_t* local_off, uint64_t scale,
                            const RemoteBlock* blocks, size_t n, int block_size,
                            uint64_t base_ptr, int device, uint64_t stream_in, int fp8_elems,
                            MoveResult* res) {
    NvtxRange nvtx(write ? "istore.write_blocks" : "istore.read_blocks");
    std::lock_guard<std::mutex> lk(mu_);
    if (n == 0) return 0;
    // Where the addresses to COMMIT go: the connection-wide list shipped by the next sync(),
    // or the caller's own list (async writes commit exactly their own blocks when THEIR
    // kernels have finished, reference: src/libinfinistore.cpp:362-395).  An address is
    // appended only after the launch that writes the block has succeeded; every error return
    // below leaves the sink exactly as the last successful launch left it.
    std::vector<uint64_t>& sink = (res && res->commits) ? *res->commits : pending_commit_;
    std::vector<uint64_t> batch_commits;
    int kd = device;
    if (device < 0 && fp8_elems) {
        fail("the fp8 KV path needs a CUDA tensor");
        return -1;
    }
    if (device < 0) {
        // host tensor: memcpy when every target segment is host memory, else a kernel on
        // the connection's default device reads / writes the (pinned) host tensor
        bool all_host_segs = true;
        size_t live = 0;
        uint64_t max_off = 0;
        for (size_t i = 0; i < n; ++i) {
            if (write && is_fake_block(blocks[i])) continue;  // dedup: first writer wins
            const uint32_t seg = addr_seg(blocks[i].remote_addr);
            if (seg >= segs_.size() && refresh_pool_map() != 0) return -1;
            if (seg >= segs_.size()) {
                fail("block refers to unknown segment");
                return -1;
            }
            if (segs_[seg].kind != kSegHostShm) all_host_segs = false;
            max_off = std::max(max_off, local_off[i] * scale);
            ++live;
        }
        if (live == 0) return 0;
        if (all_host_segs) {
            for (size_t i = 0; i < n; ++i) {
                if (write && is_fake_block(blocks[i])) continue;
                auto m = mapping(addr_seg(blocks[i].remote_addr), -1);
                if (!m || !m->host_ptr) return -1;
                uint8_t* pool = m->host_ptr + addr_off(blocks[i].remote_addr);
                uint8_t* local = reinterpret_cast<uint8_t*>(base_ptr + local_off[i] * scale);
                if (write)
                    std::memcpy(pool, local, size_t(block_size));
                else
                    std::memcpy(local, pool, size_t(block_size));
                if (write) sink.push_back(blocks[i].remote_addr);
                stats_.host_copies++;
            }
            (write ? stats_.bytes_written : stats_.bytes_read) += live * uint64_t(block_size);
            return 0;
        }
        if (!fabric::cuda_available()) {
            fail("a CUDA device is required to reach an HBM pool");
            return -1;
        }
        kd = cfg_.device >= 0 ? cfg_.device : std::max(default_device_, 0);
        auto mr = mrs_.find(base_ptr);
        const size_t span = mr != mrs_.end() ? mr->second : size_t(max_off) + size_t(block_size);
        if (ensure_host_registered(base_ptr, span, kd, mr == mrs_.end()) != 0) {
            fail("cannot pin the host tensor for the GPU data path");
            return -1;
        }
    }

    // --- kernel path on `kd`
    DevCtx* ctx = dev_ctx(kd);
    if (!ctx) return -1;
    DeviceGuard g(kd);
    if (cfg_.doorbell && write && n == 1 && device >= 0 && !fp8_elems && !res &&
        !is_fake_block(blocks[0])) {
        // latency mode: one block, handed to the persistent worker (no launch, no event)
        const RemoteBlock& rb = blocks[0];
        const uint32_t seg = addr_seg(rb.remote_addr);
        uint8_t* segbase =
            (seg < segs_.size() || refresh_pool_map() == 0) ? seg_dev_ptr(ctx, seg) : nullptr;
        const bool replicated = seg < ctx->seg_mc.size() && ctx->seg_mc[seg];
        if (segbase && !replicated && doorbell_ready(ctx, stream_in, size_t(block_size))) {
            KeyHash kh{0, 0};  // h1 == 0: not allocated through this connection, not indexed
            if (!segs_[0].index_slots || !pending_hash_.take(rb.remote_addr, &kh)) kh = KeyHash{0, 0};
            const uint64_t q[6] = {base_ptr + local_off[0] * scale,
                                   reinterpret_cast<uint64_t>(segbase) + addr_off(rb.remote_addr),
                                   kh.h1, kh.h2, rb.remote_addr,
                                   uint64_t(rb.gen) | (uint64_t(uint32_t(block_size)) << 32)};
            if (doorbell_post(ctx, kernels::kDoorbellWrite, q) != 0) return -1;
            sink.push_back(rb.remote_addr);
            stats_.calls++;
            stats_.bytes_written += uint64_t(block_size);
            return 0;
        }
    }
    if (doorbell_quiesce(ctx) != 0) return -1;
    const uint64_t t_pick0 = now_ns();
    cudaStream_t stream = ctx->pick(reinterpret_cast<cudaStream_t>(stream_in), write, streams_);
    stats_.ns_streams += now_ns() - t_pick0;
    stats_.calls++;
    if (res) {
        res->stream = stream;
        res->device = kd;
    }

    // the device index lives in segment 0
    kernels::IndexBucket* table = nullptr;
    uint64_t table_mask = 0;
    if (write && !segs_.empty() && segs_[0].index_slots) {
        if (uint8_t* p0 = seg_dev_ptr(ctx, 0)) {
            table = reinterpret_cast<kernels::IndexBucket*>(p0 + segs_[0].index_off);
            table_mask = kernels::index_bucket_mask(segs_[0].index_slots);
        }
    }

    size_t i = 0;
    while (i < n) {
        const uint64_t t_build0 = now_ns();
        const size_t batch_cap = std::min(kMaxBatch, n - i);
        const size_t at_desc = ctx->ring_alloc(batch_cap * sizeof(kernels::CopyDesc));
        auto* descs = reinterpret_cast<kernels::CopyDesc*>(ctx->ring_h + at_desc);
        size_t at_rec = 0;
        kernels::IndexEntry* recs = nullptr;
        if (table) {
            at_rec = ctx->ring_alloc(batch_cap * sizeof(kernels::IndexEntry));
            recs = reinterpret_cast<kernels::IndexEntry*>(ctx->ring_h + at_rec);
        }
        uint32_t m = 0;
        batch_commits.clear();
        uint32_t n_mc = 0;  // blocks of this batch that live in the NVLS-replicated region
        bool can_publish = table != nullptr;
        bool all_remote = true;
        bool all_local = !ctx->seg_remote.empty() && !ctx->seg_remote[0];  // index table (segment 0)
        uint64_t align_or = 0;
        const size_t nseg = ctx->seg_ptr.size();
        uint8_t* const* seg_ptr = ctx->seg_ptr.data();
        for (; i < n && m < batch_cap; ++i) {
            const RemoteBlock& rb = blocks[i];
            if (write && is_fake_block(rb)) continue;
            const uint32_t seg = addr_seg(rb.remote_addr);
            uint8_t* segbase = seg < nseg ? seg_ptr[seg] : nullptr;
            if (!segbase) {
                segbase = seg_dev_ptr(ctx, seg);
                if (!segbase && !(seg < ctx->seg_mc.size() && ctx->seg_mc[seg])) return -1;
                seg_ptr = ctx->seg_ptr.data();
            }
            if (write && seg < ctx->seg_mc.size() && ctx->seg_mc[seg]) {
                segbase = ctx->seg_mc[seg];  // replicated block: write through the multicast VA
                ++n_mc;
            } else if (!segbase) {
                fail("no local replica of the NVLS region on device " + std::to_string(kd));
                return -1;
            }
            const uint64_t pool = reinterpret_cast<uint64_t>(segbase) + addr_off(rb.remote_addr);
            const uint64_t local = base_ptr + local_off[i] * scale;
            all_remote = all_remote && ctx->seg_remote[seg];
            all_local = all_local && !ctx->seg_remote[seg];
            align_or |= local;
            descs[m].src = write ? local : pool;
            descs[m].dst = write ? pool : local;
            if (write) {
                if (recs) {
                    KeyHash kh;
                    if (pending_hash_.take(rb.remote_addr, &kh))
                        recs[m] = kernels::IndexEntry{kh.h1, kh.h2, rb.remote_addr, rb.gen,
                                                      uint32_t(block_size)};
                    else
                        can_publish = false;  // not allocated through this connection
                }
                batch_commits.push_back(rb.remote_addr);
            }
            ++m;
        }
        if (m == 0) break;
        if (n_mc && n_mc != m) {
            fail("a write batch must not mix replicated and ordinary blocks");
            return -1;
        }
        kernels::CopyLaunch L;
        L.multicast = n_mc != 0;
        L.descs = reinterpret_cast<const kernels::CopyDesc*>(ctx->ring_d + at_desc);
        L.descs_host = descs;
        L.n = m;
        L.bytes = uint32_t(block_size);
        L.align_or = align_or;
        L.status = ctx->status_d;
        L.variant = copy_variant_;
        L.stage_bytes = pipe_stage_;
        L.ring_bytes = pipe_ring_;
        L.max_ctas = max_ctas_ ? max_ctas_ : (all_remote ? 2 * kernels::sm_count() : 0);
        L.all_local = all_local && !L.multicast;
        if (L.multicast && fp8_elems) {
            fail("the fp8 path does not write to the replicated region");
            return -1;
        }
        if (can_publish) {
            L.recs = reinterpret_cast<const kernels::IndexEntry*>(ctx->ring_d + at_rec);
            L.table = table;
            L.table_mask = table_mask;
            bool shards_local = true;
            L.shards = index_shards(ctx, &shards_local);
            L.all_local = L.all_local && shards_local;
            all_local = all_local && shards_local;
            L.done = reinterpret_cast<uint32_t*>(ctx->zeros + ctx->zeros_alloc(size_t(m) * 12));
        }
        const uint64_t t_launch0 = now_ns();
        stats_.ns_build += t_launch0 - t_build0;
        cudaError_t e;
        if (fp8_elems) {
            kernels::Fp8Launch F;
            F.descs = L.descs;
            F.n = m;
            F.elems = uint32_t(fp8_elems);
            F.recs = L.recs;
            F.table = L.table;
            F.table_mask = L.table_mask;
            F.shards = L.shards;
            F.done = L.done;
            F.status = L.status;
            F.max_ctas = L.max_ctas;
            F.all_local = all_local;
            F.aligned16 = (align_or & 15) == 0;
            // TMA-pipelined flavour over NVLink (write 622 vs 537, read 705 vs 507 GB/s of fp8
            // bytes), per-thread flavour on local HBM (1705 vs 1223): profiles/r2_lab_fp8_2gpu.json
            F.variant = all_local ? 1 : 0;
            e = write ? kernels::launch_kv_write_fp8(F, stream) : kernels::launch_kv_read_fp8(F, stream);
        } else {
            e = kernels::launch_kv_copy(L, stream);
        }
        stats_.ns_launch += now_ns() - t_launch0;
        if (e != cudaSuccess) {
            fail(std::string("page mover launch failed: ") + cudaGetErrorString(e));
            return -1;
        }
        ctx->mark(stream);
        if (res) res->launched = true;
        sink.insert(sink.end(), batch_commits.begin(), batch_commits.end());
        stats_.kernel_launches++;
        (write ? stats_.bytes_written : stats_.bytes_read) += uint64_t(m) * uint64_t(block_size);
    }
    return 0;
}

int Connection::w_rdma(const uint64_t* offsets, size_t noffsets, uint64_t scale, int block_size,
                       const RemoteBlock* blocks, size_t nblocks, uint64_t base_ptr, int device,
                       uint64_t stream, MoveResult* res) {
    if (noffsets != nblocks) {
        fail("w_rdma: offsets and remote blocks differ in length");
        return -1;
    }
    return move_blocks(true, offsets, scale, blocks, nblocks, block_size, base_ptr, device,
                       stream, 0, res);
}

int Connection::w_rdma_fp8(const uint64_t* offsets, size_t noffsets, uint64_t scale, int elems,
                           const RemoteBlock* blocks, size_t nblocks, uint64_t base_ptr,
                           int device, uint64_t stream) {
    if (noffsets != nblocks || elems <= 0 || elems % 128) {
        fail("w_rdma_fp8: page size must be a positive multiple of 128 elements");
        return -1;
    }
    return move_blocks(true, offsets, scale, blocks, nblocks,
                       int(kernels::fp8_block_bytes(uint32_t(elems), 128)), base_ptr, device,
                       stream, elems);
}

int Connection::r_rdma_fp8(const std::vector<KeyOffset>& blocks, int elems, uint64_t base_ptr,
                           int device, uint64_t stream) {
    if (blocks.empty()) return 0;
    if (elems <= 0 || elems % 128 || device < 0) {
        fail("r_rdma_fp8: needs a CUDA tensor and pages of a multiple of 128 elements");
        return -1;
    }
    const int bytes = int(kernels::fp8_block_bytes(uint32_t(elems), 128));
    if (device_lookup_ && server_hbm_ && device_index_usable())
        return read_via_device_index(blocks, bytes, base_ptr, device, stream, elems);
    std::vector<RemoteBlock> rb;
    const int r = lookup_blocks(kOpReadLookup, blocks, bytes, rb);
    if (r != 0) return r;
    std::vector<uint64_t> offs(blocks.size());
    for (size_t i = 0; i < blocks.size(); ++i) offs[i] = blocks[i].offset;
    return move_blocks(false, offs.data(), 1, rb.data(), rb.size(), bytes, base_ptr, device,
                       stream, elems);
}

int Connection::r_rdma(const std::vector<KeyOffset>& blocks, int block_size, uint64_t base_ptr,
                       int device, uint64_t stream, MoveResult* res) {
    if (blocks.empty()) return 0;
    if (device_lookup_ && server_hbm_ && device >= 0 && device_index_usable())
        return read_via_device_index(blocks, block_size, base_ptr, device, stream, 0, res);
    std::vector<RemoteBlock> rb;
    const int r = lookup_blocks(kOpReadLookup, blocks, block_size, rb);
    if (r != 0) return r;
    std::vector<uint64_t> offs(blocks.size());
    for (size_t i = 0; i < blocks.size(); ++i) offs[i] = blocks[i].offset;
    return move_blocks(false, offs.data(), 1, rb.data(), rb.size(), block_size, base_ptr, device,
                       stream, 0, res);
}

int Connection::rw_local(char op, const std::vector<KeyOffset>& blocks, int block_size,
                         uint64_t base_ptr, int device, uint64_t stream) {
    if (blocks.empty()) return 0;
    if (op != kOpLocalRead && op != kOpLocalWrite) return -1;
    if (op == kOpLocalRead && device_lookup_ && server_hbm_ && device >= 0 && device_index_usable())
        return read_via_device_index(blocks, block_size, base_ptr, device, stream);
    std::vector<RemoteBlock> rb;
    const int r = lookup_blocks(op, blocks, block_size, rb);
    if (r != 0) return r;
    if (op == kOpLocalWrite && server_hbm_) {
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t i = 0; i < blocks.size(); ++i) {
            if (is_fake_block(rb[i])) continue;
            pending_hash_.put(rb[i].remote_addr,
                              hash_key(reinterpret_cast<const uint8_t*>(blocks[i].key.data()),
                                       blocks[i].key.size()));
        }
    }
    std::vector<uint64_t> offs(blocks.size());
    for (size_t i = 0; i < blocks.size(); ++i) offs[i] = blocks[i].offset;
    return move_blocks(op == kOpLocalWrite, offs.data(), 1, rb.data(), rb.size(), block_size,
                       base_ptr, device, stream);
}

// Pack keys into the pinned ring so that the lookup kernel can hash them: each key starts
// on an 8-byte boundary and is zero padded to a multiple of 8.
static size_t pack_keys(const std::string_view* keys, size_t n, uint8_t* bytes, uint32_t* off,
                        uint32_t* len) {
    size_t at = 0;
    for (size_t i = 0; i < n; ++i) {
        const std::string_view k = keys[i];
        off[i] = uint32_t(at);
        len[i] = uint32_t(k.size());
        const size_t padded = align_up(k.size() ? k.size() : 1, 8);
        std::memcpy(bytes + at, k.data(), k.size());
        std::memset(bytes + at + k.size(), 0, padded - k.size());
        at += padded;
    }
    return at;
}

int Connection::read_via_device_index(const std::vector<KeyOffset>& blocks, int block_size,
                                      uint64_t base_ptr, int device, uint64_t stream_in,
                                      int fp8_elems, MoveResult* res) {
    NvtxRange nvtx("istore.read_via_device_index");
    std::lock_guard<std::mutex> lk(mu_);
    DevCtx* ctx = dev_ctx(device);
    if (!ctx) return -1;
    auto m0 = mapping(0, device);
    if (!m0 || !m0->dev_ptr || !segs_[0].index_slots) {
        fail("the server exposes no device index");
        return -1;
    }
    DeviceGuard g(device);
    if (cfg_.doorbell && blocks.size() == 1 && !fp8_elems && !res &&
        doorbell_ready(ctx, stream_in, size_t(block_size))) {
        // latency mode: the key is hashed here, the worker probes, copies and re-checks
        const KeyHash kh = hash_key(reinterpret_cast<const uint8_t*>(blocks[0].key.data()),
                                    blocks[0].key.size());
        const uint64_t q[6] = {base_ptr + blocks[0].offset, 0, kh.h1, kh.h2, 0,
                               uint64_t(uint32_t(block_size)) << 32};
        if (doorbell_post(ctx, kernels::kDoorbellRead, q) != 0) return -1;
        stats_.calls++;
        stats_.bytes_read += uint64_t(block_size);
        return 0;
    }
    if (doorbell_quiesce(ctx) != 0) return -1;
    const uint64_t t_pick0 = now_ns();
    cudaStream_t stream = ctx->pick(reinterpret_cast<cudaStream_t>(stream_in), false, streams_);
    stats_.ns_streams += now_ns() - t_pick0;
    stats_.calls++;
    if (res) {
        res->stream = stream;
        res->device = device;
    }
    for (size_t base = 0; base < blocks.size(); base += kMaxBatch) {
        const uint64_t t_build0 = now_ns();
        const size_t n = std::min(kMaxBatch, blocks.size() - base);
        size_t key_bytes = 0;
        std::vector<std::string_view> kp(n);
        for (size_t i = 0; i < n; ++i) {
            kp[i] = blocks[base + i].key;
            key_bytes += align_up(std::max<size_t>(kp[i].size(), 1), 8);
        }
        const size_t at_bytes = ctx->ring_alloc(key_bytes);
        const size_t at_off = ctx->ring_alloc(n * 4);
        const size_t at_len = ctx->ring_alloc(n * 4);
        const size_t at_dst = ctx->ring_alloc(n * 8);
        pack_keys(kp.data(), n, ctx->ring_h + at_bytes,
                  reinterpret_cast<uint32_t*>(ctx->ring_h + at_off),
                  reinterpret_cast<uint32_t*>(ctx->ring_h + at_len));
        auto* dst = reinterpret_cast<uint64_t*>(ctx->ring_h + at_dst);
        for (size_t i = 0; i < n; ++i) dst[i] = blocks[base + i].offset;

        const uint32_t nsegs =
            uint32_t(std::min<size_t>(segs_.size(), kernels::LookupLaunch::kMaxSegs));
        uint64_t seg_base[kernels::LookupLaunch::kMaxSegs] = {0};
        bool all_remote = true;
        for (uint32_t s = 0; s < nsegs; ++s) {
            seg_base[s] = reinterpret_cast<uint64_t>(seg_dev_ptr(ctx, s));
            all_remote = all_remote && ctx->seg_remote[s];
        }
        // fp8 pages pulled over NVLink: one CTA per SM.  With both link directions busy, 296
        // CTAs of 8 compute warps pull 954 GB/s (2 GPUs, fp8 bytes), 148 pull 1085
        // (bench/configs.py fp8 --max-ctas 148, round 2); writes prefer two per SM.
        const int grid_cap = max_ctas_ ? max_ctas_
                             : all_remote ? (fp8_elems ? 1 : 2) * kernels::sm_count()
                                          : 0;
        uint64_t align_or = base_ptr;
        for (size_t i = 0; i < n; ++i) align_or |= blocks[base + i].offset;
        const uint64_t t_launch0 = now_ns();
        stats_.ns_build += t_launch0 - t_build0;
        cudaError_t e;
        // The fused kernels resolve a key in every CTA that moves a piece of its block, reading
        // the key bytes from the pinned ring each time: right for a batch of blocks (items are
        // whole blocks, or chunks when that balances the grid better - kernels/balance.h),
        // wasteful when a few LARGE blocks are split over many CTAs (measured: 4 MB
        // single-block read 83 us vs 42 us) - then resolve each key once with the lookup kernel
        // and feed the descriptors to kv_copy.
        // A handful of blocks (<= 4 MB in all): latency matters, not bandwidth.  One launch of
        // the ld/st flavour, which splits a block into 32 KB chunks over CTAs (each resolves its
        // block's key itself - a few redundant probes) and re-checks the entries in the same
        // kernel: lookup + copy + validate would be three launches (+7..15 us per read).
        const bool small_batch = n * size_t(block_size) <= (4u << 20);
        const bool whole_blocks = !small_batch && uint32_t(block_size) <= (1u << 20) && n >= 32;
        kernels::ReadFusedLaunch R;
        R.align_or = align_or;
        const bool fp8_fused = fp8_elems && kernels::fp8_read_fused_supported(R, uint32_t(fp8_elems));
        if (fp8_fused || (!fp8_elems && (whole_blocks || small_batch))) {
            // one kernel: hash + probe + move
            R.key_bytes = ctx->ring_d + at_bytes;
            R.key_off = reinterpret_cast<const uint32_t*>(ctx->ring_d + at_off);
            R.key_len = reinterpret_cast<const uint32_t*>(ctx->ring_d + at_len);
            R.dst_off = reinterpret_cast<const uint64_t*>(ctx->ring_d + at_dst);
            R.dst_base = base_ptr;
            R.n = uint32_t(n);
            R.bytes = uint32_t(block_size);
            R.align_or = copy_variant_ == kernels::kCopyLdSt ? (align_or | 16) : align_or;
            R.table = reinterpret_cast<const kernels::IndexBucket*>(m0->dev_ptr + segs_[0].index_off);
            R.table_mask = kernels::index_bucket_mask(segs_[0].index_slots);
            R.shards = index_shards(ctx, nullptr);
            R.nsegs = nsegs;
            for (uint32_t s = 0; s < nsegs; ++s) R.seg_base[s] = seg_base[s];
            R.status = ctx->status_d;
            R.max_ctas = grid_cap;
            // always: a purge (or an eviction) may free a block while a device-path read is
            // copying it; the post-copy tag check turns that into a reported miss
            R.validate = true;
            R.variant = small_batch ? int(kernels::kCopyLdSt256) : copy_variant_;
            R.stage_bytes = pipe_stage_;
            R.ring_bytes = pipe_ring_;
            // fp8 pages: the resolver rides in the dequantising TMA pipeline (one launch
            // instead of lookup + read + validate: 32 calls of 512 pages were launch-bound)
            e = fp8_fused ? kernels::launch_kv_fp8_read_fused(R, uint32_t(fp8_elems), stream)
                          : kernels::launch_kv_read_fused(R, stream);
            stats_.kernel_launches += 1;
        } else {
            kernels::LookupLaunch Q;
            Q.key_bytes = ctx->ring_d + at_bytes;
            Q.key_off = reinterpret_cast<const uint32_t*>(ctx->ring_d + at_off);
            Q.key_len = reinterpret_cast<const uint32_t*>(ctx->ring_d + at_len);
            Q.n = uint32_t(n);
            Q.table = reinterpret_cast<const kernels::IndexBucket*>(m0->dev_ptr + segs_[0].index_off);
            Q.table_mask = kernels::index_bucket_mask(segs_[0].index_slots);
            Q.shards = index_shards(ctx, nullptr);
            Q.nsegs = nsegs;
            for (uint32_t s = 0; s < nsegs; ++s) Q.seg_base[s] = seg_base[s];
            auto* out = reinterpret_cast<kernels::CopyDesc*>(
                ctx->scratch + ctx->scratch_alloc(n * sizeof(kernels::CopyDesc)));
            Q.out_descs = out;
            Q.dst_off = reinterpret_cast<const uint64_t*>(ctx->ring_d + at_dst);
            Q.dst_base = base_ptr;
            Q.need_bytes = uint32_t(block_size);
            Q.status = ctx->status_d;
            // optimistic read: the entries are re-checked after the copy (purge / eviction)
            Q.found_at = reinterpret_cast<kernels::LookupLaunch::FoundAt*>(
                    ctx->scratch + ctx->scratch_alloc(n * sizeof(kernels::LookupLaunch::FoundAt)));
            e = kernels::launch_index_lookup(Q, stream);
            if (e == cudaSuccess && fp8_elems) {
                kernels::Fp8Launch F;
                F.descs = out;
                F.n = uint32_t(n);
                F.elems = uint32_t(fp8_elems);
                F.status = ctx->status_d;
                F.max_ctas = grid_cap;
                F.aligned16 = (align_or & 15) == 0;
                F.variant = all_remote ? 0 : 1;
                e = kernels::launch_kv_read_fp8(F, stream);
            } else if (e == cudaSuccess) {
                kernels::CopyLaunch L;
                L.descs = out;
                L.n = uint32_t(n);
                L.bytes = uint32_t(block_size);
                L.align_or = align_or;
                L.status = ctx->status_d;
                L.variant = copy_variant_;
                L.stage_bytes = pipe_stage_;
                L.ring_bytes = pipe_ring_;
                L.max_ctas = grid_cap;
                e = kernels::launch_kv_copy(L, stream);
            }
            stats_.kernel_launches += 2;
            if (e == cudaSuccess && Q.found_at) {
                // the server evicts: the entries must still be the ones the lookup resolved
                kernels::ValidateLaunch V;
                V.found_at = Q.found_at;
                V.n = uint32_t(n);
                V.table = Q.table;
                V.shards = Q.shards;
                V.status = ctx->status_d;
                e = kernels::launch_index_validate(V, stream);
                stats_.kernel_launches += 1;
            }
        }
        stats_.ns_launch += now_ns() - t_launch0;
        if (e != cudaSuccess) {
            fail(std::string("device-index read failed to launch: ") + cudaGetErrorString(e));
            return -1;
        }
        ctx->mark(stream);
        if (res) res->launched = true;
        stats_.bytes_read += uint64_t(n) * uint64_t(block_size);
    }
    return 0;
}

// Device-addressable copy descriptors {mapped pool address (0 = miss), dst_base + offset} for
// blocks[base, base + n): resolved on the GPU by the lookup kernel (rb == nullptr, the
// descriptors land in device scratch) or taken from a server lookup (rb, pinned ring).
const kernels::CopyDesc* Connection::resolve_descs(DevCtx* ctx, const std::vector<KeyOffset>& blocks,
                                                   size_t base, size_t n, int block_size,
                                                   uint64_t dst_base,
                                                   const std::vector<RemoteBlock>* rb,
                                                   void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (rb) {
        const size_t at_desc = ctx->ring_alloc(n * sizeof(kernels::CopyDesc));
        auto* descs = reinterpret_cast<kernels::CopyDesc*>(ctx->ring_h + at_desc);
        for (size_t i = 0; i < n; ++i) {
            const RemoteBlock& b = (*rb)[base + i];
            uint8_t* segbase = seg_dev_ptr(ctx, addr_seg(b.remote_addr));
            if (!segbase) return nullptr;
            descs[i].src = reinterpret_cast<uint64_t>(segbase) + addr_off(b.remote_addr);
            descs[i].dst = dst_base + blocks[base + i].offset;
        }
        return reinterpret_cast<const kernels::CopyDesc*>(ctx->ring_d + at_desc);
    }
    auto m0 = mapping(0, ctx->device);
    if (!m0 || !m0->dev_ptr) return nullptr;
    size_t key_bytes = 0;
    std::vector<std::string_view> kp(n);
    for (size_t i = 0; i < n; ++i) {
        kp[i] = blocks[base + i].key;
        key_bytes += align_up(std::max<size_t>(kp[i].size(), 1), 8);
    }
    const size_t at_bytes = ctx->ring_alloc(key_bytes);
    const size_t at_off = ctx->ring_alloc(n * 4);
    const size_t at_len = ctx->ring_alloc(n * 4);
    const size_t at_dst = ctx->ring_alloc(n * 8);
    pack_keys(kp.data(), n, ctx->ring_h + at_bytes, reinterpret_cast<uint32_t*>(ctx->ring_h + at_off),
              reinterpret_cast<uint32_t*>(ctx->ring_h + at_len));
    auto* dst = reinterpret_cast<uint64_t*>(ctx->ring_h + at_dst);
    for (size_t i = 0; i < n; ++i) dst[i] = blocks[base + i].offset;
    kernels::LookupLaunch Q;
    Q.key_bytes = ctx->ring_d + at_bytes;
    Q.key_off = reinterpret_cast<const uint32_t*>(ctx->ring_d + at_off);
    Q.key_len = reinterpret_cast<const uint32_t*>(ctx->ring_d + at_len);
    Q.n = uint32_t(n);
    Q.table = reinterpret_cast<const kernels::IndexBucket*>(m0->dev_ptr + segs_[0].index_off);
    Q.table_mask = kernels::index_bucket_mask(segs_[0].index_slots);
    Q.shards = index_shards(ctx, nullptr);
    Q.nsegs = uint32_t(std::min<size_t>(segs_.size(), kernels::LookupLaunch::kMaxSegs));
    for (uint32_t sgi = 0; sgi < Q.nsegs; ++sgi)
        Q.seg_base[sgi] = reinterpret_cast<uint64_t>(seg_dev_ptr(ctx, sgi));
    auto* out = reinterpret_cast<kernels::CopyDesc*>(
        ctx->scratch + ctx->scratch_alloc(n * sizeof(kernels::CopyDesc)));
    Q.out_descs = out;
    Q.dst_off = reinterpret_cast<const uint64_t*>(ctx->ring_d + at_dst);
    Q.dst_base = dst_base;
    Q.need_bytes = uint32_t(block_size);
    Q.status = ctx->status_d;
    if (kernels::launch_index_lookup(Q, stream) != cudaSuccess) return nullptr;
    stats_.kernel_launches++;
    return out;
}

// The same pages into several destination tensors (TP ranks / beams sharing a prefix): every
// pool block crosses NVLink once and is fanned out inside a thread-block cluster
// (kernels/kv_pipe.cu: cp.async.bulk ... .multicast::cluster).  bases[r] is the base pointer
// of destination r; every destination uses the same page offsets.
int Connection::r_rdma_multi(const std::vector<KeyOffset>& blocks, int block_size,
                             const std::vector<uint64_t>& bases, int device, uint64_t stream_in) {
    if (blocks.empty() || bases.empty()) return 0;
    if (device < 0 || !server_hbm_) {
        fail("read_cache_multi needs CUDA destinations and an HBM pool");
        return -1;
    }
    if (bases.size() == 1) return r_rdma(blocks, block_size, bases[0], device, stream_in);
    const bool via_index = device_lookup_ && device_index_usable();
    std::vector<RemoteBlock> rb;
    if (!via_index) {
        const int r = lookup_blocks(kOpReadLookup, blocks, block_size, rb);
        if (r != 0) return r;
    }
    NvtxRange nvtx("istore.read_multi");
    std::lock_guard<std::mutex> lk(mu_);
    DevCtx* ctx = dev_ctx(device);
    if (!ctx) return -1;
    DeviceGuard g(device);
    if (doorbell_quiesce(ctx) != 0) return -1;
    cudaStream_t stream = ctx->pick(reinterpret_cast<cudaStream_t>(stream_in), false, streams_);
    stats_.calls++;
    uint64_t align_or = 0;
    for (uint64_t b : bases) align_or |= b;
    for (size_t base = 0; base < blocks.size(); base += kMaxBatch) {
        const size_t n = std::min(kMaxBatch, blocks.size() - base);
        for (size_t i = 0; i < n; ++i) align_or |= blocks[base + i].offset;
        const kernels::CopyDesc* descs_d =
            resolve_descs(ctx, blocks, base, n, block_size, bases[0], via_index ? nullptr : &rb, stream);
        if (!descs_d) {
            fail("multi-destination read: cannot resolve the blocks");
            return -1;
        }
        // Local pool: thread-block clusters (multicast bulk load into 2 or 4 CTAs, one
        // destination each; an odd last destination shares a 2-cluster with its predecessor,
        // which is rewritten with the same bytes).  Pool behind NVLink: one load, K stores
        // per CTA (the fan-out flavour of the TMA pipeline) - the fabric still carries every
        // page once.
        bool src_local = true;
        for (size_t sgi = 0; sgi < ctx->seg_remote.size(); ++sgi)
            if (ctx->seg_ptr[sgi] && ctx->seg_remote[sgi]) src_local = false;
        cudaError_t e = cudaSuccess;
        size_t r = 0;
        while (e == cudaSuccess && r < bases.size()) {
            const size_t left = bases.size() - r;
            if (src_local) {
                kernels::McastLaunch M;
                M.descs = descs_d;
                M.n = uint32_t(n);
                M.bytes = uint32_t(block_size);
                M.align_or = align_or;
                M.src_local = true;
                M.status = r == 0 ? ctx->status_d : nullptr;  // count a miss once
                const size_t first = left >= 2 ? r : r - 1;
                M.ndst = left >= 4 ? 4 : 2;
                for (int j = 0; j < M.ndst; ++j)
                    M.delta[j] = int64_t(bases[first + size_t(j)]) - int64_t(bases[0]);
                e = kernels::launch_kv_pipe_mcast(M, stream);
                r = first + size_t(M.ndst);
            } else {
                kernels::CopyLaunch L;
                L.descs = descs_d;
                L.n = uint32_t(n);
                L.bytes = uint32_t(block_size);
                L.align_or = align_or;
                L.status = r == 0 ? ctx->status_d : nullptr;
                L.variant = kernels::kCopyTma;
                L.max_ctas = max_ctas_;
                L.fan_n = int(std::min<size_t>(left, 4));
                for (int j = 0; j < L.fan_n; ++j)
                    L.fan_delta[j] = int64_t(bases[r + size_t(j)]) - int64_t(bases[0]);
                e = kernels::launch_kv_pipe_copy(L, stream);
                r += size_t(L.fan_n);
            }
            stats_.kernel_launches++;
        }
        if (e != cudaSuccess) {
            fail(std::string("multi-destination read failed to launch: ") + cudaGetErrorString(e));
            return -1;
        }
        ctx->mark(stream);
        stats_.bytes_read += uint64_t(n) * uint64_t(block_size) * bases.size();
    }
    return 0;
}

// read_cache fused with the layout swizzle of the attention consumer: pages are stored
// token-major ([tok][head][dim], as the prefill wrote them) and land head-major in a paged KV
// cache [page][head][tok][dim]; blocks[i].offset is the destination PAGE INDEX.  The
// transposition is done by the TMA unit (4-D tensor-map store, kernels/kv_pipe.cu).
int Connection::r_rdma_hnd(const std::vector<KeyOffset>& blocks, int tokens, int heads, int dim,
                           int elem_size, uint64_t base_ptr, uint64_t num_pages, int device,
                           uint64_t stream_in) {
    if (blocks.empty()) return 0;
    if (device < 0 || !server_hbm_) {
        fail("read_cache_hnd needs a CUDA destination and an HBM pool");
        return -1;
    }
    const int block_size = tokens * heads * dim * elem_size;
    for (const KeyOffset& b : blocks)
        if (b.offset >= num_pages) {
            fail("read_cache_hnd: page index beyond the destination tensor");
            return -1;
        }
    const bool via_index = device_lookup_ && device_index_usable();
    std::vector<RemoteBlock> rb;
    if (!via_index) {
        const int r = lookup_blocks(kOpReadLookup, blocks, block_size, rb);
        if (r != 0) return r;
    }
    NvtxRange nvtx("istore.read_hnd");
    std::lock_guard<std::mutex> lk(mu_);
    DevCtx* ctx = dev_ctx(device);
    if (!ctx) return -1;
    DeviceGuard g(device);
    if (doorbell_quiesce(ctx) != 0) return -1;
    cudaStream_t stream = ctx->pick(reinterpret_cast<cudaStream_t>(stream_in), false, streams_);
    stats_.calls++;
    for (size_t base = 0; base < blocks.size(); base += kMaxBatch) {
        const size_t n = std::min(kMaxBatch, blocks.size() - base);
        // dst_base 0: the descriptor's dst field carries the page index
        const kernels::CopyDesc* descs_d =
            resolve_descs(ctx, blocks, base, n, block_size, 0, via_index ? nullptr : &rb, stream);
        if (!descs_d) {
            fail("read_cache_hnd: cannot resolve the blocks");
            return -1;
        }
        kernels::HndLaunch H;
        H.descs = descs_d;
        H.n = uint32_t(n);
        H.tokens = uint32_t(tokens);
        H.heads = uint32_t(heads);
        H.dim = uint32_t(dim);
        H.elem_size = uint32_t(elem_size);
        H.dst_base = base_ptr;
        H.num_pages = uint32_t(num_pages);
        H.status = ctx->status_d;
        H.max_ctas = max_ctas_;
        const cudaError_t e = kernels::launch_kv_pipe_hnd(H, stream);
        if (e != cudaSuccess) {
            fail(std::string("layout-swizzling read failed to launch: ") + cudaGetErrorString(e));
            return -1;
        }
        stats_.kernel_launches++;
        ctx->mark(stream);
        stats_.bytes_read += uint64_t(n) * uint64_t(block_size);
    }
    return 0;
}

int Connection::match_via_device_index(const std::vector<std::string_view>& keys, bool exist_only) {
    std::lock_guard<std::mutex> lk(mu_);
    const int device = cfg_.device >= 0 ? cfg_.device : std::max(default_device_, 0);
    DevCtx* ctx = dev_ctx(device);
    if (!ctx) return -3;
    auto m0 = mapping(0, device);
    if (!m0 || !m0->dev_ptr || !segs_[0].index_slots) return -3;
    const size_t n = keys.size();
    size_t key_bytes = 0;
    std::vector<std::string_view> kp(n);
    for (size_t i = 0; i < n; ++i) {
        kp[i] = keys[i];
        key_bytes += align_up(std::max<size_t>(keys[i].size(), 1), 8);
    }
    if (key_bytes + n * 8 + 4096 > kRingBytes / 2) return -3;  // too large: use the control plane
    DeviceGuard g(device);
    if (doorbell_quiesce(ctx) != 0) return -3;
    const size_t at_bytes = ctx->ring_alloc(key_bytes);
    const size_t at_off = ctx->ring_alloc(n * 4);
    const size_t at_len = ctx->ring_alloc(n * 4);
    pack_keys(kp.data(), n, ctx->ring_h + at_bytes,
              reinterpret_cast<uint32_t*>(ctx->ring_h + at_off),
              reinterpret_cast<uint32_t*>(ctx->ring_h + at_len));
    kernels::LookupLaunch Q;
    Q.key_bytes = ctx->ring_d + at_bytes;
    Q.key_off = reinterpret_cast<const uint32_t*>(ctx->ring_d + at_off);
    Q.key_len = reinterpret_cast<const uint32_t*>(ctx->ring_d + at_len);
    Q.n = uint32_t(n);
    Q.table = reinterpret_cast<const kernels::IndexBucket*>(m0->dev_ptr + segs_[0].index_off);
    Q.table_mask = kernels::index_bucket_mask(segs_[0].index_slots);
    Q.shards = index_shards(ctx, nullptr);
    const size_t words = (n + 31) / 32;
    Q.present = reinterpret_cast<uint32_t*>(ctx->scratch + ctx->scratch_alloc(words * 4));
    Q.ticket = reinterpret_cast<uint32_t*>(ctx->zeros + ctx->zeros_alloc(4));
    Q.status = ctx->status_d;
    Q.want_match = true;
    Q.accept_claimed = !exist_only;  // C3: reserved-but-uncommitted keys count for match only
    // The launch is ordered after this connection's writes on the same stream, so keys
    // written just before (even without sync) are visible, as in the reference.
    cudaStream_t stream = ctx->pick(nullptr, false, std::max(streams_, 1));
    const cudaError_t e = kernels::launch_index_lookup(Q, stream);
    if (e != cudaSuccess) {
        fail(std::string("match kernel failed to launch: ") + cudaGetErrorString(e));
        return -3;
    }
    stats_.kernel_launches++;
    // external streams may hold this connection's writes: wait for them too
    ctx->mark(stream);
    ctx->wait_all();
    return int32_t(ctx->status_h[kernels::kStatMatch]);
}

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
    uint64_t sig = 1469598103934665603ull;
    auto mix_in = [&](uint64_t v) { sig = (sig ^ v) * 1099511628211ull; };
    for (uint32_t s = 0; s < L.nsegs; ++s) {
        uint8_t* base = nullptr;
        if (segs_[s].kind == kSegDeviceIpc) base = seg_dev_ptr(ctx, s);
        L.seg_base[s] = reinterpret_cast<uint64_t>(base);
        mix_in(L.seg_base[s]);
    }
    if (segs_[0].index_slots && L.seg_base[0]) {
        L.table = reinterpret_cast<kernels::IndexBucket*>(L.seg_base[0] + segs_[0].index_off);
        L.table_mask = kernels::index_bucket_mask(segs_[0].index_slots);
        L.shards = index_shards(ctx, nullptr);
    }
    mix_in(segs_.size());
    mix_in(reinterpret_cast<uint64_t>(L.table));
    mix_in(L.shards.n);
    L.epoch = ++db.epoch;
    L.first_seq = db.next_serve;
    const cudaError_t e = kernels::launch_kv_doorbell(L, db.stream);
    if (e != cudaSuccess) {
        fail(std::string("doorbell worker failed to launch: ") + cudaGetErrorString(e));
        return -1;
    }
    db.signature = sig;
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

int Connection::drain_devices(bool* device_error) {
    std::lock_guard<std::mutex> lk(mu_);
    int rc = 0;
    if (device_error) *device_error = false;
    struct AtExit {
        Connection* c;
        ~AtExit() { c->release_temporary_host_regs(); }
    } at_exit{this};
    for (auto& kv : devs_) {
        DevCtx& ctx = *kv.second;
        if (!ctx.dirty) continue;
        if (ctx.db && ctx.db->posted > ctx.db->collected) {
            DeviceGuard g(ctx.device);
            if (doorbell_wait(&ctx) != 0) {
                rc = -1;
                if (device_error) *device_error = true;
            }
        }
        ctx.wait_all();
        if (ctx.db) {  // no kernel is adding to the status words now
            ctx.status_h[kernels::kStatMiss] += ctx.db->misses;
            ctx.status_h[kernels::kStatStale] += ctx.db->stale;
            ctx.status_h[kernels::kStatPublishFail] += ctx.db->publish_failures;
            ctx.db->misses = ctx.db->stale = ctx.db->publish_failures = 0;
        }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            fail(std::string("device error during transfer: ") + cudaGetErrorString(e));
            rc = -1;
            if (device_error) *device_error = true;
        }
        if (ctx.status_h[kernels::kStatMiss]) {
            fail("read: " + std::to_string(ctx.status_h[kernels::kStatMiss]) +
                 " key(s) not found in the device index" +
                 (ctx.status_h[kernels::kStatStale]
                      ? " (" + std::to_string(ctx.status_h[kernels::kStatStale]) +
                            " evicted while being read)"
                      : std::string()));
            ctx.status_h[kernels::kStatMiss] = 0;
            ctx.status_h[kernels::kStatStale] = 0;
            rc = -kKeyNotFound;
        }
        if (ctx.status_h[kernels::kStatPublishFail]) {
            LOG_WARN("device index is full: %u block(s) are only reachable through the server; "
                     "reads fall back to server lookups",
                     ctx.status_h[kernels::kStatPublishFail]);
            publish_failures_ += ctx.status_h[kernels::kStatPublishFail];
            index_incomplete_.store(true, std::memory_order_relaxed);
            ctx.status_h[kernels::kStatPublishFail] = 0;
        }
    }
    return rc;
}

// ---------------------------------------------------------------- async API

void Connection::post(Task&& t) {
    {
        std::lock_guard<std::mutex> lk(q_mu_);
        queue_.push_back(std::move(t));
        ++inflight_async_;
    }
    q_cv_.notify_one();
}

void Connection::worker() {
    for (;;) {
        Task t;
        {
            std::unique_lock<std::mutex> lk(q_mu_);
            q_cv_.wait(lk, [this] { return !queue_.empty(); });
            t = std::move(queue_.front());
            queue_.pop_front();
        }
        if (t.kind == Task::kStop) return;
        if (t.kind == Task::kAllocate) {
            std::vector<RemoteBlock> out;
            std::vector<std::string_view> kv(t.keys.begin(), t.keys.end());
            if (allocate(kv, t.block_size, out) != 0) out.clear();
            if (t.alloc_cb) t.alloc_cb(std::move(out));
        } else {
            int status = t.status;
            if (t.event) {
                DeviceGuard g(t.device);
                if (cudaEventSynchronize(t.event) != cudaSuccess) status = -1;
                cudaEventDestroy(t.event);
            }
            if (t.commit && !t.commits.empty()) {
                if (status == 0) {
                    if (send_commit(t.commits.data(), t.commits.size()) != 0) status = -1;
                } else {
                    // this write's kernels failed: its blocks must never become visible
                    discard_blocks(t.commits.data(), t.commits.size());
                }
            }
            if (t.done_cb) t.done_cb(status);
        }
        {
            std::lock_guard<std::mutex> lk(q_mu_);
            --inflight_async_;
        }
        idle_cv_.notify_all();
    }
}

int Connection::allocate_async(const std::vector<std::string>& keys, int block_size,
                               std::function<void(std::vector<RemoteBlock>)> cb) {
    Task t;
    t.kind = Task::kAllocate;
    t.keys = keys;
    t.block_size = block_size;
    t.alloc_cb = std::move(cb);
    post(std::move(t));
    return 0;
}

int Connection::w_rdma_async(const std::vector<uint64_t>& offsets, int block_size,
                             const RemoteBlock* blocks, size_t nblocks, uint64_t base_ptr,
                             int device, uint64_t stream, std::function<void(int)> cb) {
    // The task owns the addresses of exactly the blocks this call wrote; its completion
    // commits those and no others (later writes of the connection may still be in flight on
    // other streams).  The event is recorded on the stream this call launched on.
    Task t;
    t.kind = Task::kWaitEvent;
    t.commit = true;
    MoveResult res;
    res.commits = &t.commits;
    const int r = w_rdma(offsets.data(), offsets.size(), 1, block_size, blocks, nblocks, base_ptr,
                         device, stream, &res);
    t.status = r;
    t.done_cb = std::move(cb);
    if (res.launched && res.device >= 0) {
        DeviceGuard g(res.device);
        if (cudaEventCreateWithFlags(&t.event, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventRecord(t.event, static_cast<cudaStream_t>(res.stream)) == cudaSuccess) {
            t.device = res.device;
        } else {
            // no event: fall back to waiting for the whole stream before the commit
            if (t.event) cudaEventDestroy(t.event);
            t.event = nullptr;
            (void)cudaGetLastError();
            if (cudaStreamSynchronize(static_cast<cudaStream_t>(res.stream)) != cudaSuccess) t.status = -1;
        }
    }
    post(std::move(t));
    return r;
}

int Connection::r_rdma_async(const std::vector<KeyOffset>& blocks, int block_size,
                             uint64_t base_ptr, int device, uint64_t stream,
                             std::function<void(int)> cb) {
    MoveResult res;
    const int r = r_rdma(blocks, block_size, base_ptr, device, stream, &res);
    Task t;
    t.kind = Task::kWaitEvent;
    t.status = r;
    t.done_cb = std::move(cb);
    if (r == 0 && res.launched && res.device >= 0) {
        DeviceGuard g(res.device);
        if (cudaEventCreateWithFlags(&t.event, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventRecord(t.event, static_cast<cudaStream_t>(res.stream)) == cudaSuccess) {
            t.device = res.device;
        } else {
            if (t.event) cudaEventDestroy(t.event);
            t.event = nullptr;
            (void)cudaGetLastError();
            if (cudaStreamSynchronize(static_cast<cudaStream_t>(res.stream)) != cudaSuccess) t.status = -1;
        }
    }
    post(std::move(t));
    return r;
}

}  // namespace istore
