// Server + client over loop-back TCP inside ONE native binary (host-memory pool, CPU buffers),
// built with AddressSanitizer / UBSan: the control-plane code paths the Python suite drives,
// but with every heap access checked.  Build + run: tools/build_native.py (build_loopback_test).
#include <arpa/inet.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "ctrl/client.h"
#include "ctrl/server.h"
#include "core/log.h"

using namespace istore;

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        ++g_checks;                                                              \
        if (!(cond)) {                                                           \
            ++g_failed;                                                          \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
        }                                                                        \
    } while (0)

static std::unique_ptr<Connection> connect_to(int port, int timeout_ms = 5000) {
    ClientConfig cc;
    cc.host_addr = "127.0.0.1";
    cc.service_port = port;
    cc.timeout_ms = timeout_ms;
    auto c = std::make_unique<Connection>();
    if (c->init_connection(cc) != 0 || c->setup_rdma(cc) != 0) return nullptr;
    return c;
}

static void store_round_trips(int port, Server& srv) {
    auto c = connect_to(port);
    CHECK(c != nullptr);
    if (!c) return;
    const int n = 64, bs = 4096;
    std::vector<uint8_t> src(size_t(n) * bs), dst(size_t(n) * bs, 0);
    std::mt19937 rng{7};
    for (auto& b : src) b = uint8_t(rng());
    CHECK(c->register_mr(reinterpret_cast<uint64_t>(src.data()), src.size(), -1) > 0);
    CHECK(c->register_mr(reinterpret_cast<uint64_t>(dst.data()), dst.size(), -1) > 0);
    std::vector<std::string> names;
    for (int i = 0; i < n; ++i) names.push_back("loop/key/" + std::to_string(i));
    std::vector<std::string_view> keys(names.begin(), names.end());
    std::vector<RemoteBlock> blocks;
    CHECK(c->allocate(keys, bs, blocks) == 0 && blocks.size() == size_t(n));
    std::vector<uint64_t> offs(static_cast<size_t>(n), 0);
    for (int i = 0; i < n; ++i) offs[size_t(i)] = uint64_t(i) * bs;
    CHECK(c->w_rdma(offs.data(), offs.size(), 1, bs, blocks.data(), blocks.size(),
                    reinterpret_cast<uint64_t>(src.data()), -1, 0) == 0);
    CHECK(c->check_exist(names[3]) == 1);  // not visible before sync
    CHECK(c->sync_rdma() >= 0);
    CHECK(c->check_exist(names[3]) == 0);
    std::vector<KeyOffset> rb;
    for (int i = n - 1; i >= 0; --i) rb.push_back(KeyOffset{names[size_t(i)], uint64_t(i) * bs});
    CHECK(c->r_rdma(rb, bs, reinterpret_cast<uint64_t>(dst.data()), -1, 0) == 0);
    CHECK(c->sync_rdma() >= 0);
    CHECK(src == dst);
    // dedup: the same keys again are fake blocks, a read of a missing key is an error
    std::vector<RemoteBlock> again;
    CHECK(c->allocate(keys, bs, again) == 0 && is_fake_block(again[0]) && is_fake_block(again[63]));
    std::vector<KeyOffset> missing{KeyOffset{"loop/none", 0}};
    CHECK(c->r_rdma(missing, bs, reinterpret_cast<uint64_t>(dst.data()), -1, 0) < 0);
    std::vector<std::string_view> probe{names[0], names[1], "x", "y"};
    CHECK(c->get_match_last_index(probe) == 1);
    CHECK(c->touch(keys) >= 0);
    CHECK(srv.stats().keys == uint64_t(n));
    c->close();
}

// The async API: callbacks run on the connection's completion thread.
static void async_api(int port) {
    auto c = connect_to(port);
    CHECK(c != nullptr);
    if (!c) return;
    const int n = 16, bs = 4096;
    std::vector<uint8_t> src(size_t(n) * bs, 0x77), dst(size_t(n) * bs, 0);
    CHECK(c->register_mr(reinterpret_cast<uint64_t>(src.data()), src.size(), -1) > 0);
    CHECK(c->register_mr(reinterpret_cast<uint64_t>(dst.data()), dst.size(), -1) > 0);
    std::vector<std::string> names;
    for (int i = 0; i < n; ++i) names.push_back("async/" + std::to_string(i));
    std::mutex mu;
    std::condition_variable cv;
    std::vector<RemoteBlock> blocks;
    int stage = 0, status = -99;
    c->allocate_async(names, bs, [&](std::vector<RemoteBlock> v) {
        std::lock_guard<std::mutex> lk(mu);
        blocks = std::move(v);
        stage = 1;
        cv.notify_all();
    });
    {
        std::unique_lock<std::mutex> lk(mu);
        CHECK(cv.wait_for(lk, std::chrono::seconds(10), [&] { return stage == 1; }));
    }
    CHECK(blocks.size() == size_t(n));
    std::vector<uint64_t> offs(static_cast<size_t>(n), 0);
    for (int i = 0; i < n; ++i) offs[size_t(i)] = uint64_t(i) * bs;
    c->w_rdma_async(offs, bs, blocks.data(), blocks.size(), reinterpret_cast<uint64_t>(src.data()),
                    -1, 0, [&](int st) {
                        std::lock_guard<std::mutex> lk(mu);
                        status = st;
                        stage = 2;
                        cv.notify_all();
                    });
    {
        std::unique_lock<std::mutex> lk(mu);
        CHECK(cv.wait_for(lk, std::chrono::seconds(10), [&] { return stage == 2; }));
    }
    CHECK(status == 0);
    CHECK(c->sync_rdma() >= 0);
    std::vector<KeyOffset> rb;
    for (int i = 0; i < n; ++i) rb.push_back(KeyOffset{names[size_t(i)], uint64_t(i) * bs});
    c->r_rdma_async(rb, bs, reinterpret_cast<uint64_t>(dst.data()), -1, 0, [&](int st) {
        std::lock_guard<std::mutex> lk(mu);
        status = st;
        stage = 3;
        cv.notify_all();
    });
    {
        std::unique_lock<std::mutex> lk(mu);
        CHECK(cv.wait_for(lk, std::chrono::seconds(10), [&] { return stage == 3; }));
    }
    CHECK(status == 0 && src == dst);
    c->close();
}

// Several threads share ONE connection (the layer-wise upload pattern: a compute thread and
// an upload thread, example/demo_prefill.py): every call is thread-safe.
static void shared_connection_threads(int port) {
    auto c = connect_to(port);
    CHECK(c != nullptr);
    if (!c) return;
    const int bs = 4096, per = 8, rounds = 12, nthreads = 4;
    std::vector<std::vector<uint8_t>> bufs(nthreads, std::vector<uint8_t>(size_t(per) * bs));
    std::vector<std::vector<uint8_t>> outs(nthreads, std::vector<uint8_t>(size_t(per) * bs));
    for (int t = 0; t < nthreads; ++t) {
        CHECK(c->register_mr(reinterpret_cast<uint64_t>(bufs[size_t(t)].data()), bufs[size_t(t)].size(), -1) > 0);
        CHECK(c->register_mr(reinterpret_cast<uint64_t>(outs[size_t(t)].data()), outs[size_t(t)].size(), -1) > 0);
    }
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t] {
            std::vector<uint8_t>& src = bufs[size_t(t)];
            std::vector<uint8_t>& dst = outs[size_t(t)];
            std::vector<uint64_t> offs(static_cast<size_t>(per), 0);
            for (int i = 0; i < per; ++i) offs[size_t(i)] = uint64_t(i) * bs;
            for (int r = 0; r < rounds; ++r) {
                std::memset(src.data(), t * 16 + r, src.size());
                std::vector<std::string> names;
                for (int i = 0; i < per; ++i)
                    names.push_back("shared/" + std::to_string(t) + "/" + std::to_string(r) + "/" + std::to_string(i));
                std::vector<std::string_view> keys(names.begin(), names.end());
                std::vector<RemoteBlock> blocks;
                if (c->allocate(keys, bs, blocks) != 0) { ++bad; continue; }
                if (c->w_rdma(offs.data(), offs.size(), 1, bs, blocks.data(), blocks.size(),
                              reinterpret_cast<uint64_t>(src.data()), -1, 0) != 0) ++bad;
                if (c->sync_rdma() < 0) ++bad;
                std::vector<KeyOffset> rb;
                for (int i = 0; i < per; ++i) rb.push_back(KeyOffset{names[size_t(i)], uint64_t(i) * bs});
                if (c->r_rdma(rb, bs, reinterpret_cast<uint64_t>(dst.data()), -1, 0) != 0) ++bad;
                if (c->sync_rdma() < 0) ++bad;
                if (src != dst) ++bad;
                if (c->check_exist(names[0]) != 0) ++bad;
            }
        });
    }
    for (auto& x : th) x.join();
    CHECK(bad.load() == 0);
    c->close();
}

static void eviction_and_dead_writers(int port, Server& srv) {
    auto w = connect_to(port);
    CHECK(w != nullptr);
    if (!w) return;
    const int bs = 16384;
    std::vector<uint8_t> buf(size_t(bs) * 8, 0x5a);
    CHECK(w->register_mr(reinterpret_cast<uint64_t>(buf.data()), buf.size(), -1) > 0);
    std::vector<uint64_t> offs(8);
    for (int i = 0; i < 8; ++i) offs[size_t(i)] = uint64_t(i) * bs;
    const uint64_t evicted_before = srv.stats().evicted;
    for (int round = 0; round < 40; ++round) {  // far more than the pool holds
        std::vector<std::string> names;
        for (int i = 0; i < 8; ++i) names.push_back("ev/" + std::to_string(round) + "/" + std::to_string(i));
        std::vector<std::string_view> keys(names.begin(), names.end());
        std::vector<RemoteBlock> blocks;
        if (w->allocate(keys, bs, blocks) != 0) continue;
        CHECK(w->w_rdma(offs.data(), 8, 1, bs, blocks.data(), 8,
                        reinterpret_cast<uint64_t>(buf.data()), -1, 0) == 0);
        CHECK(w->sync_rdma() >= 0);
    }
    CHECK(srv.stats().evicted > evicted_before);
    // a writer that disappears with reservations and a staged commit
    {
        auto d = connect_to(port);
        CHECK(d != nullptr);
        std::vector<std::string> names{"dead/a", "dead/b"};
        std::vector<std::string_view> keys(names.begin(), names.end());
        std::vector<RemoteBlock> blocks;
        CHECK(d && d->allocate(keys, bs, blocks) == 0);
        d->close();
    }
    for (int i = 0; i < 500 && srv.stats().inflight; ++i)
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    CHECK(srv.stats().inflight == 0);
    w->close();
}

static void garbage_on_the_wire(int port, Server& srv) {
    std::mt19937 rng{99};
    for (int round = 0; round < 200; ++round) {
        const int fd = socket(AF_INET, SOCK_STREAM, 0);
        sockaddr_in a{};
        a.sin_family = AF_INET;
        a.sin_port = htons(uint16_t(port));
        inet_pton(AF_INET, "127.0.0.1", &a.sin_addr);
        if (connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0) {
            close(fd);
            continue;
        }
        std::vector<uint8_t> junk(1 + rng() % 300);
        for (auto& b : junk) b = uint8_t(rng());
        if (round % 2 == 0 && junk.size() >= 9) {  // valid magic + known op, random body
            const uint32_t magic = kMagic;
            std::memcpy(junk.data(), &magic, 4);
            junk[4] = uint8_t("RWSEDATCMPUH"[rng() % 12]);
            const uint32_t len = uint32_t(junk.size() - 9);
            std::memcpy(junk.data() + 5, &len, 4);
        }
        (void)!send(fd, junk.data(), junk.size(), MSG_NOSIGNAL);
        timeval tv{0, 20000};
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        uint8_t sink[512];
        (void)!recv(fd, sink, sizeof(sink), 0);
        close(fd);
    }
    CHECK(srv.running());
    auto c = connect_to(port);
    CHECK(c && c->check_exist("still-alive") == 1);
}

int main() {
    set_log_level("error");  // the fuzz section would print hundreds of expected warnings
    ServerConfig sc;
    sc.service_port = 0;
    sc.host = "127.0.0.1";
    sc.pool_backend = "host";
    sc.prealloc_bytes = 64 * 16384;
    sc.minimal_allocate_size = 4;
    sc.evict = true;
    sc.evict_ratio = 0.1;
    Server srv(sc);
    std::string err;
    if (srv.start(&err) != 0) {
        std::fprintf(stderr, "server start failed: %s\n", err.c_str());
        return 2;
    }
    const int port = srv.port();
    store_round_trips(port, srv);
    srv.purge();
    async_api(port);
    srv.purge();
    shared_connection_threads(port);
    srv.purge();
    eviction_and_dead_writers(port, srv);
    garbage_on_the_wire(port, srv);
    // checkpoint / resume through the native API
    {
        auto c = connect_to(port);
        std::vector<std::string> names{"ck/a", "ck/b"};
        std::vector<std::string_view> keys(names.begin(), names.end());
        std::vector<RemoteBlock> blocks;
        std::vector<uint8_t> buf(2 * 4096, 0x33);
        std::vector<uint64_t> offs{0, 4096};
        CHECK(c && c->register_mr(reinterpret_cast<uint64_t>(buf.data()), buf.size(), -1) > 0);
        CHECK(c->allocate(keys, 4096, blocks) == 0);
        CHECK(c->w_rdma(offs.data(), 2, 1, 4096, blocks.data(), 2,
                        reinterpret_cast<uint64_t>(buf.data()), -1, 0) == 0);
        CHECK(c->sync_rdma() >= 0);
        const std::string path = "/tmp/istore_loopback.ckpt";
        CHECK(srv.dump(path, &err) >= 2);
        srv.purge();
        CHECK(c->check_exist("ck/a") == 1);
        CHECK(srv.load(path, &err) >= 2);
        CHECK(c->check_exist("ck/a") == 0 && c->check_exist("ck/b") == 0);
        unlink(path.c_str());
    }
    srv.stop();
    std::printf("%d checks, %d failed\n", g_checks, g_failed);
    return g_failed ? 1 : 0;
}
