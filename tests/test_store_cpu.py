"""End-to-end store semantics on CPU loop-back (host-memory pool, no GPU).

Ports the behaviour the reference's integration tests assert
(infinistore/test_infinistore.py:61-417) to the CPU plumbing configuration of BASELINE.json
("server + client write/read 16 keys x 4 KB on CPU loopback"), plus the contract items of
SURVEY §2.5 and the reference defects that must NOT be reproduced (D1-D5, D10-D12).
"""
import asyncio
import os
import multiprocessing as mp
import random
import socket
import string
import struct
import time

import numpy as np
import pytest
import torch

import infinistore_b200 as ist
from conftest import make_conn


def rand_key(n=10):
    return "".join(random.choice(string.ascii_letters + string.digits) for _ in range(n))


# ------------------------------------------------------------------ reference test ports
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("new_connection", [True, False])
def test_basic_read_write_cache(host_server, dtype, new_connection):
    _, port = host_server
    conn = make_conn(port)
    key = rand_key()
    src = torch.arange(4096).to(dtype)
    conn.register_mr(src)
    blocks = conn.allocate_rdma([key], 4096 * src.element_size())
    conn.rdma_write_cache(src, [0], 4096, blocks)
    conn.sync()
    if new_connection:
        conn = make_conn(port)
    dst = torch.zeros(4096, dtype=dtype)
    conn.register_mr(dst)
    conn.read_cache(dst, [(key, 0)], 4096)
    conn.sync()
    assert torch.equal(src, dst)


def test_plumbing_config_16_keys_x_4kb(host_server):
    """BASELINE.json config 1."""
    _, port = host_server
    conn = make_conn(port)
    keys = [rand_key() for _ in range(16)]
    src = torch.randn(16 * 1024)  # 16 x 4 KB of fp32
    dst = torch.zeros_like(src)
    conn.register_mr(src)
    conn.register_mr(dst)
    blocks = conn.allocate_rdma(keys, 4096)
    conn.rdma_write_cache(src, [i * 1024 for i in range(16)], 1024, blocks)
    conn.sync()
    conn.read_cache(dst, [(k, i * 1024) for i, k in enumerate(keys)], 1024)
    conn.sync()
    assert torch.equal(src, dst)
    assert conn.stats()["host_copies"] == 32


def test_batch_read_write_cache(host_server):
    _, port = host_server
    conn = make_conn(port)
    nblocks, bs = 10, 4096
    src = torch.arange(nblocks * bs, dtype=torch.float32)
    conn.register_mr(src)
    for _ in range(3):
        keys = [rand_key() for _ in range(nblocks)]
        blocks = [(keys[i], i * bs) for i in range(nblocks)]
        remote = conn.allocate_rdma(keys, bs * 4)
        conn.rdma_write_cache(src, [i * bs for i in range(nblocks)], bs, remote)
        conn.sync()
        dst = torch.zeros(nblocks * bs, dtype=torch.float32)
        conn.register_mr(dst)
        conn.read_cache(dst, blocks, bs)
        conn.sync()
        assert torch.equal(src, dst)


def _client_proc(port, q):
    try:
        conn = make_conn(port)
        key = rand_key()
        src = torch.arange(4096, dtype=torch.float32)
        conn.register_mr(src)
        blocks = conn.allocate_rdma([key], 4096 * 4)
        conn.rdma_write_cache(src, [0], 4096, blocks)
        conn.sync()
        conn = make_conn(port)
        dst = torch.zeros(4096, dtype=torch.float32)
        conn.read_cache(dst, [(key, 0)], 4096)
        conn.sync()
        q.put(bool(torch.equal(src, dst)))
    except Exception as e:  # pragma: no cover
        q.put(repr(e))


def test_multiple_clients(host_server):
    _, port = host_server
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_client_proc, args=(port, q)) for _ in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert results == [True, True]


def test_key_check(host_server):
    _, port = host_server
    conn = make_conn(port)
    key = rand_key(5)
    src = torch.randn(4096)
    conn.register_mr(src)
    blocks = conn.allocate_rdma([key], 4096 * 4)
    assert not conn.check_exist(key)  # reserved but not committed (C3)
    conn.rdma_write_cache(src, [0], 4096, blocks)
    conn.sync()
    assert conn.check_exist(key)
    assert not conn.check_exist("never-written")


def test_get_match_last_index(host_server):
    _, port = host_server
    conn = make_conn(port)
    src = torch.randn(4096)
    conn.register_mr(src)
    blocks = conn.allocate_rdma(["key1", "key2", "key3"], 1024 * 4)
    conn.rdma_write_cache(src, [0, 1024, 2048], 1024, blocks)
    # no sync on purpose: match does not require `committed` (C3), and the non-monotone
    # input must give exactly what the reference's binary search gives (C7)
    assert conn.get_match_last_index(["A", "B", "C", "key1", "D", "E"]) == 3
    assert conn.get_match_last_index(["key1", "key2", "key3", "X"]) == 2
    assert conn.get_match_last_index(["key1"]) == 0
    with pytest.raises(Exception, match="can't find a match"):
        conn.get_match_last_index(["A", "B"])


def test_key_not_found(host_server):
    _, port = host_server
    conn = make_conn(port)
    dst = torch.zeros(4096)
    with pytest.raises(Exception):
        conn.read_cache(dst, [("not_exist_key", 0)], 4096)
    # a reserved-but-uncommitted key is not readable either
    conn.allocate_rdma(["pending"], 4096 * 4)
    with pytest.raises(Exception):
        conn.read_cache(dst, [("pending", 0)], 4096)


def test_deduplicate_first_writer_wins(host_server):
    srv, port = host_server
    conn = make_conn(port)
    key = "duplicate_key"
    src = torch.arange(4096, dtype=torch.float32)
    conn.register_mr(src)
    conn.rdma_write_cache(src, [0], 4096, conn.allocate_rdma([key], 4096 * 4))
    conn.sync()
    used = srv.stats()["used_bytes"]
    src2 = torch.randn(4096)
    conn.register_mr(src2)
    again = conn.allocate_rdma([key], 4096 * 4)
    assert again["rkey"][0] == 0 and again["remote_addr"][0] == 0  # fake block sentinel
    conn.rdma_write_cache(src2, [0], 4096, again)  # silently skipped
    conn.sync()
    assert srv.stats()["used_bytes"] == used  # dedup does not leak pool space (ref. defect D1)
    dst = torch.zeros(4096)
    conn.read_cache(dst, [(key, 0)], 4096)
    conn.sync()
    assert torch.equal(src, dst) and not torch.equal(src2, dst)


def test_server_counts_hits_misses_and_dedup(host_server):
    srv, port = host_server
    conn = make_conn(port)
    src = torch.randn(2 * 1024)
    dst = torch.zeros(1024)
    conn.register_mr(src)
    conn.rdma_write_cache(src, [0, 1024], 1024, conn.allocate_rdma(["h-a", "h-b"], 4096))
    conn.sync()
    again = conn.allocate_rdma(["h-a", "h-c"], 4096)  # one duplicate, one fresh
    assert again["rkey"][0] == 0 and again["rkey"][1] != 0
    conn.read_cache(dst, [("h-a", 0)], 1024)
    conn.read_cache(dst, [("h-b", 0)], 1024)
    with pytest.raises(Exception):
        conn.read_cache(dst, [("h-missing", 0)], 1024)
    conn.sync()
    st = srv.stats()
    assert (st["lookup_hits"], st["lookup_misses"], st["dedup_skips"]) == (2, 1, 1)


def test_partial_dedup_batch_completes(host_server):
    """Reference defect D2: a batch whose LAST block is a duplicate must still complete."""
    _, port = host_server
    conn = make_conn(port)
    src = torch.randn(2 * 1024)
    conn.register_mr(src)
    conn.rdma_write_cache(src, [1024], 1024, conn.allocate_rdma(["tail"], 4096))
    conn.sync()
    blocks = conn.allocate_rdma(["head", "tail"], 4096)
    assert blocks["rkey"].tolist()[1] == 0

    async def run():
        await conn.rdma_write_cache_async(src, [0, 1024], 1024, blocks)

    asyncio.run(asyncio.wait_for(run(), 10))
    conn.sync()
    assert conn.check_exist("head")


def test_async_api(host_server):
    _, port = host_server
    cfg = ist.ClientConfig(host_addr="127.0.0.1", service_port=port,
                           connection_type=ist.TYPE_RDMA)
    conn = ist.InfinityConnection(cfg)

    async def run():
        await conn.connect_async()
        key = rand_key(5)
        src = torch.randn(4096)
        dst = torch.zeros(4096)
        await asyncio.to_thread(lambda: (conn.register_mr(src), conn.register_mr(dst)))
        remote = await conn.allocate_rdma_async([key], 4096 * 4)
        await conn.rdma_write_cache_async(src, [0], 4096, remote)
        await conn.read_cache_async(dst, [(key, 0)], 4096)
        assert torch.equal(src, dst)

    asyncio.run(asyncio.wait_for(run(), 20))


def test_local_connection_requires_cuda_tensor_and_localhost(host_server):
    _, port = host_server
    conn = make_conn(port, connection_type=ist.TYPE_LOCAL_GPU)
    with pytest.raises(Exception, match="CUDA"):
        conn.local_gpu_write_cache(torch.zeros(16), [("k", 0)], 16)
    cfg = ist.ClientConfig(host_addr="10.1.2.3", service_port=port,
                           connection_type=ist.TYPE_LOCAL_GPU)
    with pytest.raises(Exception, match="localhost"):
        ist.InfinityConnection(cfg).connect()


def test_config_verify_matches_reference_rules():
    with pytest.raises(Exception):
        ist.ClientConfig(host_addr="127.0.0.1", service_port=1).verify()  # no connection type
    with pytest.raises(Exception):
        ist.ClientConfig(connection_type=ist.TYPE_RDMA, host_addr="", service_port=1).verify()
    with pytest.raises(Exception):
        ist.ClientConfig(connection_type=ist.TYPE_RDMA, host_addr="h", service_port=0).verify()
    with pytest.raises(Exception):  # the reference's own tests use ports > 65535 (defect D8)
        ist.ClientConfig(connection_type=ist.TYPE_RDMA, host_addr="h", service_port=92345).verify()
    with pytest.raises(Exception):
        ist.ClientConfig(connection_type=ist.TYPE_RDMA, host_addr="h", service_port=1,
                         link_type="foo").verify()
    with pytest.raises(Exception):
        ist.ServerConfig(service_port=1, manage_port=2, minimal_allocate_size=8).verify()
    with pytest.raises(Exception):
        ist.ServerConfig(service_port=1, manage_port=0).verify()
    ist.ServerConfig(service_port=1, manage_port=2).verify()
    c = ist.ServerConfig()
    assert (c.prealloc_size, c.minimal_allocate_size, c.num_stream, c.auto_increase) == \
        (16, 64, 1, False)


def test_noncontiguous_tensor_rejected(host_server):
    _, port = host_server
    conn = make_conn(port)
    t = torch.zeros(64, 64).t()
    with pytest.raises(Exception, match="contiguous"):
        conn.register_mr(t)


# ------------------------------------------------------------------ pool behaviour
def _server(**kw):
    from infinistore_b200 import _infinistore as m

    cfg = m.ServerConfig()
    cfg.service_port = 0
    cfg.host = "127.0.0.1"
    cfg.pool_backend = "host"
    cfg.minimal_allocate_size = 16
    for k, v in kw.items():
        setattr(cfg, k, v)
    srv = m.Server(cfg)
    return srv, srv.start()


def test_out_of_memory_is_reported_and_nothing_is_half_allocated():
    srv, port = _server(prealloc_bytes=8 * 16384)
    try:
        conn = make_conn(port)
        with pytest.raises(Exception):
            conn.allocate_rdma([f"k{i}" for i in range(9)], 16384)  # 507, not a timeout (D3)
        assert srv.kvmap_len() == 0 and srv.stats()["used_bytes"] == 0  # (D4)
        ok = conn.allocate_rdma([f"k{i}" for i in range(8)], 16384)
        assert len(ok) == 8
        with pytest.raises(Exception):
            conn.allocate_rdma(["one-more"], 16384)
        assert srv.purge() == 8
        assert len(conn.allocate_rdma(["one-more"], 16384)) == 1  # purge returns the space (D10)
    finally:
        srv.stop()


def test_full_pool_evicts_least_recently_used_when_enabled():
    """The reference answers 507 forever once the pool is full; with --evict the
    least recently used committed blocks make room."""
    srv, port = _server(prealloc_bytes=8 * 16384, evict=True, evict_ratio=0.25)
    try:
        conn = make_conn(port)
        src = torch.arange(12 * 4096, dtype=torch.float32)
        dst = torch.zeros(4096)
        conn.register_mr(src)
        keys = [f"k{i}" for i in range(8)]
        conn.rdma_write_cache(src, [i * 4096 for i in range(8)], 4096,
                              conn.allocate_rdma(keys, 16384))
        conn.sync()
        assert srv.stats()["used_bytes"] == 8 * 16384
        # touch k0: a read makes it the most recently used block
        conn.read_cache(dst, [("k0", 0)], 4096)
        conn.sync()
        assert torch.equal(dst, src[:4096])
        # the pool is full: two more blocks push out the two oldest ones (25 % per round)
        more = conn.allocate_rdma(["k8", "k9"], 16384)
        assert len(more) == 2 and (more["rkey"] != 0).all()
        conn.rdma_write_cache(src, [8 * 4096, 9 * 4096], 4096, more)
        conn.sync()
        st = srv.stats()
        assert st["evicted"] == 2 and st["keys"] == 8 and st["used_bytes"] == 8 * 16384
        assert not conn.check_exist("k1") and not conn.check_exist("k2")
        assert conn.check_exist("k0") and conn.check_exist("k3") and conn.check_exist("k9")
        with pytest.raises(Exception):
            conn.read_cache(dst, [("k1", 0)], 4096)
        conn.read_cache(dst, [("k9", 0)], 4096)
        conn.sync()
        assert torch.equal(dst, src[9 * 4096:10 * 4096])
        # an evicted key can be written again (it is a new block)
        again = conn.allocate_rdma(["k1"], 16384)
        assert again["rkey"][0] != 0
        conn.rdma_write_cache(src, [11 * 4096], 4096, again)
        conn.sync()
        conn.read_cache(dst, [("k1", 0)], 4096)
        conn.sync()
        assert torch.equal(dst, src[11 * 4096:12 * 4096])
    finally:
        srv.stop()


def test_touch_refreshes_recency_for_readers_the_server_does_not_see():
    srv, port = _server(prealloc_bytes=8 * 16384, evict=True, evict_ratio=0.25)
    try:
        conn = make_conn(port)
        src = torch.randn(10 * 4096)
        conn.register_mr(src)
        keys = [f"k{i}" for i in range(8)]
        conn.rdma_write_cache(src, [i * 4096 for i in range(8)], 4096,
                              conn.allocate_rdma(keys, 16384))
        conn.sync()
        # the oldest two would be evicted next; a cache manager says they were just used
        assert conn.touch(["k0", "k1", "not-there"]) == 2
        conn.rdma_write_cache(src, [8 * 4096, 9 * 4096], 4096,
                              conn.allocate_rdma(["k8", "k9"], 16384))
        conn.sync()
        assert conn.check_exist("k0") and conn.check_exist("k1")
        assert not conn.check_exist("k2") and not conn.check_exist("k3")
    finally:
        srv.stop()
    # a store that does not evict keeps no recency order: the hint is a no-op
    srv, port = _server(prealloc_bytes=8 * 16384)
    try:
        conn = make_conn(port)
        conn.rdma_write_cache(src, [0], 4096, conn.allocate_rdma(["x"], 16384))
        conn.sync()
        assert conn.touch(["x"]) == 0
    finally:
        srv.stop()


def test_eviction_skips_blocks_a_reader_holds_and_uncommitted_ones():
    srv, port = _server(prealloc_bytes=4 * 16384, evict=True, evict_ratio=0.25)
    try:
        reader, writer = make_conn(port), make_conn(port)
        src = torch.randn(8 * 4096)
        writer.register_mr(src)
        writer.rdma_write_cache(src, [0, 4096], 4096, writer.allocate_rdma(["a", "b"], 16384))
        writer.sync()
        dst = torch.zeros(4096)
        reader.read_cache(dst, [("a", 0)], 4096)  # pinned until the reader's next sync
        pending = writer.allocate_rdma(["c", "d"], 16384)  # reserved, never committed: pool full
        assert len(pending) == 2
        # only "b" can go: "a" is leased, "c"/"d" are in flight
        one = writer.allocate_rdma(["e"], 16384)
        assert one["rkey"][0] != 0 and srv.stats()["evicted"] == 1
        assert not writer.check_exist("b")
        with pytest.raises(Exception):
            writer.allocate_rdma(["f"], 16384)  # nothing evictable left: 507
        reader.sync()  # releases the lease on "a"
        assert torch.equal(dst, src[:4096])
        assert writer.allocate_rdma(["f"], 16384)["rkey"][0] != 0
        assert srv.stats()["evicted"] == 2 and not writer.check_exist("a")
    finally:
        srv.stop()


def test_concurrent_clients_on_an_evicting_store_never_read_foreign_bytes():
    """Four writer/reader threads hammer a pool that holds a fraction of what they write.
    A read may miss (the block was evicted) but a read that succeeds returns the writer's
    bytes: leases keep blocks alive while they are read, staged commits of one connection
    never commit another connection's blocks."""
    import threading

    srv, port = _server(prealloc_bytes=96 * 16384, evict=True, evict_ratio=0.1)
    errors, stats = [], {"hits": 0, "misses": 0}
    lock = threading.Lock()

    def worker(tid):
        try:
            conn = make_conn(port)
            rng = np.random.default_rng(tid)
            src = torch.zeros(8 * 4096)
            dst = torch.zeros(4096)
            conn.register_mr(src)
            mine = []
            for it in range(60):
                keys = [f"t{tid}-i{it}-b{b}" for b in range(8)]
                for b in range(8):
                    src[b * 4096:(b + 1) * 4096] = float(tid * 100000 + it * 8 + b)
                try:
                    blocks = conn.allocate_rdma(keys, 16384)
                except Exception:
                    continue  # everything evictable is leased or in flight right now
                conn.rdma_write_cache(src, [b * 4096 for b in range(8)], 4096, blocks)
                conn.sync()
                mine.extend((k, float(tid * 100000 + it * 8 + b)) for b, k in enumerate(keys))
                for _ in range(4):
                    k, want = mine[int(rng.integers(0, len(mine)))]
                    try:
                        conn.read_cache(dst, [(k, 0)], 4096)
                        conn.sync()
                    except Exception:
                        with lock:
                            stats["misses"] += 1
                        continue
                    with lock:
                        stats["hits"] += 1
                    if not bool((dst == want).all()):
                        errors.append((k, want, float(dst[0])))
            conn.close()
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    try:
        threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(120)
        assert not errors, errors[:3]
        st = srv.stats()
        assert st["evicted"] > 0 and stats["hits"] > 0 and stats["misses"] > 0
        assert st["used_bytes"] <= 96 * 16384 and st["inflight"] == 0
    finally:
        srv.stop()


def test_auto_increase_adds_segments_and_clients_map_them():
    srv, port = _server(prealloc_bytes=8 * 16384, auto_increase=True)
    try:
        conn = make_conn(port)
        src = torch.randn(20 * 4096)
        dst = torch.zeros_like(src)
        conn.register_mr(src)
        keys = [f"k{i}" for i in range(20)]
        blocks = conn.allocate_rdma(keys, 16384)
        assert srv.stats()["segments"] >= 3
        assert len(set(blocks["rkey"].tolist())) >= 3
        conn.rdma_write_cache(src, [i * 4096 for i in range(20)], 4096, blocks)
        conn.sync()
        conn.read_cache(dst, [(k, i * 4096) for i, k in enumerate(keys)], 4096)
        conn.sync()
        assert torch.equal(src, dst)
    finally:
        srv.stop()


def test_block_spanning_several_granules_and_short_reads():
    srv, port = _server(prealloc_bytes=64 * 16384)
    try:
        conn = make_conn(port)
        src = torch.randn(40000 // 4)
        conn.register_mr(src)
        b = conn.allocate_rdma(["big"], 40000)  # 3 granules
        assert srv.stats()["used_bytes"] == 3 * 16384
        conn.rdma_write_cache(src, [0], 10000, b)
        conn.sync()
        dst = torch.zeros(10000)
        conn.read_cache(dst, [("big", 0)], 10000)
        conn.sync()
        assert torch.equal(src, dst)
        with pytest.raises(Exception):  # never read past what was written
            conn.read_cache(torch.zeros(20000), [("big", 0)], 20000)
    finally:
        srv.stop()


def test_dead_writer_does_not_leave_reserved_keys(host_server):
    srv, port = host_server
    conn = make_conn(port)
    conn.allocate_rdma(["orphan"], 4096)
    assert srv.kvmap_len() == 1
    conn.close()
    deadline = time.time() + 5
    while srv.kvmap_len() and time.time() < deadline:
        time.sleep(0.01)
    assert srv.kvmap_len() == 0 and srv.stats()["used_bytes"] == 0
    conn2 = make_conn(port)
    assert len(conn2.allocate_rdma(["orphan"], 4096)) == 1  # writable again


def test_purge_while_reader_holds_lease(host_server):
    srv, port = host_server
    conn = make_conn(port)
    src = torch.randn(4096)
    conn.register_mr(src)
    conn.rdma_write_cache(src, [0], 4096, conn.allocate_rdma(["p"], 16384))
    conn.sync()
    dst = torch.zeros(4096)
    conn.read_cache(dst, [("p", 0)], 4096)  # lookup pins the block until the next sync
    assert srv.purge() == 1
    # the leased block left the map but keeps its pool space until the reader's sync(): a new
    # writer can never be handed memory that an in-flight copy is still reading
    assert srv.stats()["used_bytes"] == 16384
    conn.sync()
    assert srv.stats()["used_bytes"] == 0
    assert not conn.check_exist("p")


def test_purge_does_not_hand_leased_space_to_a_new_writer():
    """A pool of exactly two blocks, both leased by a reader: after purge() an allocation
    must fail (507) until the reader's sync() releases them (purge dropped leases)."""
    from infinistore_b200 import _infinistore as m

    cfg = m.ServerConfig()
    cfg.service_port = 0
    cfg.host = "127.0.0.1"
    cfg.pool_backend = "host"
    cfg.prealloc_bytes = 2 * 16384
    cfg.minimal_allocate_size = 16
    srv = m.Server(cfg)
    port = srv.start()
    try:
        writer, reader, other = make_conn(port), make_conn(port), make_conn(port)
        src = torch.randn(2 * 4096)
        writer.register_mr(src)
        writer.rdma_write_cache(src, [0, 4096], 4096, writer.allocate_rdma(["a", "b"], 16384))
        writer.sync()
        dst = torch.zeros(2 * 4096)
        reader.read_cache(dst, [("a", 0), ("b", 4096)], 4096)  # leases until reader.sync()
        assert srv.purge() == 2
        with pytest.raises(Exception):
            other.allocate_rdma(["c"], 16384)  # the space is still the reader's
        reader.sync()
        assert torch.equal(src, dst)
        assert len(other.allocate_rdma(["c"], 16384)) == 1
    finally:
        srv.stop()


def test_async_write_commits_only_its_own_blocks(host_server):
    """An async write's completion commits exactly the blocks it wrote - not blocks of other
    writes of the connection that are still waiting for their sync()."""
    _, port = host_server
    conn = make_conn(port)
    probe = make_conn(port)
    src = torch.randn(2 * 1024)
    conn.register_mr(src)
    blocks = conn.allocate_rdma(["sync-side", "async-side"], 4096)
    conn.rdma_write_cache(src, [0], 1024, blocks[:1])  # pending until conn.sync()

    async def run():
        await conn.rdma_write_cache_async(src, [1024], 1024, blocks[1:])

    asyncio.run(asyncio.wait_for(run(), 10))
    assert probe.check_exist("async-side")
    assert not probe.check_exist("sync-side")  # its commit still waits for sync()
    conn.sync()
    assert probe.check_exist("sync-side")


def test_read_miss_does_not_discard_commits_of_the_same_window(host_server):
    _, port = host_server
    conn = make_conn(port)
    src = torch.randn(1024)
    conn.register_mr(src)
    conn.rdma_write_cache(src, [0], 1024, conn.allocate_rdma(["w-ok"], 4096))
    dst = torch.zeros(1024)
    with pytest.raises(Exception):
        conn.read_cache(dst, [("absent-key", 0)], 1024)
        conn.sync()
    conn.sync()
    assert conn.check_exist("w-ok")


# ------------------------------------------------------------------ protocol hardening
def _raw(port):
    s = socket.create_connection(("127.0.0.1", port), timeout=5)
    return s


def _closed(s):
    """True when the peer closed the connection (orderly EOF or reset)."""
    try:
        return s.recv(1) == b""
    except (ConnectionResetError, BrokenPipeError):
        return True


def test_unknown_op_gets_400_and_server_survives(host_server):
    srv, port = host_server
    s = _raw(port)
    s.sendall(struct.pack("<IcI", 0xDEADBEEF, b"Z", 0))
    assert struct.unpack("<i", s.recv(4))[0] == 400
    assert _closed(s)
    # the reactor is not stuck (reference defect D11 spins forever here)
    conn = make_conn(port)
    assert not conn.check_exist("x")


def test_bad_magic_closes_connection(host_server):
    _, port = host_server
    s = _raw(port)
    s.sendall(struct.pack("<IcI", 0x12345678, b"C", 1) + b"x")
    assert _closed(s)


def test_oversized_body_is_refused(host_server):
    _, port = host_server
    s = _raw(port)
    s.sendall(struct.pack("<IcI", 0xDEADBEEF, b"M", 0x7FFFFFFF))
    assert struct.unpack("<i", s.recv(4))[0] == 400
    assert _closed(s)


def test_exchange_body_must_be_30_bytes(host_server):
    _, port = host_server
    s = _raw(port)
    s.sendall(struct.pack("<IcI", 0xDEADBEEF, b"E", 64) + b"\0" * 64)
    assert struct.unpack("<i", s.recv(4))[0] == 400


def test_garbage_flatbuffer_gets_400_and_connection_stays_usable(host_server):
    _, port = host_server
    s = _raw(port)
    junk = b"\xff" * 40
    s.sendall(struct.pack("<IcI", 0xDEADBEEF, b"D", len(junk)) + junk)
    assert struct.unpack("<i", s.recv(4))[0] == 400
    s.sendall(struct.pack("<IcI", 0xDEADBEEF, b"C", 3) + b"abc")
    code, val = struct.unpack("<ii", s.recv(8))
    assert (code, val) == (200, 1)


def test_request_split_across_tcp_segments(host_server):
    from infinistore_b200 import _infinistore as m

    _, port = host_server
    s = _raw(port)
    s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
    body = m.testing.encode_match_request(["a", "b", "c"])
    msg = struct.pack("<IcI", 0xDEADBEEF, b"M", len(body)) + body
    for i in range(len(msg)):  # one byte per segment
        s.sendall(msg[i:i + 1])
        if i % 7 == 0:
            time.sleep(0.001)
    code, idx = struct.unpack("<ii", s.recv(8))
    assert (code, idx) == (200, -1)
    # two requests in one segment + SYNC with an uninitialised body_size (as the
    # reference client sends it)
    s.sendall(msg + struct.pack("<IcI", 0xDEADBEEF, b"S", 0xCCCCCCCC))
    data = b""
    while len(data) < 16:
        data += s.recv(16 - len(data))
    assert struct.unpack("<iiiI", data) == (200, -1, 200, 0)


def _raw_request(s, op, body=b""):
    s.sendall(struct.pack("<IcI", 0xDEADBEEF, op, len(body)) + body)


def _recv_exact(s, n):
    data = b""
    while len(data) < n:
        chunk = s.recv(n - len(data))
        assert chunk, "connection closed"
        data += chunk
    return data


def test_staged_commit_becomes_visible_only_at_sync(host_server):
    """'U' ships the commit list ahead of the writer's drain; nothing is visible until the
    writer's next 'S'; block_size -1 discards; a dead writer's staged list dies with it."""
    from infinistore_b200 import _infinistore as m

    srv, port = host_server
    enc = m.testing.encode_remote_meta
    s = _raw(port)
    s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
    _raw_request(s, b"D", enc(["st-a", "st-b", "st-c"], 4096, 0, [], "D"))
    code, n = struct.unpack("<iI", _recv_exact(s, 8))
    assert code == 200
    blocks = m.testing.decode_allocate_response(_recv_exact(s, n))
    addrs = [int(a) for a in blocks["remote_addr"]]
    observer = make_conn(port)
    # staged: still invisible
    _raw_request(s, b"U", enc([], 0, 0, addrs[:2], "U"))
    _raw_request(s, b"C", b"st-a")  # a round trip on the same connection orders us after 'U'
    assert struct.unpack("<ii", _recv_exact(s, 8)) == (200, 1)
    assert not observer.check_exist("st-a") and srv.stats()["inflight"] == 3
    # discard (the writer's kernels failed): the staged blocks are released right away - their
    # data may be incomplete, so they must never become visible - and the keys are free again
    _raw_request(s, b"U", enc([], -1, 0, [], "U"))
    _raw_request(s, b"C", b"st-a")
    assert struct.unpack("<ii", _recv_exact(s, 8)) == (200, 1)
    assert srv.stats()["inflight"] == 1
    # stage the remaining block, then SYNC applies exactly that
    _raw_request(s, b"U", enc([], 0, 0, addrs[2:], "U"))
    _raw_request(s, b"S")
    assert struct.unpack("<iI", _recv_exact(s, 8)) == (200, 0)
    assert observer.check_exist("st-c") and not observer.check_exist("st-b")
    assert srv.stats()["inflight"] == 0
    # a writer that dies with a staged list commits nothing
    _raw_request(s, b"D", enc(["st-d"], 4096, 0, [], "D"))
    code, n = struct.unpack("<iI", _recv_exact(s, 8))
    assert code == 200
    late = [int(a) for a in m.testing.decode_allocate_response(_recv_exact(s, n))["remote_addr"]]
    _raw_request(s, b"U", enc([], 0, 0, late, "U"))
    _raw_request(s, b"C", b"st-d")
    assert struct.unpack("<ii", _recv_exact(s, 8)) == (200, 1)
    s.close()
    deadline = time.time() + 5
    while srv.stats()["inflight"] and time.time() < deadline:
        time.sleep(0.01)
    assert srv.stats()["inflight"] == 0 and srv.kvmap_len() == 1
    assert not observer.check_exist("st-b") and not observer.check_exist("st-d")
    # malformed 'U' gets no reply and does not desynchronise the stream
    s2 = _raw(port)
    _raw_request(s2, b"U", b"\xff" * 24)
    _raw_request(s2, b"C", b"st-c")
    assert struct.unpack("<ii", _recv_exact(s2, 8)) == (200, 0)


def test_fault_injection_dropped_request_surfaces_as_error(host_server):
    srv, port = host_server
    conn = make_conn(port, timeout_ms=2000)
    srv.inject_drop_after(1)
    with pytest.raises(Exception):
        conn.check_exist("k")
    conn2 = make_conn(port)
    assert not conn2.check_exist("k")


def test_reply_timeout_drops_the_connection_instead_of_desynchronising_it(host_server):
    """A reply that arrives after the client's deadline must never be taken for the answer to
    the next request: the client closes the connection on a timeout and fails fast after."""
    srv, port = host_server
    writer = make_conn(port)
    src = torch.randn(1024)
    writer.register_mr(src)
    writer.rdma_write_cache(src, [0], 1024, writer.allocate_rdma(["slow-key"], 4096))
    writer.sync()
    conn = make_conn(port, timeout_ms=200)
    assert conn.check_exist("slow-key") is True
    srv.inject_delay(700, 1)  # the next request is answered long after the 200 ms deadline
    t0 = time.time()
    with pytest.raises(Exception):
        conn.check_exist("slow-key")
    assert time.time() - t0 < 0.65  # the client gave up at its own deadline
    time.sleep(0.8)  # the late reply (exists = 0) is on the wire now
    # without the fix this call would read that stale reply and report "missing-key exists"
    with pytest.raises(Exception):
        conn.check_exist("missing-key")
    conn2 = make_conn(port, timeout_ms=2000)
    assert conn2.check_exist("slow-key") and not conn2.check_exist("missing-key")


def test_client_that_never_reads_replies_is_disconnected():
    srv, port = _server(prealloc_bytes=64 << 20, max_pending_reply_bytes=256 << 10)
    try:
        s = _raw(port)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 4096)
        msg = struct.pack("<IcII", 0xDEADBEEF, b"P", 4, 0)  # pool map: a ~170-byte reply each
        sent = 0
        s.settimeout(5)
        try:
            for _ in range(20000):  # unread replies pile up: the server gives up on us
                s.sendall(msg * 64)
                sent += 64
        except (BrokenPipeError, ConnectionResetError, socket.timeout):
            pass
        assert sent < 20000 * 64  # we were cut off
        assert srv.running()
        conn = make_conn(port)
        assert conn.check_exist("anything") is False
    finally:
        srv.stop()


def test_many_keys_single_request(host_server):
    _, port = host_server
    conn = make_conn(port)
    n = 2000
    keys = [f"{i:08d}-" + rand_key(27) for i in range(n)]
    src = torch.randn(n * 64)
    dst = torch.zeros_like(src)
    conn.register_mr(src)
    blocks = conn.allocate_rdma(keys, 256)
    assert len(set(blocks["remote_addr"].tolist())) == n
    conn.rdma_write_cache(src, [i * 64 for i in range(n)], 64, blocks)
    conn.sync()
    conn.read_cache(dst, [(k, i * 64) for i, k in enumerate(keys)], 64)
    conn.sync()
    assert torch.equal(src, dst)
    assert conn.get_match_last_index(keys) == n - 1


def test_checkpoint_dump_and_load(tmp_path):
    """Checkpoint / resume: dump a store, start a fresh one, load, read back."""
    srv, port = _server(prealloc_bytes=64 * 16384)
    path = str(tmp_path / "store.ckpt")
    try:
        conn = make_conn(port)
        src = torch.randn(12 * 1024)
        conn.register_mr(src)
        keys = [f"ck-{i}" for i in range(12)]
        conn.rdma_write_cache(src, [i * 1024 for i in range(12)], 1024, conn.allocate_rdma(keys, 4096))
        conn.allocate_rdma(["reserved-only"], 4096)  # uncommitted: must not be dumped
        conn.sync()
        assert srv.dump(path) == 12
    finally:
        srv.stop()
    srv2, port2 = _server(prealloc_bytes=64 * 16384)
    try:
        conn = make_conn(port2)
        live = torch.full((1024,), 7.0)
        conn.register_mr(live)
        conn.rdma_write_cache(live, [0], 1024, conn.allocate_rdma(["ck-3"], 4096))  # newer copy wins
        conn.sync()
        assert srv2.load(path) == 11 and srv2.kvmap_len() == 12
        dst = torch.zeros(12 * 1024)
        conn.read_cache(dst, [(k, i * 1024) for i, k in enumerate(keys)], 1024)
        conn.sync()
        want = src.clone()
        want[3 * 1024:4 * 1024] = 7.0
        assert torch.equal(dst, want)
        assert not conn.check_exist("reserved-only")
        with pytest.raises(Exception):
            srv2.load(str(tmp_path / "missing.ckpt"))
    finally:
        srv2.stop()


def test_huge_key_lists_are_chunked_transparently():
    """150k keys x ~40 bytes exceed the 4 MiB message cap: allocate / lookup chunk them."""
    srv, port = _server(prealloc_bytes=160000 * 16384)
    try:
        conn = make_conn(port)
        n = 150000
        keys = [f"{i:07d}-" + "k" * 32 for i in range(n)]
        blocks = conn.allocate_rdma(keys, 64)
        assert len(blocks) == n and len(set(blocks["remote_addr"].tolist())) == n
        assert srv.stats()["ops"]["ALLOCATE"] >= 2
        src = torch.arange(n * 16, dtype=torch.float32)
        conn.register_mr(src)
        conn.rdma_write_cache(src, np.arange(n) * 16, 16, blocks)
        conn.sync()
        dst = torch.zeros_like(src)
        conn.read_cache(dst, list(zip(keys, (np.arange(n) * 16).tolist())), 16)
        conn.sync()
        assert torch.equal(src, dst)
    finally:
        srv.stop()


@pytest.mark.parametrize("seed", [1, 21])
def test_randomised_soak_against_a_model(seed):
    """tools/soak_cpu.py for a bounded number of operations (seed 21 forces auto-increase)."""
    import subprocess
    import sys as _sys
    from conftest import ROOT

    env = dict(os.environ, **({"SOAK_AUTO": "1"} if seed == 21 else {}))
    r = subprocess.run([_sys.executable, os.path.join(ROOT, "tools", "soak_cpu.py"), str(seed),
                        "120", "2000"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and " OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_posted_commit_skips_the_round_trip_and_still_commits(host_server):
    """posted_commit=True: sync() sends the commit list one-way (reference semantics); the
    same connection sees its writes at once (TCP order), other connections shortly after."""
    srv, port = host_server
    conn = make_conn(port, posted_commit=True)
    src = torch.randn(4 * 1024)
    conn.register_mr(src)
    keys = [f"posted-{i}" for i in range(4)]
    before = conn.stats()["ctrl_requests"]
    conn.rdma_write_cache(src, [i * 1024 for i in range(4)], 1024, conn.allocate_rdma(keys, 4096))
    conn.sync()
    assert conn.stats()["ctrl_requests"] - before == 2  # ALLOCATE + one-way COMMIT, no SYNC
    assert conn.check_exist(keys[0])  # ordered behind the commit on the same connection
    other = make_conn(port)
    deadline = time.time() + 5
    while not other.check_exist(keys[3]) and time.time() < deadline:
        time.sleep(0.001)
    assert other.check_exist(keys[3])
    dst = torch.zeros(4 * 1024)
    conn.read_cache(dst, [(k, i * 1024) for i, k in enumerate(keys)], 1024)  # takes leases
    conn.sync()  # leases held: this sync does the round trip that releases them
    assert torch.equal(src, dst)
    assert srv.stats()["inflight"] == 0


def test_doorbell_mode_is_inert_without_a_gpu_pool(host_server):
    """doorbell=True only changes how single blocks reach an HBM pool from a CUDA tensor; with
    CPU tensors and the host pool every call takes the ordinary path."""
    _, port = host_server
    conn = make_conn(port, doorbell=True, doorbell_idle_us=50)
    assert conn.config.doorbell and conn.config.doorbell_idle_us == 50
    src = torch.randn(1024)
    dst = torch.zeros(1024)
    conn.register_mr(src)
    conn.rdma_write_cache(src, [0], 1024, conn.allocate_rdma(["db-cpu"], 4096))
    conn.sync()
    conn.read_cache(dst, [("db-cpu", 0)], 1024)
    conn.sync()
    assert torch.equal(src, dst)
    assert conn.stats()["doorbell_ops"] == 0 and conn.stats()["doorbell_launches"] == 0


def test_one_connection_shared_by_four_threads(host_server):
    """The native connection serialises its control plane (one transaction at a time), its
    data-plane bookkeeping and sync(); four threads that share ONE InfinityConnection - the
    GIL is released inside every native call - allocate, write, sync and read their own keys
    concurrently.  A sync() of one thread may commit the finished writes of another; it must
    never commit an unfinished one or lose one."""
    import threading

    srv, port = host_server
    conn = make_conn(port)
    errors = []

    def worker(tid):
        try:
            rng = np.random.default_rng(100 + tid)
            src = torch.zeros(8 * 1024)
            dst = torch.zeros(8 * 1024)
            conn.register_mr(src)
            conn.register_mr(dst)
            for it in range(40):
                n = int(rng.integers(1, 9))
                keys = [f"shared-t{tid}-i{it}-b{b}" for b in range(n)]
                for b in range(n):
                    src[b * 1024:(b + 1) * 1024] = float(tid * 1000000 + it * 10 + b)
                blocks = conn.allocate_rdma(keys, 4096)
                conn.rdma_write_cache(src, [b * 1024 for b in range(n)], 1024, blocks)
                conn.sync()
                assert conn.get_match_last_index(keys) == n - 1
                dst.zero_()
                conn.read_cache(dst, [(k, b * 1024) for b, k in enumerate(keys)], 1024)
                conn.sync()
                if not torch.equal(dst[:n * 1024], src[:n * 1024]):
                    errors.append(("mismatch", tid, it))
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors[:3]
    st = srv.stats()
    assert st["inflight"] == 0 and st["keys"] == srv.kvmap_len()
    conn.close()


def test_blocks_as_keys_plus_offset_array(host_server):
    """read_cache / local calls also take (keys, offsets) with an integer ndarray of offsets
    instead of a list of (key, offset) pairs."""
    _, port = host_server
    conn = make_conn(port)
    n, page = 12, 1024
    src = torch.randn(n * page)
    dst = torch.zeros_like(src)
    conn.register_mr(src)
    conn.register_mr(dst)
    keys = [f"split-{i}" for i in range(n)]
    offs = np.arange(n, dtype=np.int64) * page
    conn.rdma_write_cache(src, offs, page, conn.allocate_rdma(keys, page * 4))
    conn.sync()
    conn.read_cache(dst, (keys, offs[::-1].copy()), page)      # reversed placement
    conn.sync()
    assert torch.equal(dst.view(n, page), src.view(n, page).flip(0))
    dst.zero_()
    conn.read_cache(dst, (keys, offs.astype(np.int32)), page)  # any integer dtype
    conn.sync()
    assert torch.equal(dst, src)
    with pytest.raises(Exception):
        conn.read_cache(dst, (keys, offs[:-1]), page)           # lengths differ
    with pytest.raises(Exception):
        conn.read_cache(dst, (keys, -offs - 1), page)           # negative offset
    conn.read_cache(dst, [(keys[0], 0), (keys[1], page)], page)  # the pair form is untouched
    conn.sync()
