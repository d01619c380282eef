"""Sharded store, page layouts and the paged KV cache on CPU loop-back."""
import random

import pytest
import torch

import infinistore_b200 as ist
from infinistore_b200 import _infinistore as native
from infinistore_b200.models import PagedKVCache, KVLayout, chain_hashes, get_layout, page_key
from infinistore_b200.parallel import ShardedConnection, shard_of, start_shard_server
from conftest import make_conn


@pytest.fixture()
def three_shards():
    servers = []
    for _ in range(3):
        cfg = native.ServerConfig()
        cfg.service_port = 0
        cfg.host = "127.0.0.1"
        cfg.pool_backend = "host"
        cfg.prealloc_bytes = 32 << 20
        cfg.minimal_allocate_size = 16
        s = native.Server(cfg)
        s.start()
        servers.append(s)
    yield servers
    for s in servers:
        s.stop()


def test_sharded_connection_routes_and_round_trips(three_shards):
    cfgs = [ist.ClientConfig(host_addr="127.0.0.1", service_port=s.port(),
                             connection_type=ist.TYPE_RDMA) for s in three_shards]
    conn = ShardedConnection(cfgs)
    conn.connect()
    n, page = 90, 1024
    keys = [f"blk-{i:04d}" for i in range(n)]
    src = torch.randn(n * page)
    dst = torch.zeros_like(src)
    conn.register_mr(src)
    conn.register_mr(dst)
    blocks = conn.allocate_rdma(keys, page * 4)
    assert len(blocks) == n
    conn.rdma_write_cache(src, [i * page for i in range(n)], page, blocks)
    conn.sync()
    per_shard = [s.kvmap_len() for s in three_shards]
    assert sum(per_shard) == n and all(c > 0 for c in per_shard)
    expect = [0, 0, 0]
    for k in keys:
        expect[shard_of(k, 3)] += 1
    assert per_shard == expect
    order = list(range(n))
    random.Random(3).shuffle(order)
    conn.read_cache(dst, [(keys[i], i * page) for i in order], page)
    conn.sync()
    assert torch.equal(src, dst)
    assert conn.check_exist(keys[5]) and not conn.check_exist("nope")
    assert conn.get_match_last_index(keys[:10] + ["x", "y"]) == 9
    with pytest.raises(Exception):
        conn.get_match_last_index(["x", "y"])
    assert conn.touch(keys[:7]) == 0  # routed to every shard; these shards do not evict
    conn.close()


def test_layouts():
    l8 = get_layout("llama-3-8b")
    assert l8.page_bytes == 128 * 8 * 128 * 2 == 256 * 1024  # BASELINE config 3 page
    assert l8.with_tp(8).page_bytes == 32 * 1024
    assert get_layout("llama-3-70b", page_tokens=16).page_elems == 16 * 8 * 128
    assert get_layout("qwen2.5-7b", dtype=torch.float16).page_bytes == 128 * 4 * 128 * 2
    assert l8.token_bytes_all_layers == 2 * 32 * 8 * 128 * 2


def test_chain_hashes_are_prefix_monotone():
    toks = list(range(1000))
    a = chain_hashes(toks, 128)
    b = chain_hashes(toks[:512] + [7] * 488, 128)
    assert len(a) == 7 and a[:4] == b[:4] and a[4] != b[4]
    assert chain_hashes(toks, 128, salt="other-model")[0] != a[0]
    assert page_key("m", 3, "K", 1, "abc") == "m/L3/K/tp1/abc"


def test_paged_kv_cache_layerwise_round_trip(host_server):
    _, port = host_server
    layout = KVLayout("tiny", layers=3, kv_heads=2, head_dim=16, page_tokens=8,
                      dtype=torch.float32)
    prefill = PagedKVCache(layout, num_pages=6, device="cpu")
    decode = PagedKVCache(layout, num_pages=6, device="cpu")
    prefill.data.normal_()
    conn_p = make_conn(port)
    conn_d = make_conn(port)
    hashes = chain_hashes(list(range(8 * 4)), 8)  # 4 full pages
    pages = [5, 0, 3, 1]                           # scattered physical pages
    assert decode.cached_prefix_pages(conn_d, hashes) == 0
    for layer in range(layout.layers):             # "layer by layer" upload
        assert prefill.write_layer(conn_p, layer, pages, hashes) == 8
    conn_p.sync()
    assert decode.cached_prefix_pages(conn_d, hashes + ["f" * 32]) == 4
    dst_pages = [2, 4, 1, 0]
    for layer in range(layout.layers):
        decode.read_layer(conn_d, layer, dst_pages, hashes)
    conn_d.sync()
    for layer in range(layout.layers):
        for kv in (0, 1):
            for s, d in zip(pages, dst_pages):
                assert torch.equal(prefill.page(layer, kv, s), decode.page(layer, kv, d))
    # the same prefix uploaded again is deduplicated, not overwritten
    prefill.data.zero_()
    prefill.write_layer(conn_p, 0, pages, hashes)
    conn_p.sync()
    decode.data.zero_()
    decode.read_layer(conn_d, 0, dst_pages, hashes)
    conn_d.sync()
    assert float(decode.data[0].abs().sum()) > 0


def test_start_shard_server_helper():
    srv = start_shard_server(0, 0, 8 << 20, granule_kb=16)
    try:
        conn = make_conn(srv.port())
        assert len(conn.allocate_rdma(["a"], 100)) == 1
    finally:
        srv.stop()


def test_paged_kv_touch_prefix_keeps_a_hot_prefix_from_eviction():
    layout = KVLayout("tiny", layers=2, kv_heads=1, head_dim=16, page_tokens=8,
                      dtype=torch.float32)
    # room for exactly two prefixes of 2 pages (2 layers x K,V x 2 pages = 8 blocks each)
    srv = start_shard_server(0, 0, 16 * 16384, granule_kb=16, evict=True, evict_ratio=0.5)
    try:
        conn = make_conn(srv.port())
        cache = PagedKVCache(layout, num_pages=4, device="cpu")
        cache.data.normal_()
        hot = chain_hashes(list(range(16)), 8, salt="hot")
        cold = chain_hashes(list(range(16)), 8, salt="cold")
        new = chain_hashes(list(range(16)), 8, salt="new")
        for hashes in (hot, cold):
            for layer in range(layout.layers):
                cache.write_layer(conn, layer, [0, 1], hashes)
        conn.sync()
        assert srv.stats()["used_bytes"] == 16 * 16384
        assert cache.touch_prefix(conn, hot) == 8  # the older prefix is the hot one
        for layer in range(layout.layers):
            cache.write_layer(conn, layer, [2, 3], new)
        conn.sync()
        assert cache.cached_prefix_pages(conn, hot) == 2
        assert cache.cached_prefix_pages(conn, cold) == 0
        assert cache.cached_prefix_pages(conn, new) == 2
    finally:
        srv.stop()


def test_examples_run_on_cpu(host_server, monkeypatch, capsys):
    import sys
    from infinistore_b200.example import client, client_async

    _, port = host_server
    monkeypatch.setattr(sys, "argv", ["client", "--service-port", str(port)])
    client.main()
    monkeypatch.setattr(sys, "argv", ["client_async", "--service-port", str(port), "--iterations", "2"])
    client_async.main()
    out = capsys.readouterr().out
    assert "fabric cpu->cpu" in out and "iteration 1: ok" in out


def test_head_major_cache_geometry_and_shared_key_scheme():
    """HeadMajorKVCache (decode side) and PagedKVCache (prefill side) must agree on keys - the
    decode instance fetches what prefill uploaded - and the head-major page is the permuted
    token-major page."""
    from infinistore_b200.models import HeadMajorKVCache, layer_keys, read_layer_multi

    layout = KVLayout("geo", layers=3, kv_heads=4, head_dim=8, page_tokens=16, tp=2)
    a = PagedKVCache(layout, num_pages=5, device="cpu", tp_rank=1)
    h = HeadMajorKVCache(layout, num_pages=5, device="cpu", tp_rank=1)
    assert layout.heads_per_rank == 2 and layout.page_elems == 16 * 2 * 8
    assert h.data.shape == (3, 2, 5, 2, 16, 8) and a.data.shape == (3, 2, 5, 16 * 2 * 8)
    hashes = chain_hashes(list(range(16 * 3)), 16, salt="geo")
    for layer in range(3):
        for kv in (0, 1):
            assert a.keys(layer, kv, hashes) == h.keys(layer, kv, hashes) == \
                layer_keys(layout, 1, layer, kv, hashes)
    assert a.keys(0, 0, hashes)[0] != a.keys(0, 1, hashes)[0] != a.keys(1, 0, hashes)[0]
    assert "/tp1/" in a.keys(2, 1, hashes)[0] and "/L2/V/" in a.keys(2, 1, hashes)[0]
    # page_token_major is a VIEW of the head-major page
    h.data[1, 0, 3].copy_(torch.arange(2 * 16 * 8, dtype=layout.dtype).view(2, 16, 8))
    tm = h.page_token_major(1, 0, 3)
    assert tm.shape == (16, 2, 8) and tm[5, 1, 7] == h.data[1, 0, 3, 1, 5, 7]
    # read_layer_multi insists on caches that really are replicas of one another
    b = PagedKVCache(layout, num_pages=6, device="cpu", tp_rank=1)
    with pytest.raises(ValueError):
        read_layer_multi(None, [a, b], 0, [0], hashes[:1])
    c = PagedKVCache(layout, num_pages=5, device="cpu", tp_rank=0)
    with pytest.raises(ValueError):
        read_layer_multi(None, [a, c], 0, [0], hashes[:1])


def test_timing_utilities_without_a_gpu():
    """percentile() and the clock sampler bench.py uses (NVML, falling back to nvidia-smi);
    on a box with neither the sampler reports nothing rather than failing the run."""
    import time

    from infinistore_b200.utils import ClockSampler, percentile

    vals = sorted([5.0, 1.0, 3.0, 2.0, 4.0])
    assert percentile(vals, 0) == 1.0 and percentile(vals, 50) == 3.0 and percentile(vals, 100) == 5.0
    assert percentile(vals, 99) == 5.0 and percentile([], 50) != percentile([], 50)   # nan
    s = ClockSampler(0)
    s.start()
    time.sleep(0.15)
    out = s.stop()
    assert set(out) == {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}
    assert out["source"] in ("nvml", "nvidia-smi") and isinstance(out["reasons"], list)
    assert (out["samples"] == 0) == (out["sm_mhz"] is None)
