#!/usr/bin/env python
"""BASELINE config 2: write+read sweep over block sizes 4 KB - 16 MB THROUGH THE STORE API
(allocate outside the timing, 32 calls per phase, sync after each phase) plus single-block
latency p50/p99 (one write + sync, one read + sync), for a pool on the same GPU (--pool 0)
or on a peer GPU over NVLink (--pool 1)."""
import argparse
import json
import os
import sys
import time
import uuid

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infinistore_b200 as ist  # noqa: E402
from infinistore_b200.parallel import start_shard_server  # noqa: E402


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(q / 100 * (len(v) - 1))))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pool", type=int, default=0, help="GPU hosting the pool (client is GPU 0)")
    ap.add_argument("--total-mb", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--lat-samples", type=int, default=300)
    ap.add_argument("--host-lookup", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    total = a.total_mb << 20
    srv = start_shard_server(a.pool, 0, (a.iters + 2) * total + (1 << 30), granule_kb=16)
    conn = ist.InfinityConnection(ist.ClientConfig(
        host_addr="127.0.0.1", service_port=srv.port(), connection_type=ist.TYPE_RDMA, device=0,
        device_lookup=not a.host_lookup))
    conn.connect()
    src = torch.randn(total // 2, device="cuda:0").to(torch.bfloat16)
    src = torch.cat([src, src])[: total // 2].contiguous()
    dst = torch.zeros_like(src)
    conn.register_mr(src)
    conn.register_mr(dst)
    rows = []
    for kb in (4, 16, 64, 128, 256, 1024, 4096, 16384):
        bs = kb << 10
        elems = bs // 2
        n = total // bs
        layers = min(32, n)
        per = n // layers
        offs = np.arange(n, dtype=np.int64) * elems
        tw = tr = 0.0
        for it in range(a.iters + 1):
            keys = [uuid.uuid4().hex for _ in range(n)]
            remote = conn.allocate_rdma(keys, bs)
            blocks = list(zip(keys, offs.tolist()))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for l in range(layers):
                s = slice(l * per, (l + 1) * per)
                conn.rdma_write_cache(src, offs[s], elems, remote[s])
            conn.sync()
            t1 = time.perf_counter()
            for l in range(layers):
                s = slice(l * per, (l + 1) * per)
                conn.read_cache(dst, blocks[s], elems)
            conn.sync()
            t2 = time.perf_counter()
            if it:
                tw += t1 - t0
                tr += t2 - t1
            srv.purge()
        assert torch.equal(src, dst), kb
        # single-block latency
        lw, lr = [], []
        keys = [uuid.uuid4().hex for _ in range(a.lat_samples)]
        remote = conn.allocate_rdma(keys, bs)
        one = np.zeros(1, dtype=np.int64)
        for i in range(a.lat_samples):
            t0 = time.perf_counter()
            conn.rdma_write_cache(src, one, elems, remote[i:i + 1])
            conn.sync()
            t1 = time.perf_counter()
            conn.read_cache(dst, [(keys[i], 0)], elems)
            conn.sync()
            t2 = time.perf_counter()
            if i >= 20:
                lw.append((t1 - t0) * 1e6)
                lr.append((t2 - t1) * 1e6)
        srv.purge()
        row = {"block_kb": kb, "blocks_per_call": per,
               "write_GBps": round(total * a.iters / tw / 1e9, 1),
               "read_GBps": round(total * a.iters / tr / 1e9, 1),
               "write_sync_us_p50": round(pct(lw, 50), 1), "write_sync_us_p99": round(pct(lw, 99), 1),
               "read_sync_us_p50": round(pct(lr, 50), 1), "read_sync_us_p99": round(pct(lr, 99), 1)}
        rows.append(row)
        print(row, flush=True)
    conn.close()
    srv.stop()
    os.makedirs("gpurun_out", exist_ok=True)
    name = f"gpurun_out/api_sweep_pool{a.pool}{'_hostlookup' if a.host_lookup else ''}.json"
    json.dump({"pool_gpu": a.pool, "lookup": "host" if a.host_lookup else "device", "rows": rows},
              open(name, "w"), indent=1)


if __name__ == "__main__":
    main()
