#!/usr/bin/env python
"""Steady-state churn: a pool that is always full, every new page evicts an old one.

The reference has no eviction (a full pool answers 507 until /purge), so this is the
operating point it cannot reach at all.  Measures the write rate through the public API
(allocate + write per call, sync every 32 calls; eviction work happens inside allocate) with
eviction on, against the same loop into a pool large enough never to fill; then reads the
most recent pages (hits, with post-copy validation) and the oldest ones (all evicted)."""
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


def run(pool_gb, total_gb, evict, block_kb, per_call, device_lookup=True):
    bs = block_kb << 10
    elems = bs // 2
    srv = start_shard_server(0, 0, int(pool_gb * (1 << 30)), granule_kb=64, evict=evict,
                             evict_ratio=0.05)
    conn = ist.InfinityConnection(ist.ClientConfig(
        host_addr="127.0.0.1", service_port=srv.port(), connection_type=ist.TYPE_RDMA, device=0,
        device_lookup=device_lookup))
    conn.connect()
    src = torch.randn(per_call * 32 * elems // 2, device="cuda:0").to(torch.bfloat16)
    src = torch.cat([src, src]).contiguous()
    conn.register_mr(src)
    calls = int(total_gb * (1 << 30)) // (per_call * bs)
    offs = [np.arange(per_call, dtype=np.int64) * elems + (c % 32) * per_call * elems
            for c in range(32)]
    run_id = uuid.uuid4().hex[:12]  # key generation stays outside the timed loop
    all_keys = [[f"{run_id}/{c:06d}/{i:04d}" for i in range(per_call)] for c in range(calls)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_alloc = 0.0
    for c in range(calls):
        keys = all_keys[c]
        ta = time.perf_counter()
        remote = conn.allocate_rdma(keys, bs)
        t_alloc += time.perf_counter() - ta
        conn.rdma_write_cache(src, offs[c % 32], elems, remote)
        if c % 32 == 31:
            conn.sync()
    conn.sync()
    dt = time.perf_counter() - t0
    st = srv.stats()
    out = {"evict": evict, "pool_gb": pool_gb, "written_gb": calls * per_call * bs / 2**30,
           "block_kb": block_kb, "write_GBps": round(calls * per_call * bs / dt / 1e9, 1),
           "allocate_share": round(t_alloc / dt, 3), "evicted_blocks": st["evicted"],
           "keys_at_end": st["keys"], "used_gb": round(st["used_bytes"] / 2**30, 2)}
    if evict:
        # newest pages: all present, read back bit-exact through the validated device path
        dst = torch.zeros(per_call * elems, dtype=torch.bfloat16, device="cuda:0")
        conn.register_mr(dst)
        hits = 0
        recent = range(calls - 16, calls)
        queries = [list(zip(all_keys[c], (np.arange(per_call) * elems).tolist())) for c in recent]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for q in queries:
            conn.read_cache(dst, q, elems)
            hits += per_call
        conn.sync()
        out["read_recent_GBps"] = round(hits * bs / (time.perf_counter() - t0) / 1e9, 1)
        c = calls - 1
        lo = (c % 32) * per_call * elems
        assert torch.equal(dst, src[lo:lo + per_call * elems])
        # oldest pages: evicted, for the server map and the device index alike
        missing = sum(0 if conn.check_exist(k) else 1 for k in all_keys[0][:64])
        out["oldest_64_missing"] = missing
    conn.close()
    srv.stop()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pool-gb", type=float, default=4)
    ap.add_argument("--total-gb", type=float, default=24)
    ap.add_argument("--block-kb", type=int, default=128)
    ap.add_argument("--per-call", type=int, default=256)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    res = {"churn": run(a.pool_gb, a.total_gb, True, a.block_kb, a.per_call),
           "no_eviction_big_pool": run(a.total_gb + 2, a.total_gb, False, a.block_kb, a.per_call)}
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/evict_churn.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
