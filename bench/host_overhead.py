#!/usr/bin/env python
"""Where does the host time of one write/read call go?  Times the public API calls with
the GPU idle-waited in between, so the numbers are pure issue cost (Python + binding +
descriptor build + launch), per call and per block."""
import json
import os
import sys
import time
import uuid

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infinistore_b200 as ist  # noqa: E402
from infinistore_b200 import _infinistore as native  # noqa: E402


def main():
    torch.cuda.set_device(0)
    cfg = native.ServerConfig()
    cfg.service_port = 0
    cfg.host = "127.0.0.1"
    cfg.pool_backend = "hbm"
    cfg.pool_devices = [0]
    cfg.prealloc_bytes = 8 << 30
    cfg.minimal_allocate_size = 64
    srv = native.Server(cfg)
    port = srv.start()
    out = {}
    for lookup in (True, False):
        conn = ist.InfinityConnection(ist.ClientConfig(
            host_addr="127.0.0.1", service_port=port, connection_type=ist.TYPE_RDMA,
            device_lookup=lookup))
        conn.connect()
        elems = 64 * 1024
        for per_call in (16, 256, 4096):
            n = per_call * 8
            src = torch.randn(n * elems // 2, device="cuda:0").to(torch.bfloat16)
            src = torch.cat([src, src])
            dst = torch.zeros_like(src)
            conn.register_mr(src)
            conn.register_mr(dst)
            keys = [str(uuid.uuid4()) for _ in range(n)]
            offs = [i * elems for i in range(n)]
            blocks = list(zip(keys, offs))
            t0 = time.perf_counter()
            remote = conn.allocate_rdma(keys, elems * 2)
            t_alloc = time.perf_counter() - t0
            torch.cuda.synchronize()
            tw = []
            for c in range(8):
                a, b = c * per_call, (c + 1) * per_call
                t0 = time.perf_counter()
                conn.rdma_write_cache(src, offs[a:b], elems, remote[a:b])
                tw.append(time.perf_counter() - t0)
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            conn.sync()
            t_sync_w = time.perf_counter() - t0
            tr = []
            for c in range(8):
                a, b = c * per_call, (c + 1) * per_call
                t0 = time.perf_counter()
                conn.read_cache(dst, blocks[a:b], elems)
                tr.append(time.perf_counter() - t0)
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            conn.sync()
            t_sync_r = time.perf_counter() - t0
            assert torch.equal(src, dst)
            st = conn.stats()
            med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
            row = {"alloc_us_per_key": t_alloc / n * 1e6, "write_call_us": med(tw) * 1e6,
                   "write_ns_per_block": med(tw) / per_call * 1e9, "read_call_us": med(tr) * 1e6,
                   "read_ns_per_block": med(tr) / per_call * 1e9, "sync_after_write_us": t_sync_w * 1e6,
                   "sync_after_read_us": t_sync_r * 1e6,
                   "cxx_build_us_per_call": st["ns_build"] / max(st["calls"], 1) / 1e3,
                   "cxx_streams_us_per_call": st["ns_streams"] / max(st["calls"], 1) / 1e3,
                   "cxx_launch_us_per_call": st["ns_launch"] / max(st["calls"], 1) / 1e3}
            out[f"lookup={'device' if lookup else 'host'} blocks/call={per_call}"] = {
                k: round(v, 2) for k, v in row.items()}
            label = f"lookup={'device' if lookup else 'host'} blocks/call={per_call}"
            print(lookup, per_call, out[label], flush=True)
            srv.purge()
        conn.close()
    srv.stop()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/host_overhead.json", "w"), indent=1)


if __name__ == "__main__":
    main()
