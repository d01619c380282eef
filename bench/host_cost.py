"""Where does the host time of a read_cache / rdma_write_cache call go?  One GPU, local pool,
the flagship's shape (1024 pages of 128 KB per call, 32 calls per phase).

    python bench/host_cost.py            # prints one JSON line, writes gpurun_out/host_cost.json
"""
import json
import os
import sys
import time
import uuid

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infinistore_b200 as ist  # noqa: E402
from infinistore_b200 import _infinistore as native  # noqa: E402
from infinistore_b200.parallel import start_shard_server  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    layers, per, elems = 32, 1024, 65536
    n = layers * per
    srv = start_shard_server(0, 0, 3 * n * elems * 2 + (256 << 20), granule_kb=128)
    conn = ist.InfinityConnection(ist.ClientConfig(
        host_addr="127.0.0.1", service_port=srv.port(), connection_type=ist.TYPE_RDMA,
        device=0, device_lookup=True, log_level="warning"))
    conn.connect()
    src = torch.randn(n * elems, device=dev).to(torch.bfloat16)
    dst = torch.zeros_like(src)
    conn.register_mr(src)
    conn.register_mr(dst)
    offs = np.arange(n, dtype=np.int64) * elems
    out = {}
    for rep in range(3):
        tag = uuid.uuid4().hex[:12]
        keys = [f"{tag}-{i}" for i in range(n)]
        remote = conn.allocate_rdma(keys, elems * 2)
        blocks = list(zip(keys, offs.tolist()))
        torch.cuda.synchronize()
        s0 = conn.stats()
        t0 = time.perf_counter()
        for l in range(layers):
            conn.rdma_write_cache(src, offs[l * per:(l + 1) * per], elems, remote[l * per:(l + 1) * per])
        t1 = time.perf_counter()
        conn.sync()
        t2 = time.perf_counter()
        s1 = conn.stats()
        # pieces of a read call, in isolation
        ta = time.perf_counter()
        for l in range(layers):
            _ = blocks[l * per:(l + 1) * per]
        tb = time.perf_counter()
        for l in range(layers):
            native.testing.parse_blocks(blocks[l * per:(l + 1) * per])
        tc = time.perf_counter()
        for l in range(layers):
            conn.read_cache(dst, blocks[l * per:(l + 1) * per], elems)
        td = time.perf_counter()
        conn.sync()
        te = time.perf_counter()
        s2 = conn.stats()
        us = lambda a, b: round((b - a) / layers * 1e6, 1)  # noqa: E731
        ns = lambda k, a, b: round((b[k] - a[k]) / layers / 1e3, 1)  # noqa: E731
        out = {"per_call_us": {
            "write_issue": us(t0, t1), "write_native_build": ns("ns_build", s0, s1),
            "write_native_streams": ns("ns_streams", s0, s1), "write_native_launch": ns("ns_launch", s0, s1),
            "python_slice_1024": us(ta, tb), "slice_plus_parse_1024": us(tb, tc),
            "read_issue": us(tc, td), "read_native_build": ns("ns_build", s1, s2),
            "read_native_streams": ns("ns_streams", s1, s2), "read_native_launch": ns("ns_launch", s1, s2)},
            "phase_ms": {"write_sync_wait": round((t2 - t1) * 1e3, 3),
                         "read_sync_wait": round((te - td) * 1e3, 3)}}
        assert torch.equal(src, dst)
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/host_cost.json", "w") as f:
        json.dump(out, f, indent=1)
    conn.close()
    srv.stop()


if __name__ == "__main__":
    main()
