#!/usr/bin/env python
"""Comparison baselines: the reference's data-movement patterns re-created with library
calls (the reference itself cannot be built on this image).

  localgpu  : the reference's LOCAL_GPU path (src/infinistore.cpp:570-804): per block one
              cudaMemcpyAsync GPU -> pinned host pool (write) and back (read), a fresh
              stream + event per request.  Ceiling: PCIe Gen5 x16.
  peercopy  : "a naive NVLink port": per block one cudaMemcpyAsync GPU0 -> GPU1.
  nccl      : the "only calls NCCL" baseline: send/recv of each request's pages as one
              contiguous tensor between two ranks (run with torchrun, 2 ranks).
Workload = bench.py's: --size-mb of pages in --block-kb blocks, --layers requests per phase.
Prints one JSON line per baseline; wall-clock around write phase + read phase like the
reference's benchmark."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infinistore_b200 import _infinistore as native  # noqa: E402


def run_memcpy(label, src, pool, dst, nblocks, bs, layers, iters, device):
    per = nblocks // layers
    sp = [src.data_ptr() + i * bs for i in range(nblocks)]
    pp = [pool.data_ptr() + i * bs for i in range(nblocks)]
    dp = [dst.data_ptr() + i * bs for i in range(nblocks)]
    tw = tr = 0.0
    for it in range(iters + 1):
        t0 = time.perf_counter()
        for l in range(layers):
            a, b = l * per, (l + 1) * per
            native.baseline.memcpy_blocks(pp[a:b], sp[a:b], bs, True, device)
        t1 = time.perf_counter()
        for l in range(layers):
            a, b = l * per, (l + 1) * per
            native.baseline.memcpy_blocks(dp[a:b], pp[a:b], bs, True, device)
        t2 = time.perf_counter()
        if it:  # first iteration is warm-up
            tw += t1 - t0
            tr += t2 - t1
    torch.cuda.synchronize()
    ok = bool(torch.equal(src.cpu(), dst.cpu()))
    size = nblocks * bs * iters
    return {"baseline": label, "write_GBps": round(size / tw / 1e9, 2),
            "read_GBps": round(size / tr / 1e9, 2),
            "write_read_GBps": round(2 * size / (tw + tr) / 1e9, 2), "verified": ok,
            "block_kb": bs >> 10, "layers": layers}


def run_nccl(a, nblocks, bs, layers, iters):
    import torch.distributed as dist

    rank = int(os.environ["RANK"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    per = nblocks // layers
    src = torch.empty(nblocks * bs, dtype=torch.uint8, device=dev).random_(0, 255)
    pool = torch.empty_like(src)
    peer = 1 - rank
    tw = tr = 0.0
    for it in range(iters + 1):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for l in range(layers):  # "write": rank 0's pages -> rank 1's pool
            seg = slice(l * per * bs, (l + 1) * per * bs)
            if rank == 0:
                dist.send(src[seg], peer)
            else:
                dist.recv(pool[seg], peer)
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        for l in range(layers):  # "read": back
            seg = slice(l * per * bs, (l + 1) * per * bs)
            if rank == 1:
                dist.send(pool[seg], peer)
            else:
                dist.recv(src[seg], peer)
        torch.cuda.synchronize()
        dist.barrier()
        t2 = time.perf_counter()
        if it:
            tw += t1 - t0
            tr += t2 - t1
    size = nblocks * bs * iters
    if rank == 0:
        print(json.dumps({"baseline": "nccl send/recv (one tensor per request)",
                          "write_GBps": round(size / tw / 1e9, 2), "read_GBps": round(size / tr / 1e9, 2),
                          "write_read_GBps": round(2 * size / (tw + tr) / 1e9, 2),
                          "block_kb": bs >> 10, "layers": layers}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size-mb", type=int, default=1024)
    ap.add_argument("--block-kb", type=int, default=128)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--nccl", action="store_true")
    a = ap.parse_args()
    bs = a.block_kb << 10
    nblocks = (a.size_mb << 20) // bs
    if a.nccl:
        run_nccl(a, nblocks, bs, a.layers, a.iters)
        return
    out = []
    src = torch.empty(nblocks * bs, dtype=torch.uint8, device="cuda:0").random_(0, 255)
    dst = torch.zeros_like(src)
    host_pool = torch.empty(nblocks * bs, dtype=torch.uint8).pin_memory()
    out.append(run_memcpy("localgpu (reference LOCAL_GPU pattern: per-block cudaMemcpyAsync to pinned host)",
                          src, host_pool, dst, nblocks, bs, a.layers, a.iters, 0))
    pool0 = torch.empty_like(src)
    out.append(run_memcpy("samegpu (per-block cudaMemcpyAsync D2D, pool on the same GPU)",
                          src, pool0, dst, nblocks, bs, a.layers, a.iters, 0))
    if torch.cuda.device_count() >= 2:
        native.enable_peer_access(0, 1)
        native.enable_peer_access(1, 0)
        pool1 = torch.empty(nblocks * bs, dtype=torch.uint8, device="cuda:1")
        dst.zero_()
        out.append(run_memcpy("peercopy (per-block cudaMemcpyAsync GPU0 <-> GPU1)",
                              src, pool1, dst, nblocks, bs, a.layers, a.iters, 0))
    for r in out:
        print(json.dumps(r))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/baselines_{a.block_kb}k.json", "w"), indent=1)


if __name__ == "__main__":
    main()
