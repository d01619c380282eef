#!/usr/bin/env python
"""The multi-GPU configurations of BASELINE.json beyond the flagship (bench.py = config 2).

  fanin   (config 3)  N-1 client GPUs -> 1 pool GPU, Llama-3-8B KV pages (256 KiB per K or V
                      page per layer), layer-wise writes then reads.     torchrun, N ranks
  bcast   (config 4)  1 writer -> all GPUs through NVLS multicast, 1 MB blocks, plus
                      get_match_last_index on the device.                 single process
  fp8     (config 5)  fp8 KV path: write fused with the bf16->e4m3 cast, read fused with the
                      dequantising gather, 64 KB (fp8) blocks, ring over N GPUs.  torchrun

Every number is wall time around (issue + sync) taken as the max over ranks; one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import uuid

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import infinistore_b200 as ist  # noqa: E402
from infinistore_b200 import _infinistore as native  # noqa: E402
from infinistore_b200.models import get_layout  # noqa: E402
from infinistore_b200.parallel import PrefixBroadcaster, nvls_available, start_shard_server  # noqa: E402


def dist_setup():
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    return dist, rank, world, local, dev


def allmax(dist, dev, x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def base_port():
    return 25000 + int(os.environ.get("MASTER_PORT", "0")) % 2000


# --------------------------------------------------------------------------- config 3
def fanin(a):
    dist, rank, world, local, dev = dist_setup()
    layout = get_layout("llama-3-8b")           # 256 KiB pages
    pages, layers, elems = a.pages, layout.layers, layout.page_elems
    nblk = pages * 2 * layers                    # K and V pages of every layer
    per_client = nblk * layout.page_bytes
    port = base_port()
    server = None
    if rank == 0:
        server = start_shard_server(local, port, (world - 1) * per_client * (a.iters + 2) + (256 << 20),
                                    granule_kb=64)
    dist.barrier()
    tw = tr = 0.0
    ok = True
    if rank > 0:
        conn = ist.InfinityConnection(ist.ClientConfig(
            host_addr="127.0.0.1", service_port=port, connection_type=ist.TYPE_RDMA,
            device=local, device_lookup=True))
        conn.connect()
        src = torch.randn(nblk * elems, device=dev).to(torch.bfloat16)
        dst = torch.zeros_like(src)
        conn.register_mr(src)
        conn.register_mr(dst)
        per_layer = pages * 2
        offs = np.arange(nblk, dtype=np.int64) * elems
    for it in range(a.iters + 1):
        if rank > 0:
            keys = [f"r{rank}/{it}/{i}/{uuid.uuid4().hex[:8]}" for i in range(nblk)]
            remote = conn.allocate_rdma(keys, layout.page_bytes)
            blocks = list(zip(keys, offs.tolist()))
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        if rank > 0:
            for l in range(layers):
                s = slice(l * per_layer, (l + 1) * per_layer)
                conn.rdma_write_cache(src, offs[s], elems, remote[s])
            conn.sync()
        t1 = time.perf_counter()
        dist.barrier()
        t2 = time.perf_counter()
        if rank > 0:
            for l in range(layers):
                s = slice(l * per_layer, (l + 1) * per_layer)
                conn.read_cache(dst, blocks[s], elems)
            conn.sync()
        t3 = time.perf_counter()
        dist.barrier()
        if it:
            tw += t1 - t0
            tr += t3 - t2
    if rank > 0:
        ok = bool(torch.equal(src, dst))
        conn.close()
    tw, tr = allmax(dist, dev, tw), allmax(dist, dev, tr)
    okall = allmax(dist, dev, 0.0 if ok else 1.0) == 0.0
    total = (world - 1) * per_client * a.iters
    if rank == 0:
        print(json.dumps({"config": "fanin (BASELINE config 3)", "clients": world - 1, "model": "llama-3-8b",
                          "page_kib": layout.page_bytes >> 10, "bytes_per_client_per_iter": per_client,
                          "write_GBps_aggregate": round(total / tw / 1e9, 1),
                          "read_GBps_aggregate": round(total / tr / 1e9, 1),
                          "roofline_GBps": {"pool_ingress_write": 711, "pool_egress_read": 779},
                          "verified": okall}))
        server.stop()
    dist.destroy_process_group()


# --------------------------------------------------------------------------- config 4
def bcast(a):
    ndev = torch.cuda.device_count()
    out = {"config": "bcast (BASELINE config 4)", "gpus": ndev}
    if ndev < 2 or not nvls_available():
        out["unavailable"] = "needs >= 2 GPUs with NVLS multicast"
        print(json.dumps(out))
        return
    bs, nblk = 1 << 20, a.blocks
    bc = PrefixBroadcaster(list(range(ndev)), nblk * bs)
    src = torch.randint(0, 255, (nblk * bs,), dtype=torch.uint8, device="cuda:0")
    offs = [i * bs for i in range(nblk)]
    for ctas in (0, 148, 296):
        for _ in range(2):
            bc.broadcast(src, offs, offs, bs, max_ctas=ctas)
        torch.cuda.synchronize(0)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.device(0):
                e0.record()
                bc.broadcast(src, offs, offs, bs, max_ctas=ctas)
                e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[f"ctas={ctas or 'auto'}"] = {
            "ms": round(ms, 3), "writer_egress_GBps": round(nblk * bs / ms / 1e6, 1),
            "delivered_GBps": round((ndev - 1) * nblk * bs / ms / 1e6, 1)}
    for d in range(ndev):
        torch.cuda.synchronize(d)
    out["verified"] = all(bool(torch.equal(bc.replica(d)[:nblk * bs].cpu(), src.cpu()))
                          for d in range(ndev))
    # unicast comparison: the same blocks pushed to each peer one after the other
    from infinistore_b200 import ops

    peers = []
    for d in range(1, ndev):
        native.enable_peer_access(0, d)
        peers.append(torch.empty(nblk * bs, dtype=torch.uint8, device=f"cuda:{d}"))
    descs = [ops.make_descs([src.data_ptr() + o for o in offs], [p.data_ptr() + o for o in offs], "cuda:0")
             for p in peers]
    with torch.cuda.device(0):
        for _ in range(2):
            for d in descs:
                ops.kv_copy(d, bs)
        torch.cuda.synchronize(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for d in descs:
            ops.kv_copy(d, bs)
        e1.record()
        e1.synchronize()
    ms = e0.elapsed_time(e1)
    out["unicast_to_each_peer"] = {"ms": round(ms, 3),
                                   "delivered_GBps": round((ndev - 1) * nblk * bs / ms / 1e6, 1)}
    # get_match_last_index on the device: keys published in a local index table
    table = ops.new_index_table(1 << 18, "cuda:0")
    nkeys = a.keys
    keys = [b"prefix-%08d" % i for i in range(nkeys)]
    pool = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
    pub = ops.PublishArgs(table, keys[: nkeys // 2], [(1 << 44)] * (nkeys // 2),
                          list(range(1, nkeys // 2 + 1)), 64)
    d = ops.make_descs([pool.data_ptr()] * (nkeys // 2), [pool.data_ptr()] * (nkeys // 2), "cuda:0")
    ops.kv_copy(d, 64, publish=pub)
    torch.cuda.synchronize(0)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        _, _, match = ops.index_lookup(table, keys, want_match=True)
    dt = (time.perf_counter() - t0) / reps
    out["match_last_index"] = {"keys": nkeys, "result": match, "expected": nkeys // 2 - 1,
                               "ms_per_call_incl_pack_and_sync": round(dt * 1e3, 3)}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/config4_bcast.json", "w"), indent=1)


# --------------------------------------------------------------------------- config 5
def fp8(a):
    dist, rank, world, local, dev = dist_setup()
    elems = 65536                                  # 128 KB bf16 page -> 64 KB e4m3 (+ scales)
    nblk = (a.size_mb << 20) // (elems * 2)
    port = base_port()
    server = start_shard_server(local, port + rank, (a.iters + 2) * nblk * 80 * 1024 + (256 << 20),
                                granule_kb=16)
    dist.barrier()
    conn = ist.InfinityConnection(ist.ClientConfig(
        host_addr="127.0.0.1", service_port=port + (rank + 1) % world,
        connection_type=ist.TYPE_RDMA, device=local, device_lookup=True))
    conn.connect()
    src = (torch.randn(nblk * elems, device=dev) * 2).to(torch.bfloat16)
    dst = torch.zeros_like(src)
    conn.register_mr(src)
    conn.register_mr(dst)
    layers = 32
    per = nblk // layers
    offs = np.arange(nblk, dtype=np.int64) * elems
    nbytes = conn.fp8_page_bytes(elems)
    tw = tr = 0.0
    for it in range(a.iters + 1):
        keys = [uuid.uuid4().hex for _ in range(nblk)]
        remote = conn.allocate_rdma(keys, nbytes)
        blocks = list(zip(keys, offs.tolist()))
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for l in range(layers):
            s = slice(l * per, (l + 1) * per)
            conn.rdma_write_cache_fp8(src, offs[s], elems, remote[s])
        conn.sync()
        t1 = time.perf_counter()
        for l in range(layers):
            s = slice(l * per, (l + 1) * per)
            conn.read_cache_fp8(dst, blocks[s], elems)
        conn.sync()
        t2 = time.perf_counter()
        dist.barrier()
        if it:
            tw += t1 - t0
            tr += t2 - t1
    err = (dst.float() - src.float()).abs().max().item() / src.float().abs().max().item()
    tw, tr = allmax(dist, dev, tw), allmax(dist, dev, tr)
    err = allmax(dist, dev, err)
    conn.close()
    dist.barrier()
    server.stop()
    if rank == 0:
        fp8_bytes = world * nblk * nbytes * a.iters
        bf16_bytes = world * nblk * elems * 2 * a.iters
        print(json.dumps({"config": "fp8 ring (BASELINE config 5)", "gpus": world, "fp8_block_kb": nbytes / 1024,
                          "write_GBps_fp8_bytes": round(fp8_bytes / tw / 1e9, 1),
                          "read_GBps_fp8_bytes": round(fp8_bytes / tr / 1e9, 1),
                          "write_GBps_bf16_equiv": round(bf16_bytes / tw / 1e9, 1),
                          "read_GBps_bf16_equiv": round(bf16_bytes / tr / 1e9, 1),
                          "max_rel_err_vs_bf16": round(err, 4)}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["fanin", "bcast", "fp8"])
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--pages", type=int, default=8, help="fanin: 128-token pages per client")
    ap.add_argument("--blocks", type=int, default=256, help="bcast: 1 MB blocks")
    ap.add_argument("--keys", type=int, default=4096, help="bcast: keys for match_last_index")
    ap.add_argument("--size-mb", type=int, default=1024, help="fp8: bf16 MB per GPU per iteration")
    a = ap.parse_args()
    {"fanin": fanin, "bcast": bcast, "fp8": fp8}[a.config](a)


if __name__ == "__main__":
    main()
