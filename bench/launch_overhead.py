#!/usr/bin/env python
"""Per-launch cost of the page mover at the batch size the store actually uses
(256 x 128 KB = 32 MB per call): where do the microseconds beyond bytes/bandwidth go?
Attribution matrix: descriptors in device memory vs pinned host memory, in-band publish on
or off, local pool vs peer pool.  Back-to-back launches in one stream, CUDA events."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infinistore_b200 import _infinistore as native  # noqa: E402
from infinistore_b200 import ops  # noqa: E402


def run(pool_dev, label, nblk, bs, launches, rows, variants):
    run_dev = "cuda:0"
    total = nblk * launches
    src = torch.empty(total * bs, dtype=torch.uint8, device=run_dev).random_(0, 255)
    pool = torch.empty(total * bs, dtype=torch.uint8, device=pool_dev)
    table = torch.zeros(4 * 65536 * 4, dtype=torch.int64, device=pool_dev)  # 262144 slots
    ideal_us = nblk * bs / (3.0e6 if pool_dev == run_dev else 0.71e6)
    for variant in variants:
        for desc_loc in ("device", "host"):
            for publish in (False, True):
                for direction in ("write", "read"):
                    if direction == "read" and publish:
                        continue
                    descs_all = []
                    pubs = []
                    for l in range(launches):
                        ids = range(l * nblk, (l + 1) * nblk)
                        a = [src.data_ptr() + i * bs for i in ids]
                        b = [pool.data_ptr() + i * bs for i in ids]
                        d = ops.make_descs(a, b, run_dev) if direction == "write" else \
                            ops.make_descs(b, a, run_dev)
                        if desc_loc == "host":
                            h = torch.empty(d.shape, dtype=torch.int64).pin_memory()
                            h.copy_(d.cpu())
                            # UVA: the pinned host pointer is directly usable by the kernel
                            d = h
                        descs_all.append(d)
                        if publish:
                            keys = [b"k-%d-%d" % (l, i) for i in range(nblk)]
                            pubs.append(ops.PublishArgs(
                                table, keys, [(1 << 44) | (i * bs) for i in ids],
                                list(range(1, nblk + 1)), bs))

                    def launch_all():
                        for l in range(launches):
                            d = descs_all[l]
                            p = pubs[l] if publish else None
                            native.kernels.kv_copy(
                                d.data_ptr(), nblk, bs, ops.VARIANTS[variant], 0,
                                ops._stream(torch.device(run_dev)),
                                p.recs.data_ptr() if p else 0, p.table.data_ptr() if p else 0,
                                p.mask if p else 0, p.done.data_ptr() if p else 0, 0, 0)

                    with torch.cuda.device(run_dev):
                        for _ in range(2):
                            table.zero_()
                            launch_all()
                        torch.cuda.synchronize()
                        ts = []
                        for _ in range(5):
                            table.zero_()
                            torch.cuda.synchronize()
                            e0 = torch.cuda.Event(enable_timing=True)
                            e1 = torch.cuda.Event(enable_timing=True)
                            e0.record()
                            launch_all()
                            e1.record()
                            e1.synchronize()
                            ts.append(e0.elapsed_time(e1) * 1e3 / launches)
                    us = sorted(ts)[len(ts) // 2]
                    rows.append({"pool": label, "variant": variant, "descs": desc_loc,
                                 "publish": publish, "dir": direction, "us_per_launch": round(us, 2),
                                 "gbps": round(nblk * bs / us / 1e3, 1),
                                 "overhead_us": round(us - ideal_us, 2)})
                    print(f"{label:6s} {variant:8s} descs={desc_loc:6s} publish={int(publish)} "
                          f"{direction:5s} {us:7.2f} us/launch {nblk * bs / us / 1e3:7.1f} GB/s "
                          f"(+{us - ideal_us:5.1f} us over ideal)", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=256)
    ap.add_argument("--block-kb", type=int, default=128)
    ap.add_argument("--launches", type=int, default=32)
    ap.add_argument("--out", default="gpurun_out/launch_overhead.json")
    a = ap.parse_args()
    rows = []
    run("cuda:0", "local", a.blocks, a.block_kb << 10, a.launches, rows, ["ldst", "ldst256", "tma"])
    if torch.cuda.device_count() >= 2:
        assert native.enable_peer_access(0, 1) and native.enable_peer_access(1, 0)
        run("cuda:1", "peer", a.blocks, a.block_kb << 10, a.launches, rows, ["ldst", "tma"])
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
