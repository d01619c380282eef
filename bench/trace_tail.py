#!/usr/bin/env python
"""Where does the epilogue of a write kernel spend its time?  Per-CTA %globaltimer stamps
from the control warp: entry, claims done, copy warps done, stores performed (fence), commit."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infinistore_b200 import _infinistore as native  # noqa: E402
from infinistore_b200 import ops  # noqa: E402


def run(pool_dev, label, nblk=256, bs=128 << 10, ctas=0, all_local=False, debug=0):
    dev = "cuda:0"
    src = torch.empty(nblk * bs, dtype=torch.uint8, device=dev).random_(0, 255)
    pool = torch.empty(nblk * bs, dtype=torch.uint8, device=pool_dev)
    table = torch.zeros(65536 * 4, dtype=torch.int64, device=pool_dev)
    d = ops.make_descs([src.data_ptr() + i * bs for i in range(nblk)],
                       [pool.data_ptr() + i * bs for i in range(nblk)], dev)
    out = {}
    for it in range(4):
        table.zero_()
        keys = [b"t-%d-%d" % (it, i) for i in range(nblk)]
        p = ops.PublishArgs(table, keys, [(1 << 44) | (i * bs) for i in range(nblk)],
                            list(range(1, nblk + 1)), bs)
        trace = torch.zeros(2048 * 8, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        with torch.cuda.device(dev):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            native.kernels.kv_copy(d.data_ptr(), nblk, bs, ops.VARIANTS["ldst256"], ctas,
                                   ops._stream(torch.device(dev)), p.recs.data_ptr(),
                                   p.table.data_ptr(), p.mask, p.done.data_ptr(), 0, 0,
                                   trace.data_ptr(), all_local, debug)
            e1.record()
            e1.synchronize()
        t = trace.cpu().numpy().reshape(-1, 8)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        rel = (t[:, :5] - t0) / 1e3
        out = {"event_us": round(e0.elapsed_time(e1) * 1e3, 1), "ctas": int(len(t)),
               "entry_max": rel[:, 0].max(), "claim_done_med": np.median(rel[:, 1]),
               "claim_done_max": rel[:, 1].max(), "copy_done_med": np.median(rel[:, 2]),
               "copy_done_max": rel[:, 2].max(), "fence_done_med": np.median(rel[:, 3]),
               "fence_done_max": rel[:, 3].max(), "commit_max": rel[:, 4].max(),
               "fence_cost_med": np.median(rel[:, 3] - rel[:, 2]),
               "fence_cost_max": (rel[:, 3] - rel[:, 2]).max()}
    out = {k: (round(float(v), 2) if not isinstance(v, int) else v) for k, v in out.items()}
    shown = ("event_us", "fence_done_max", "commit_max", "claim_done_max", "copy_done_max")
    print(label, {a: b for a, b in out.items() if a in shown}, flush=True)
    return out


res = {"local": run("cuda:0", "local-sys"), "local_gpu_scope": run("cuda:0", "local-gpu", all_local=True)}
for dbg in (1, 2, 4, 7):
    res[f"local_gpu_dbg{dbg}"] = run("cuda:0", f"local-gpu dbg={dbg}", all_local=True, debug=dbg)
if torch.cuda.device_count() >= 2:
    native.enable_peer_access(0, 1)
    res["peer"] = run("cuda:1", "peer ")
    for dbg in (1, 2, 4, 7):
        res[f"peer_dbg{dbg}"] = run("cuda:1", f"peer dbg={dbg}", debug=dbg)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/trace_tail.json", "w"), indent=1)
