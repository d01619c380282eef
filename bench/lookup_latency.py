#!/usr/bin/env python
"""Device-index lookup kernel alone: CUDA-event time per launch for n keys (hits / misses)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from infinistore_b200 import ops  # noqa: E402

DEV = "cuda:0"


def main():
    n_keys = 1 << 16
    table = ops.new_index_table(1 << 18, DEV)
    keys = [b"lat/%07d" % i for i in range(n_keys)]
    pool = torch.zeros(64, dtype=torch.uint8, device=DEV)
    for lo in range(0, n_keys, 8192):
        ks = keys[lo:lo + 8192]
        wd = ops.make_descs([pool.data_ptr()] * len(ks), [pool.data_ptr()] * len(ks), DEV)
        pub = ops.PublishArgs(table, ks, [(1 << 44) | ((lo + i) * 64) for i in range(len(ks))],
                              list(range(lo + 1, lo + len(ks) + 1)), 64)
        ops.kv_copy(wd, 64, publish=pub)
        torch.cuda.synchronize()
    res = {}
    for n in (1, 128, 4096):
        for kind in ("hit", "miss"):
            q = keys[:n] if kind == "hit" else [b"absent/%07d" % i for i in range(n)]
            kb, ko, kl = ops.pack_keys(q, DEV)
            slots = table.numel() * 8 // 32
            mask = slots // 8 - 1 if hasattr(ops, "index_bucket_mask") else slots - 1
            descs = torch.zeros((n, 2), dtype=torch.int64, device=DEV)
            doff = torch.zeros(n, dtype=torch.int64, device=DEV)
            present = torch.zeros((n + 31) // 32, dtype=torch.int32, device=DEV)
            status = torch.zeros(8, dtype=torch.int32, device=DEV)
            ticket = torch.zeros(1, dtype=torch.int32, device=DEV)
            for mode, pres in (("read", 0), ("match", present.data_ptr())):
                ts = []
                for rep in range(40):
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    ops.K.index_lookup(kb.data_ptr(), ko.data_ptr(), kl.data_ptr(), n,
                                       table.data_ptr(), mask, [0x1000000], descs.data_ptr(),
                                       doff.data_ptr(), 0, 1, pres, status.data_ptr(),
                                       ticket.data_ptr(), False, ops._stream(torch.device(DEV)))
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                res[f"n{n}_{kind}_{mode}_us_p50"] = round(ts[len(ts) // 2], 2)
            if kind == "hit":
                assert (descs[:, 0] != 0).all()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
