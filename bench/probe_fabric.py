#!/usr/bin/env python
"""What the box allows: device count, P2P matrix, VMM / POSIX-fd / fabric handles, NVLS."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infinistore_b200 import _infinistore as native  # noqa: E402

out = {"devices": torch.cuda.device_count(), "name": torch.cuda.get_device_name(0)}
n = out["devices"]
out["p2p"] = [[bool(torch.cuda.can_device_access_peer(i, j)) if i != j else True
               for j in range(n)] for i in range(n)]
p = native.nvls_probe(0)
out["nvls_probe"] = {k: getattr(p, k) for k in ("driver_ok", "multicast_supported", "vmm_supported",
                                                 "posix_fd_supported", "fabric_handle_supported",
                                                 "granularity", "detail")}
if n >= 2 and p.multicast_supported:
    try:
        g = native.NvlsGroup.create(list(range(n)), 64 << 20)
        out["nvls_group"] = {"size": g.size(), "bytes": g.bytes(), "mc_ptr": hex(g.mc_ptr(0))}
    except Exception as e:  # noqa: BLE001
        out["nvls_group"] = {"error": str(e)}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_fabric.json", "w"), indent=1)
