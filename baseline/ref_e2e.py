"""End-to-end step of the REFERENCE through its public Python API (runs in a clean interpreter
with only baseline/_ref on the path): per step the pages come from pinned host memory (H2D,
layer by layer), are written with `local_gpu_write_cache`, read back with `read_cache` and a
slice of the result goes back to the host.  Same shape as bench.py's e2e for the b200 arm.
The reference's server copies on its own streams, unordered with the client's: the client
must `torch.cuda.synchronize()` before every write call (as the reference's tests do,
infinistore/test_infinistore.py) - that is part of its end-to-end cost."""
import argparse
import json
import time
import uuid

import infinistore
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--service-port", type=int, required=True)
    ap.add_argument("--size-mb", type=int, default=1024)
    ap.add_argument("--block-kb", type=int, default=128)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--gpu", type=int, default=0)
    a = ap.parse_args()
    cfg = infinistore.ClientConfig(host_addr="127.0.0.1", service_port=a.service_port,
                                   log_level="warning")
    cfg.connection_type = infinistore.TYPE_LOCAL_GPU
    conn = infinistore.InfinityConnection(cfg)
    conn.connect()
    dev = f"cuda:{a.gpu}"
    torch.cuda.set_device(a.gpu)
    elems = a.block_kb * 1024 // 4
    nblocks = a.size_mb * 1024 // a.block_kb
    layers = a.layers
    while nblocks % layers and layers > 1:
        layers //= 2
    per = nblocks // layers
    with infinistore.DisableTorchCaching():
        src = torch.zeros(nblocks * elems, device=dev, dtype=torch.float32)
        dst = torch.zeros(nblocks * elems, device=dev, dtype=torch.float32)
    host_src = torch.rand(nblocks * elems, dtype=torch.float32).pin_memory()
    host_out = torch.empty(elems, dtype=torch.float32).pin_memory()
    offs = [i * elems for i in range(nblocks)]
    secs = 0.0
    ok = True
    for step in range(a.steps + 1):  # first one is warm-up
        keys = [str(uuid.uuid4()) for _ in range(nblocks)]
        blocks = list(zip(keys, offs))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for l in range(layers):
            s = slice(l * per * elems, (l + 1) * per * elems)
            src[s].copy_(host_src[s], non_blocking=True)
            torch.cuda.synchronize()  # the server's copy is not ordered with our stream
            conn.local_gpu_write_cache(src, blocks[l * per:(l + 1) * per], elems)
        conn.sync()
        for l in range(layers):
            conn.read_cache(dst, blocks[l * per:(l + 1) * per], elems)
        conn.sync()
        host_out.copy_(dst[:elems], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if step:
            secs += dt
        ok = ok and bool(torch.equal(host_out, host_src[:elems]))
    print(json.dumps({"e2e_secs": secs, "steps": a.steps, "h2d_bytes_per_step": nblocks * elems * 4,
                      "d2h_bytes_per_step": elems * 4, "verified": ok}))


if __name__ == "__main__":
    main()
