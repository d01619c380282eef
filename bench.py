#!/usr/bin/env python
"""Flagship benchmark: KV-block write+read throughput through the store.

Metric (BASELINE.json): KV-block write/read GB/s, 128 KB paged-KV blocks, synthetic random
blocks.  Methodology mirrors the reference's own benchmark (infinistore/benchmark.py:
132-208): `--size-mb` of KV pages split into `--block-kb` blocks with fresh UUID keys every
step, block allocation outside the timed region, writes then reads issued in `--layers`
batches, one `sync()` after each phase, read-back verified against the source.

Topology at N GPUs (one process per GPU, torchrun): every rank hosts one pool shard of
the store (server thread + HBM pool on its GPU) and is a client of the shard on GPU
(rank+1) % N, so for N >= 2 every byte crosses NVLink exactly once per direction and the
per-GPU work is fixed (weak scaling).  N = 1: the pool is on the same GPU (HBM to HBM).

One JSON line on rank 0.  `value` = aggregate (write+read) payload GB/s over all ranks,
timed on the device with CUDA events (max over ranks).  `e2e` = the same metric through
the public API including, per step, the host->device copy of the step's pages from pinned
memory (overlapped layer by layer with the writes) and a device->host read of the result.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import uuid

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "kv_block_write_read_GBps"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--size-mb", type=int, default=1024, help="KV bytes written+read per GPU per step")
    p.add_argument("--block-kb", type=int, default=128)
    p.add_argument("--layers", type=int, default=32, help="batches per phase (reference --steps)")
    p.add_argument("--variant", default="auto", choices=["auto", "ldst", "tma", "ldst256"])
    p.add_argument("--max-ctas", type=int, default=0)
    p.add_argument("--streams", type=int, default=4, help="internal launch streams per connection")
    p.add_argument("--host-lookup", action="store_true",
                   help="resolve read keys through the server instead of the HBM index")
    p.add_argument("--base-port", type=int, default=0)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--pool-gb", type=int, default=48,
                   help="HBM pool per GPU; steps beyond its capacity run in further epochs "
                        "(purge + re-allocate outside the timed region)")
    return p.parse_args()


def reference_arm(args):
    """The reference cannot be built on this image: `pip install` of /root/reference succeeds
    only for its Python files; the native module needs infiniband/verbs.h, libuv,
    flatbuffers and boost, none of which exist here."""
    why = None
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "infinistore")):
        why = ("baseline/_ref/infinistore is absent: the offline pip install of /root/reference "
               "only yields its .py files and its native module cannot be built (no libibverbs, "
               "libuv, flatbuffers, boost headers on this image)")
    else:
        # Import the reference in a clean interpreter (cwd and PYTHONPATH outside this repo),
        # so that nothing of this repo - in particular its `infinistore` alias package - can
        # stand in for the reference's own native module.
        env = dict(os.environ, PYTHONPATH=ref)
        r = subprocess.run([sys.executable, "-c", "import infinistore._infinistore"],
                           cwd="/tmp", env=env, capture_output=True, text=True, timeout=300)
        if r.returncode != 0:
            last = (r.stderr.strip().splitlines() or ["import failed"])[-1]
            why = ("reference native module not buildable offline (needs libibverbs/libuv/"
                   "flatbuffers/boost headers): " + last)
    if why is None:
        why = "reference imported but needs an active mlx5 RDMA port, absent on this box"
    print(json.dumps({"impl": "reference", "unavailable": why[:300]}))


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs.  Uses
    NVML in-process (a `nvidia-smi` subprocess every 200 ms takes driver-wide locks long
    enough to perturb a launch-latency-sensitive loop); falls back to nvidia-smi."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = 0
        self._stop_ev = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            idx = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis:
                idx = int(vis.split(",")[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nvml = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for name, bit in (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(
            ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
             "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5
        ).stdout.strip().split(",")
        if len(out) >= 6:
            self.samples.append(int(float(out[0])))
            self.max_mhz = int(float(out[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown",
                                "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                if v.strip().lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_ev.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:  # noqa: BLE001
                pass
            self._stop_ev.wait(0.1 if self._nvml is not None else 0.5)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz or None,
                "reasons": sorted(self.reasons), "samples": len(s),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch

    import infinistore_b200 as ist
    from infinistore_b200 import _infinistore as native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (run it through gpurun)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    block_bytes = args.block_kb * 1024
    size_bytes = args.size_mb << 20
    nblocks = size_bytes // block_bytes
    layers = args.layers
    while nblocks % layers != 0 and layers > 1:
        layers //= 2
    per_layer = nblocks // layers
    elems = block_bytes // 2  # bf16 KV pages
    total_steps = args.steps + args.warmup
    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 5))

    # ---- one pool shard per rank
    base_port = args.base_port or (23000 + (int(os.environ.get("MASTER_PORT", "0")) % 2000))
    scfg = native.ServerConfig()
    scfg.service_port = base_port + rank
    scfg.host = "127.0.0.1"
    scfg.pool_backend = "hbm"
    scfg.pool_devices = [local_rank]
    scfg.minimal_allocate_size = max(16, min(args.block_kb, 64))
    epoch_cap = max(1, ((args.pool_gb << 30) - (64 << 20)) // size_bytes - 1)
    pool_steps = min(max(total_steps, e2e_steps + 1), epoch_cap)
    scfg.prealloc_bytes = (pool_steps + 1) * size_bytes + (64 << 20)
    scfg.log_level = "warning"
    server = native.Server(scfg)
    server.start()
    barrier()

    peer = (rank + 1) % world
    ccfg = ist.ClientConfig(host_addr="127.0.0.1", service_port=base_port + peer,
                            connection_type=ist.TYPE_RDMA, log_level="warning",
                            device=local_rank, device_lookup=not args.host_lookup,
                            copy_variant=args.variant, max_ctas=args.max_ctas,
                            streams=args.streams)
    conn = ist.InfinityConnection(ccfg)
    conn.connect()

    # ---- synthetic KV pages (bf16), larger than L2 (126 MB) so nothing is served from cache
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    src = torch.randn(nblocks * elems, device=dev, dtype=torch.float32, generator=gen).to(torch.bfloat16)
    dst = torch.zeros_like(src)
    conn.register_mr(src)
    conn.register_mr(dst)
    offsets = [i * elems for i in range(nblocks)]
    stream = torch.cuda.Stream(device=dev)

    def fresh_step():
        keys = [str(uuid.uuid4()) for _ in range(nblocks)]
        remote = conn.allocate_rdma(keys, block_bytes)
        return keys, remote

    host_t = {"issue_write": 0.0, "sync_write": 0.0, "issue_read": 0.0, "sync_read": 0.0}
    phase_events = []

    def run_step(keys, remote, blocks, record=False):
        t0 = time.perf_counter()
        if record:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record(stream)
        for l in range(layers):
            a, b = l * per_layer, (l + 1) * per_layer
            conn.rdma_write_cache(src, offsets_np[a:b], elems, remote[a:b])
        if record:
            e[1].record(stream)
        t1 = time.perf_counter()
        conn.sync()
        t2 = time.perf_counter()
        if record:
            e[2].record(stream)
        for l in range(layers):
            a, b = l * per_layer, (l + 1) * per_layer
            conn.read_cache(dst, blocks[a:b], elems)
        if record:
            e[3].record(stream)
            phase_events.append(e)
        t3 = time.perf_counter()
        conn.sync()
        t4 = time.perf_counter()
        if record:
            host_t["issue_write"] += t1 - t0
            host_t["sync_write"] += t2 - t1
            host_t["issue_read"] += t3 - t2
            host_t["sync_read"] += t4 - t3

    offsets_np = np.asarray(offsets, dtype=np.int64)

    def fresh_prepared():
        keys, remote = fresh_step()
        return keys, remote, list(zip(keys, offsets))  # (key, offset) list built like the reference

    def new_epoch(nsteps):
        """Empty this rank's pool shard and reserve blocks for `nsteps` steps (untimed)."""
        barrier()          # nobody is still reading the shard we are about to purge
        server.purge()
        barrier()
        return [fresh_prepared() for _ in range(nsteps)]  # allocation is outside the timing

    def run_epochs(nsteps, record):
        """Run `nsteps` steps; returns the device time (ms) of the timed regions: each epoch
        is bracketed by barrier + synchronize on both sides and timed with CUDA events."""
        total_ms = 0.0
        remaining = nsteps
        while remaining > 0:
            n_ep = min(remaining, epoch_cap)
            prepared = new_epoch(n_ep)
            dst.zero_()
            torch.cuda.synchronize()
            barrier()
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for s in range(n_ep):
                run_step(*prepared[s], record=record)
            ev1.record(stream)
            torch.cuda.synchronize()
            barrier()
            total_ms += ev0.elapsed_time(ev1)
            remaining -= n_ep
        return total_ms

    with torch.cuda.stream(stream):
        run_epochs(args.warmup, record=False)
        torch.cuda.synchronize()
        assert args.warmup == 0 or torch.equal(src, dst), "read-back mismatch after warm-up"
        launches0 = conn.stats()["kernel_launches"]
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms = run_epochs(args.steps, record=True)
        clocks = sampler.stop()
        launches = conn.stats()["kernel_launches"] - launches0
        # phase wall time = issue + sync (kernels run on the connection's internal streams)
        w_ms = (host_t["issue_write"] + host_t["sync_write"]) / args.steps * 1e3
        r_ms = (host_t["issue_read"] + host_t["sync_read"]) / args.steps * 1e3
        breakdown = {"write_phase_ms": round(w_ms, 3), "read_phase_ms": round(r_ms, 3),
                     "write_phase_GBps": round(nblocks * block_bytes / w_ms / 1e6, 1),
                     "read_phase_GBps": round(nblocks * block_bytes / r_ms / 1e6, 1),
                     **{"host_" + k + "_ms": round(v / args.steps * 1e3, 3) for k, v in host_t.items()},
                     "cpus": os.cpu_count(), "epochs": -(-args.steps // epoch_cap)}
    ok = bool(torch.equal(src, dst))

    ms_max = allmax(ms)
    ms_per_step = ms_max / args.steps
    bytes_per_step = 2 * nblocks * block_bytes  # write + read
    value = world * bytes_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- end to end: pinned host pages -> H2D (layer-wise, overlapped) -> write -> sync ->
    #      read -> sync -> D2H of the result
    e2e = None
    if e2e_steps:
        host_src = torch.empty(nblocks * elems, dtype=torch.bfloat16).pin_memory()
        host_src.copy_(src.cpu())
        host_out = torch.empty(elems + 1, dtype=torch.bfloat16).pin_memory()
        e2e_steps = min(e2e_steps, epoch_cap - 1) if epoch_cap > 1 else 1
        e2e_prepared = new_epoch(e2e_steps + 1)
        h2d_bytes = nblocks * block_bytes
        d2h_bytes = (elems + 1) * 2

        def e2e_step(keys, remote, blocks):
            with torch.cuda.stream(stream):
                for l in range(layers):
                    a, b = l * per_layer, (l + 1) * per_layer
                    src[a * elems:b * elems].copy_(host_src[a * elems:b * elems], non_blocking=True)
                    conn.rdma_write_cache(src, offsets_np[a:b], elems, remote[a:b])
                conn.sync()
                for l in range(layers):
                    a, b = l * per_layer, (l + 1) * per_layer
                    conn.read_cache(dst, blocks[a:b], elems)
                conn.sync()
                same = (dst[-elems:] == src[-elems:]).all().to(torch.bfloat16).reshape(1)
                host_out[:elems].copy_(dst[:elems], non_blocking=True)
                host_out[elems:].copy_(same, non_blocking=True)
                stream.synchronize()
            return float(host_out[elems].item())

        e2e_step(*e2e_prepared[0])  # warm-up
        barrier()
        t0 = time.perf_counter()
        good = 1.0
        for s in range(1, e2e_steps + 1):
            good = min(good, e2e_step(*e2e_prepared[s]))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt_max = allmax(dt)
        ok = ok and good == 1.0
        e2e = {"value": round(world * bytes_per_step * e2e_steps / dt_max / 1e9, 2), "unit": "GB/s",
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
               "steps": e2e_steps, "ms_per_step": round(dt_max / e2e_steps * 1e3, 3)}

    all_ok = allsum(0.0 if ok else 1.0) == 0.0
    total_launches = int(allsum(float(launches)))
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    if world == 1:
        # local HBM->HBM copy: every payload byte is read once and written once
        roof = peaks.get("hbm_gbs", 6650.0) / 2
        roof_name = "measured HBM copy bandwidth / 2 (read+write per payload byte)"
    else:
        # ring placement: every GPU pushes and is pushed to (write phase), pulls and is pulled
        # from (read phase) at the same time; measured with both directions busy
        # (bench/bidir.py, profiles/r1_nvlink_bidirectional_roofline.json): 703 / 667 GB/s
        roof = world * 2.0 / (1.0 / 703.0 + 1.0 / 667.0)
        roof_name = ("measured NVLink bandwidth with both directions busy: 703 GB/s push, "
                     "667 GB/s pull per GPU (unidirectional: 711 / 779; nominal 900)")

    conn.close()
    server.stop()
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic random KV pages, fresh UUID keys per step",
            "verified": all_ok, "impl": "b200",
            "config": {"model": "paged-KV blocks", "block_kb": args.block_kb,
                       "bytes_per_gpu_per_step": bytes_per_step, "global_batch": nblocks * world,
                       "seq_len": None, "layers_per_phase": layers,
                       "parallelism": f"{world} pool shards, ring placement (rank r -> GPU (r+1)%N)",
                       "l2": "working set 2 GiB per GPU per step >> 126 MB L2 (no flush needed)",
                       "variant": args.variant, "streams": args.streams, "lookup": "host" if args.host_lookup else "device-index",
                       "phase_sync": True},
            "roofline": {"gbps": round(roof, 1), "what": roof_name,
                         "fraction": round(value / roof, 3)},
            "clocks": clocks, "gpu_launches": total_launches, "e2e": e2e,
            "breakdown": breakdown,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
