#!/usr/bin/env python
"""One launch (a few) of every data-plane kernel for `ncu`, on PEER memory when a second GPU
is visible (single process: this is what ncu can profile; bench.py's ranks are separate
processes and are never wrapped in ncu).

    ncu --set full --section Nvlink --clock-control none --import-source on \\
        -k regex:'kv_|istore' -o gpurun_out/r2_prof python bench/ncu_driver.py

Kernels, in launch order (128 KB pages unless noted, 512 pages = 64 MiB per launch - long
enough for steady state, short enough for ~40 replays):
  raw launchers : kv_pipe_copy push / pull / local, kv_copy_ldst<32> push / pull
  store API     : write (kv_pipe_copy + publish), read (kv_pipe_read, fused lookup),
                  read_cache_hnd (kv_pipe_hnd), read_cache_multi (fan-out or cluster multicast),
                  fp8 write / read (kv_fp8_pipe), get_match_last_index (kv_index_lookup)
  NVLS          : kv_bcast_nvls + kv_read_when_ready (2 GPUs with multicast)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infinistore_b200 as ist  # noqa: E402
from infinistore_b200 import _infinistore as native  # noqa: E402
from infinistore_b200 import ops  # noqa: E402
from infinistore_b200.parallel import PrefixBroadcaster, nvls_available  # noqa: E402

two = torch.cuda.device_count() >= 2
peer = "cuda:1" if two else "cuda:0"
if two:
    assert native.enable_peer_access(0, 1) and native.enable_peer_access(1, 0)
torch.cuda.set_device(0)
bs, n = 128 << 10, 512
loc_a = torch.empty(n * bs, dtype=torch.uint8, device="cuda:0").random_(0, 255)
loc_b = torch.empty(n * bs, dtype=torch.uint8, device="cuda:0")
rem = torch.empty(n * bs, dtype=torch.uint8, device=peer).random_(0, 255)


def descs(src, dst):
    return ops.make_descs([src.data_ptr() + i * bs for i in range(n)],
                          [dst.data_ptr() + i * bs for i in range(n)], "cuda:0")


push, pull, local = descs(loc_a, rem), descs(rem, loc_b), descs(loc_a, loc_b)
for d in (push, pull, local):
    ops.kv_copy(d, bs, variant="tma")
for d in (push, pull):
    ops.kv_copy(d, bs, variant="ldst256")
torch.cuda.synchronize()

# ---- through the store: pool on the peer GPU
cfg = native.ServerConfig()
cfg.service_port = 0
cfg.host = "127.0.0.1"
cfg.pool_backend = "hbm"
cfg.pool_devices = [1 if two else 0]
cfg.prealloc_bytes = 1 << 30
cfg.minimal_allocate_size = 64
srv = native.Server(cfg)
port = srv.start()
conn = ist.InfinityConnection(ist.ClientConfig(host_addr="127.0.0.1", service_port=port,
                                               connection_type=ist.TYPE_RDMA, device=0,
                                               device_lookup=True, streams=0))
conn.connect()
tokens, heads, dim = 64, 8, 128          # 128 KB bf16 pages, token-major
elems = tokens * heads * dim
src = torch.randn(n * elems, device="cuda:0").to(torch.bfloat16)
dst = torch.zeros_like(src)
conn.register_mr(src)
conn.register_mr(dst)
keys = [f"ncu-{i}" for i in range(n)]
offs = [i * elems for i in range(n)]
conn.rdma_write_cache(src, offs, elems, conn.allocate_rdma(keys, elems * 2))   # pipe copy + publish
conn.sync()
blocks = list(zip(keys, offs))
conn.read_cache(dst, blocks, elems)                                            # fused pipe read
conn.sync()
assert torch.equal(src, dst)
hnd = torch.zeros((n, heads, tokens, dim), device="cuda:0", dtype=torch.bfloat16)
conn.read_cache_hnd(hnd, [(k, i) for i, k in enumerate(keys)])                 # TMA tensor store
conn.sync()
assert torch.equal(hnd, src.view(n, tokens, heads, dim).permute(0, 2, 1, 3))
d2 = [torch.zeros_like(src) for _ in range(4)]
conn.read_cache_multi(d2, blocks, elems)                                       # 1 fetch, 4 stores
conn.sync()
assert all(torch.equal(d, src) for d in d2)
fkeys = [f"ncu-fp8-{i}" for i in range(n)]
conn.rdma_write_cache_fp8(src, offs, elems, conn.allocate_rdma(fkeys, conn.fp8_page_bytes(elems)))
conn.sync()
conn.read_cache_fp8(dst, list(zip(fkeys, offs)), elems)
conn.sync()
assert conn.get_match_last_index(keys + ["absent"] * 3584) == n - 1             # 4096-key lookup
conn.close()
srv.stop()

if two and nvls_available():
    nb = 128
    bc = PrefixBroadcaster([0, 1], nb * (1 << 20), flag_slots=nb)
    s8 = torch.randint(0, 255, (nb << 20,), dtype=torch.uint8, device="cuda:0")
    o = [i << 20 for i in range(nb)]
    out = torch.zeros(nb << 20, dtype=torch.uint8, device="cuda:1")
    want = bc.expected_flags(list(range(nb)), 1 << 20)
    bc.broadcast(s8, o, o, 1 << 20, flag_ids=list(range(nb)))
    torch.cuda.synchronize(0)
    bc.read_when_ready(1, out, o, o, 1 << 20, list(range(nb)), expect=want)
    torch.cuda.synchronize(1)
    assert torch.equal(out.cpu(), s8.cpu())
print("ncu driver ok")
