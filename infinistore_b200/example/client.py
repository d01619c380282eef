"""Blocking API walk-through: every source/destination placement, both connection types
(counterpart of the reference's example/client.py).

    python -m infinistore_b200.server --service-port 22345 &
    python -m infinistore_b200.example.client --service-port 22345
"""
import argparse
import time
import uuid

import torch

import infinistore_b200 as infinistore


def run(conn, src_device, dst_device, local=False):
    n, page = 16, 4096
    src = torch.randn(n * page, device=src_device, dtype=torch.float32)
    dst = torch.zeros(n * page, device=dst_device, dtype=torch.float32)
    keys = [str(uuid.uuid4()) for _ in range(n)]
    blocks = [(k, i * page) for i, k in enumerate(keys)]
    t0 = time.time()
    if local:
        conn.local_gpu_write_cache(src, blocks, page)
    else:
        conn.register_mr(src)
        conn.register_mr(dst)
        remote = conn.allocate_rdma(keys, page * 4)       # bytes
        conn.rdma_write_cache(src, [i * page for i in range(n)], page, remote)  # elements
    conn.sync()
    t1 = time.time()
    conn.read_cache(dst, blocks, page)
    conn.sync()
    t2 = time.time()
    assert torch.equal(src.cpu(), dst.cpu())
    print(f"{'local' if local else 'fabric'} {src_device}->{dst_device}: write {1e3 * (t1 - t0):.2f} ms, "
          f"read {1e3 * (t2 - t1):.2f} ms, prefix match {conn.get_match_last_index(keys)}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--server", default="127.0.0.1")
    ap.add_argument("--service-port", type=int, default=22345)
    a = ap.parse_args()
    cfg = infinistore.ClientConfig(host_addr=a.server, service_port=a.service_port,
                                   connection_type=infinistore.TYPE_RDMA)
    conn = infinistore.InfinityConnection(cfg)
    conn.connect()
    devices = ["cpu"] + ([f"cuda:{i}" for i in range(min(torch.cuda.device_count(), 2))]
                         if torch.cuda.is_available() else [])
    for s in devices:
        for d in devices:
            run(conn, s, d)
    conn.close()
    if torch.cuda.is_available():
        cfg.connection_type = infinistore.TYPE_LOCAL_GPU
        conn = infinistore.InfinityConnection(cfg)
        conn.connect()
        run(conn, "cuda:0", "cuda:0", local=True)
        conn.close()


if __name__ == "__main__":
    main()
