"""asyncio API: connect_async / allocate_rdma_async / rdma_write_cache_async /
read_cache_async (counterpart of the reference's example/client_async.py)."""
import argparse
import asyncio
import uuid

import torch

import infinistore_b200 as infinistore


async def main_async(a):
    cfg = infinistore.ClientConfig(host_addr=a.server, service_port=a.service_port,
                                   connection_type=infinistore.TYPE_RDMA)
    conn = infinistore.InfinityConnection(cfg)
    await conn.connect_async()
    device = "cuda:0" if torch.cuda.is_available() else "cpu"
    src = torch.randn(4096, device=device)
    dst = torch.zeros(4096, device=device)
    await asyncio.to_thread(lambda: (conn.register_mr(src), conn.register_mr(dst)))
    for it in range(a.iterations):
        keys = [str(uuid.uuid4()) for _ in range(4)]
        remote = await conn.allocate_rdma_async(keys, 1024 * 4)
        # two writes in flight at once
        await asyncio.gather(
            conn.rdma_write_cache_async(src, [0, 1024], 1024, remote[:2]),
            conn.rdma_write_cache_async(src, [2048, 3072], 1024, remote[2:]))
        await conn.read_cache_async(dst, [(k, i * 1024) for i, k in enumerate(keys)], 1024)
        assert torch.equal(src, dst)
        print(f"iteration {it}: ok")
    conn.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--server", default="127.0.0.1")
    ap.add_argument("--service-port", type=int, default=22345)
    ap.add_argument("--iterations", type=int, default=3)
    asyncio.run(main_async(ap.parse_args()))


if __name__ == "__main__":
    main()
