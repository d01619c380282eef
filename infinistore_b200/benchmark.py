"""Throughput benchmark with the reference's CLI and methodology.

Mirrors infinistore/benchmark.py:12-215: ``--size`` MB of fp32 split into ``--block-size``
KB blocks, fresh UUID keys per iteration, allocation outside the timed region, writes then
reads issued in ``--steps`` batches ("layers"), one ``sync()`` after each phase, host clock
around the loops, ``size * iterations / sum(elapsed)`` MB/s, and a final ``torch.equal``.
Extras: ``--device-lookup``, ``--variant``, ``--json`` and a CPU mode (``--cpu``) so the
plumbing config (16 keys x 4 KB on loop-back) runs without a GPU.

    python -m infinistore_b200.benchmark --service-port 22345 --size 128 --block-size 32 --rdma
"""
import argparse
import json
import time
import uuid

import torch

from . import ClientConfig, InfinityConnection, TYPE_LOCAL_GPU, TYPE_RDMA


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--rdma", action="store_true", help="use the fabric (RDMA-type) connection")
    p.add_argument("--server", default="127.0.0.1", type=str)
    p.add_argument("--service-port", type=int, default=22345)
    p.add_argument("--dev-name", default="mlx5_1", type=str, help="ignored")
    p.add_argument("--iteration", type=int, default=1)
    p.add_argument("--block-size", type=int, default=32, help="KB")
    p.add_argument("--size", type=int, default=128, help="MB")
    p.add_argument("--src-gpu", type=int, default=0)
    p.add_argument("--dst-gpu", type=int, default=1)
    p.add_argument("--ib-port", type=int, default=1, help="ignored")
    p.add_argument("--link-type", default="IB", type=str, help="ignored")
    p.add_argument("--steps", type=int, default=32, help="batches per phase (layers)")
    p.add_argument("--cpu", action="store_true", help="CPU tensors (needs --rdma)")
    p.add_argument("--device-lookup", action="store_true", help="resolve keys in the HBM index")
    p.add_argument("--variant", default="auto", choices=["auto", "ldst", "tma", "ldst256"])
    p.add_argument("--posted-commit", action="store_true", help="one-way commit in sync()")
    p.add_argument("--doorbell", action="store_true",
                   help="latency mode: single blocks through the persistent worker CTA")
    p.add_argument("--json", action="store_true", help="print one JSON line with the result")
    return p.parse_args(argv)


def run(args):
    config = ClientConfig(
        host_addr=args.server,
        service_port=args.service_port,
        dev_name=args.dev_name,
        ib_port=args.ib_port,
        link_type=args.link_type,
        log_level="warning",
        device_lookup=args.device_lookup,
        posted_commit=args.posted_commit,
        doorbell=args.doorbell,
        copy_variant=args.variant,
    )
    config.connection_type = TYPE_RDMA if args.rdma else TYPE_LOCAL_GPU
    conn = InfinityConnection(config)
    conn.connect()

    if args.cpu:
        src_device = dst_device = "cpu"
    else:
        ngpu = torch.cuda.device_count()
        src_device = f"cuda:{min(args.src_gpu, ngpu - 1)}"
        dst_device = f"cuda:{min(args.dst_gpu, ngpu - 1)}"

    block_size = args.block_size * 1024 // 4  # fp32 elements
    num_of_blocks = args.size * 1024 * 1024 // (args.block_size * 1024)
    src_tensor = torch.rand(num_of_blocks * block_size, device=src_device, dtype=torch.float32)
    dst_tensor = torch.rand(num_of_blocks * block_size, device=dst_device, dtype=torch.float32)
    if not args.cpu:
        torch.cuda.synchronize(src_tensor.device)
        torch.cuda.synchronize(dst_tensor.device)
    if args.rdma:
        conn.register_mr(src_tensor)
        conn.register_mr(dst_tensor)

    write_sum = 0.0
    read_sum = 0.0
    for _ in range(args.iteration):
        keys = [str(uuid.uuid4()) for _ in range(num_of_blocks)]
        offset_blocks = [i * block_size for i in range(num_of_blocks)]
        blocks = list(zip(keys, offset_blocks))
        if args.rdma:
            remote_addrs = conn.allocate_rdma(keys, block_size * 4)
        steps = args.steps
        while len(blocks) % steps != 0 and steps > 1:
            steps = int(steps / 2)
        n = len(blocks) // steps

        start = time.time()
        for i in range(steps):
            if args.rdma:
                conn.rdma_write_cache(src_tensor, offset_blocks[i * n: i * n + n], block_size,
                                      remote_addrs[i * n: i * n + n])
            else:
                conn.local_gpu_write_cache(src_tensor, blocks[i * n: i * n + n], block_size)
        conn.sync()
        mid = time.time()
        write_sum += mid - start
        for i in range(steps):
            conn.read_cache(dst_tensor, blocks[i * n: i * n + n], block_size)
        conn.sync()
        read_sum += time.time() - mid

    total_mb = args.size * args.iteration
    result = {
        "size_mb": total_mb,
        "block_kb": args.block_size,
        "connection": config.connection_type,
        "write_mb_s": total_mb / write_sum,
        "read_mb_s": total_mb / read_sum,
        "kernel_launches": conn.stats()["kernel_launches"],
    }
    if args.json:
        print(json.dumps(result))
    else:
        print("size: {} MB, block size: {} KB, connection type: {}".format(
            total_mb, args.block_size, config.connection_type))
        print("write cache: {:.2f} MB/s, read cache: {:.2f} MB/s".format(
            result["write_mb_s"], result["read_mb_s"]))
    assert torch.equal(src_tensor.cpu(), dst_tensor.cpu())
    conn.close()
    return result


def main(argv=None):
    run(parse_args(argv))


if __name__ == "__main__":
    main()
