"""Warm-up: one write+read round trip per visible GPU through LOCAL_GPU.

Pays CUDA context creation, pool mapping (IPC open / peer enable) and kernel module load
up front, and doubles as a smoke test (reference: infinistore/warmup.py:7-72).
"""
import argparse
import time

import torch

from . import ClientConfig, InfinityConnection, Logger, TYPE_LOCAL_GPU


def warm_up(args):
    num_devices = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if num_devices == 0:
        Logger.warn("warm up skipped: no CUDA device")
        return 0
    src, dst = [], []
    for i in range(num_devices):
        Logger.info(f"detect device GPU:{i} ...")
        src.append(torch.randn(4096, device=f"cuda:{i}", dtype=torch.float32))
        dst.append(torch.zeros(4096, device=f"cuda:{i}", dtype=torch.float32))
        torch.cuda.synchronize(i)

    config = ClientConfig(
        host_addr="127.0.0.1",
        service_port=args.service_port,
        connection_type=TYPE_LOCAL_GPU,
        log_level="error",
    )
    time.sleep(args.start_delay)  # let the server come up when launched from server.py
    conn = InfinityConnection(config)
    conn.connect()
    tag = int(time.time() * 1000)
    for i in range(num_devices):
        Logger.info(f"warming up device GPU:{i}")
        key = f"warmup_{tag}_{i}"
        conn.local_gpu_write_cache(src[i], [(key, 0)], 4096)
        conn.sync()
        conn.read_cache(dst[i], [(key, 0)], 4096)
        conn.sync()
        assert torch.allclose(src[i], dst[i]), f"device {i} data mismatch"
    conn.close()
    Logger.info("warm up finished")
    return num_devices


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--service-port", required=True, type=int, help="which port to connect to")
    parser.add_argument("--start-delay", required=False, type=int, default=0,
                        help="delay before starting warm up, used in script")
    warm_up(parser.parse_args(argv))


if __name__ == "__main__":
    main()
