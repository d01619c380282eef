"""Layer-by-layer KV upload overlapped with prefill compute (the usage pattern of the
reference's design doc, docs/source/design.rst:56-63, and example/demo_prefill.py).

The reference needs a CUDA event per layer, an upload thread and `local_gpu_write_cache`
from that thread.  Here `write_layer` is simply called right after layer l's KV exists: the
page mover is enqueued behind the producing kernels (it waits for the caller's stream) and
runs on the connection's own streams, so layer l+1's compute overlaps layer l's upload with
no extra thread.  Reports the prefill time with and without the upload.
"""
import argparse
import time

import torch

import infinistore_b200 as infinistore
from infinistore_b200.models import PagedKVCache, chain_hashes, get_layout


def prefill(layout, cache, x, weights, pages, conn=None, hashes=None):
    """Toy transformer stack: per layer one big GEMM producing K/V pages."""
    elems = layout.page_elems
    for layer in range(layout.layers):
        h = torch.relu(x @ weights[layer])
        kv = h[:, : 2 * elems * len(pages) // x.shape[0]].reshape(2, len(pages), elems)
        for p_i, p in enumerate(pages):
            cache.data[layer, 0, p].copy_(kv[0, p_i])
            cache.data[layer, 1, p].copy_(kv[1, p_i])
        if conn is not None:
            cache.write_layer(conn, layer, pages, hashes)
        x = h[:, : x.shape[1]]
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--service-port", type=int, default=22345)
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--pages", type=int, default=16, help="128-token pages in the prompt")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    layout = get_layout(a.model)
    cache = PagedKVCache(layout, num_pages=a.pages, device=dev)
    pages = list(range(a.pages))
    rows = 2048
    need = 2 * layout.page_elems * a.pages // rows
    width = max(4096, need)
    x = torch.randn(rows, 4096, device=dev, dtype=torch.bfloat16)
    weights = [torch.randn(4096, width, device=dev, dtype=torch.bfloat16) * 0.02
               for _ in range(layout.layers)]
    conn = infinistore.InfinityConnection(infinistore.ClientConfig(
        host_addr="127.0.0.1", service_port=a.service_port,
        connection_type=infinistore.TYPE_RDMA, device_lookup=True))
    conn.connect()

    def timed(with_upload, tag):
        hashes = chain_hashes(list(range(a.pages * 128)), 128, salt=f"{tag}-{time.time()}")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prefill(layout, cache, x, weights, pages, conn if with_upload else None, hashes)
        torch.cuda.synchronize()
        if with_upload:
            conn.sync()
        return time.perf_counter() - t0

    timed(False, "warm")
    timed(True, "warm")
    base = min(timed(False, "b") for _ in range(3))
    up = min(timed(True, f"u{i}") for i in range(3))
    mb = 2 * layout.layers * a.pages * layout.page_bytes / 1e6
    print(f"prefill {base * 1e3:.2f} ms; with layer-wise upload of {mb:.0f} MB: {up * 1e3:.2f} ms "
          f"(+{100 * (up - base) / base:.1f}%)")
    conn.close()


if __name__ == "__main__":
    main()
