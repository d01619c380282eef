"""Decode-side counterpart of demo_prefill.py: find the cached prefix, then pull it layer by
layer into the layout the attention kernel wants while the previous layer computes.

A prefill instance has uploaded token-major pages (``PagedKVCache.write_layer``).  The decode
instance asks the store how much of its prompt is cached (``get_match_last_index`` on the
prefix-chained page hashes, resolved in the HBM index on the GPU) and fetches those pages
into a HEAD-major cache: ``read_cache_hnd`` transposes inside the read (4-D tensor-map TMA
store), so no repack kernel runs here.  Layer l+1's read is enqueued before layer l's
attention stand-in, on the connection's own streams: the fetch overlaps the compute, the
pattern of the reference's design doc (docs/source/design.rst:56-63) on the consumer side.

    python -m infinistore.server --service-port 22345 --prealloc-size 8 &
    python -m infinistore_b200.example.demo_decode --service-port 22345
"""
import argparse
import time

import torch

import infinistore_b200 as infinistore
from infinistore_b200.models import HeadMajorKVCache, PagedKVCache, chain_hashes, get_layout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--service-port", type=int, default=22345)
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--pages", type=int, default=32, help="128-token pages of cached prefix")
    a = ap.parse_args()
    if not torch.cuda.is_available():
        print("demo_decode needs a CUDA device (the layout-swizzling read is a GPU kernel)")
        return
    dev = torch.device("cuda:0")
    layout = get_layout(a.model)
    mk = lambda: infinistore.InfinityConnection(infinistore.ClientConfig(  # noqa: E731
        host_addr="127.0.0.1", service_port=a.service_port,
        connection_type=infinistore.TYPE_RDMA, device_lookup=True))
    prefill_conn, decode_conn = mk(), mk()
    prefill_conn.connect()
    decode_conn.connect()

    # ---- "prefill instance": upload a prompt's pages, token-major
    tokens = list(range(a.pages * layout.page_tokens))
    hashes = chain_hashes(tokens, layout.page_tokens, salt=f"demo-decode-{time.time()}")
    produced = PagedKVCache(layout, num_pages=a.pages, device=dev)
    produced.data.normal_()
    pages = list(range(a.pages))
    for layer in range(layout.layers):
        produced.write_layer(prefill_conn, layer, pages, hashes)
    prefill_conn.sync()

    # ---- "decode instance": longest cached prefix, then layer-wise fetch under compute
    cache = HeadMajorKVCache(layout, num_pages=a.pages, device=dev)
    probe = PagedKVCache(layout, num_pages=1, device=dev)
    hit = probe.cached_prefix_pages(decode_conn, hashes)
    print(f"{hit} of {a.pages} prompt pages are cached")
    q = torch.randn(layout.heads_per_rank, 64, layout.head_dim, device=dev, dtype=layout.dtype)

    def attend(layer):  # stand-in for paged attention over the head-major K pages
        k = cache.data[layer, 0, :hit]                       # [pages, heads, tokens, dim]
        return torch.einsum("hqd,phtd->hqpt", q, k).amax(dim=(2, 3))

    def run(overlap):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cache.read_layer(decode_conn, 0, pages[:hit], hashes[:hit])
        for layer in range(layout.layers):
            decode_conn.sync()                               # layer l's pages have landed
            if overlap and layer + 1 < layout.layers:
                cache.read_layer(decode_conn, layer + 1, pages[:hit], hashes[:hit])
            attend(layer)
            if not overlap and layer + 1 < layout.layers:
                torch.cuda.synchronize()
                cache.read_layer(decode_conn, layer + 1, pages[:hit], hashes[:hit])
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run(True)
    serial = min(run(False) for _ in range(3))
    overlapped = min(run(True) for _ in range(3))
    mb = 2 * layout.layers * hit * layout.page_bytes / 1e6
    ok = all(torch.equal(cache.page_token_major(l, kv, p),
                         produced.page(l, kv, p).view(layout.page_tokens, layout.heads_per_rank,
                                                      layout.head_dim))
             for l in (0, layout.layers - 1) for kv in (0, 1) for p in (0, hit - 1))
    print(f"fetched {mb:.0f} MB into head-major pages: serial {serial * 1e3:.2f} ms, "
          f"overlapped with attention {overlapped * 1e3:.2f} ms; layout check: {'ok' if ok else 'MISMATCH'}")
    prefill_conn.close()
    decode_conn.close()


if __name__ == "__main__":
    main()
