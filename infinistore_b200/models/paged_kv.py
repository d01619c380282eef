"""A paged KV cache whose pages are store blocks, with layer-wise streaming."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .kv_layout import KVLayout, page_key


def layer_keys(layout: KVLayout, tp_rank: int, layer: int, kv: int,
               page_hashes: Sequence[str]) -> List[str]:
    """Store keys of one layer's K (kv = 0) or V (kv = 1) pages."""
    kind = "K" if kv == 0 else "V"
    return [page_key(layout.name, layer, kind, tp_rank, h) for h in page_hashes]


class PagedKVCache:
    """``cache[layer, kv, page]`` is one contiguous page of ``layout.page_elems`` elements.

    The whole cache is one tensor, registered once with the connection (``register_mr``) and
    addressed by element offset, the calling convention of the reference
    (infinistore/lib.py:645-667).

    Layer-wise streaming (docs/source/design.rst:56-63): ``write_layer`` is called right
    after layer l's KV has been produced, on the producing stream; the page mover runs on
    the connection's own streams behind it, so the upload overlaps the next layer's compute.
    """

    def __init__(self, layout: KVLayout, num_pages: int, device, tp_rank: int = 0):
        self.layout = layout
        self.num_pages = num_pages
        self.tp_rank = tp_rank
        self.data = torch.zeros(layout.layers, 2, num_pages, layout.page_elems,
                                dtype=layout.dtype, device=device)
        self._registered = set()

    # ------------------------------------------------------------------ geometry
    def page_offset(self, layer: int, kv: int, page: int) -> int:
        """Element offset of a page from the start of the cache tensor."""
        return ((layer * 2 + kv) * self.num_pages + page) * self.layout.page_elems

    def page(self, layer: int, kv: int, page: int) -> torch.Tensor:
        return self.data[layer, kv, page]

    def keys(self, layer: int, kv: int, page_hashes: Sequence[str]) -> List[str]:
        return layer_keys(self.layout, self.tp_rank, layer, kv, page_hashes)

    def _ensure_registered(self, conn):
        if id(conn) not in self._registered:
            conn.register_mr(self.data)
            self._registered.add(id(conn))

    # ------------------------------------------------------------------ store I/O
    def write_layer(self, conn, layer: int, pages: Sequence[int], page_hashes: Sequence[str],
                    fp8: bool = False, stream="current") -> int:
        """Upload K and V pages of one layer.  Returns the number of blocks written."""
        self._ensure_registered(conn)
        elems = self.layout.page_elems
        nbytes = conn.fp8_page_bytes(elems) if fp8 else self.layout.page_bytes
        n = 0
        for kv in (0, 1):
            keys = self.keys(layer, kv, page_hashes)
            blocks = conn.allocate_rdma(keys, nbytes)
            offs = np.asarray([self.page_offset(layer, kv, p) for p in pages], dtype=np.int64)
            if fp8:
                conn.rdma_write_cache_fp8(self.data, offs, elems, blocks, stream=stream)
            else:
                conn.rdma_write_cache(self.data, offs, elems, blocks, stream=stream)
            n += len(keys)
        return n

    def read_layer(self, conn, layer: int, pages: Sequence[int], page_hashes: Sequence[str],
                   fp8: bool = False, stream="current") -> int:
        self._ensure_registered(conn)
        elems = self.layout.page_elems
        n = 0
        for kv in (0, 1):
            keys = self.keys(layer, kv, page_hashes)
            blocks = [(k, self.page_offset(layer, kv, p)) for k, p in zip(keys, pages)]
            if fp8:
                conn.read_cache_fp8(self.data, blocks, elems, stream=stream)
            else:
                conn.read_cache(self.data, blocks, elems, stream=stream)
            n += len(keys)
        return n

    def cached_prefix_pages(self, conn, page_hashes: Sequence[str],
                            layer: Optional[int] = None) -> int:
        """How many leading pages of this prefix the store already holds (last layer's V
        pages are written last, so probing them answers for the whole stack)."""
        if not page_hashes:
            return 0
        keys = self.keys(self.layout.layers - 1 if layer is None else layer, 1, page_hashes)
        try:
            return conn.get_match_last_index(keys) + 1
        except Exception:
            return 0

    def touch_prefix(self, conn, page_hashes: Sequence[str]) -> int:
        """Tell an evicting store (``--evict``) that these pages were just used: every layer's
        K and V block of every page becomes most-recently-used.  Needed by clients that read
        through the device index, whose reads the server never sees.  Returns the number of
        blocks refreshed."""
        keys: List[str] = []
        for layer in range(self.layout.layers):
            for kv in (0, 1):
                keys.extend(self.keys(layer, kv, page_hashes))
        return conn.touch(keys) if keys else 0


def read_layer_multi(conn, caches: Sequence[PagedKVCache], layer: int, pages: Sequence[int],
                     page_hashes: Sequence[str], stream="current") -> int:
    """The same prefix pages into several caches of ONE GPU (beams, or tensor-parallel
    consumers that replicate KV heads): every page crosses the fabric once and is stored to
    all destinations by the read kernel (``InfinityConnection.read_cache_multi``: fan-out
    stores for a pool behind NVLink, a thread-block cluster with TMA multicast for a local
    pool).  The caches share layout, page numbering and TP rank.  Returns the number of
    blocks read per cache."""
    first = caches[0]
    for c in caches:
        if c.layout != first.layout or c.num_pages != first.num_pages or c.tp_rank != first.tp_rank:
            raise ValueError("read_layer_multi: caches of one layout, size and TP rank")
    n = 0
    for kv in (0, 1):
        keys = first.keys(layer, kv, page_hashes)
        blocks = [(k, first.page_offset(layer, kv, p)) for k, p in zip(keys, pages)]
        conn.read_cache_multi([c.data for c in caches], blocks, first.layout.page_elems,
                              stream=stream)
        n += len(keys)
    return n


class HeadMajorKVCache:
    """Decode-side cache in the layout a paged-attention kernel streams:
    ``data[layer, kv]`` is ``[num_pages, heads, page_tokens, head_dim]`` - one head's tokens
    of a page are contiguous.  Prefill writes pages token-major (``[tokens, heads, dim]``,
    :class:`PagedKVCache`); ``read_layer`` fetches them with
    ``InfinityConnection.read_cache_hnd``, whose store is a 4-D tensor-map TMA: the
    transposition happens inside the read, the page crosses HBM once and no permute kernel
    runs on the decode GPU.  The reference hands opaque bytes to the consumer
    (infinistore/lib.py:377-379), which then repacks them itself."""

    def __init__(self, layout: KVLayout, num_pages: int, device, tp_rank: int = 0):
        self.layout = layout
        self.num_pages = num_pages
        self.tp_rank = tp_rank
        self.data = torch.zeros(layout.layers, 2, num_pages, layout.heads_per_rank,
                                layout.page_tokens, layout.head_dim, dtype=layout.dtype,
                                device=device)

    def keys(self, layer: int, kv: int, page_hashes: Sequence[str]) -> List[str]:
        return layer_keys(self.layout, self.tp_rank, layer, kv, page_hashes)

    def page_token_major(self, layer: int, kv: int, page: int) -> torch.Tensor:
        """A page as prefill laid it out, ``[tokens, heads, dim]`` (a permuted view)."""
        return self.data[layer, kv, page].permute(1, 0, 2)

    def read_layer(self, conn, layer: int, pages: Sequence[int], page_hashes: Sequence[str],
                   stream="current") -> int:
        """Fetch K and V pages of one layer into head-major pages ``pages``."""
        n = 0
        for kv in (0, 1):
            keys = self.keys(layer, kv, page_hashes)
            conn.read_cache_hnd(self.data[layer, kv], list(zip(keys, [int(p) for p in pages])),
                                stream=stream)
            n += len(keys)
        return n
