"""A paged KV cache whose pages are store blocks, with layer-wise streaming."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .kv_layout import KVLayout, page_key


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
        kind = "K" if kv == 0 else "V"
        return [page_key(self.layout.name, layer, kind, self.tp_rank, h) for h in page_hashes]

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

    def cached_prefix_pages(self, conn, page_hashes: Sequence[str], layer: int = 0) -> int:
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
