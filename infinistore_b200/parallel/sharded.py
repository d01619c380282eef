"""Client of several pool shards (one store server per shard)."""
from __future__ import annotations

import zlib
from typing import List, Sequence, Tuple

import numpy as np
import torch

from ..lib import ClientConfig, InfinityConnection


def shard_of(key: str, nshards: int) -> int:
    """Stable key -> shard routing (crc32 of the UTF-8 key)."""
    return zlib.crc32(key.encode()) % nshards


class ShardedBlocks:
    """Result of ``ShardedConnection.allocate_rdma``: per-shard block arrays plus, for
    every shard, the positions of its keys in the caller's key list."""

    def __init__(self, positions: List[np.ndarray], blocks: List[np.ndarray], n: int):
        self.positions = positions
        self.blocks = blocks
        self.n = n

    def __len__(self):
        return self.n


class ShardedConnection:
    """Same call surface as ``InfinityConnection`` over N servers.

    A key lives on shard ``crc32(key) % N``.  Batch calls are split per shard and issued
    back to back (the kernels of different shards run concurrently on the client GPU);
    ``sync()`` waits for all shards.
    """

    def __init__(self, configs: Sequence[ClientConfig]):
        if not configs:
            raise ValueError("at least one shard is required")
        self.conns = [InfinityConnection(c) for c in configs]

    @property
    def nshards(self) -> int:
        return len(self.conns)

    def connect(self):
        for c in self.conns:
            c.connect()

    def close(self):
        for c in self.conns:
            c.close()

    def register_mr(self, cache: torch.Tensor):
        return [c.register_mr(cache) for c in self.conns][0]

    def _split(self, keys: Sequence[str]):
        n = self.nshards
        ids = np.fromiter((shard_of(k, n) for k in keys), dtype=np.int64, count=len(keys))
        return [np.nonzero(ids == s)[0] for s in range(n)]

    def allocate_rdma(self, keys: Sequence[str], page_size_in_bytes: int) -> ShardedBlocks:
        positions = self._split(keys)
        blocks = []
        for s, pos in enumerate(positions):
            if len(pos) == 0:
                blocks.append(np.empty(0))
                continue
            blocks.append(self.conns[s].allocate_rdma([keys[i] for i in pos], page_size_in_bytes))
        return ShardedBlocks(positions, blocks, len(keys))

    def rdma_write_cache(self, cache: torch.Tensor, offsets, page_size: int,
                         blocks: ShardedBlocks, stream="current"):
        offs = np.asarray(offsets, dtype=np.int64)
        for s, pos in enumerate(blocks.positions):
            if len(pos):
                self.conns[s].rdma_write_cache(cache, offs[pos], page_size, blocks.blocks[s],
                                               stream=stream)
        return 0

    def read_cache(self, cache: torch.Tensor, blocks: List[Tuple[str, int]], page_size: int,
                   stream="current"):
        n = self.nshards
        per = [[] for _ in range(n)]
        for kb in blocks:
            per[shard_of(kb[0], n)].append(kb)
        for s, lst in enumerate(per):
            if lst:
                self.conns[s].read_cache(cache, lst, page_size, stream=stream)

    rdma_read_cache = read_cache

    def sync(self):
        for c in self.conns:
            c.sync()

    def check_exist(self, key: str) -> bool:
        return self.conns[shard_of(key, self.nshards)].check_exist(key)

    def touch(self, keys: Sequence[str]) -> int:
        """Recency hint (see ``InfinityConnection.touch``), routed to the owning shards."""
        n = 0
        for s, pos in enumerate(self._split(keys)):
            if len(pos):
                n += self.conns[s].touch([keys[i] for i in pos])
        return n

    def get_match_last_index(self, keys: List[str]) -> int:
        """The reference's binary search (src/infinistore.cpp:1092-1108) with each probe
        answered by the shard owning that key."""
        left, right = 0, len(keys)
        n = self.nshards
        while left < right:
            mid = left + (right - left) // 2
            conn = self.conns[shard_of(keys[mid], n)]
            try:
                present = conn.get_match_last_index([keys[mid]]) == 0
            except Exception:
                present = False
            if present:
                left = mid + 1
            else:
                right = mid
        if left - 1 < 0:
            raise Exception("can't find a match")
        return left - 1

    def stats(self):
        out = {}
        for c in self.conns:
            for k, v in c.stats().items():
                out[k] = out.get(k, 0) + v
        return out
