"""One writer -> every GPU replication of KV blocks through NVLS multicast.

Use case: a shared prompt prefix that all decode GPUs will read.  The reference serves it
with N independent unicast reads of the same blocks (src/infinistore.cpp:424-533); here the
writer stores each 16-byte vector once to a multicast address and the NVSwitch replicates
it into a replica region on every GPU, which the readers then consume at local-HBM speed.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from .. import _infinistore
from ..ops import make_descs, _stream


def nvls_available(device: int = 0) -> bool:
    if not _infinistore.cuda_available():
        return False
    return bool(_infinistore.nvls_probe(device).multicast_supported)


class _RawCuda:
    """Expose raw device memory to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1",
                                         "data": (ptr, False), "version": 2}


class PrefixBroadcaster:
    """Replica region of `bytes_per_gpu` on each of `devices`, bound to one multicast object.

    ``broadcast(src, src_offsets, slot_offsets, nbytes)`` copies pages of a tensor on the
    writer GPU into the same byte offsets of every replica with a single kernel;
    ``replica(i)`` is a uint8 torch view of GPU devices[i]'s local copy.
    (Single-process flavour: this process drives all listed GPUs, like the store server
    that owns pool segments on several devices.)
    """

    def __init__(self, devices: Sequence[int], bytes_per_gpu: int):
        self.devices = list(devices)
        self.group = _infinistore.NvlsGroup.create(self.devices, bytes_per_gpu)
        self.bytes = self.group.bytes()
        self._keep = []

    def replica(self, i: int) -> torch.Tensor:
        raw = _RawCuda(self.group.uc_ptr(i), self.bytes)
        self._keep.append(raw)
        return torch.as_tensor(raw, device=torch.device("cuda", self.devices[i]))

    def broadcast(self, src: torch.Tensor, src_offsets: Sequence[int], slot_offsets: Sequence[int],
                  nbytes: int, writer: int = 0, max_ctas: int = 0) -> None:
        """Offsets are in bytes.  Launches on the current stream of the writer device."""
        dev = torch.device("cuda", self.devices[writer])
        assert src.device == dev and nbytes % 16 == 0
        mc = self.group.mc_ptr(writer)
        descs = make_descs([src.data_ptr() + o for o in src_offsets],
                           [mc + o for o in slot_offsets], dev)
        with torch.cuda.device(dev):
            _infinistore.kernels.kv_bcast_nvls(descs.data_ptr(), descs.shape[0], nbytes, max_ctas,
                                               _stream(dev))
        self._last = descs  # keep alive until the kernel ran
