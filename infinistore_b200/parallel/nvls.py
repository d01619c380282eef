"""One writer -> every GPU replication of KV blocks through NVLS multicast.

Use case: a shared prompt prefix that all decode GPUs will read.  The reference serves it
with N independent unicast reads of the same blocks (src/infinistore.cpp:424-533); here the
writer stores each 16-byte vector once to a multicast address and the NVSwitch replicates
it into a replica region on every GPU, which the readers then consume at local-HBM speed.
"""
from __future__ import annotations

from typing import List, Sequence

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

    In-band readiness (``flag_slots > 0``): the region ends with one u32 flag per slot, also
    replicated.  ``broadcast(..., flag_ids=[...])`` makes the writer kernel bump a block's
    flag in every replica (``multimem.red.release.sys``) as its chunks land, and
    ``read_when_ready(i, dst, ...)`` launches a reader kernel on GPU devices[i] that waits for
    each block's flag in ITS OWN replica (``ld.acquire.sys``) before copying the block out -
    the reader can be launched before the writer, nothing synchronises on the host and
    nothing crosses the fabric while it waits.  (Readers that share the writer's GPU must be
    launched after the broadcast on the same stream: a spinning grid would starve it.)
    (Single-process flavour: this process drives all listed GPUs, like the store server
    that owns pool segments on several devices.)
    """

    def __init__(self, devices: Sequence[int], bytes_per_gpu: int, flag_slots: int = 0):
        self.devices = list(devices)
        self.data_bytes = (bytes_per_gpu + 255) // 256 * 256
        self.flag_slots = int(flag_slots)
        self.group = _infinistore.NvlsGroup.create(self.devices,
                                                   self.data_bytes + 4 * self.flag_slots)
        self.bytes = self.group.bytes()
        self._keep = []
        self._flag_epoch = [0] * self.flag_slots  # value a slot's flag has after its last broadcast
        if self.flag_slots:
            for i in range(len(self.devices)):
                self.replica(i)[self.data_bytes:self.data_bytes + 4 * self.flag_slots].zero_()
                torch.cuda.synchronize(self.devices[i])

    def replica(self, i: int) -> torch.Tensor:
        raw = _RawCuda(self.group.uc_ptr(i), self.bytes)
        self._keep.append(raw)
        return torch.as_tensor(raw, device=torch.device("cuda", self.devices[i]))

    def broadcast(self, src: torch.Tensor, src_offsets: Sequence[int], slot_offsets: Sequence[int],
                  nbytes: int, writer: int = 0, max_ctas: int = 0,
                  flag_ids: Sequence[int] = None) -> None:
        """Offsets are in bytes.  Launches on the current stream of the writer device.
        ``flag_ids[i]`` (optional): flag slot that announces block i in every replica."""
        dev = torch.device("cuda", self.devices[writer])
        assert src.device == dev and nbytes % 16 == 0
        mc = self.group.mc_ptr(writer)
        flags_mc = 0
        if flag_ids is not None:
            consecutive = list(range(flag_ids[0], flag_ids[0] + len(flag_ids)))
            assert self.flag_slots and list(flag_ids) == consecutive, \
                "flag ids: a contiguous run of slots, one per block"
            flags_mc = mc + self.data_bytes + 4 * flag_ids[0]
            per = _infinistore.kernels.bcast_chunks_per_block(nbytes)
            for f in flag_ids:
                self._flag_epoch[f] += per
        descs = make_descs([src.data_ptr() + o for o in src_offsets],
                           [mc + o for o in slot_offsets], dev)
        with torch.cuda.device(dev):
            _infinistore.kernels.kv_bcast_nvls(descs.data_ptr(), descs.shape[0], nbytes, max_ctas,
                                               _stream(dev), flags_mc)
        self._last = descs  # keep alive until the kernel ran

    def read_when_ready(self, reader: int, dst: torch.Tensor, slot_offsets: Sequence[int],
                        dst_offsets: Sequence[int], nbytes: int, flag_ids: Sequence[int],
                        expect: Sequence[int] = None, status: torch.Tensor = None) -> None:
        """On GPU devices[reader]: wait for each block's flag in the LOCAL replica, then copy the
        block from the local replica into ``dst`` (byte offsets).  ``expect``: flag values that
        mean "complete" (default: what the broadcasts announced so far for these slots - call
        ``expected_flags`` BEFORE launching the writer to overlap the two)."""
        dev = torch.device("cuda", self.devices[reader])
        assert dst.device == dev and nbytes % 16 == 0 and self.flag_slots
        assert list(flag_ids) == list(range(flag_ids[0], flag_ids[0] + len(flag_ids)))
        want = list(expect) if expect is not None else [self._flag_epoch[f] for f in flag_ids]
        assert len(set(want)) == 1, "one ready value per launch"
        uc = self.group.uc_ptr(reader)
        descs = make_descs([uc + o for o in slot_offsets], [dst.data_ptr() + o for o in dst_offsets], dev)
        with torch.cuda.device(dev):
            _infinistore.kernels.kv_read_when_ready(
                descs.data_ptr(), descs.shape[0], nbytes, uc + self.data_bytes + 4 * flag_ids[0],
                want[0], 0, _stream(dev), status.data_ptr() if status is not None else 0)
        self._keep.append(descs)

    def expected_flags(self, flag_ids: Sequence[int], nbytes: int) -> List[int]:
        """Flag values the NEXT broadcast of these slots will produce."""
        per = _infinistore.kernels.bcast_chunks_per_block(nbytes)
        return [self._flag_epoch[f] + per for f in flag_ids]
