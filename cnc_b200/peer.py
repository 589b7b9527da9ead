"""Buffers every rank of the box can load from and store to (CUDA IPC over NVLink / NVSwitch), and the three kernels the
data-parallel step runs on them (csrc/peer.cu): barrier, reduce of the owned rows straight out of the peers' gradient
buffers, push of the owned bit-plane words into the peers' planes.  `dp.ShardedTableAdam(exchange="peer")` is the user.

The control plane (who maps what) is torch.distributed: one `all_gather_object` of the 64-byte handles per allocation, at
construction time.  Nothing here touches the data path of a single-GPU run.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.distributed as dist

from ._lib import check, lib, stream


def _ok(rc: int) -> None:
    """result of a call that launches nothing (allocation, mapping)"""
    if rc != 0:
        raise RuntimeError(f"cnc_b200 [{rc}]: {lib().cnc_last_error().decode()}")


_TYPESTR = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4", torch.uint32: "<u4"}


class _Raw:
    """device memory that torch did not allocate, as seen through __cuda_array_interface__"""

    def __init__(self, address: int, numel: int, dtype: torch.dtype, owner):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": _TYPESTR[dtype], "data": (address, False), "version": 2}
        self._owner = owner   # the mapping outlives every tensor cut from it


class PeerMemory:
    """`nbytes` of zeroed device memory on every rank of `group`, each mapped into all of them.

    `ptrs[k]` is rank k's buffer as THIS process addresses it (k == rank: the own allocation).  Collective: every rank
    constructs it at the same point.  `control_group` (optional) carries the handles when `group` cannot move python
    objects between these processes (tests that put two ranks on one GPU use a gloo group for it)."""

    def __init__(self, nbytes: int, group=None, device=None, control_group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        L = lib()
        with torch.cuda.device(self.device):
            own = C.c_void_p()
            _ok(L.cnc_peer_alloc(self.nbytes, C.byref(own)))
            self._own = own.value
            hb = L.cnc_peer_handle_bytes()
            buf = (C.c_uint8 * hb)()
            _ok(L.cnc_peer_export(self._own, buf))
            handles: List[bytes] = [None] * self.world
            dist.all_gather_object(handles, (bytes(buf), self.nbytes), group=control_group if control_group is not None else group)
            self.ptrs: List[int] = []
            self._imported: List[int] = []
            for k, (h, nb) in enumerate(handles):
                if nb != self.nbytes:
                    raise RuntimeError(f"PeerMemory: rank {k} allocated {nb} bytes, this rank {self.nbytes}")
                if k == self.rank:
                    self.ptrs.append(self._own)
                    continue
                p = C.c_void_p()
                _ok(L.cnc_peer_import((C.c_uint8 * hb).from_buffer_copy(h), C.byref(p)))
                self.ptrs.append(p.value)
                self._imported.append(p.value)
        self._array = (C.c_void_p * self.world)(*self.ptrs)

    def tensor(self, offset_bytes: int, numel: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """a torch view of the own buffer (no copy)"""
        size = numel * torch.empty(0, dtype=dtype).element_size()
        if offset_bytes < 0 or offset_bytes + size > self.nbytes:
            raise ValueError("PeerMemory.tensor: out of range")
        if numel == 0:
            return torch.empty(0, dtype=dtype, device=self.device)
        return torch.as_tensor(_Raw(self._own + offset_bytes, numel, dtype, self), device=self.device)

    def pointer_array(self, offset_bytes: int = 0):
        """ctypes array of the world's pointers, each advanced by `offset_bytes` (argument of the peer kernels)"""
        if offset_bytes == 0:
            return self._array
        return (C.c_void_p * self.world)(*[p + offset_bytes for p in self.ptrs])

    def close(self) -> None:
        """unmap / free (every tensor cut from the buffer must be dead); synchronises the device first"""
        if self._own is None:
            return
        L = lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._imported:
                L.cnc_peer_unmap(p)
            L.cnc_peer_free(self._own)
        self._own, self._imported, self.ptrs = None, [], []


class PeerSignals:
    """the signal pads of the barrier kernel: `barrier(slot)` = every rank's stream reaches this point before any goes on"""

    def __init__(self, group=None, device=None, control_group=None, timeout_ms: int = 20000):
        self.mem = PeerMemory(lib().cnc_peer_pad_bytes(), group=group, device=device, control_group=control_group)
        self.epochs = {}
        self.timeout_ms = timeout_ms

    def barrier(self, slot: int) -> None:
        e = self.epochs.get(slot, 0) + 1
        self.epochs[slot] = e
        m = self.mem
        check(lib().cnc_peer_barrier(m.pointer_array(), m.rank, m.world, slot, e & 0xFFFFFFFF, self.timeout_ms, stream()))


    def minimum(self, slot: int, value: int, out: torch.Tensor) -> None:
        """out[0] (uint32 / int32 device word) = min over the ranks of `value` (0 <= value < 2^32); uses pad slots slot..slot+2"""
        e = self.epochs.get(slot, 0) + 1
        self.epochs[slot] = e
        m = self.mem
        check(lib().cnc_peer_min(m.pointer_array(), m.rank, m.world, slot, e & 0xFFFFFFFF, min(int(value), 0xFFFFFFFF), out.data_ptr(),
                                 self.timeout_ms, stream()))


def reduce_rows(srcs, world: int, lo: int, count: int, scale: float, out: torch.Tensor, blocks: int = 0) -> None:
    """out[:count] = scale * sum over the ranks of their buffer[lo : lo + count] (fp32, rank order); `srcs` = pointer array"""
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= count
    check(lib().cnc_peer_reduce(srcs, world, lo, count, scale, out.data_ptr(), blocks, stream()))


def push_words(mem: PeerMemory, segments) -> None:
    """store words [off, off + n) of the own arena at the same place of every peer's arena; segments = [(off_words, n_words)]"""
    for i in range(0, len(segments), 8):
        seg = segments[i:i + 8]
        offs = (C.c_int64 * len(seg))(*[s[0] for s in seg])
        cnts = (C.c_int64 * len(seg))(*[s[1] for s in seg])
        check(lib().cnc_peer_push(mem.pointer_array(), mem.rank, mem.world, offs, cnts, len(seg), stream()))


def peer_capable(group=None) -> bool:
    """True when every rank of the group sits on a CUDA device of this machine that can address all the others' memory"""
    if not (dist.is_initialized() and torch.cuda.is_available()):
        return False
    world, me = dist.get_world_size(group), torch.cuda.current_device()
    if world > 8:
        return False
    devs = [None] * world
    dist.all_gather_object(devs, (me, _hostname()), group=group)
    ok = len({h for _, h in devs}) == 1 and len({d for d, _ in devs}) == world
    if ok:
        ok = all(d == me or torch.cuda.can_device_access_peer(me, d) for d, _ in devs)
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok), group=group)
    return all(flags)


def _hostname() -> str:
    import socket

    return socket.gethostname()
