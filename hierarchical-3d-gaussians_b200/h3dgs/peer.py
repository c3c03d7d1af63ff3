"""Peer memory of the tile-sharded multi-GPU mode (one process per GPU on one NVLink / NVSwitch box): one exportable
allocation per rank (libh3dgs.so: h3dgs_peer_alloc, CUDA IPC), carved into named regions, mapped into every other
rank's address space; plus the device-side barrier (h3dgs_peer_barrier, an ordinary kernel on the current stream).

torch.distributed is used once, at construction, to exchange the 64-byte IPC handles; nothing here runs per step
except the barrier launch.  The kernels that read / write this memory are the blend kernels in peer mode
(h3dgs_raster_args.peer_count, include/h3dgs.h) and the L1 loss kernel."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def _align(n, a=256):
    return (int(n) + a - 1) // a * a


class _CudaView:
    """CUDA array interface over raw device memory -> torch.as_tensor gives a tensor VIEW (no copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class PeerArena:
    def __init__(self, regions, world, rank, device, group=None):
        """regions: {name: nbytes}.  A `flags` region (barrier) is added.  Collective over `group`."""
        self.L = _lib.lib()
        self.world, self.rank, self.device, self.group = int(world), int(rank), torch.device(device), group
        if not (1 < self.world <= _lib.MAX_PEERS):
            raise ValueError(f"peer mode supports 2..{_lib.MAX_PEERS} ranks")
        self.offsets, off = {}, 0
        for name, nbytes in [("flags", self.L.h3dgs_peer_flag_bytes())] + list(regions.items()):
            self.offsets[name] = (off, int(nbytes))
            off = _align(off + int(nbytes))
        self.total = off
        with torch.cuda.device(self.device):
            base = C.c_void_p(0)
            _lib.check(self.L.h3dgs_peer_alloc(self.total, C.byref(base)))
            handle = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
            _lib.check(self.L.h3dgs_peer_export(base, handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.bases = []
            for r in range(self.world):
                if r == self.rank:
                    self.bases.append(int(base.value))
                else:
                    p = C.c_void_p(0)
                    buf = (C.c_ubyte * _lib.IPC_HANDLE_BYTES).from_buffer_copy(handles[r])
                    _lib.check(self.L.h3dgs_peer_open(buf, C.byref(p)))
                    self.bases.append(int(p.value))
        self._local = torch.as_tensor(_CudaView(self.bases[self.rank], self.total), device=self.device)
        self._flag_ptrs = (C.c_void_p * self.world)(*[self.ptr("flags", r) for r in range(self.world)])
        dist.barrier(group=group)              # every rank has mapped every arena before anyone writes into one

    def ptr(self, name, r=None):
        return self.bases[self.rank if r is None else r] + self.offsets[name][0]

    def ptrs(self, name):
        return [self.ptr(name, r) for r in range(self.world)]

    def tensor(self, name, dtype, shape):
        """torch view of the LOCAL region"""
        off, nbytes = self.offsets[name]
        return self._local[off:off + nbytes].view(dtype).view(shape)

    def barrier(self):
        """device-side barrier among the ranks, enqueued on the current stream (capturable)"""
        _lib.check(self.L.h3dgs_peer_barrier(self.world, self.rank, self.ptr("flags"), self._flag_ptrs,
                                             torch.cuda.current_stream(self.device).cuda_stream))

    def timed_out(self):
        return bool(self.L.h3dgs_peer_barrier_status(self.ptr("flags"), torch.cuda.current_stream(self.device).cuda_stream))

    def close(self):
        if getattr(self, "bases", None):
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            for r, b in enumerate(self.bases):
                if r != self.rank:
                    self.L.h3dgs_peer_close(C.c_void_p(b))
            dist.barrier(group=self.group)
            self._local = None
            self.L.h3dgs_peer_free(C.c_void_p(self.bases[self.rank]))
            self.bases = None


def probe(world, rank, device, group=None):
    """Can the ranks of `group` map each other's memory (CUDA IPC + peer access) and run the device-side barrier?
    Collective; every rank takes part in every exchange whatever happened locally, and all ranks return the same
    answer: (True, "") or (False, reason).  bench.py falls back to the NCCL form of the sharded step on False."""
    L = _lib.lib()
    dev = torch.device(device)
    err, base, opened = "", C.c_void_p(0), []
    handle = bytes(_lib.IPC_HANDLE_BYTES)
    nbytes = _align(L.h3dgs_peer_flag_bytes()) + 256
    try:
        with torch.cuda.device(dev):
            _lib.check(L.h3dgs_peer_alloc(nbytes, C.byref(base)))
            buf = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
            _lib.check(L.h3dgs_peer_export(base, buf))
            handle = bytes(buf)
    except Exception as e:          # noqa: BLE001
        err = f"rank {rank}: alloc/export failed: {e}"
    handles = [None] * world
    dist.all_gather_object(handles, (handle, err), group=group)
    err = err or next((h[1] for h in handles if h[1]), "")
    ptrs = [0] * world
    if not err:
        try:
            with torch.cuda.device(dev):
                for r in range(world):
                    if r == rank:
                        ptrs[r] = int(base.value)
                    else:
                        p = C.c_void_p(0)
                        _lib.check(L.h3dgs_peer_open((C.c_ubyte * _lib.IPC_HANDLE_BYTES).from_buffer_copy(handles[r][0]), C.byref(p)))
                        ptrs[r] = int(p.value); opened.append(p)
        except Exception as e:      # noqa: BLE001
            err = f"rank {rank}: open failed: {e}"
    errs = [None] * world
    dist.all_gather_object(errs, err, group=group)
    err = next((e for e in errs if e), "")
    if not err:
        try:
            with torch.cuda.device(dev):
                fp = (C.c_void_p * world)(*ptrs)
                s = torch.cuda.current_stream(dev).cuda_stream
                for _ in range(3):
                    _lib.check(L.h3dgs_peer_barrier(world, rank, ptrs[rank], fp, s))
                if L.h3dgs_peer_barrier_status(ptrs[rank], s):
                    err = f"rank {rank}: device-side barrier timed out"
        except Exception as e:      # noqa: BLE001
            err = f"rank {rank}: barrier failed: {e}"
    dist.all_gather_object(errs, err, group=group)
    err = next((e for e in errs if e), "")
    torch.cuda.synchronize(dev)
    dist.barrier(group=group)
    for p in opened:
        L.h3dgs_peer_close(p)
    dist.barrier(group=group)
    if base.value:
        L.h3dgs_peer_free(base)
    return (not err), err
