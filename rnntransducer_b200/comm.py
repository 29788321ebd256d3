"""Gradient all-reduce over NVLink peer memory for the one-process-per-GPU path.

``PeerAllReduce`` is host-side plumbing around ``rnntb200_comm_*`` (csrc/comm.cu): every rank allocates one
buffer, the ranks swap its cudaIpc handle through ``torch.distributed`` (whatever backend the process group
has), map each other's buffers, and from then on ``all_reduce_mean_`` is ONE kernel launch on the caller's
stream -- no NCCL call, no host synchronisation, capturable inside the step's CUDA graph.  It reduces what DDP
would reduce for this path: the gradients our kernels produce for ``fc.weight`` / ``fc.bias`` (reference
train.py:45-48, model.py:59).  Single node, <= 8 ranks, fp32.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import torch
import torch.distributed as dist

from . import _lib


class PeerAllReduce:
    def __init__(self, max_floats: int, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError("PeerAllReduce: at most 8 ranks (one NVSwitch domain)")
        self.max_floats = int(max_floats)
        self.device = torch.device("cuda", torch.cuda.current_device())
        nbytes = self.lib.rnntb200_comm_buffer_bytes(self.max_floats, self.world)
        own = ctypes.c_void_p()
        _lib.check(self.lib.rnntb200_comm_alloc(nbytes, ctypes.byref(own)), "rnntb200_comm_alloc")
        self.own = own.value
        handle = ctypes.create_string_buffer(64)
        _lib.check(self.lib.rnntb200_comm_export(self.own, handle), "rnntb200_comm_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.rank, bytes(handle.raw)), group=group)
        self.peers = (ctypes.c_void_p * self.world)()
        self._imported = []
        for r, h in handles:
            if r == self.rank:
                self.peers[r] = self.own
            else:
                p = ctypes.c_void_p()
                _lib.check(self.lib.rnntb200_comm_import(h, ctypes.byref(p)), f"rnntb200_comm_import (rank {r})")
                self.peers[r] = p.value
                self._imported.append(p.value)
        dist.barrier(group=group)  # every buffer is zeroed and mapped before anybody launches

    def all_reduce_mean_(self, tensors: Sequence[torch.Tensor], average: bool = True) -> None:
        """In place: every tensor <- mean (or sum) over ranks.  fp32, contiguous, CUDA, <= 4 tensors."""
        if self.own is None:
            raise RuntimeError("PeerAllReduce: used after close()")
        n = len(tensors)
        if not 1 <= n <= 4:
            raise RuntimeError("PeerAllReduce: between 1 and 4 tensors per call")
        segs = (ctypes.c_void_p * n)()
        sizes = (ctypes.c_int * n)()
        for i, t in enumerate(tensors):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise RuntimeError("PeerAllReduce: contiguous fp32 CUDA tensors only")
            segs[i], sizes[i] = t.data_ptr(), t.numel()
        _lib.check(self.lib.rnntb200_comm_allreduce(
            self.peers, self.rank, self.world, segs, sizes, n, self.max_floats,
            1.0 / self.world if average else 1.0, torch.cuda.current_stream().cuda_stream), "rnntb200_comm_allreduce")

    def close(self) -> None:
        """Collective: unmap the peers' buffers, then free our own once nobody maps it any more."""
        if self.own is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._imported:
            self.lib.rnntb200_comm_release(p)
        dist.barrier(group=self.group)
        self.lib.rnntb200_comm_free(self.own)
        self.own, self._imported = None, []
