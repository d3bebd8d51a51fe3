"""Peer-memory all-reduce of the per-step ``[loss | gradients]`` buffer
(libvibo_b200.so ``vibo_comm_*``; csrc/vibo_comm.cu).

One kernel over NVLink peer memory (CUDA IPC), capturable in a CUDA graph, sum
in rank order (bit-identical on every rank).  ``torch.distributed`` is used
only once, to exchange the 64-byte IPC handles.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class PeerAllReduce:
    def __init__(self, max_floats: int, device, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device)
        self.handle = C.c_void_p()
        mine = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            rc = self.lib.vibo_comm_create(self.rank, self.world, int(max_floats), C.byref(self.handle), mine)
            self._check(rc, "vibo_comm_create")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine.raw), group=group)
            blob = C.create_string_buffer(b"".join(handles), 64 * self.world)
            rc = self.lib.vibo_comm_connect(self.handle, blob)
            self._check(rc, "vibo_comm_connect")
        # nobody may arrive in a region that is not mapped everywhere yet
        dist.barrier(group=group)
        self.max_floats = int(max_floats)

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.vibo_comm_last_error()
            raise _lib.ViboError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")

    def all_reduce_(self, flat: torch.Tensor):
        """In-place sum over ranks of a contiguous float32 CUDA tensor, on the current stream."""
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
        rc = self.lib.vibo_comm_allreduce(self.handle, C.c_void_p(flat.data_ptr()), flat.numel(),
                                          C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream))
        self._check(rc, "vibo_comm_allreduce")
        return flat

    def all_reduce_adam_(self, flat, skip, param, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8):
        """all_reduce_ followed, in the same kernel, by the Adam step on ``flat[skip:]`` (the summed
        gradients of ``param``); ``step`` is the device-side int64 update count."""
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
        assert param.numel() == flat.numel() - skip
        ptr = lambda x: C.c_void_p(x.data_ptr())
        rc = self.lib.vibo_comm_allreduce_adam(self.handle, ptr(flat), flat.numel(), int(skip), ptr(param),
                                               ptr(exp_avg), ptr(exp_avg_sq), ptr(step), C.c_float(lr),
                                               C.c_float(betas[0]), C.c_float(betas[1]), C.c_float(eps),
                                               C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream))
        self._check(rc, "vibo_comm_allreduce_adam")
        return flat

    def status(self):
        self._check(self.lib.vibo_comm_status(self.handle), "vibo_comm_status")

    def close(self):
        if self.handle:
            dist.barrier(group=self.group)  # peers may still be reading this rank's region
            self.lib.vibo_comm_destroy(self.handle)
            self.handle = C.c_void_p()
