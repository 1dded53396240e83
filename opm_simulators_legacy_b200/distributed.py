"""Row-partitioned multi-GPU solve: one process per GPU (torchrun), cells split into contiguous
k-slabs, block-Jacobi ILU0 per GPU, halo exchange of x with ncclSend/ncclRecv and dot products
with ncclAllReduce inside libopmgpu.so.  torch.distributed is only the
bootstrap: it carries the ncclUniqueId from rank 0 to the other ranks.

The reference's counterpart is the MPI branch of ISTLSolver::solve
(opm/autodiff/ISTLSolver.hpp:286-298: OverlappingSchwarzOperator + per-subdomain ILU0).
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import _lib as L
from .solver import GpuLinearSolver, make_params


def slab_offsets(dims, world):
    """Contiguous k-slabs: rank g owns k in [g*nz/G, (g+1)*nz/G)."""
    nx, ny, nz = dims
    return np.array([(nz * g // world) * nx * ny for g in range(world + 1)], dtype=np.int64)


def local_rows(rowptr, colidx, vals, rhs, lo, hi):
    """A rank's rows of a global BCRS system: local rowptr, GLOBAL column ids, values, rhs."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    k0, k1 = rowptr[lo], rowptr[hi]
    return ((rowptr[lo:hi + 1] - k0).astype(np.int32), np.asarray(colidx[k0:k1], dtype=np.int64),
            np.ascontiguousarray(vals[k0:k1]), np.ascontiguousarray(rhs[lo:hi]))


class DistributedSolver(GpuLinearSolver):
    """A GpuLinearSolver whose handle owns one slab of a global Cartesian system."""

    def __init__(self, system, device: int):
        import torch
        import torch.distributed as dist
        self.lib = L.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            raw = (C.c_char * 128)()
            rc = self.lib.opmgpu_nccl_unique_id(raw)
            if rc != L.OK:
                raise RuntimeError("opmgpu_nccl_unique_id: " + self.lib.opmgpu_last_error(None).decode())
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        raw = (C.c_char * 128).from_buffer_copy(bytes(idbuf.cpu().numpy().tobytes()))
        self.h = C.c_void_p()
        rc = self.lib.opmgpu_create_distributed(int(device), self.rank, self.world, raw, C.byref(self.h))
        if rc != L.OK:
            raise RuntimeError("opmgpu_create_distributed: " + self.lib.opmgpu_last_error(None).decode())
        self.last = None
        self.offsets = slab_offsets(system.dims, self.world)
        lo, hi = int(self.offsets[self.rank]), int(self.offsets[self.rank + 1])
        rp, cg, v, b = local_rows(system.rowptr.numpy(), system.colidx.numpy(), system.vals.numpy(),
                                  system.rhs.numpy(), lo, hi)
        self.N, self.nnzb = hi - lo, cg.size
        self.use_torch_stream()
        t0 = time.perf_counter()
        self._check(self.lib.opmgpu_set_pattern_bcrs_distributed(
            self.h, self.N, self.nnzb, rp.ctypes.data_as(C.POINTER(C.c_int)),
            cg.ctypes.data_as(C.POINTER(C.c_longlong)), self.offsets.ctypes.data_as(C.POINTER(C.c_longlong))))
        self.analysis_ms = (time.perf_counter() - t0) * 1e3
        self.vals = torch.from_numpy(v).cuda()
        self.rhs = torch.from_numpy(b).cuda()
        self.x = torch.zeros_like(self.rhs)
        self.lo, self.hi = lo, hi

    def solve(self, params=None, raise_on_failure=True):
        return self.solve_bcrs_dev(self.vals, self.rhs, self.x, params=params or make_params(),
                                   raise_on_failure=raise_on_failure)
