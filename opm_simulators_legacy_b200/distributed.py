"""Row-partitioned multi-GPU solve: one process per GPU (torchrun), cells split into slabs
along the grid axis with the weakest coupling, block-Jacobi ILU0 per GPU, halo exchange of x with ncclSend/ncclRecv and dot products
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


def weakest_axis(dims, rowptr, colidx, vals, world):
    """Grid axis (0 = i, 1 = j, 2 = k) whose couplings are the weakest: block-Jacobi ILU0 drops
    the couplings a slab boundary cuts, so cutting the strong direction (the vertical one in a
    reservoir grid: thin cells, large areas) costs iterations.  Measured by the mean |a_00| of
    the off-diagonal blocks along each axis; axes shorter than the number of ranks are skipped."""
    nx, ny, nz = dims
    rowptr = np.asarray(rowptr, dtype=np.int64)
    rows = np.repeat(np.arange(rowptr.size - 1, dtype=np.int64), np.diff(rowptr))
    delta = np.asarray(colidx, dtype=np.int64) - rows
    a00 = np.abs(np.asarray(vals).reshape(-1, 9)[:, 0])
    best, best_w = 2, np.inf
    for axis, stride, n in ((0, 1, nx), (1, nx, ny), (2, nx * ny, nz)):
        if n < world or (axis == 0 and nx == 1) or (axis == 1 and ny == 1):
            continue
        m = delta == stride if stride != 1 or nx > 1 else np.zeros_like(delta, dtype=bool)
        if not m.any():
            continue
        w = float(a00[m].mean())
        if w < best_w * (1 - 1e-12) or (abs(w - best_w) <= 1e-12 * best_w and axis > best):
            best, best_w = axis, w            # ties: the slower-running axis keeps slabs more contiguous
    return best


def slab_partition(dims, world, axis):
    """Slabs along `axis`: rank g owns the cells whose coordinate c along the axis satisfies
    g*n/G <= c < (g+1)*n/G.  Returns (perm, offsets): perm[new] = natural row, the rows of a
    rank are contiguous in the new numbering and keep their natural relative order (so the
    rank's diagonal block is again a Cartesian stencil in natural ordering)."""
    nx, ny, nz = dims
    n = dims[axis]
    cell = np.arange(nx * ny * nz, dtype=np.int64)
    coord = (cell % nx, (cell // nx) % ny, cell // (nx * ny))[axis]
    bounds = np.array([n * g // world for g in range(world + 1)], dtype=np.int64)
    owner = np.searchsorted(bounds, coord, side="right") - 1
    perm = np.argsort(owner, kind="stable").astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=world))]).astype(np.int64)
    return perm, offsets


def local_rows(rowptr, colidx, vals, rhs, lo, hi):
    """A rank's rows of a global BCRS system: local rowptr, GLOBAL column ids, values, rhs."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    k0, k1 = rowptr[lo], rowptr[hi]
    return ((rowptr[lo:hi + 1] - k0).astype(np.int32), np.asarray(colidx[k0:k1], dtype=np.int64),
            np.ascontiguousarray(vals[k0:k1]), np.ascontiguousarray(rhs[lo:hi]))


def local_rows_permuted(rowptr, colidx, vals, rhs, perm, lo, hi):
    """The same for a renumbered system (new row q = natural row perm[q]): the rank's rows
    perm[lo:hi] with column ids in the NEW global numbering.  Entries keep their natural order
    inside a row (the SpMV then sums in the reference's order: bit-identical results); the
    columns the rank owns still ascend, because a rank's cells keep their relative order."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    colidx = np.asarray(colidx, dtype=np.int64)
    inv = np.empty(perm.size, dtype=np.int64)
    inv[perm] = np.arange(perm.size, dtype=np.int64)
    rows = perm[lo:hi]
    lens = rowptr[rows + 1] - rowptr[rows]
    rp = np.concatenate([[0], np.cumsum(lens)])
    ent = np.repeat(rowptr[rows] - rp[:-1], lens) + np.arange(rp[-1], dtype=np.int64)     # BCRS slots of the rows
    cols = inv[colidx[ent]]
    vals = np.asarray(vals).reshape(-1, 9)
    return (rp.astype(np.int32), np.ascontiguousarray(cols), np.ascontiguousarray(vals[ent]),
            np.ascontiguousarray(np.asarray(rhs).reshape(-1, 3)[rows]))


class DistributedSolver(GpuLinearSolver):
    """A GpuLinearSolver whose handle owns one slab of a global Cartesian system."""

    def __init__(self, system, device: int, axis="auto"):
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
        g_rp, g_ci, g_v, g_b = system.rowptr.numpy(), system.colidx.numpy(), system.vals.numpy(), system.rhs.numpy()
        self.axis = weakest_axis(system.dims, g_rp, g_ci, g_v, self.world) if axis == "auto" else int(axis)
        # perm[new row] = natural row; every rank computes the same partition
        self.perm, self.offsets = slab_partition(system.dims, self.world, self.axis)
        lo, hi = int(self.offsets[self.rank]), int(self.offsets[self.rank + 1])
        if self.axis == 2:
            rp, cg, v, b = local_rows(g_rp, g_ci, g_v, g_b, lo, hi)        # k-slabs are contiguous as they are
        else:
            rp, cg, v, b = local_rows_permuted(g_rp, g_ci, g_v, g_b, self.perm, lo, hi)
        self.N, self.nnzb = hi - lo, cg.size
        self.nnzb_diag = int(((cg >= lo) & (cg < hi)).sum())              # blocks the rank's ILU0 is built on
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

    def to_natural(self, x_new_order):
        """Rows gathered from all ranks (new numbering) -> natural cell order."""
        out = np.empty_like(x_new_order)
        out[self.perm] = x_new_order
        return out

    def solve(self, params=None, raise_on_failure=True):
        return self.solve_bcrs_dev(self.vals, self.rhs, self.x, params=params or make_params(),
                                   raise_on_failure=raise_on_failure)
