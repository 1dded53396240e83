"""World-size-2 CPU tests (gloo) of the multi-GPU host logic: the row partition, the ghost /
halo plan and the block-Jacobi diagonal block, exercised as a distributed SpMV and a
distributed block-Jacobi-preconditioned residual check against the global oracle."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opm_simulators_legacy_b200 import _lib
from opm_simulators_legacy_b200.distributed import local_rows, slab_offsets
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian


def _partition(rp, cg, offsets, world, rank):
    lib = _lib.load()
    f = lib.opmgpu_debug_partition
    ip, lp = C.POINTER(C.c_int), C.POINTER(C.c_longlong)
    f.argtypes = [C.c_int, ip, lp, lp, C.c_int, C.c_int, ip, lp, ip, ip, ip, ip, ip]
    f.restype = C.c_int
    n, nnz = rp.size - 1, cg.size
    full = np.zeros(nnz, dtype=np.int32); ghosts = np.zeros(nnz, dtype=np.int64); cnt = np.zeros(world, dtype=np.int32)
    rpd = np.zeros(n + 1, dtype=np.int32); cid = np.zeros(nnz, dtype=np.int32); src = np.zeros(nnz, dtype=np.int32)
    nd = C.c_int()
    ng = f(n, rp.ctypes.data_as(ip), cg.ctypes.data_as(lp), offsets.ctypes.data_as(lp), world, rank,
           full.ctypes.data_as(ip), ghosts.ctypes.data_as(lp), cnt.ctypes.data_as(ip), rpd.ctypes.data_as(ip),
           cid.ctypes.data_as(ip), src.ctypes.data_as(ip), C.byref(nd))
    return full, ghosts[:ng], cnt, rpd, cid[:nd.value], src[:nd.value]


def _worker(rank, world, port, dims, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_py as O
        s = synth_blackoil_jacobian(*dims, perm="lognormal")
        rp_g, ci_g, v_g, b_g = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
        offsets = slab_offsets(dims, world)
        lo, hi = int(offsets[rank]), int(offsets[rank + 1])
        rp, cg, v, b = local_rows(rp_g, ci_g, v_g, b_g, lo, hi)
        full, ghosts, cnt, rpd, cid, src = _partition(rp, cg, offsets, world, rank)
        n = hi - lo
        # every ghost is owned by exactly the rank the count vector says, in ascending order
        owners = np.searchsorted(offsets, ghosts, side="right") - 1
        assert (np.bincount(owners, minlength=world) == cnt).all() and (np.diff(ghosts) > 0).all() and cnt[rank] == 0
        # halo exchange: ask the owners for the ghost rows of x (what ncclSend/ncclRecv do on the GPU)
        wants = [None] * world
        dist.all_gather_object(wants, ghosts.tolist())
        x_loc = s.xstar.numpy()[lo:hi]
        sends = [x_loc[np.array(wants[p], dtype=np.int64)[(np.array(wants[p], dtype=np.int64) >= lo) &
                                                          (np.array(wants[p], dtype=np.int64) < hi)] - lo] for p in range(world)]
        got = [None] * world
        dist.all_gather_object(got, [a.tolist() for a in sends])
        ghost_vals = np.concatenate([np.array(got[p][rank], dtype=np.float64).reshape(-1, 3) for p in range(world)]) \
            if ghosts.size else np.zeros((0, 3))
        x_ext = np.concatenate([x_loc, ghost_vals])
        rp_ext = np.concatenate([rp, np.full(ghosts.size, rp[-1], dtype=np.int32)])      # ghost rows are empty
        y_loc = O.spmv(rp_ext, full, v, x_ext)[:n]
        y_ref = O.spmv(rp_g, ci_g, v_g, s.xstar.numpy())[lo:hi]
        assert np.array_equal(y_loc, y_ref)
        # the diagonal block is exactly the local rows restricted to local columns
        sel = (cg >= lo) & (cg < hi)
        assert np.array_equal(src, np.nonzero(sel)[0]) and np.array_equal(cid, (cg[sel] - lo).astype(np.int32))
        assert rpd[-1] == sel.sum()
        # block-Jacobi ILU0 of the diagonal block is usable: applying it reduces the local residual of b
        lu, bad = O.ilu0_factor(rpd, cid, v[src])
        assert bad == -1
        z = O.ilu0_apply(rpd, cid, lu, 1.0, b)
        z_ext_parts = [None] * world
        dist.all_gather_object(z_ext_parts, z.tolist())
        z_glob = np.concatenate([np.array(p, dtype=np.float64).reshape(-1, 3) for p in z_ext_parts])
        r_loc = b - O.spmv(rp_g, ci_g, v_g, z_glob)[lo:hi]
        tot = torch.tensor([float((r_loc ** 2).sum()), float((b ** 2).sum())], dtype=torch.float64)
        dist.all_reduce(tot)                                  # what ncclAllReduce does for the dot products
        assert float(tot[0]) < float(tot[1])
        q.put((rank, "ok"))
    except Exception as e:                                    # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("dims", [(6, 5, 8), (10, 10, 3)])
def test_row_partition_and_halo_plan_world2(dims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


@pytest.mark.parametrize("dims,world,axis", [((6, 5, 8), 2, 1), ((10, 9, 4), 3, 0), ((7, 8, 6), 4, 1), ((6, 5, 8), 2, 2)])
def test_slab_partition_along_any_axis(dims, world, axis):
    """Slabs along i or j need a renumbering (rows of a rank contiguous, natural relative order):
    the renumbered rows assemble to the same operator (bit-identical SpMV), own columns ascend, and the
    rank's diagonal block is again a Cartesian stencil (nx x ny x nz of the slab)."""
    from oracle import oracle_py as O
    from opm_simulators_legacy_b200.distributed import local_rows_permuted, slab_partition, weakest_axis
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp_g, ci_g, v_g, b_g = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    perm, offsets = slab_partition(dims, world, axis)
    assert sorted(perm.tolist()) == list(range(s.N)) and offsets[0] == 0 and offsets[-1] == s.N
    x = s.xstar.numpy()
    y_ref = O.spmv(rp_g, ci_g, v_g, x)
    rps, cols, vals = [np.zeros(1, dtype=np.int64)], [], []
    for r in range(world):
        lo, hi = int(offsets[r]), int(offsets[r + 1])
        rp, cg, v, b = local_rows_permuted(rp_g, ci_g, v_g, b_g, perm, lo, hi)
        assert np.array_equal(b, b_g[perm[lo:hi]])
        for i in range(hi - lo):
            own = cg[rp[i]:rp[i + 1]]
            assert (np.diff(own[(own >= lo) & (own < hi)]) > 0).all()
        rps.append(rp[1:].astype(np.int64) + rps[-1][-1]); cols.append(cg); vals.append(v)
        # the diagonal block of the slab is a Cartesian stencil in natural ordering
        sel = (cg >= lo) & (cg < hi)
        rows = np.repeat(np.arange(hi - lo), np.diff(rp))
        rpd = np.concatenate([[0], np.cumsum(np.bincount(rows[sel], minlength=hi - lo))]).astype(np.int32)
        cid = (cg[sel] - lo).astype(np.int32)
        lu, bad = O.ilu0_factor(rpd, cid, v[sel])
        rc, out, info = _host_apply_local(rpd, cid, lu, b)
        want = list(dims)
        want[axis] = dims[axis] * (r + 1) // world - dims[axis] * r // world
        assert rc == 0 and tuple(info[:3]) == tuple(want)
        assert np.array_equal(out, O.ilu0_apply(rpd, cid, lu, 0.9, b))
    y_new = O.spmv(np.concatenate(rps).astype(np.int32), np.concatenate(cols).astype(np.int32), np.concatenate(vals), x[perm])
    assert np.array_equal(y_new, y_ref[perm])          # entries keep the natural order inside a row
    # the vertical couplings of the synthetic reservoir grid are the strong ones: never cut k
    assert weakest_axis(dims, rp_g, ci_g, v_g, 2) in (0, 1)


def _host_apply_local(rp, ci, lu, d):
    lib = _lib.load()
    f = lib.opmgpu_debug_host_program_apply
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    f.argtypes = [C.c_int, ip, ip, dp, C.c_int, C.c_double, dp, dp, ip]
    f.restype = C.c_int
    rp = np.ascontiguousarray(rp, dtype=np.int32); ci = np.ascontiguousarray(ci, dtype=np.int32)
    lu = np.ascontiguousarray(lu); d = np.ascontiguousarray(d)
    out = np.zeros_like(d); info = np.zeros(8, dtype=np.int32)
    rc = f(len(rp) - 1, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), lu.ctypes.data_as(dp), 7, 0.9,
           d.ctypes.data_as(dp), out.ctypes.data_as(dp), info.ctypes.data_as(ip))
    return rc, out, info
