"""CPU-side checks: the C-ABI library loads and exports every declared symbol, fails loudly
without a GPU, the host analysis (sweep programs) is exact, and the host mirror's well
elimination / recovery matches a dense reference.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from opm_simulators_legacy_b200 import _lib
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs
from opm_simulators_legacy_b200.solver import (ADB, eliminateVariable, recoverVariable, make_params)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "opm_gpu_solver.h")).read()
    declared = sorted(set(re.findall(r"\b(opmgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(set(_lib.EXPORTS)) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_test_hooks_and_experiment_symbols():
    """include/opm_gpu_solver_testhooks.h: the shipping library exports its test hooks and none of
    the experiment entry points; the experiments build exports both and the whole C ABI."""
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "opm_gpu_solver_testhooks.h")).read()
    declared = sorted(set(re.findall(r"\b(opmgpu_debug_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(_lib.TEST_HOOKS + _lib.EXP_HOOKS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.TEST_HOOKS:
        assert hasattr(lib, name), name
    for name in _lib.EXP_HOOKS:
        assert not hasattr(lib, name), f"{name} must not be in the shipping library"
    exp = _lib.load_experiments()
    for name in _lib.TEST_HOOKS + _lib.EXP_HOOKS + _lib.EXPORTS:
        assert hasattr(exp, name), name


def test_default_params_are_the_reference_defaults():
    p = make_params()
    assert p.linear_solver_reduction == 1e-2 and p.linear_solver_maxiter == 150
    assert p.ilu_relaxation == 0.9 and p.linear_solver_ignoreconvergencefailure == 0
    q = make_params({"linear_solver_reduction": "1e-3", "linear_solver_maxiter": 50, "unrelated_key": 1})
    assert q.linear_solver_reduction == 1e-3 and q.linear_solver_maxiter == 50
    g = make_params({"newton_use_gmres": "true", "linear_solver_restart": 25})
    assert g.newton_use_gmres == 1 and g.linear_solver_restart == 25 and p.newton_use_gmres == 0 and p.linear_solver_restart == 40
    with pytest.raises(ValueError):
        make_params({"linear_solver_use_amg": True})


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.opmgpu_create(0, C.byref(h))
    assert rc != 0 and not h
    assert b"no CPU fallback" in lib.opmgpu_last_error(None) or b"CUDA" in lib.opmgpu_last_error(None)
    from opm_simulators_legacy_b200.solver import GpuLinearSolver
    with pytest.raises(RuntimeError):
        GpuLinearSolver(0)


def _host_apply(rp, ci, lu, P, w, d):
    lib = _lib.load()
    f = lib.opmgpu_debug_host_program_apply
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    f.argtypes = [C.c_int, ip, ip, dp, C.c_int, C.c_double, dp, dp, ip]
    f.restype = C.c_int
    rp = np.ascontiguousarray(rp, dtype=np.int32); ci = np.ascontiguousarray(ci, dtype=np.int32)
    lu = np.ascontiguousarray(lu); d = np.ascontiguousarray(d)
    out = np.zeros_like(d); info = np.zeros(8, dtype=np.int32)
    rc = f(len(rp) - 1, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), lu.ctypes.data_as(dp), P, w,
           d.ctypes.data_as(dp), out.ctypes.data_as(dp), info.ctypes.data_as(ip))
    return rc, out, info


@pytest.mark.parametrize("dims", [(10, 10, 3), (24, 20, 12), (64, 1, 1), (30, 17, 1), (7, 6, 40)])
@pytest.mark.parametrize("P", [1, 7, 148])
def test_sweep_programs_are_exact_on_cartesian_grids(oracle, dims, P):
    """The per-CTA step records (tiles, window slots, pushed results, tail lists) interpreted
    sequentially on the host must reproduce ParallelOverlappingILU0::apply bit for bit."""
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    for w in (0.9, 1.0):
        rc, out, info = _host_apply(rp, ci, lu, P, w, b)
        assert rc == 0
        assert tuple(info[:3]) == dims if min(dims[:2]) > 1 or dims[1] == 1 else True
        assert np.array_equal(out, oracle.ilu0_apply(rp, ci, lu, w, b))


@pytest.mark.parametrize("dims,P", [((40, 40, 20), 148), ((100, 100, 4), 148), ((24, 20, 12), 16)])
def test_sweep_programs_with_thread_block_clusters(oracle, dims, P, monkeypatch):
    """Cluster-major tile numbering and results delivered through distributed shared memory
    (dep codes >= kCxBase, push ids with kPushDsmem): interpreted programs stay bit-exact."""
    monkeypatch.setenv("OPMGPU_TEST_CLUSTER_CTAS", f"{P},{P // 4 * 4 - 4 if P > 8 else 0},{P // 8 * 8 - 8 if P > 16 else 0}")
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    rc, out, info = _host_apply(rp, ci, lu, P, 0.9, b)
    assert rc == 0 and np.array_equal(out, oracle.ilu0_apply(rp, ci, lu, 0.9, b))


@pytest.mark.parametrize("dims,P", [((40, 40, 12), 7), ((60, 50, 8), 16), ((33, 47, 5), 5)])
def test_programs_with_several_tiles_per_cta(oracle, dims, P):
    """More columns than one pass per CTA: tiles dealt out in wavefront order, several per CTA,
    ascending in the lower sweep / factorisation and descending in the upper sweep."""
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    rc, out, info = _host_apply(rp, ci, lu, P, 0.9, b)
    assert rc == 0 and info[4] <= 96 and np.array_equal(out, oracle.ilu0_apply(rp, ci, lu, 0.9, b))
    rc, lu2, bad2, info2 = _host_factor(rp, ci, v, P)
    assert rc == 0 and bad2 == -1 and np.array_equal(lu2, lu)


@pytest.mark.parametrize("N,extra,dense,P", [(700, 3, 12, 5), (700, 3, 12, 148), (3000, 2, 300, 3), (500, 6, 40, 64)])
def test_sweep_programs_are_exact_on_general_patterns(oracle, N, extra, dense, P):
    rp, ci, v = random_bcrs(N, extra, seed=N + dense, dense_group=dense)
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    assert bad == -1
    d = np.random.default_rng(1).standard_normal((N, 3))
    rc, out, info = _host_apply(rp, ci, lu, P, 0.9, d)
    assert rc == 0 and np.array_equal(out, oracle.ilu0_apply(rp, ci, lu, 0.9, d))


def _host_factor(rp, ci, v, P):
    lib = _lib.load()
    f = lib.opmgpu_debug_host_factor_program
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    f.argtypes = [C.c_int, ip, ip, dp, C.c_int, dp, ip, ip]
    f.restype = C.c_int
    rp = np.ascontiguousarray(rp, dtype=np.int32); ci = np.ascontiguousarray(ci, dtype=np.int32)
    v = np.ascontiguousarray(v)
    lu = np.zeros_like(v); bad = C.c_int(0); info = np.zeros(4, dtype=np.int32)
    rc = f(len(rp) - 1, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), v.ctypes.data_as(dp), P,
           lu.ctypes.data_as(dp), C.byref(bad), info.ctypes.data_as(ip))
    return rc, lu, bad.value, info


@pytest.mark.parametrize("dims", [(10, 10, 3), (24, 20, 12), (64, 1, 1), (30, 17, 1), (7, 6, 40), (3, 3, 3)])
@pytest.mark.parametrize("P", [1, 7, 148])
def test_factor_program_is_exact_on_cartesian_grids(oracle, dims, P):
    """The pipelined factorisation records (A_ii / A_ij / A_ji per row, pivot window, pushed
    pivots) interpreted on the host must reproduce bilu0_decomposition bit for bit."""
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp, ci, v = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy()
    lu_ref, bad_ref = oracle.ilu0_factor(rp, ci, v)
    rc, lu, bad, info = _host_factor(rp, ci, v, P)
    if rc == -2:
        pytest.skip("pattern has no pipelined factorisation program (falls back to the tile kernel)")
    assert rc == 0 and bad == bad_ref == -1
    assert np.array_equal(lu, lu_ref)


def test_factor_program_reports_singular_row_and_rejects_fill(oracle):
    s = synth_blackoil_jacobian(8, 7, 5, perm="lognormal")
    rp, ci, v = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy().copy()
    v[rp[0]] = 0.0                       # row 0 has no lower blocks: its pivot is A_00 itself
    rc, lu, bad, info = _host_factor(rp, ci, v, 5)
    assert rc == 0 and bad == oracle.ilu0_factor(rp, ci, v)[1] == 0
    # general pattern (fill outside the diagonal): no pipelined program, the tile kernel is used
    rp2, ci2, v2 = random_bcrs(300, 3, seed=5, dense_group=10)
    assert _host_factor(rp2, ci2, v2, 7)[0] == -2


def test_well_elimination_and_recovery_match_dense_solve():
    """eliminateVariable / recoverVariable (NewtonIterationUtilities.cpp:45-184): solving the
    Schur-reduced system and recovering must equal solving the full system."""
    rng = np.random.default_rng(3)
    N, nw = 40, 3
    sizes = [N, N, N, nw * 3, nw]
    nt = sum(sizes)
    A = rng.standard_normal((nt, nt)) * 0.1 + np.eye(nt) * 3.0
    A[:3 * N, 3 * N:] *= (rng.random((3 * N, nt - 3 * N)) < 0.1)       # sparse well couplings
    A[3 * N:, :3 * N] *= (rng.random((nt - 3 * N, 3 * N)) < 0.1)
    b = rng.standard_normal(nt)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    eqs = []
    for e in range(5):
        jac = [sp.csc_matrix(A[offs[e]:offs[e + 1], offs[v]:offs[v + 1]]) for v in range(5)]
        eqs.append(ADB(b[offs[e]:offs[e + 1]].copy(), jac))
    elim = [eqs[3]]
    red = eliminateVariable(eqs, 3)
    elim.append(red[3])
    red = eliminateVariable(red, 3)
    assert len(red) == 3 and all(len(e.jac) == 3 for e in red)
    Ared = np.block([[red[e].jac[v].toarray() for v in range(3)] for e in range(3)])
    bred = np.concatenate([red[e].value for e in range(3)])
    x = np.linalg.solve(Ared, bred)
    x = recoverVariable(elim[1], x, 3)
    x = recoverVariable(elim[0], x, 3)
    assert np.abs(x - np.linalg.solve(A, b)).max() <= 1e-10
