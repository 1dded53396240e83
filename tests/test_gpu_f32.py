"""GPU parity tests of the single-precision instance (the reference's Impl<3,float>, selected by
LinearisedBlackoilResidual::singlePrecision: NewtonIterationBlackoilInterleaved.cpp:467-487,
BlackoilModelBase_impl.hpp:284), against the float build of the CPU oracle (oracle_py.f32).

Tolerances.  Matrix values, right-hand side and x are rounded to float once on both sides; SpMV,
ILU0 factors and ILU0 sweeps then follow the same operation order in float and must be
BIT-IDENTICAL (the C ABI exchanges doubles: every returned value must be a float widened exactly).
Scalar products are float tree reductions on the GPU and a sequential float sum in dune, so the
BiCGStab iterates agree to float accuracy, not bitwise: iteration counts must be equal at the
reference's tolerance, the increment must agree within rel 1e-3 (two solutions of a 1e-2-reduction
solve in float).
"""
import numpy as np
import pytest

from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs
from opm_simulators_legacy_b200.solver import (GpuLinearSolver, NewtonIterationBlackoilGPU, ADB,
                                               LinearisedBlackoilResidual)

pytestmark = pytest.mark.gpu

CASES = {
    "c1_spe1_shape": dict(dims=(10, 10, 3), perm="homogeneous"),
    "small_lognormal": dict(dims=(24, 20, 12), perm="lognormal"),
    "plane_2d": dict(dims=(30, 17, 1), perm="lognormal"),
    "mid_lognormal": dict(dims=(40, 40, 20), perm="lognormal"),
    "wide_multi_tile": dict(dims=(120, 125, 6), perm="lognormal"),
}


def _np(s):
    return s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()


@pytest.fixture(scope="module")
def f32_solver():
    s = GpuLinearSolver(0)
    s.set_precision(True)
    assert s.single_precision()
    yield s
    s.close()


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    cfg = CASES[request.param]
    return synth_blackoil_jacobian(*cfg["dims"], perm=cfg["perm"])


def _is_float_valued(a):
    return np.array_equal(a, a.astype(np.float32).astype(np.float64))


def test_f32_spmv_bit_exact(f32_solver, oracle, case):
    rp, ci, v, b = _np(case)
    f32_solver.set_pattern(rp, ci)
    f32_solver.set_values(v)
    x = case.xstar.numpy()
    got = f32_solver.spmv(x)
    ref = oracle.f32.spmv(rp, ci, v, x)
    assert ref.dtype == np.float32 and _is_float_valued(got)
    assert np.array_equal(got, ref.astype(np.float64))
    # and it is not the double result in disguise
    assert not np.array_equal(got, oracle.spmv(rp, ci, v, x))


def test_f32_ilu0_factor_and_apply_bit_exact(f32_solver, oracle, case):
    rp, ci, v, b = _np(case)
    f32_solver.set_pattern(rp, ci)
    f32_solver.set_values(v)
    assert f32_solver.ilu0_factor() == -1
    lu_ref, bad = oracle.f32.ilu0_factor(rp, ci, v)
    assert bad == -1
    lu = f32_solver.ilu0_factors()
    assert _is_float_valued(lu)
    assert np.array_equal(lu, lu_ref.astype(np.float64))
    for w in (0.9, 1.0):
        got = f32_solver.ilu0_apply(w, b)
        assert np.array_equal(got, oracle.f32.ilu0_apply(rp, ci, lu_ref, w, b).astype(np.float64))


def test_f32_solve_iteration_parity(f32_solver, oracle, case):
    rp, ci, v, b = _np(case)
    f32_solver.set_pattern(rp, ci)
    x, res = f32_solver.solve_bcrs(v, b)
    x_ref, ref = oracle.f32.solve_bcrs(rp, ci, v, b)
    assert res["iterations"] == ref["iterations"] and res["half_steps"] == ref["half_steps"]
    assert res["converged"] == 1 and res["reduction"] < 1e-2
    assert _is_float_valued(x)
    scale = np.abs(x_ref).max(0)
    assert (np.abs(x - x_ref).max(0) <= 1e-3 * scale).all()
    # against the double instance the float one is a different solve of the same system
    x64, ref64 = oracle.solve_bcrs(rp, ci, v, b)
    assert (np.abs(x - x64).max(0) <= 5e-2 * np.abs(x64).max(0)).all()


def test_f32_equal_half_steps(f32_solver, oracle, case):
    rp, ci, v, b = _np(case)
    f32_solver.set_pattern(rp, ci)
    for hs in (1, 2):
        x, res = f32_solver.solve_bcrs(v, b, raise_on_failure=False, linear_solver_reduction=1e-30, max_half_steps=hs)
        x_ref, ref = oracle.f32.solve_bcrs(rp, ci, v, b, reduction=1e-30, max_half_steps=hs)
        if ref["half_steps"] < hs:
            continue
        assert res["half_steps"] == ref["half_steps"]
        scale = np.abs(x_ref).max(0)
        assert (np.abs(x - x_ref).max(0) <= 1e-4 * scale).all()


def test_f32_csc_blocks_path_matches_oracle(f32_solver, oracle, case):
    blocks = case.csc_blocks()
    rhs = case.rhs_eqmajor_unscaled.numpy()
    dx, res = f32_solver.solve_from_csc_blocks(case.N, blocks, case.matbalscale, rhs)
    dx_ref, ref = oracle.f32.solve_from_csc_blocks(case.N, blocks, case.matbalscale, rhs)
    assert res["iterations"] == ref["iterations"]
    sc = np.abs(dx_ref.reshape(3, -1)).max(1).repeat(case.N)
    assert (np.abs(dx - dx_ref) <= 1e-3 * sc).all()
    assert _is_float_valued(dx)


def test_f32_general_pattern_with_dense_well_coupling(f32_solver, oracle):
    # Schur fill of a multi-perforation well: rows with many couplings (tile factorisation kernel,
    # sweeps with tail lists) in the float instance
    rp, ci, v = random_bcrs(700, extra_per_row=3, seed=7, dense_group=12)
    b = np.random.default_rng(3).standard_normal((700, 3))
    f32_solver.set_pattern(rp, ci)
    f32_solver.set_values(v)
    assert f32_solver.ilu0_factor() == -1
    lu_ref, bad = oracle.f32.ilu0_factor(rp, ci, v)
    assert bad == -1
    assert np.array_equal(f32_solver.ilu0_factors(), lu_ref.astype(np.float64))
    got = f32_solver.ilu0_apply(0.9, b)
    assert np.array_equal(got, oracle.f32.ilu0_apply(rp, ci, lu_ref, 0.9, b).astype(np.float64))
    assert np.array_equal(f32_solver.spmv(b), oracle.f32.spmv(rp, ci, v, b).astype(np.float64))
    xs, res = f32_solver.solve_bcrs(v, b)
    xr, ref = oracle.f32.solve_bcrs(rp, ci, v, b)
    assert res["iterations"] == ref["iterations"]


def test_switching_precision_on_one_handle(oracle):
    s = synth_blackoil_jacobian(16, 12, 8, perm="lognormal")
    rp, ci, v, b = _np(s)
    g = GpuLinearSolver(0)
    g.set_pattern(rp, ci)
    x64a, r64a = g.solve_bcrs(v, b)
    g.set_precision(True)
    with pytest.raises(ValueError):
        g.spmv(b)                      # values belong to the other instance
    x32, r32 = g.solve_bcrs(v, b)
    g.set_precision(False)
    x64b, r64b = g.solve_bcrs(v, b)
    assert np.array_equal(x64a, x64b) and r64a["iterations"] == r64b["iterations"]
    assert not np.array_equal(x64a, x32)
    xr, ref = oracle.f32.solve_bcrs(rp, ci, v, b)
    assert r32["iterations"] == ref["iterations"]
    # restarted GMRES exists for the double instance only
    g.set_precision(True)
    with pytest.raises(ValueError):
        g.solve_bcrs(v, b, newton_use_gmres=True)
    g.close()


def test_newton_iteration_blackoil_gpu_honours_single_precision(oracle):
    import scipy.sparse as sp
    s = synth_blackoil_jacobian(12, 10, 6, perm="lognormal")
    blocks = s.csc_blocks()
    N = s.N
    rhs = s.rhs_eqmajor_unscaled.numpy()
    eqs = []
    for p1 in range(3):
        jac = [sp.csc_matrix((blocks[p1 * 3 + p2][2], blocks[p1 * 3 + p2][1], blocks[p1 * 3 + p2][0]), shape=(N, N))
               for p2 in range(3)]
        eqs.append(ADB(rhs[p1 * N:(p1 + 1) * N].copy(), jac))
    solver = NewtonIterationBlackoilGPU({})
    res64 = LinearisedBlackoilResidual(eqs, matbalscale=s.matbalscale, singlePrecision=False)
    res32 = LinearisedBlackoilResidual(eqs, matbalscale=s.matbalscale, singlePrecision=True)
    dx32 = solver.computeNewtonIncrement(res32)
    it32 = solver.iterations()
    dx64 = solver.computeNewtonIncrement(res64)
    ref32, r32 = oracle.f32.solve_from_csc_blocks(N, blocks, s.matbalscale, rhs)
    ref64, r64 = oracle.solve_from_csc_blocks(N, blocks, s.matbalscale, rhs)
    assert it32 == r32["iterations"] and solver.iterations() == r64["iterations"]
    sc = np.abs(ref64.reshape(3, -1)).max(1).repeat(N)
    assert (np.abs(dx64 - ref64) <= 1e-8 * sc).all()
    assert (np.abs(dx32 - ref32) <= 1e-3 * sc).all()
    assert np.array_equal(dx32, dx32.astype(np.float32).astype(np.float64))


# ---- committed golden vectors of the float instance (tests/golden/f32, make_golden.py) --------------
import glob
import os

GOLDEN_F32 = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "f32", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN_F32, ids=[os.path.basename(p)[:-4] for p in GOLDEN_F32])
def test_f32_cuda_path_reproduces_golden_vectors(f32_solver, path):
    g = np.load(path)
    rp, ci, v, b = g["rowptr"], g["colidx"], g["vals"], g["rhs"]
    f32_solver.set_pattern(rp, ci)
    f32_solver.set_values(v)
    assert np.array_equal(f32_solver.spmv(g["x_probe"]), g["spmv"].astype(np.float64))
    assert f32_solver.ilu0_factor() == -1
    assert np.array_equal(f32_solver.ilu0_factors(), g["lu"].astype(np.float64))
    assert np.array_equal(f32_solver.ilu0_apply(0.9, b), g["apply_w09"].astype(np.float64))
    assert np.array_equal(f32_solver.ilu0_apply(1.0, b), g["apply_w1"].astype(np.float64))
    x, res = f32_solver.solve_bcrs(v, b)
    assert res["iterations"] == int(g["iterations"]) and res["half_steps"] == int(g["half_steps"])
    assert np.abs(x - g["x"]).max() <= 1e-3 * np.abs(g["x"]).max()


# ---- error contract and robustness of the float instance ------------------------------------------
def test_f32_zero_rhs_and_error_contract(f32_solver, oracle):
    from opm_simulators_legacy_b200.solver import LinearSolverProblem, NumericalIssue
    s = synth_blackoil_jacobian(8, 8, 4, perm="lognormal")
    rp, ci, v, b = _np(s)
    f32_solver.set_pattern(rp, ci)
    x, res = f32_solver.solve_bcrs(v, np.zeros_like(b))
    assert res["iterations"] == 0 and res["converged"] == 1 and not x.any()
    # not converged -> LinearSolverProblem, iterations still reported (ISTLSolver.hpp:358-368);
    # the float oracle runs into maxiter at the same count
    with pytest.raises(LinearSolverProblem):
        f32_solver.solve_bcrs(v, b, linear_solver_reduction=1e-14, linear_solver_maxiter=2)
    _, ref = oracle.f32.solve_bcrs(rp, ci, v, b, reduction=1e-14, maxiter=2)
    assert f32_solver.last["iterations"] == ref["iterations"] == 2 and f32_solver.last["converged"] == 0
    # singular pivot block -> NumericalIssue naming the row the float oracle names
    v2 = v.copy()
    row = 37
    d = np.searchsorted(ci[rp[row]:rp[row + 1]], row) + rp[row]
    v2[d] = 0.0
    for k in range(rp[row], rp[row + 1]):
        if ci[k] < row:
            v2[k] = 0.0
    _, bad = oracle.f32.ilu0_factor(rp, ci, v2)
    with pytest.raises(NumericalIssue):
        f32_solver.solve_bcrs(v2, b)
    assert f32_solver.last["bad_row"] == bad == row


def test_f32_nan_and_inf_end_the_solve(f32_solver):
    """NaN / Inf in values or right-hand side: "not converged" or a singular-block error, promptly,
    no watchdog trip (the float containers' empty marker is never a result), handle usable after."""
    import time
    from opm_simulators_legacy_b200.solver import LinearSolverProblem, NumericalIssue
    s = synth_blackoil_jacobian(24, 20, 12, perm="lognormal")
    rp, ci, v, b = _np(s)
    f32_solver.set_pattern(rp, ci)
    x_ok, res_ok = f32_solver.solve_bcrs(v, b)
    for what in ("rhs_nan", "rhs_inf", "val_nan_offdiag", "val_nan_diag"):
        v2, b2 = v.copy(), b.copy()
        if what == "rhs_nan":
            b2[777, 1] = np.nan
        elif what == "rhs_inf":
            b2[777, 1] = np.inf
        elif what == "val_nan_offdiag":
            v2[rp[900] + 1, 5] = np.nan
        else:
            d = np.searchsorted(ci[rp[900]:rp[901]], 900) + rp[900]
            v2[d, 0] = np.nan
        t0 = time.time()
        with pytest.raises((LinearSolverProblem, NumericalIssue)):
            f32_solver.solve_bcrs(v2, b2, linear_solver_maxiter=20)
        assert time.time() - t0 < 5.0, what
        assert "watchdog" not in f32_solver.error(), (what, f32_solver.error())
        x, res = f32_solver.solve_bcrs(v, b)
        assert res["iterations"] == res_ok["iterations"] and np.array_equal(x, x_ok), what


def test_f32_full_size_properties():
    """C3 (1M cells) in the float instance: size-independent properties -- the solve converges at the
    reference's tolerance, the true residual (in double) is reduced accordingly, results are
    deterministic run to run, the float SpMV (128-row tiles) agrees with the double one to float accuracy."""
    import torch
    s = synth_blackoil_jacobian(100, 100, 100, perm="lognormal")
    g = GpuLinearSolver(0)
    try:
        g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
        vals = s.vals.cuda(); rhs = s.rhs.cuda()
        y64 = torch.zeros_like(rhs); y32 = torch.zeros_like(rhs)
        g.set_values_dev(vals)
        g.spmv_dev(rhs, y64)
        g.set_precision(True)
        g.set_values_dev(vals)
        g.spmv_dev(rhs, y32)
        torch.cuda.synchronize()
        assert float((y32 - y64).abs().max()) <= 1e-5 * float(y64.abs().max())
        x1 = torch.zeros_like(rhs); x2 = torch.zeros_like(rhs)
        r1 = g.solve_bcrs_dev(None, rhs, x1)
        r2 = g.solve_bcrs_dev(None, rhs, x2)
        assert r1["converged"] == 1 and r1["iterations"] == r2["iterations"] and torch.equal(x1, x2)
        g.set_precision(False)
        g.set_values_dev(vals)
        ax = torch.zeros_like(rhs)
        g.spmv_dev(x1, ax)
        torch.cuda.synchronize()
        assert float((rhs - ax).norm() / rhs.norm()) <= 1e-2 * 1.01
    finally:
        g.close()
