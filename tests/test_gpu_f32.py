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
