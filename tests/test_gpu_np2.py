"""GPU parity tests of block size np = 2 (the reference's Impl<2,Scalar>: two-phase decks,
NewtonIterationBlackoilInterleaved.cpp:467-487) against the oracle built with -DORACLE_BS=2.
The two-phase systems are the water / oil rows and (p, sw) columns of the synthetic three-phase
Jacobians.  Bars as for np = 3: SpMV, ILU0 factors and apply bit-identical, iteration counts equal,
increment within rel 1e-8 (double) / 1e-3 (float instance)."""
import numpy as np
import pytest

from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs
from opm_simulators_legacy_b200.solver import (GpuLinearSolver, NewtonIterationBlackoilGPU, ADB,
                                               LinearisedBlackoilResidual)

pytestmark = pytest.mark.gpu

CASES = {"c1_shape": dict(dims=(10, 10, 3), perm="homogeneous"), "small_lognormal": dict(dims=(24, 20, 12), perm="lognormal"),
         "plane_2d": dict(dims=(30, 17, 1), perm="lognormal"), "mid_lognormal": dict(dims=(40, 40, 20), perm="lognormal")}
SEL = [0, 1, 3, 4]                 # entries [eq][var] of a 3x3 block kept in the 2x2 block


BOOST = 1.02                       # the water / oil sub-system alone is close to singular: stronger diagonal blocks


def two_phase(s):
    rp, ci = s.rowptr.numpy(), s.colidx.numpy()
    v = np.ascontiguousarray(s.vals.numpy()[:, SEL])
    rows = np.repeat(np.arange(s.N), np.diff(rp))
    v[rows == ci] *= BOOST
    return rp, ci, v, np.ascontiguousarray(s.rhs.numpy()[:, :2])


@pytest.fixture(scope="module")
def np2_solver():
    g = GpuLinearSolver(0)
    g.set_block_size(2)
    yield g
    g.close()


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    cfg = CASES[request.param]
    return synth_blackoil_jacobian(*cfg["dims"], perm=cfg["perm"])


@pytest.mark.parametrize("single", [False, True], ids=["f64", "f32"])
def test_np2_kernels_bit_exact(np2_solver, oracle, case, single):
    rp, ci, v, b = two_phase(case)
    O = oracle.instance(single, 2)
    g = np2_solver
    g.set_precision(single)
    g.set_pattern(rp, ci)
    x = np.ascontiguousarray(case.xstar.numpy()[:, :2])
    assert np.array_equal(g.spmv_np(2, v, x), O.spmv(rp, ci, v, x).astype(np.float64))
    lu_ref, bad = O.ilu0_factor(rp, ci, v)
    assert bad == -1
    for w in (0.9, 1.0):
        lu, got, badg = g.ilu0_np(2, v, w, b)
        assert badg == -1 and np.array_equal(lu, lu_ref.astype(np.float64))
        assert np.array_equal(got, O.ilu0_apply(rp, ci, lu_ref, w, b).astype(np.float64))
    g.set_precision(False)


@pytest.mark.parametrize("single", [False, True], ids=["f64", "f32"])
def test_np2_solve_iteration_parity(np2_solver, oracle, case, single):
    rp, ci, v, b = two_phase(case)
    O = oracle.instance(single, 2)
    g = np2_solver
    g.set_precision(single)
    g.set_pattern(rp, ci)
    x, res = g.solve_bcrs_np(2, v, b)
    x_ref, ref = O.solve_bcrs(rp, ci, v, b)
    assert res["iterations"] == ref["iterations"] and res["half_steps"] == ref["half_steps"] and res["converged"] == 1
    tol = 1e-3 if single else 1e-8
    assert (np.abs(x - x_ref).max(0) <= tol * np.abs(x_ref).max(0)).all()
    if not single:
        x, res = g.solve_bcrs_np(2, v, b, linear_solver_reduction=1e-10, linear_solver_maxiter=400)
        assert res["converged"] == 1 and res["reduction"] < 1e-10
        # against a direct solve of the 2N x 2N system
        import scipy.sparse as sp
        import scipy.sparse.linalg as spl
        A = sp.bsr_matrix((v.reshape(-1, 2, 2), ci, rp)).tocsc()
        xs = spl.spsolve(A, b.reshape(-1)).reshape(-1, 2)
        assert (np.abs(x - xs).max(0) <= 1e-6 * np.abs(xs).max(0)).all()
    g.set_precision(False)


def _blocks2(s):
    b9 = s.csc_blocks()
    out = []
    for p1 in range(2):
        for p2 in range(2):
            cp, ri, val = b9[p1 * 3 + p2]
            val = np.array(val, dtype=np.float64, copy=True)
            cols = np.repeat(np.arange(len(cp) - 1), np.diff(cp))
            val[np.asarray(ri) == cols] *= BOOST
            out.append((cp, ri, val))
    return out


@pytest.mark.parametrize("single", [False, True], ids=["f64", "f32"])
def test_np2_csc_blocks_path(oracle, case, single):
    blocks = _blocks2(case)
    N = case.N
    rhs = case.rhs_eqmajor_unscaled.numpy()[:2 * N]
    scale = list(case.matbalscale)[:2]
    O = oracle.instance(single, 2)
    g = GpuLinearSolver(0)
    try:
        g.set_precision(single)
        dx, res = g.solve_from_csc_blocks_np(N, 2, blocks, scale, rhs)
        dx_ref, ref = O.solve_from_csc_blocks(N, blocks, scale, rhs)
        assert res["iterations"] == ref["iterations"]
        sc = np.abs(dx_ref.reshape(2, -1)).max(1).repeat(N)
        assert (np.abs(dx - dx_ref) <= (1e-3 if single else 1e-8) * sc).all()
        dx2, res2 = g.solve_from_csc_blocks_np(N, 2, blocks, scale, rhs)      # cached pattern
        assert res2["ms_analysis"] == 0.0 and np.array_equal(dx, dx2)
        # the same handle goes back to three phases
        dx3, res3 = g.solve_from_csc_blocks(N, case.csc_blocks(), case.matbalscale, case.rhs_eqmajor_unscaled.numpy())
        ref3 = oracle.instance(single, 3).solve_from_csc_blocks(N, case.csc_blocks(), case.matbalscale, case.rhs_eqmajor_unscaled.numpy())[1]
        assert res3["iterations"] == ref3["iterations"]
    finally:
        g.close()


def test_np2_general_pattern_and_errors(np2_solver, oracle):
    rp, ci, v9 = random_bcrs(500, extra_per_row=3, seed=5, dense_group=10)
    v = np.ascontiguousarray(v9[:, SEL])
    b = np.random.default_rng(2).standard_normal((500, 2))
    g = np2_solver
    g.set_pattern(rp, ci)
    assert np.array_equal(g.spmv_np(2, v, b), oracle.np2.spmv(rp, ci, v, b))
    lu_ref, _ = oracle.np2.ilu0_factor(rp, ci, v)
    lu, got, bad = g.ilu0_np(2, v, 0.9, b)
    assert np.array_equal(lu, lu_ref) and np.array_equal(got, oracle.np2.ilu0_apply(rp, ci, lu_ref, 0.9, b))
    x, res = g.solve_bcrs_np(2, v, b, linear_solver_reduction=1e-8)
    x_ref, ref = oracle.np2.solve_bcrs(rp, ci, v, b, reduction=1e-8)
    assert res["iterations"] == ref["iterations"] and np.abs(x - x_ref).max() <= 1e-8 * np.abs(x_ref).max()
    # singular pivot: the row the oracle names
    v2 = v.copy()
    row = 41
    d = np.searchsorted(ci[rp[row]:rp[row + 1]], row) + rp[row]
    v2[d] = 0.0
    for k in range(rp[row], rp[row + 1]):
        if ci[k] < row:
            v2[k] = 0.0
    _, bad_ref = oracle.np2.ilu0_factor(rp, ci, v2)
    assert g.ilu0_np(2, v2)[2] == bad_ref == row
    # block sizes outside the reference's range 2..6, and one the pattern was not prepared for
    with pytest.raises(ValueError):
        g.set_block_size(7)
    with pytest.raises(ValueError):
        g.set_block_size(1)
    with pytest.raises(ValueError):
        g.solve_bcrs_np(5, np.zeros((len(ci), 25)), np.zeros((500, 5)))
    # a pattern that was not prepared for np = 2
    h = GpuLinearSolver(0)
    h.set_pattern(rp, ci)
    with pytest.raises(ValueError):
        h.solve_bcrs_np(2, v, b)
    h.close()


def test_newton_iteration_blackoil_gpu_two_phase(oracle):
    import scipy.sparse as sp
    s = synth_blackoil_jacobian(12, 10, 6, perm="lognormal")
    blocks = _blocks2(s)
    N = s.N
    rhs = s.rhs_eqmajor_unscaled.numpy()[:2 * N]
    eqs = []
    for p1 in range(2):
        jac = [sp.csc_matrix((blocks[p1 * 2 + p2][2], blocks[p1 * 2 + p2][1], blocks[p1 * 2 + p2][0]), shape=(N, N)) for p2 in range(2)]
        eqs.append(ADB(rhs[p1 * N:(p1 + 1) * N].copy(), jac))
    solver = NewtonIterationBlackoilGPU({})
    for single in (False, True):
        res = LinearisedBlackoilResidual(eqs, matbalscale=s.matbalscale, singlePrecision=single)
        dx = solver.computeNewtonIncrement(res)
        ref, r = oracle.instance(single, 2).solve_from_csc_blocks(N, blocks, list(s.matbalscale)[:2], rhs)
        assert solver.iterations() == r["iterations"] and dx.size == 2 * N
        sc = np.abs(ref.reshape(2, -1)).max(1).repeat(N)
        assert (np.abs(dx - ref) <= (1e-3 if single else 1e-8) * sc).all()
