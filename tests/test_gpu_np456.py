"""GPU parity tests of block sizes np = 4, 5, 6 (the reference's Impl<np,Scalar> for the polymer /
solvent extensions, NewtonIterationBlackoilInterleaved.cpp:467-487 and .hpp:73) against the oracle
built with -DORACLE_BS=np.  Systems: random np x np blocks on Cartesian and general patterns, block
rows diagonally dominant (jacobian.block_system_np).  Bars as for np = 2 / 3: SpMV, ILU0 factors and
apply bit-identical, iteration counts equal, increment within rel 1e-8 (double) / 1e-3 (float; 1e-2 where a pivot block was made ill-conditioned on purpose)."""
import numpy as np
import pytest

from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs, block_system_np
from opm_simulators_legacy_b200.solver import (GpuLinearSolver, NewtonIterationBlackoilGPU, ADB,
                                               LinearisedBlackoilResidual)

pytestmark = pytest.mark.gpu

PATTERNS = {"c1_shape": (10, 10, 3), "small": (24, 20, 12), "plane_2d": (30, 17, 1)}


@pytest.fixture(scope="module", params=[4, 5, 6])
def bs(request):
    return request.param


@pytest.fixture(scope="module")
def solver(bs):
    g = GpuLinearSolver(0)
    g.set_block_size(bs)
    yield g
    g.close()


@pytest.fixture(scope="module", params=list(PATTERNS))
def pattern(request):
    s = synth_blackoil_jacobian(*PATTERNS[request.param], perm="homogeneous")
    return s.rowptr.numpy(), s.colidx.numpy()


@pytest.mark.parametrize("single", [False, True], ids=["f64", "f32"])
def test_np456_kernels_bit_exact_and_solve_parity(solver, oracle, bs, pattern, single):
    rp, ci = pattern
    v, b, xstar = block_system_np(rp, ci, bs, seed=bs)
    # kernels: the first block row's pivot needs dune's row swap (np = 5, 6) -- a no-op for the 4x4 closed form
    vp = v.copy()
    d0 = rp[0] + int(np.searchsorted(ci[rp[0]:rp[1]], 0))
    vp[d0, 0] = 0.0
    O = oracle.instance(single, bs)
    g = solver
    g.set_precision(single)
    g.set_pattern(rp, ci)
    assert np.array_equal(g.spmv_np(bs, vp, xstar), O.spmv(rp, ci, vp, xstar).astype(np.float64))
    lu_ref, bad = O.ilu0_factor(rp, ci, vp)
    assert bad == -1
    for w in (0.9, 1.0):
        lu, got, badg = g.ilu0_np(bs, vp, w, b)
        assert badg == -1 and np.array_equal(lu, lu_ref.astype(np.float64))
        assert np.array_equal(got, O.ilu0_apply(rp, ci, lu_ref, w, b).astype(np.float64))
    # solves: the well-conditioned system (the zeroed pivot entry makes the first block ill-conditioned,
    # which amplifies the summation-order differences of the scalar products beyond the 1e-8 bar)
    x, res = g.solve_bcrs_np(bs, v, b)
    x_ref, ref = O.solve_bcrs(rp, ci, v, b)
    assert res["converged"] == 1
    if single:      # float scalar products are summed in another order: a residual next to the threshold may need one more half step
        assert abs(res["half_steps"] - ref["half_steps"]) <= 1 and abs(res["iterations"] - ref["iterations"]) <= 1
    else:
        assert res["iterations"] == ref["iterations"] and res["half_steps"] == ref["half_steps"]
    tol = 1e-2 if single else 1e-8
    assert (np.abs(x - x_ref).max(0) <= tol * np.abs(x_ref).max(0)).all()
    if not single:
        x, res = g.solve_bcrs_np(bs, v, b, linear_solver_reduction=1e-11, linear_solver_maxiter=400)
        assert res["converged"] == 1
        assert np.abs(x - xstar).max() <= 1e-7 * np.abs(xstar).max()
    g.set_precision(False)


def test_np456_general_pattern_and_singular_block(solver, oracle, bs):
    rp, ci, _ = random_bcrs(400, extra_per_row=3, seed=6, dense_group=8)
    v, b, xstar = block_system_np(rp, ci, bs, seed=11)
    O = oracle.instance(False, bs)
    g = solver
    g.set_precision(False)
    g.set_pattern(rp, ci)
    lu_ref, _ = O.ilu0_factor(rp, ci, v)
    lu, got, bad = g.ilu0_np(bs, v, 0.9, b)
    assert np.array_equal(lu, lu_ref) and np.array_equal(got, O.ilu0_apply(rp, ci, lu_ref, 0.9, b))
    x, res = g.solve_bcrs_np(bs, v, b, linear_solver_reduction=1e-8)
    x_ref, ref = O.solve_bcrs(rp, ci, v, b, reduction=1e-8)
    assert res["iterations"] == ref["iterations"] and np.abs(x - x_ref).max() <= 1e-8 * np.abs(x_ref).max()
    # singular pivot: the row the oracle names
    v2 = v.copy()
    row = 37
    for k in range(rp[row], rp[row + 1]):
        if ci[k] <= row:
            v2[k] = 0.0
    _, bad_ref = O.ilu0_factor(rp, ci, v2)
    assert g.ilu0_np(bs, v2)[2] == bad_ref == row


def _csc_blocks_np(rp, ci, v, bs):
    """The bs*bs scalar CSC blocks (Eigen layout) of a BCRS block matrix."""
    import scipy.sparse as sp
    N = rp.size - 1
    out = []
    for p1 in range(bs):
        for p2 in range(bs):
            m = sp.csr_matrix((v[:, p1 * bs + p2], ci, rp), shape=(N, N)).tocsc()
            m.sort_indices()
            out.append((m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data.copy()))
    return out


@pytest.mark.parametrize("single", [False, True], ids=["f64", "f32"])
def test_np456_csc_front_end_and_drop_in(oracle, bs, single):
    import scipy.sparse as sp
    s = synth_blackoil_jacobian(12, 10, 6, perm="homogeneous")
    rp, ci = s.rowptr.numpy(), s.colidx.numpy()
    v, b, _ = block_system_np(rp, ci, bs, seed=20 + bs)
    N = s.N
    blocks = _csc_blocks_np(rp, ci, v, bs)
    scale = [1.0 + 0.1 * p for p in range(bs)]
    rhs = np.ascontiguousarray(b.T).reshape(-1)            # equation-major
    O = oracle.instance(single, bs)
    dx_ref, ref = O.solve_from_csc_blocks(N, blocks, scale, rhs)
    g = GpuLinearSolver(0)
    try:
        g.set_precision(single)
        dx, res = g.solve_from_csc_blocks_np(N, bs, blocks, scale, rhs)
        assert abs(res["iterations"] - ref["iterations"]) <= (1 if single else 0)
        sc = np.abs(dx_ref.reshape(bs, -1)).max(1).repeat(N)
        assert (np.abs(dx - dx_ref) <= (1e-2 if single else 1e-8) * sc).all()
        dx2, res2 = g.solve_from_csc_blocks_np(N, bs, blocks, scale, rhs)      # cached pattern
        assert res2["ms_analysis"] == 0.0 and np.array_equal(dx, dx2)
    finally:
        g.close()
    # the drop-in: LinearisedBlackoilResidual with bs material-balance equations
    eqs = []
    for p1 in range(bs):
        jac = [sp.csc_matrix((blocks[p1 * bs + p2][2], blocks[p1 * bs + p2][1], blocks[p1 * bs + p2][0]), shape=(N, N)) for p2 in range(bs)]
        eqs.append(ADB(rhs[p1 * N:(p1 + 1) * N].copy(), jac))
    solver = NewtonIterationBlackoilGPU({})
    r = LinearisedBlackoilResidual(eqs, matbalscale=scale, singlePrecision=single)
    dx3 = solver.computeNewtonIncrement(r)
    assert abs(solver.iterations() - ref["iterations"]) <= (1 if single else 0) and dx3.size == bs * N
    assert (np.abs(dx3 - dx_ref) <= (1e-2 if single else 1e-8) * sc).all()
