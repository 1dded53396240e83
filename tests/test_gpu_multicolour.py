"""GPU parity tests of the multicolour ILU0 variant (OPMGPU_ILU_MULTICOLOUR; include/opm_gpu_solver.h).

The variant is the reference's ILU0 (Dune::bilu0_decomposition, Opm::ParallelOverlappingILU0::apply)
of the symmetric permutation P A P^T that sorts the rows by colour -- the kind of reordering the
reference's own ilu_redblack option asks for (opm/autodiff/ISTLSolver.hpp:207-209).  It is a
DIFFERENT preconditioner than the natural-order one: its iteration counts are compared with the
oracle run on P A P^T, never with the natural-order counts.

Tolerances.  Factors and applies: BIT-IDENTICAL to the oracle on the permuted system (both
precisions).  Solves: equal iteration / half-step counts at the reference's tolerance and increments
within rel 1e-6 (double) / 3e-2 (float), and a true residual within the reference's tolerance.  The
1e-8 bar of the natural-order path does not apply: BiCGStab runs in the caller's ordering, so SpMV
rows AND scalar products are summed in another order than in the oracle's permuted run, and ~15
iterations amplify that (measured 5e-8 in double, 4e-3 in float).  Float solves may differ by one
half step next to the threshold.
"""
import numpy as np
import pytest

from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs
from opm_simulators_legacy_b200.solver import GpuLinearSolver, NumericalIssue, multicolour_order
from test_multicolour_cpu import permute_bcrs

pytestmark = pytest.mark.gpu


def _stencil(dims, perm):
    s = synth_blackoil_jacobian(*dims, perm=perm)
    return s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy().reshape(-1, 3)


def _general(N, seed, dense):
    rp, ci, v = random_bcrs(N, extra_per_row=3, seed=seed, dense_group=dense)
    rng = np.random.default_rng(seed + 100)
    return rp, ci, v, rng.standard_normal((N, 3))


def _diagonal(N):
    rng = np.random.default_rng(5)
    v = rng.standard_normal((N, 9)) + 4.0 * np.eye(3).reshape(1, 9)
    return np.arange(N + 1, dtype=np.int32), np.arange(N, dtype=np.int32), v, rng.standard_normal((N, 3))


CASES = {
    "c1_spe1_shape": lambda: _stencil((10, 10, 3), "homogeneous"),
    "small_lognormal": lambda: _stencil((24, 20, 12), "lognormal"),
    "slab_1d": lambda: _stencil((64, 1, 1), "homogeneous"),
    "odd_sizes": lambda: _stencil((33, 17, 5), "lognormal"),        # colour boundary inside a 64-row tile
    "mid_lognormal": lambda: _stencil((40, 40, 20), "lognormal"),
    "general_sparse": lambda: _general(3000, 3, 0),
    "general_dense_group": lambda: _general(2500, 4, 40),          # a clique: >= 40 colours, most of them tiny
    "diagonal_one_colour": lambda: _diagonal(777),
    "single_row": lambda: _diagonal(1),
}


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    return CASES[request.param]()


@pytest.fixture(scope="module", params=["f64", "f32"])
def mc_solver(request):
    s = GpuLinearSolver(0)
    s.set_precision(request.param == "f32")
    s.set_ilu_ordering(True)
    s.f32 = request.param == "f32"
    yield s
    s.close()


def _permuted(rp, ci, v, n2p):
    prp, pci, pv, order = permute_bcrs(rp, ci, v, n2p)
    return prp, pci, pv, order, np.argsort(n2p)


def test_multicolour_factor_and_apply_bit_exact(mc_solver, oracle, case):
    rp, ci, v, b = case
    orc = oracle.f32 if mc_solver.f32 else oracle
    mc_solver.set_pattern(rp, ci)
    nc, n2p = mc_solver.ilu_permutation()
    nc_ref, colour, n2p_ref = multicolour_order(rp, ci)
    assert nc == nc_ref and np.array_equal(n2p, n2p_ref)
    assert mc_solver.num_levels() == (nc, nc)
    mc_solver.set_values(v)
    assert mc_solver.ilu0_factor() == -1
    prp, pci, pv, order, p2n = _permuted(rp, ci, v, n2p)
    lu_ref, bad = orc.ilu0_factor(prp, pci, pv)
    assert bad == -1
    lu = mc_solver.ilu0_factors()                       # slots of the caller's pattern
    assert np.array_equal(lu[order], lu_ref.astype(np.float64))
    for w in (0.9, 1.0):
        got = mc_solver.ilu0_apply(w, b.reshape(-1))
        ref = orc.ilu0_apply(prp, pci, lu_ref, w, b[p2n].reshape(-1)).astype(np.float64)
        back = np.empty_like(ref)
        back[p2n] = ref
        assert np.array_equal(got.reshape(-1, 3), back)


def test_multicolour_solve_iteration_parity_with_oracle_on_permuted_system(mc_solver, oracle, case):
    rp, ci, v, b = case
    orc = oracle.f32 if mc_solver.f32 else oracle
    mc_solver.set_pattern(rp, ci)
    nc, n2p = mc_solver.ilu_permutation()
    prp, pci, pv, order, p2n = _permuted(rp, ci, v, n2p)
    x, res = mc_solver.solve_bcrs(v, b.reshape(-1))
    xp, ref = orc.solve_bcrs(prp, pci, pv, b[p2n].reshape(-1))
    assert res["converged"] == 1 and res["reduction"] < 1e-2
    if mc_solver.f32:
        assert abs(res["half_steps"] - ref["half_steps"]) <= 1 and abs(res["iterations"] - ref["iterations"]) <= 1
    else:
        assert res["iterations"] == ref["iterations"] and res["half_steps"] == ref["half_steps"]
    x_ref = np.empty((rp.size - 1, 3))
    x_ref[p2n] = xp
    scale = np.abs(x_ref).max(0)
    tol = 3e-2 if mc_solver.f32 else 1e-6
    assert (np.abs(x.reshape(-1, 3) - x_ref).max(0) <= tol * scale).all()
    # the increment solves the caller's system to the reference's tolerance
    r = b.reshape(-1) - oracle.spmv(rp, ci, v, x.reshape(-1)).reshape(-1)
    assert np.linalg.norm(r) <= (1.05e-2 if mc_solver.f32 else 1.0001e-2) * np.linalg.norm(b)
    # GMRES uses the same preconditioner apply (double instance)
    if not mc_solver.f32 and rp.size > 2:
        xg, rg = mc_solver.solve_bcrs(v, b.reshape(-1), newton_use_gmres=True)
        xgp, refg = orc.solve_gmres_bcrs(prp, pci, pv, b[p2n].reshape(-1))
        assert rg["iterations"] == refg["iterations"]


def test_multicolour_csc_front_end_and_switching_back(oracle):
    s = synth_blackoil_jacobian(20, 16, 9, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    g = GpuLinearSolver(0)
    try:
        blocks, scale, rhs_eq = s.csc_blocks(), s.matbalscale, s.rhs_eqmajor_unscaled.numpy()
        dx_nat, res_nat = g.solve_from_csc_blocks(s.N, blocks, scale, rhs_eq)
        g.set_ilu_ordering(True)
        dx_mc, res_mc = g.solve_from_csc_blocks(s.N, blocks, scale, rhs_eq)
        dx_mc2, res_mc2 = g.solve_from_csc_blocks(s.N, blocks, scale, rhs_eq)       # cached pattern
        assert res_mc2["ms_analysis"] == 0 and np.array_equal(dx_mc, dx_mc2)
        # oracle on the permuted interleaved system
        orp, oci, ov = oracle.interleave(s.N, blocks, scale)
        nc, colour, n2p = multicolour_order(orp, oci)
        prp, pci, pv, order = permute_bcrs(orp, oci, ov, n2p)
        p2n = np.argsort(n2p)
        rhs_cell = (rhs_eq.reshape(3, s.N) * np.asarray(scale).reshape(3, 1)).T
        xp, ref = oracle.solve_bcrs(prp, pci, pv, rhs_cell[p2n].reshape(-1))
        assert res_mc["iterations"] == ref["iterations"]
        x_ref = np.empty((s.N, 3)); x_ref[p2n] = xp
        got = dx_mc.reshape(3, s.N).T
        assert (np.abs(got - x_ref).max(0) <= 1e-6 * np.abs(x_ref).max(0)).all()
        # both orderings solve the same system to the same tolerance; the counts are reported side by side
        assert res_nat["converged"] == 1 and res_mc["converged"] == 1
        g.set_ilu_ordering(False)
        dx_back, res_back = g.solve_from_csc_blocks(s.N, blocks, scale, rhs_eq)
        assert np.array_equal(dx_back, dx_nat) and res_back["iterations"] == res_nat["iterations"]
    finally:
        g.close()


def test_multicolour_singular_block_is_reported_in_the_callers_numbering(mc_solver):
    rp, ci, v, b = _stencil((9, 8, 7), "lognormal")
    v = v.copy()
    rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))
    bad = 301
    v[(rows == bad) & (ci == bad)] = 0.0
    v[(rows == bad) & (ci != bad)] = 0.0          # nothing is eliminated into the zero pivot
    mc_solver.set_pattern(rp, ci)
    mc_solver.set_values(v)
    assert mc_solver.ilu0_factor() == bad
    assert f"block row {bad} " in mc_solver.error()
    x, res = mc_solver.solve_bcrs(v, b.reshape(-1), raise_on_failure=False)
    assert res["status"] == 2 and res["bad_row"] == bad
    with pytest.raises(NumericalIssue):
        mc_solver.solve_bcrs(v, b.reshape(-1))


def test_multicolour_is_refused_where_it_is_not_built():
    g = GpuLinearSolver(0)
    try:
        g.set_block_size(2)
        g.set_ilu_ordering(True)
        rp, ci, v, b = _stencil((4, 4, 4), "homogeneous")
        with pytest.raises(ValueError, match="3x3"):
            g.set_pattern(rp, ci)
    finally:
        g.close()


def test_ilu_redblack_key_selects_the_variant_in_the_drop_in(oracle):
    """ilu_redblack (FlowLinearSolverParameters, ISTLSolver.hpp:207-209) behind NewtonIterationBlackoilGPU."""
    import scipy.sparse as sp
    from opm_simulators_legacy_b200.solver import NewtonIterationBlackoilGPU, ADB, LinearisedBlackoilResidual
    s = synth_blackoil_jacobian(14, 11, 7, perm="lognormal")
    blocks, N = s.csc_blocks(), s.N
    rhs = s.rhs_eqmajor_unscaled.numpy()
    eqs = []
    for p1 in range(3):
        jac = [sp.csc_matrix((blocks[p1 * 3 + p2][2], blocks[p1 * 3 + p2][1], blocks[p1 * 3 + p2][0]), shape=(N, N)) for p2 in range(3)]
        eqs.append(ADB(rhs[p1 * N:(p1 + 1) * N].copy(), jac))
    res = LinearisedBlackoilResidual(eqs, matbalscale=s.matbalscale, singlePrecision=False)
    nat = NewtonIterationBlackoilGPU({})
    rb = NewtonIterationBlackoilGPU({"ilu_redblack": "true"})
    dx_nat, dx_rb = nat.computeNewtonIncrement(res), rb.computeNewtonIncrement(res)
    orp, oci, ov = oracle.interleave(N, blocks, s.matbalscale)
    nc, colour, n2p = multicolour_order(orp, oci)
    prp, pci, pv, order = permute_bcrs(orp, oci, ov, n2p)
    p2n = np.argsort(n2p)
    rhs_cell = (rhs.reshape(3, N) * np.asarray(s.matbalscale).reshape(3, 1)).T
    xp, ref = oracle.solve_bcrs(prp, pci, pv, rhs_cell[p2n].reshape(-1))
    assert rb.iterations() == ref["iterations"]
    assert nat.iterations() == oracle.solve_from_csc_blocks(N, blocks, s.matbalscale, rhs)[1]["iterations"]
    # two preconditioners, one system: both increments reduce the true residual by linear_solver_reduction
    for dx in (dx_nat, dx_rb):
        x_cell = np.ascontiguousarray(dx.reshape(3, N).T)
        r = rhs_cell - oracle.spmv(orp, oci, ov, x_cell)
        assert np.linalg.norm(r) <= 1.0001e-2 * np.linalg.norm(rhs_cell)


# ---- k-line ordering (OPMGPU_ILU_MULTICOLOUR_LINES): Cartesian stencils only -------------------------
LINE_CASES = {
    "c1_spe1_shape": ((10, 10, 3), "homogeneous"),
    "small_lognormal": ((24, 20, 12), "lognormal"),
    "odd_sizes": ((33, 17, 5), "lognormal"),            # column counts of the colours differ, not multiples of 4
    "tall": ((7, 6, 40), "lognormal"),                  # more planes than pipeline stages, tiny tiles
    "plane_2d": ((30, 17, 1), "lognormal"),             # one plane: point red-black
    "mid_lognormal": ((40, 40, 20), "lognormal"),
    "wide": ((120, 125, 6), "lognormal"),               # 64-row tiles, several column blocks per SM
}


@pytest.fixture(scope="module", params=list(LINE_CASES))
def line_case(request):
    dims, perm = LINE_CASES[request.param]
    return dims, _stencil(dims, perm)


@pytest.fixture(scope="module", params=["f64", "f32"])
def line_solver(request):
    s = GpuLinearSolver(0)
    s.set_precision(request.param == "f32")
    s.set_ilu_ordering("lines")
    s.f32 = request.param == "f32"
    yield s
    s.close()


def test_lines_factor_apply_bit_exact_and_iteration_parity(line_solver, oracle, line_case):
    from opm_simulators_legacy_b200.solver import line_order
    dims, (rp, ci, v, b) = line_case
    g = line_solver
    orc = oracle.f32 if g.f32 else oracle
    g.set_pattern(rp, ci)
    nc, n2p = g.ilu_permutation()
    assert np.array_equal(n2p, line_order(*dims))
    g.set_values(v)
    assert g.ilu0_factor() == -1
    prp, pci, pv, order, p2n = _permuted(rp, ci, v, n2p)
    lu_ref, bad = orc.ilu0_factor(prp, pci, pv)
    assert bad == -1
    assert np.array_equal(g.ilu0_factors()[order], lu_ref.astype(np.float64))
    for w in (0.9, 1.0):
        got = g.ilu0_apply(w, b.reshape(-1))
        ref = orc.ilu0_apply(prp, pci, lu_ref, w, b[p2n].reshape(-1)).astype(np.float64)
        back = np.empty_like(ref)
        back[p2n] = ref
        assert np.array_equal(got.reshape(-1, 3), back)
    x, res = g.solve_bcrs(v, b.reshape(-1))
    xp, ref = orc.solve_bcrs(prp, pci, pv, b[p2n].reshape(-1))
    assert res["converged"] == 1 and res["reduction"] < 1e-2
    if g.f32:
        assert abs(res["half_steps"] - ref["half_steps"]) <= 1
    else:
        assert res["iterations"] == ref["iterations"] and res["half_steps"] == ref["half_steps"]
    x_ref = np.empty((rp.size - 1, 3))
    x_ref[p2n] = xp
    tol = 3e-2 if g.f32 else 1e-6
    assert (np.abs(x.reshape(-1, 3) - x_ref).max(0) <= tol * np.abs(x_ref).max(0)).all()


def test_lines_refuse_patterns_that_are_not_cartesian_stencils():
    g = GpuLinearSolver(0)
    try:
        g.set_ilu_ordering("lines")
        rp, ci, v, b = _general(500, 3, 6)
        with pytest.raises(ValueError, match="Cartesian"):
            g.set_pattern(rp, ci)
        # a stencil with one extra same-colour coupling (a well's Schur fill between two cells of one column colour)
        s = synth_blackoil_jacobian(8, 8, 4, perm="homogeneous")
        rp, ci = s.rowptr.numpy(), s.colidx.numpy()
        import scipy.sparse as sp
        A = sp.csr_matrix((np.ones(ci.size), ci, rp)).tolil()
        A[0, 2] = 1; A[2, 0] = 1                           # cells (0,0,0) and (2,0,0): same colour
        A = A.tocsr(); A.sort_indices()
        with pytest.raises(ValueError, match="Cartesian"):
            g.set_pattern(A.indptr.astype(np.int32), A.indices.astype(np.int32))
        g.set_ilu_ordering(True)                            # the point colouring takes any pattern
        g.set_pattern(A.indptr.astype(np.int32), A.indices.astype(np.int32))
    finally:
        g.close()
