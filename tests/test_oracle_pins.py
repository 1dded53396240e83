"""Pins for the CPU oracle itself (no GPU).  The reference's tests assert no result at this
boundary (tests/test_linearsolver.cpp:109-121 computes `exact` and drops it), so the oracle is
pinned by (1) the reference's own fixtures turned into known-answer tests, (2) scipy as an
independent implementation, (3) defining properties of ILU0, (4) the committed golden vectors."""
import ctypes
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def block_replicate(rowptr, colidx, data):
    """Scalar CSR -> 3x3 BCRS with each scalar a times a fixed well-conditioned 3x3 matrix."""
    B = np.array([[2.0, 0.5, 0.0], [0.25, 1.5, 0.5], [0.0, 0.75, 3.0]])
    vals = np.asarray(data)[:, None] * B.reshape(1, 9)
    return np.asarray(rowptr, dtype=np.int32), np.asarray(colidx, dtype=np.int32), vals


def laplacian_5pt(N):
    """createLaplacian, tests/test_linearsolver.cpp:50-87."""
    rowptr, col, dat = [0], [], []
    for row in range(N * N):
        x, y = row % N, row // N
        if y > 0: col.append(row - N); dat.append(-1.0)
        if x > 0: col.append(row - 1); dat.append(-1.0)
        col.append(row); dat.append(4.0)
        if x < N - 1: col.append(row + 1); dat.append(-1.0)
        if y < N - 1: col.append(row + N); dat.append(-1.0)
        rowptr.append(len(col))
    return rowptr, col, dat


def c_rand_vector(n):
    """x = (rand() % 100) / 10 with the C library's default seed, tests/test_linearsolver.cpp:92-94."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    return np.array([(libc.rand() % 100) / 10.0 for _ in range(n)])


def test_reference_fixture_laplacian_known_answer(oracle):
    rp, ci, v = block_replicate(*laplacian_5pt(4))
    x_exact = c_rand_vector(16 * 3).reshape(16, 3)
    b = oracle.spmv(rp, ci, v, x_exact)
    x, res = oracle.solve_bcrs(rp, ci, v, b, reduction=1e-12, maxiter=200)
    assert res["converged"] == 1
    assert np.abs(x - x_exact).max() <= 1e-9 * np.abs(x_exact).max()


def test_reference_fixture_1d_laplacian_known_answer(oracle):
    """Owner rows [-1 2 -1], N = 100 (tests/DuneIstlTestHelpers.hpp:132-148).  Block tridiagonal:
    ILU0 is the exact LU, so with w = 1 BiCGStab must converge in one iteration."""
    N = 100
    rowptr, col, dat = [0], [], []
    for r in range(N):
        if r > 0: col.append(r - 1); dat.append(-1.0)
        col.append(r); dat.append(2.0)
        if r < N - 1: col.append(r + 1); dat.append(-1.0)
        rowptr.append(len(col))
    rp, ci, v = block_replicate(rowptr, col, dat)
    x_exact = np.linspace(1.0, 2.0, 3 * N).reshape(N, 3)
    b = oracle.spmv(rp, ci, v, x_exact)
    x, res = oracle.solve_bcrs(rp, ci, v, b, reduction=1e-10, relax=1.0)
    assert res["iterations"] == 1 and res["converged"] == 1
    assert np.abs(x - x_exact).max() <= 1e-10 * np.abs(x_exact).max()
    # with the reference's default relaxation 0.9 the preconditioned operator is 0.9 I: still one iteration
    x, res = oracle.solve_bcrs(rp, ci, v, b, reduction=1e-10, relax=0.9)
    assert res["iterations"] == 1


@pytest.mark.parametrize("dims,perm", [((10, 10, 3), "homogeneous"), ((14, 11, 6), "lognormal")])
def test_against_scipy(oracle, dims, perm):
    s = synth_blackoil_jacobian(*dims, perm=perm)
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    A = sp.bsr_matrix((v.reshape(-1, 3, 3), ci, rp), shape=(3 * s.N, 3 * s.N))
    x = s.xstar.numpy()
    y = oracle.spmv(rp, ci, v, x)
    assert np.abs(y.ravel() - A @ x.ravel()).max() <= 1e-12 * np.abs(A).dot(np.abs(x.ravel())).max()
    xs, res = oracle.solve_bcrs(rp, ci, v, b, reduction=1e-12, maxiter=500)
    xd = spl.spsolve(A.tocsc(), b.ravel()).reshape(-1, 3)
    assert res["converged"] == 1
    assert (np.abs(xs - xd).max(0) <= 1e-6 * np.abs(xd).max(0)).all()
    # the csc-blocks entry point restates ...Interleaved.cpp:234-283 and must give the same system
    dx, r2 = oracle.solve_from_csc_blocks(s.N, s.csc_blocks(), s.matbalscale, s.rhs_eqmajor_unscaled.numpy())
    x1, r1 = oracle.solve_bcrs(rp, ci, v, b)
    assert r2["iterations"] == r1["iterations"] and np.array_equal(dx.reshape(3, -1).T, x1)
    rp2, ci2, v2 = oracle.interleave(s.N, s.csc_blocks(), s.matbalscale)
    assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and np.array_equal(v2, v)


def test_ilu0_defining_property(oracle):
    """(L U)_ij = A_ij on the sparsity pattern."""
    rp, ci, v = random_bcrs(150, extra_per_row=3, seed=5, dense_group=6)
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    assert bad == -1
    N = len(rp) - 1
    L = np.zeros((3 * N, 3 * N)); U = np.zeros((3 * N, 3 * N)); A = np.zeros((3 * N, 3 * N))
    for i in range(N):
        for k in range(rp[i], rp[i + 1]):
            j = ci[k]
            blk = lu[k].reshape(3, 3)
            A[3 * i:3 * i + 3, 3 * j:3 * j + 3] = v[k].reshape(3, 3)
            if j < i: L[3 * i:3 * i + 3, 3 * j:3 * j + 3] = blk
            elif j == i: U[3 * i:3 * i + 3, 3 * j:3 * j + 3] = np.linalg.inv(blk)
            else: U[3 * i:3 * i + 3, 3 * j:3 * j + 3] = blk
    L += np.eye(3 * N)
    P = L @ U
    mask = A != 0
    assert np.abs(P - A)[mask].max() <= 1e-10 * np.abs(A).max()
    # and the apply is w * U^-1 L^-1 d
    d = np.random.default_rng(1).standard_normal((N, 3))
    ref = 0.9 * np.linalg.solve(U, np.linalg.solve(L, d.ravel()))
    got = oracle.ilu0_apply(rp, ci, lu, 0.9, d).ravel()
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()


def test_singular_pivot_reported(oracle):
    s = synth_blackoil_jacobian(5, 4, 3)
    rp, ci, v = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy().copy()
    row = 7
    for k in range(rp[row], rp[row + 1]):
        if ci[k] <= row:
            v[k] = 0.0
    _, bad = oracle.ilu0_factor(rp, ci, v)
    assert bad == row
    _, res = oracle.solve_bcrs(rp, ci, v, s.rhs.numpy())
    assert res["status"] == 2 and res["bad_row"] == row


def test_dune_iteration_counting(oracle):
    """res.iterations = ceil(it): a stop after the first half step of iteration k reports k."""
    s = synth_blackoil_jacobian(9, 8, 5, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    x, res = oracle.solve_bcrs(rp, ci, v, b)
    assert res["iterations"] == (res["half_steps"] + 1) // 2
    _, r0 = oracle.solve_bcrs(rp, ci, v, np.zeros_like(b))
    assert r0["iterations"] == 0 and r0["converged"] == 1
    _, rmax = oracle.solve_bcrs(rp, ci, v, b, reduction=1e-30, maxiter=3)
    assert rmax["iterations"] == 3 and rmax["converged"] == 0 and rmax["status"] == 1


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(oracle, path):
    g = np.load(path)
    rp, ci, v, b = g["rowptr"], g["colidx"], g["vals"], g["rhs"]
    assert np.array_equal(oracle.spmv(rp, ci, v, g["x_probe"]), g["spmv"])
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    assert bad == -1 and np.array_equal(lu, g["lu"])
    assert np.array_equal(oracle.ilu0_apply(rp, ci, lu, 0.9, b), g["apply_w09"])
    assert np.array_equal(oracle.ilu0_apply(rp, ci, lu, 1.0, b), g["apply_w1"])
    x, res = oracle.solve_bcrs(rp, ci, v, b)
    assert res["iterations"] == int(g["iterations"]) and np.array_equal(x, g["x"])


# ---- restarted GMRES (newton_use_gmres; Dune::RestartedGMResSolver restated in oracle_gmres3) ----
def test_gmres_matches_a_direct_solve_also_across_restarts():
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian
    from oracle import oracle_py as o
    s = synth_blackoil_jacobian(12, 10, 8, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    A = sp.bsr_matrix((v.reshape(-1, 3, 3), ci, rp)).tocsc()
    xs = spla.spsolve(A, b.reshape(-1))
    for restart, tol in ((40, 1e-9), (5, 1e-6)):
        x, res = o.solve_gmres_bcrs(rp, ci, v, b, reduction=1e-12, maxiter=2000, restart=restart)
        assert res["converged"] == 1 and res["status"] == 0
        assert np.abs(x.reshape(-1) - xs).max() <= tol * np.abs(xs).max()
    # restart = 5 needs more Arnoldi steps than one long cycle
    assert o.solve_gmres_bcrs(rp, ci, v, b, reduction=1e-8, maxiter=2000, restart=5)[1]["iterations"] > \
        o.solve_gmres_bcrs(rp, ci, v, b, reduction=1e-8, maxiter=2000, restart=200)[1]["iterations"]


def test_gmres_known_answers():
    """Block-tridiagonal system: ILU0 is the exact LU, so W^-1 A = I/w and GMRES stops after one
    Arnoldi step with x = A^-1 b; maxiter reached -> status 1 with iterations == maxiter."""
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian
    from oracle import oracle_py as o
    s = synth_blackoil_jacobian(40, 1, 1, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    x, res = o.solve_gmres_bcrs(rp, ci, v, b, reduction=1e-10, relax=1.0)
    assert res["iterations"] == 1 and res["converged"] == 1
    assert np.abs(x - s.xstar.numpy()).max() <= 1e-9 * np.abs(s.xstar.numpy()).max()
    s2 = synth_blackoil_jacobian(10, 10, 6, perm="lognormal")
    rp, ci, v, b = s2.rowptr.numpy(), s2.colidx.numpy(), s2.vals.numpy(), s2.rhs.numpy()
    x, res = o.solve_gmres_bcrs(rp, ci, v, b, reduction=1e-14, maxiter=3)
    assert res["iterations"] == 3 and res["converged"] == 0 and res["status"] == 1


# ---- the float instance of the oracle (Impl<3,float>, liboracle_f32.so) -----------------------------
GOLDEN_F32 = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "f32", "*.npz")))


def test_float_oracle_is_the_same_algorithm_in_float(oracle):
    """Pins of the float build: (1) the interleaved values are the double ones rounded ONCE (the
    assignment to a float matrix, ...Interleaved.cpp:189); (2) SpMV and ILU0 apply of float-representable
    data agree with the double oracle to float accuracy and are float-valued; (3) block-tridiagonal
    system: ILU0 is the exact LU, one iteration also in float; (4) iteration count at the reference's
    tolerance equals the double instance's on a well-conditioned case."""
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian
    s = synth_blackoil_jacobian(12, 9, 7, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    blocks = s.csc_blocks()
    rp64, ci64, v64 = oracle.interleave(s.N, blocks, s.matbalscale)
    rp32, ci32, v32 = oracle.f32.interleave(s.N, blocks, s.matbalscale)
    assert v32.dtype == np.float32 and np.array_equal(rp32, rp64) and np.array_equal(ci32, ci64)
    assert np.array_equal(v32, v64.astype(np.float32))
    x = s.xstar.numpy()
    y32, y64 = oracle.f32.spmv(rp, ci, v, x), oracle.spmv(rp, ci, v, x)
    assert y32.dtype == np.float32
    assert np.abs(y32 - y64).max() <= 2e-5 * np.abs(y64).max()
    lu32, bad = oracle.f32.ilu0_factor(rp, ci, v)
    lu64, _ = oracle.ilu0_factor(rp, ci, v)
    assert bad == -1 and np.abs(lu32 - lu64).max() <= 1e-4 * np.abs(lu64).max()
    a32, a64 = oracle.f32.ilu0_apply(rp, ci, lu32, 0.9, b), oracle.ilu0_apply(rp, ci, lu64, 0.9, b)
    assert np.abs(a32 - a64).max() <= 1e-3 * np.abs(a64).max()
    x32, r32 = oracle.f32.solve_bcrs(rp, ci, v, b)
    x64, r64 = oracle.solve_bcrs(rp, ci, v, b)
    assert r32["iterations"] == r64["iterations"] and r32["converged"] == 1
    t = synth_blackoil_jacobian(40, 1, 1, perm="lognormal")
    rpt, cit, vt, bt = t.rowptr.numpy(), t.colidx.numpy(), t.vals.numpy(), t.rhs.numpy()
    xt, rt = oracle.f32.solve_bcrs(rpt, cit, vt, bt, reduction=1e-4, relax=1.0)
    assert rt["iterations"] == 1 and rt["converged"] == 1
    assert np.abs(xt - t.xstar.numpy()).max() <= 1e-3 * np.abs(t.xstar.numpy()).max()


@pytest.mark.parametrize("path", GOLDEN_F32, ids=[os.path.basename(p)[:-4] for p in GOLDEN_F32])
def test_float_oracle_reproduces_golden(oracle, path):
    g = np.load(path)
    rp, ci, v, b = g["rowptr"], g["colidx"], g["vals"], g["rhs"]
    F = oracle.f32
    assert np.array_equal(F.spmv(rp, ci, v, g["x_probe"]), g["spmv"])
    lu, bad = F.ilu0_factor(rp, ci, v)
    assert bad == -1 and np.array_equal(lu, g["lu"])
    assert np.array_equal(F.ilu0_apply(rp, ci, lu, 0.9, b), g["apply_w09"])
    assert np.array_equal(F.ilu0_apply(rp, ci, lu, 1.0, b), g["apply_w1"])
    x, res = F.solve_bcrs(rp, ci, v, b)
    assert res["iterations"] == int(g["iterations"]) and np.array_equal(x, g["x"])


# ---- block size 2 (liboracle_np2*.so: the same file with -DORACLE_BS=2) ------------------------------
def test_np2_oracle_pins(oracle):
    """scipy bsr SpMV, a direct solve at a tight tolerance, the defining property (LU)_ij = A_ij on
    the pattern, the block-tridiagonal known answer (ILU0 exact: one iteration), and the float build."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian
    s = synth_blackoil_jacobian(10, 9, 5, perm="lognormal")
    rp, ci = s.rowptr.numpy(), s.colidx.numpy()
    v = np.ascontiguousarray(s.vals.numpy()[:, [0, 1, 3, 4]])
    v[np.repeat(np.arange(s.N), np.diff(rp)) == ci] *= 1.3      # (the water / oil sub-system alone is close to singular)
    b = np.ascontiguousarray(s.rhs.numpy()[:, :2])
    N = s.N
    A = sp.bsr_matrix((v.reshape(-1, 2, 2), ci, rp), shape=(2 * N, 2 * N))
    x = np.ascontiguousarray(s.xstar.numpy()[:, :2])
    y = oracle.np2.spmv(rp, ci, v, x)
    assert np.abs(y.ravel() - A @ x.ravel()).max() <= 1e-12 * np.abs(y).max()
    xs, res = oracle.np2.solve_bcrs(rp, ci, v, b, reduction=1e-12, maxiter=400)
    assert res["converged"] == 1
    ref = spl.spsolve(A.tocsc(), b.ravel())
    assert np.abs(xs.ravel() - ref).max() <= 1e-8 * np.abs(ref).max()
    # (LU)_ij = A_ij on the pattern: L unit lower (stored blocks), U upper, diagonal stored inverted
    lu, bad = oracle.np2.ilu0_factor(rp, ci, v)
    assert bad == -1
    Lm = sp.lil_matrix((2 * N, 2 * N)); Um = sp.lil_matrix((2 * N, 2 * N))
    for i in range(N):
        for k in range(rp[i], rp[i + 1]):
            j = ci[k]; blk = lu[k].reshape(2, 2)
            if j < i:
                Lm[2 * i:2 * i + 2, 2 * j:2 * j + 2] = blk
            elif j == i:
                Um[2 * i:2 * i + 2, 2 * i:2 * i + 2] = np.linalg.inv(blk)
                Lm[2 * i:2 * i + 2, 2 * i:2 * i + 2] = np.eye(2)
            else:
                Um[2 * i:2 * i + 2, 2 * j:2 * j + 2] = blk
    P = (Lm.tocsr() @ Um.tocsr()).toarray()
    Ad = A.toarray()
    mask = np.kron((sp.csr_matrix((np.ones(len(ci)), ci, rp), shape=(N, N)).toarray() != 0), np.ones((2, 2))) != 0
    assert np.abs(P - Ad)[mask].max() <= 1e-10 * np.abs(Ad).max()
    t = synth_blackoil_jacobian(40, 1, 1, perm="lognormal")
    vt = np.ascontiguousarray(t.vals.numpy()[:, [0, 1, 3, 4]]); bt = np.ascontiguousarray(t.rhs.numpy()[:, :2])
    _, rt = oracle.np2.solve_bcrs(t.rowptr.numpy(), t.colidx.numpy(), vt, bt, reduction=1e-10, relax=1.0)
    assert rt["iterations"] == 1 and rt["converged"] == 1
    x32, r32 = oracle.np2_f32.solve_bcrs(rp, ci, v, b)
    x64, r64 = oracle.np2.solve_bcrs(rp, ci, v, b)
    assert x32.dtype == np.float32 and r32["iterations"] == r64["iterations"]


# ---- block sizes 4..6 (liboracle_np<bs>*.so: the same file with -DORACLE_BS=4 / 5 / 6) ----------------
@pytest.mark.parametrize("bs", [4, 5, 6])
def test_np456_oracle_pins(oracle, bs):
    """The block inverse of each size against numpy (4x4: OPM's closed form; 5, 6: dune's LU with
    thresholded row pivoting, including a block that needs the row swaps and a singular one), scipy
    bsr SpMV, a direct solve, the block-tridiagonal known answer (ILU0 exact), and the float build."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, block_system_np
    O, O32 = oracle.instance(False, bs), oracle.instance(True, bs)
    rng = np.random.default_rng(bs)
    one_rp, one_ci = np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32)
    # ILU0 of a 1 x 1 block matrix = the block inverse
    for trial in range(20):
        A = rng.standard_normal((bs, bs)) + (3.0 if trial % 2 else 0.0) * np.eye(bs)
        if trial >= 16:
            A[0, 0] = 0.0                       # forces dune's row swap at the first pivot (no-op for the 4x4 closed form)
        if trial == 19:
            A[2, 0] = A[1, 0] = 0.0             # and again at a later column
        lu, bad = O.ilu0_factor(one_rp, one_ci, A.reshape(1, -1))
        assert bad == -1
        assert np.abs(lu.reshape(bs, bs) @ A - np.eye(bs)).max() <= 1e-10 * np.linalg.cond(A)
        assert np.abs(lu.reshape(bs, bs) - np.linalg.inv(A)).max() <= 1e-10 * np.linalg.cond(A) * np.abs(np.linalg.inv(A)).max()
    S = np.ones((bs, bs))                        # singular: reported as the failing row
    _, bad = O.ilu0_factor(one_rp, one_ci, S.reshape(1, -1))
    assert bad == 0
    s = synth_blackoil_jacobian(9, 8, 5, perm="homogeneous")
    rp, ci = s.rowptr.numpy(), s.colidx.numpy()
    v, b, xstar = block_system_np(rp, ci, bs, seed=3)
    N = s.N
    A = sp.bsr_matrix((v.reshape(-1, bs, bs), ci, rp), shape=(bs * N, bs * N))
    y = O.spmv(rp, ci, v, xstar)
    assert y.shape == (N, bs) and np.abs(y.ravel() - A @ xstar.ravel()).max() <= 1e-12 * np.abs(y).max()
    xs, res = O.solve_bcrs(rp, ci, v, b, reduction=1e-12, maxiter=400)
    assert res["converged"] == 1
    assert np.abs(xs - xstar).max() <= 1e-8 * np.abs(xstar).max()
    ref = spl.spsolve(A.tocsc(), b.ravel())
    assert np.abs(xs.ravel() - ref).max() <= 1e-8 * np.abs(ref).max()
    # block tridiagonal: ILU0 is the exact factorisation, one iteration with w = 1
    t = synth_blackoil_jacobian(30, 1, 1, perm="homogeneous")
    vt, bt, _ = block_system_np(t.rowptr.numpy(), t.colidx.numpy(), bs, seed=4)
    _, rt = O.solve_bcrs(t.rowptr.numpy(), t.colidx.numpy(), vt, bt, reduction=1e-10, relax=1.0)
    assert rt["iterations"] == 1 and rt["converged"] == 1
    x32, r32 = O32.solve_bcrs(rp, ci, v, b)
    x64, r64 = O.solve_bcrs(rp, ci, v, b)
    assert x32.dtype == np.float32 and r32["iterations"] == r64["iterations"]
    assert np.abs(x32 - x64).max() <= 1e-2 * np.abs(x64).max()
