"""Multicolour ILU0 variant, host side (no GPU): the ordering rule exported by libopmgpu.so
(opmgpu_multicolour_order) against its definition, and the oracle on the permuted system (what the
GPU parity tests of the variant compare with).  The variant is the reference's ILU0 of P A P^T
(compare the reference's own ilu_redblack option, opm/autodiff/ISTLSolver.hpp:207-209): its iteration
counts are NOT the reference's natural-order counts and are never reported as parity."""
import numpy as np
import pytest

from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs
from opm_simulators_legacy_b200.solver import multicolour_order, line_order


def greedy_reference(rp, ci):
    """Definition: natural row order, smallest colour no already coloured neighbour (either direction) has."""
    N = rp.size - 1
    nbr = [set() for _ in range(N)]
    for i in range(N):
        for k in range(rp[i], rp[i + 1]):
            j = int(ci[k])
            if j != i:
                nbr[i].add(j); nbr[j].add(i)
    colour = -np.ones(N, dtype=np.int32)
    for i in range(N):
        used = {int(colour[j]) for j in nbr[i] if j < i}
        c = 0
        while c in used:
            c += 1
        colour[i] = c
    return colour, nbr


def permute_bcrs(rp, ci, v, n2p):
    """BCRS of P A P^T (rows and columns renumbered by n2p, columns ascending) and, for every slot of
    the permuted pattern, the slot of the original pattern it came from."""
    N = rp.size - 1
    rows = np.repeat(np.arange(N), np.diff(rp))
    pr, pc = n2p[rows], n2p[ci]
    order = np.lexsort((pc, pr))
    prp = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(np.bincount(pr, minlength=N), out=prp[1:])
    return prp, pc[order].astype(np.int32), np.ascontiguousarray(v[order]), order


def test_cartesian_stencil_gets_red_black():
    s = synth_blackoil_jacobian(7, 6, 5, perm="homogeneous")
    rp, ci = s.rowptr.numpy(), s.colidx.numpy()
    nc, colour, n2p = multicolour_order(rp, ci)
    assert nc == 2
    i, j, k = np.meshgrid(np.arange(7), np.arange(6), np.arange(5), indexing="ij")
    cell = (i + 7 * (j + 6 * k)).ravel()
    expect = np.empty(cell.size, dtype=np.int32)
    expect[cell] = ((i + j + k) % 2).ravel()
    assert np.array_equal(colour, expect)


def test_ordering_follows_its_definition_on_general_patterns():
    for seed, dense in ((1, 0), (2, 9), (3, 17)):
        rp, ci, _ = random_bcrs(300, extra_per_row=2, seed=seed, dense_group=dense)
        # make the pattern structurally nonsymmetric: drop a few strictly upper entries
        rows = np.repeat(np.arange(300), np.diff(rp))
        keep = ~((ci > rows) & ((rows * 7 + ci) % 5 == 0))
        ci2 = ci[keep]
        rp2 = np.zeros(301, dtype=np.int32)
        np.cumsum(np.bincount(rows[keep], minlength=300), out=rp2[1:])
        nc, colour, n2p = multicolour_order(rp2, ci2)
        ref, nbr = greedy_reference(rp2, ci2)
        assert np.array_equal(colour, ref) and nc == ref.max() + 1
        if dense:
            assert nc >= dense                      # a clique needs a colour per member
        for i in range(300):
            assert all(colour[j] != colour[i] for j in nbr[i])
        # rows sorted by (colour, natural index)
        p2n = np.argsort(n2p, kind="stable")
        assert np.array_equal(np.sort(n2p), np.arange(300))
        key = colour[p2n].astype(np.int64) * 300 + p2n
        assert (np.diff(key) > 0).all()


def test_oracle_on_the_permuted_system_solves_the_same_problem(oracle):
    s = synth_blackoil_jacobian(12, 10, 8, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    nc, colour, n2p = multicolour_order(rp, ci)
    p2n = np.argsort(n2p)
    prp, pci, pv, order = permute_bcrs(rp, ci, v, n2p)
    xp, resp = oracle.solve_bcrs(prp, pci, pv, b.reshape(-1, 3)[p2n], reduction=1e-10)
    xn, resn = oracle.solve_bcrs(rp, ci, v, b, reduction=1e-10)
    assert resp["converged"] and resn["converged"]
    x = np.empty_like(xp)
    x[p2n] = xp
    assert np.abs(x - xn).max() <= 1e-6 * np.abs(xn).max()
    # a different preconditioner: the counts need not agree (and do not on this system)
    assert resp["iterations"] != resn["iterations"]


def test_line_order_is_red_black_over_columns_with_natural_order_inside(oracle):
    nx, ny, nz = 7, 5, 6
    n2p = line_order(nx, ny, nz)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    nat = np.arange(nx * ny * nz)
    p2n_ref = np.lexsort((nat, (i + j) % 2))                # colour, then natural index (k-major inside a colour)
    assert np.array_equal(np.argsort(n2p), p2n_ref)
    # a column's cells keep their order and are one colour-plane stride apart
    ncols = [((i + j) % 2 == c)[:nx * ny].sum() for c in (0, 1)]
    col = 3 + nx * 2                                        # column (3, 2): colour 1
    q = n2p[col + nx * ny * np.arange(nz)]
    assert (np.diff(q) == ncols[1]).all() and q[0] >= nz * ncols[0]
    # on a reservoir-like system (strong vertical couplings) the ordering is at least as good as point red-black
    s = synth_blackoil_jacobian(14, 12, 10, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    n2p = line_order(14, 12, 10)
    p2n = np.argsort(n2p)
    prp, pci, pv, order = permute_bcrs(rp, ci, v, n2p)
    xl, rl = oracle.solve_bcrs(prp, pci, pv, b.reshape(-1, 3)[p2n])
    nc, colour, n2p_rb = multicolour_order(rp, ci)
    prp, pci, pv, order = permute_bcrs(rp, ci, v, n2p_rb)
    xr, rr = oracle.solve_bcrs(prp, pci, pv, b.reshape(-1, 3)[np.argsort(n2p_rb)])
    assert rl["converged"] and rr["converged"] and rl["half_steps"] <= rr["half_steps"]


def _host_mc(rp, ci, v, lines, w, d):
    """opmgpu_debug_host_mc_apply: the multicolour program (permutation, [L | Dinv | U] layout, update
    lists) interpreted sequentially on the host exactly as the kernels index it."""
    import ctypes as C
    from opm_simulators_legacy_b200 import _lib
    lib = _lib.load()
    f = lib.opmgpu_debug_host_mc_apply
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    f.argtypes = [C.c_int, ip, ip, dp, C.c_int, C.c_double, dp, dp, dp, ip]
    f.restype = C.c_int
    rp = np.ascontiguousarray(rp, dtype=np.int32); ci = np.ascontiguousarray(ci, dtype=np.int32)
    v = np.ascontiguousarray(v, dtype=np.float64); d = np.ascontiguousarray(d, dtype=np.float64)
    out = np.zeros_like(d); lu = np.zeros_like(v); info = np.zeros(2, dtype=np.int32)
    rc = f(rp.size - 1, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), v.ctypes.data_as(dp), int(lines), float(w),
           d.ctypes.data_as(dp), out.ctypes.data_as(dp), lu.ctypes.data_as(dp), info.ctypes.data_as(ip))
    return rc, out, lu, info


@pytest.mark.parametrize("lines", [False, True], ids=["points", "k-lines"])
@pytest.mark.parametrize("dims", [(10, 10, 3), (9, 7, 12), (33, 17, 5), (30, 17, 1), (64, 1, 1)])
def test_multicolour_program_is_exact_on_the_host(oracle, dims, lines):
    """Host-side data of both orderings against the oracle on P A P^T, bit for bit: permutation, L / U split
    (U stored in descending column order), unified factor layout, update lists of the factorisation."""
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy().reshape(-1, 3)
    n2p = line_order(*dims) if lines else multicolour_order(rp, ci)[2]
    p2n = np.argsort(n2p)
    prp, pci, pv, order = permute_bcrs(rp, ci, v, n2p)
    lu_ref, bad = oracle.ilu0_factor(prp, pci, pv)
    assert bad == -1
    for w in (0.9, 1.0):
        rc, out, lu, info = _host_mc(rp, ci, v, lines, w, b.reshape(-1))
        assert rc == 0
        assert np.array_equal(lu[order], lu_ref)
        ref = oracle.ilu0_apply(prp, pci, lu_ref, w, b[p2n].reshape(-1))
        back = np.empty_like(ref); back[p2n] = ref
        assert np.array_equal(out.reshape(-1, 3), back)


def test_multicolour_program_general_pattern_and_refusals(oracle):
    rp, ci, v = random_bcrs(600, extra_per_row=3, seed=9, dense_group=15)
    b = np.random.default_rng(1).standard_normal((600, 3))
    nc, colour, n2p = multicolour_order(rp, ci)
    p2n = np.argsort(n2p)
    prp, pci, pv, order = permute_bcrs(rp, ci, v, n2p)
    lu_ref, bad = oracle.ilu0_factor(prp, pci, pv)
    rc, out, lu, info = _host_mc(rp, ci, v, False, 0.9, b.reshape(-1))
    assert rc == 0 and info[0] == nc >= 15
    assert np.array_equal(lu[order], lu_ref)               # update lists that touch off-diagonal blocks too
    ref = oracle.ilu0_apply(prp, pci, lu_ref, 0.9, b[p2n].reshape(-1))
    back = np.empty_like(ref); back[p2n] = ref
    assert np.array_equal(out.reshape(-1, 3), back)
    # the k-line ordering refuses what is not a Cartesian stencil with vertical same-colour couplings only
    assert _host_mc(rp, ci, v, True, 0.9, b.reshape(-1))[0] == -2
    # a singular pivot is reported in the caller's numbering
    s = synth_blackoil_jacobian(6, 5, 4, perm="homogeneous")
    rp, ci, v = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy().copy()
    rows = np.repeat(np.arange(s.N), np.diff(rp))
    v[rows == 17] = 0.0
    for lines in (False, True):
        assert _host_mc(rp, ci, v, lines, 0.9, s.rhs.numpy().reshape(-1))[0] == 1 + 17
