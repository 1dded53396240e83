"""CPU checks of the column-owned sweep program (colprog.cpp): the record streams, program-order
right-hand sides and hand-over positions, interpreted sequentially on the host exactly as the
kernel indexes them, must reproduce ParallelOverlappingILU0::apply bit for bit."""
import ctypes as C

import numpy as np
import pytest

from opm_simulators_legacy_b200 import _lib
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs


def _host_col_apply(rp, ci, dims, lu, P, w, d):
    lib = _lib.load_experiments()          # the column-owned sweeps exist in the experiments build only
    f = lib.opmgpu_debug_host_col_apply
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    f.argtypes = [C.c_int, ip, ip, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_double, dp, dp, ip]
    f.restype = C.c_int
    rp = np.ascontiguousarray(rp, dtype=np.int32); ci = np.ascontiguousarray(ci, dtype=np.int32)
    lu = np.ascontiguousarray(lu); d = np.ascontiguousarray(d)
    out = np.zeros_like(d); info = np.zeros(8, dtype=np.int32)
    rc = f(len(rp) - 1, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), dims[0], dims[1], dims[2],
           lu.ctypes.data_as(dp), P, w, d.ctypes.data_as(dp), out.ctypes.data_as(dp), info.ctypes.data_as(ip))
    return rc, out, info


@pytest.mark.parametrize("dims", [(10, 10, 3), (24, 20, 12), (64, 1, 1), (30, 17, 1), (7, 6, 40), (3, 3, 3), (1, 1, 9), (33, 47, 5)])
@pytest.mark.parametrize("P", [1, 7, 148])
def test_column_program_is_exact(oracle, dims, P):
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    for w in (0.9, 1.0):
        rc, out, info = _host_col_apply(rp, ci, dims, lu, P, w, b)
        assert rc == 0, rc
        assert info[0] * info[1] <= 32 and info[2] * info[3] <= 6 and info[5] <= P and info[7] >= 3
        assert np.array_equal(out, oracle.ilu0_apply(rp, ci, lu, w, b))


@pytest.mark.parametrize("shape", ["8x4/1x3", "4x8/2x2", "5x6/3x1", "16x2/1x1", "3x3/2x3"])
def test_forced_shapes(oracle, shape, monkeypatch):
    monkeypatch.setenv("OPMGPU_COL_SHAPE", shape)
    dims = (21, 19, 11)
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    rc, out, info = _host_col_apply(rp, ci, dims, lu, 5, 0.9, b)
    assert rc == 0
    assert f"{info[0]}x{info[1]}/{info[2]}x{info[3]}" == shape
    assert np.array_equal(out, oracle.ilu0_apply(rp, ci, lu, 0.9, b))


def test_non_stencil_patterns_have_no_column_program(oracle):
    rp, ci, v = random_bcrs(500, 3, seed=5, dense_group=20)
    lu, bad = oracle.ilu0_factor(rp, ci, v)
    d = np.zeros((500, 3))
    rc, out, info = _host_col_apply(rp, ci, (500, 1, 1), lu, 8, 0.9, d)
    assert rc == -2
    # a stencil with one extra coupling (a well's Schur fill) is not a column program either
    s = synth_blackoil_jacobian(6, 5, 4, perm="lognormal")
    rp, ci = s.rowptr.numpy().copy(), s.colidx.numpy().copy()
    rc, out, info = _host_col_apply(rp, ci, (5, 6, 4), np.zeros((len(ci), 3, 3)), 8, 0.9, np.zeros((120, 3)))
    assert rc == -2
