"""ctypes face of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity status: unpinned at the dune-istl level, see oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


class OracleResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("half_steps", C.c_int),
                ("status", C.c_int), ("bad_row", C.c_int), ("reduction", C.c_double),
                ("norm0", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class OracleCsc(C.Structure):
    _fields_ = [("colptr", C.POINTER(C.c_int)), ("rowidx", C.POINTER(C.c_int)),
                ("val", C.POINTER(C.c_double))]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Oracle:
    """One instance of the oracle library: real = double (liboracle.so, the reference's
    Impl<3,double>) or real = float (liboracle_f32.so, Impl<3,float>).  Matrix values, factors and
    vectors cross the boundary as arrays of `real`; the caller-side data of the CSC front end
    (Jacobian values, scaling, equation-major vectors) and all parameters are doubles."""

    def __init__(self, so_name, real_dtype, bs=3):
        self.bs = bs                  # block size np the library was built for (ORACLE_BS)
        self.so = os.path.join(_HERE, "_build", so_name)
        self.real = np.dtype(real_dtype)
        self.c_real = C.c_float if self.real == np.float32 else C.c_double
        self._lib = None

    def lib(self):
        if self._lib is None:
            build()
            if not os.path.exists(self.so):
                build(force=True)
            try:
                self._lib = C.CDLL(self.so)
            except OSError:
                build(force=True)
                self._lib = C.CDLL(self.so)
            L = self._lib
            ip, dp, rp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(self.c_real)
            L.oracle_interleave_pattern.argtypes = [C.c_int, C.c_int, C.POINTER(OracleCsc), C.c_int, ip, C.POINTER(ip)]
            L.oracle_interleave_pattern.restype = C.c_int
            L.oracle_interleave_values.argtypes = [C.c_int, C.c_int, C.POINTER(OracleCsc), dp, ip, ip, rp]
            L.oracle_interleave_values.restype = C.c_int
            L.oracle_spmv3.argtypes = [C.c_int, ip, ip, rp, rp, rp]
            L.oracle_spmv3.restype = None
            L.oracle_ilu0_factor3.argtypes = [C.c_int, ip, ip, rp]
            L.oracle_ilu0_factor3.restype = C.c_int
            L.oracle_ilu0_apply3.argtypes = [C.c_int, ip, ip, rp, C.c_double, rp, rp]
            L.oracle_ilu0_apply3.restype = None
            L.oracle_bicgstab3.argtypes = [C.c_int, ip, ip, rp, rp, C.c_double, rp, rp, C.c_double,
                                           C.c_int, C.c_int, dp, C.c_int, C.POINTER(OracleResult)]
            L.oracle_bicgstab3.restype = None
            L.oracle_solve_bcrs3.argtypes = [C.c_int, ip, ip, rp, rp, rp, C.c_double, C.c_int,
                                             C.c_double, C.c_int, C.POINTER(OracleResult)]
            L.oracle_solve_bcrs3.restype = None
            L.oracle_solve_from_csc_blocks.argtypes = [C.c_int, C.POINTER(OracleCsc), dp, dp, dp,
                                                       C.c_double, C.c_int, C.c_double, C.c_int,
                                                       C.POINTER(OracleResult)]
            L.oracle_solve_from_csc_blocks.restype = None
            L.oracle_gmres3.argtypes = [C.c_int, ip, ip, rp, rp, C.c_double, rp, rp, C.c_double, C.c_int, C.c_int,
                                        dp, C.c_int, C.POINTER(OracleResult)]
            L.oracle_gmres3.restype = None
            L.oracle_solve_gmres_bcrs3.argtypes = [C.c_int, ip, ip, rp, rp, rp, C.c_double, C.c_int, C.c_double, C.c_int,
                                                   C.POINTER(OracleResult)]
            L.oracle_solve_gmres_bcrs3.restype = None
            L.oracle_free.argtypes = [C.c_void_p]
            L.oracle_free.restype = None
        return self._lib

    # arrays of `real` (rounded from whatever the caller has)
    def _r(self, a):
        a = np.ascontiguousarray(a, dtype=self.real)
        return a, a.ctypes.data_as(C.POINTER(self.c_real))

    def _rp(self, a):
        return a.ctypes.data_as(C.POINTER(self.c_real))

    def interleave(self, N, blocks, scale, require_full=False, np_=None):
        """-> rowptr, colidx, vals[nnzb, np*np] of the interleaved BCRS system."""
        np_ = self.bs if np_ is None else np_
        assert np_ == self.bs, "this oracle library was built for another block size"
        arr, keep = _csc_array(blocks)
        rowptr = np.zeros(N + 1, dtype=np.int32)
        out = C.POINTER(C.c_int)()
        nnzb = self.lib().oracle_interleave_pattern(N, np_, arr, int(require_full),
                                                    rowptr.ctypes.data_as(C.POINTER(C.c_int)), C.byref(out))
        colidx = np.ctypeslib.as_array(out, shape=(max(nnzb, 1),))[:nnzb].copy()
        self.lib().oracle_free(out)
        vals = np.zeros((nnzb, np_ * np_), dtype=self.real)
        sc, psc = _d(scale)
        rc = self.lib().oracle_interleave_values(N, np_, arr, psc, rowptr.ctypes.data_as(C.POINTER(C.c_int)),
                                                 colidx.ctypes.data_as(C.POINTER(C.c_int)), self._rp(vals))
        if rc != 0:
            raise ValueError("Jacobian entry outside the pressure-derivative pattern (dune would throw)")
        return rowptr, colidx, vals

    def spmv(self, rowptr, colidx, vals, x):
        rowptr, prp = _i(rowptr); colidx, pci = _i(colidx); vals, pv = self._r(vals); x, px = self._r(x)
        N = rowptr.size - 1
        y = np.zeros(N * self.bs, dtype=self.real)
        self.lib().oracle_spmv3(N, prp, pci, pv, px, self._rp(y))
        return y.reshape(N, self.bs)

    def ilu0_factor(self, rowptr, colidx, vals):
        """-> (lu[nnzb,9] with inverted diagonal blocks, bad_row or -1)."""
        rowptr, prp = _i(rowptr); colidx, pci = _i(colidx)
        lu = np.array(vals, dtype=self.real, copy=True, order="C")
        rc = self.lib().oracle_ilu0_factor3(rowptr.size - 1, prp, pci, self._rp(lu))
        return lu, rc - 1

    def ilu0_apply(self, rowptr, colidx, lu, w, d):
        rowptr, prp = _i(rowptr); colidx, pci = _i(colidx); lu, plu = self._r(lu); d, pd = self._r(d)
        N = rowptr.size - 1
        v = np.zeros(N * self.bs, dtype=self.real)
        self.lib().oracle_ilu0_apply3(N, prp, pci, plu, float(w), pd, self._rp(v))
        return v.reshape(N, self.bs)

    def solve_bcrs(self, rowptr, colidx, vals, rhs, reduction=1e-2, maxiter=150, relax=0.9, max_half_steps=-1):
        rowptr, prp = _i(rowptr); colidx, pci = _i(colidx); vals, pv = self._r(vals); rhs, pr = self._r(rhs)
        N = rowptr.size - 1
        x = np.zeros(N * self.bs, dtype=self.real)
        res = OracleResult()
        self.lib().oracle_solve_bcrs3(N, prp, pci, pv, pr, self._rp(x),
                                      float(reduction), int(maxiter), float(relax), int(max_half_steps),
                                      C.byref(res))
        return x.reshape(N, self.bs), res.as_dict()

    def bicgstab(self, rowptr, colidx, vals, lu, w, rhs, reduction=1e-2, maxiter=150, max_half_steps=-1,
                 history_cap=0):
        """lu=None -> identity preconditioner.  Returns x, result dict, |r| history."""
        rowptr, prp = _i(rowptr); colidx, pci = _i(colidx); vals, pv = self._r(vals)
        N = rowptr.size - 1
        b = np.array(rhs, dtype=self.real, copy=True).reshape(-1)
        x = np.zeros(N * self.bs, dtype=self.real)
        hist = np.zeros(max(history_cap, 1))
        plu = None
        if lu is not None:
            lu, plu = self._r(lu)
        res = OracleResult()
        self.lib().oracle_bicgstab3(N, prp, pci, pv, plu, float(w), self._rp(b),
                                    self._rp(x), float(reduction), int(maxiter),
                                    int(max_half_steps), hist.ctypes.data_as(C.POINTER(C.c_double)),
                                    int(history_cap), C.byref(res))
        return x.reshape(N, self.bs), res.as_dict(), hist[:min(history_cap, res.half_steps)]

    def solve_gmres_bcrs(self, rowptr, colidx, vals, rhs, reduction=1e-2, maxiter=150, relax=0.9, restart=40):
        """ILU0 + Dune::RestartedGMResSolver (newton_use_gmres): x, result dict."""
        rowptr, prp = _i(rowptr); colidx, pci = _i(colidx); vals, pv = self._r(vals); rhs, pr = self._r(rhs)
        N = rowptr.size - 1
        x = np.zeros(N * self.bs, dtype=self.real)
        res = OracleResult()
        self.lib().oracle_solve_gmres_bcrs3(N, prp, pci, pv, pr, self._rp(x),
                                            float(reduction), int(maxiter), float(relax), int(restart), C.byref(res))
        return x.reshape(N, self.bs), res.as_dict()

    def gmres(self, rowptr, colidx, vals, lu, w, rhs, reduction=1e-2, maxiter=150, restart=40, history_cap=0):
        """lu=None -> identity preconditioner.  Returns x, result dict, preconditioned defect history."""
        rowptr, prp = _i(rowptr); colidx, pci = _i(colidx); vals, pv = self._r(vals)
        N = rowptr.size - 1
        b = np.array(rhs, dtype=self.real, copy=True).reshape(-1)
        x = np.zeros(N * self.bs, dtype=self.real)
        hist = np.zeros(max(history_cap, 1))
        plu = None
        if lu is not None:
            lu, plu = self._r(lu)
        res = OracleResult()
        self.lib().oracle_gmres3(N, prp, pci, pv, plu, float(w), self._rp(b),
                                 self._rp(x), float(reduction), int(maxiter), int(restart),
                                 hist.ctypes.data_as(C.POINTER(C.c_double)), int(history_cap), C.byref(res))
        return x.reshape(N, self.bs), res.as_dict(), hist[:min(history_cap, res.half_steps)]

    def solve_from_csc_blocks(self, N, blocks9, matbalscale, rhs_eqmajor, reduction=1e-2, maxiter=150,
                              relax=0.9, require_full=False):
        arr, keep = _csc_array(blocks9)
        sc, psc = _d(matbalscale); rhs, pr = _d(rhs_eqmajor)
        dx = np.zeros(self.bs * N)
        res = OracleResult()
        self.lib().oracle_solve_from_csc_blocks(N, arr, psc, pr, dx.ctypes.data_as(C.POINTER(C.c_double)),
                                                float(reduction), int(maxiter), float(relax),
                                                int(require_full), C.byref(res))
        return dx, res.as_dict()


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _csc_array(blocks):
    keep = []
    arr = (OracleCsc * len(blocks))()
    for k, (cp, ri, v) in enumerate(blocks):
        cp, pcp = _i(cp); ri, pri = _i(ri); v, pv = _d(v)
        keep += [cp, ri, v]
        arr[k].colptr, arr[k].rowidx, arr[k].val = pcp, pri, pv
    return arr, keep


f64 = _Oracle("liboracle.so", np.float64)      # the reference's Impl<3,double>
f32 = _Oracle("liboracle_f32.so", np.float32)  # the reference's Impl<3,float> (singlePrecision)
np2 = _Oracle("liboracle_np2.so", np.float64, bs=2)          # Impl<2,double>: two-phase decks
np2_f32 = _Oracle("liboracle_np2_f32.so", np.float32, bs=2)  # Impl<2,float>
# Impl<4..6,Scalar> (polymer / solvent extensions of the black-oil model)
_npN = {(b, f): _Oracle(f"liboracle_np{b}{'_f32' if f else ''}.so", np.float32 if f else np.float64, bs=b)
        for b in (4, 5, 6) for f in (False, True)}


def instance(single_precision=False, np_=3):
    if np_ == 2:
        return np2_f32 if single_precision else np2
    if np_ in (4, 5, 6):
        return _npN[(np_, bool(single_precision))]
    assert np_ == 3, "the oracle restates block sizes 2..6"
    return f32 if single_precision else f64


# module-level functions: the double instance
lib = f64.lib
interleave = f64.interleave
spmv = f64.spmv
ilu0_factor = f64.ilu0_factor
ilu0_apply = f64.ilu0_apply
solve_bcrs = f64.solve_bcrs
bicgstab = f64.bicgstab
solve_gmres_bcrs = f64.solve_gmres_bcrs
gmres = f64.gmres
solve_from_csc_blocks = f64.solve_from_csc_blocks


_omp = None


def solve_bcrs_openmp(rowptr, colidx, vals, rhs, reduction=1e-2, maxiter=150, relax=0.9, nthreads=0):
    """Baseline B (oracle_omp.c): ILU0 + BiCGStab with OpenMP, level-scheduled.  Returns x, result
    dict with ms_factor / ms_solve / threads."""
    global _omp
    if _omp is None:
        build()
        _omp = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_omp.so"))
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        _omp.oracle_omp_solve_bcrs3.argtypes = [C.c_int, ip, ip, dp, dp, dp, C.c_double, C.c_int, C.c_double, C.c_int,
                                                C.POINTER(OracleResult), dp, dp, ip]
        _omp.oracle_omp_solve_bcrs3.restype = C.c_int
    rowptr, prp = _i(rowptr); colidx, pci = _i(colidx); vals, pv = _d(vals); rhs, pr = _d(rhs)
    N = rowptr.size - 1
    x = np.zeros(N * 3)
    res = OracleResult()
    msf, mss, nt = C.c_double(0), C.c_double(0), C.c_int(0)
    _omp.oracle_omp_solve_bcrs3(N, prp, pci, pv, pr, x.ctypes.data_as(C.POINTER(C.c_double)), float(reduction),
                                int(maxiter), float(relax), int(nthreads), C.byref(res), C.byref(msf), C.byref(mss), C.byref(nt))
    d = res.as_dict()
    d.update(ms_factor=msf.value, ms_solve=mss.value, threads=nt.value)
    return x.reshape(N, 3), d
