"""Host-side mirror of the reference's linear-solver boundary over the C ABI.

`NewtonIterationBlackoilGPU` has the interface of NewtonIterationBlackoilInterface
(opm/autodiff/NewtonIterationBlackoilInterface.hpp:31-52): computeNewtonIncrement(residual),
iterations(), parallelInformation(); it is constructed from the same parameter keys as
NewtonIterationBlackoilInterleaved (opm/autodiff/NewtonIterationBlackoilInterleaved.hpp:55-56)
and raises what the reference throws (LinearSolverProblem on non-convergence,
opm/autodiff/ISTLSolver.hpp:358-368).  The C++ twin that compiles against real OPM headers is
csrc/host/NewtonIterationBlackoilGPU.hpp; this Python twin exists so the parity tests read
like reference tests.  `GpuLinearSolver` is the plain handle wrapper used by tests and bench.

Everything numerical happens in libopmgpu.so on the GPU; nothing here computes a solve on
the CPU except the well Schur elimination / recovery, which the reference also keeps on the
host (opm/autodiff/NewtonIterationUtilities.cpp:45-184).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L


class LinearSolverProblem(RuntimeError):
    """Opm::LinearSolverProblem -- caught by AdaptiveTimeStepping, which chops the step
    (opm/simulators/timestepping/AdaptiveTimeStepping_impl.hpp:251-257)."""


class NumericalIssue(RuntimeError):
    """Opm::NumericalIssue / Dune::MatrixBlockError class of failures (singular ILU0 pivot,
    BiCGStab breakdown); also caught by AdaptiveTimeStepping (:258-282)."""


_PARAM_KEYS = {
    # key -> (field, default, cast); defaults are FlowLinearSolverParameters::reset() (2019.04)
    "linear_solver_reduction": ("linear_solver_reduction", 1e-2, float),
    "linear_solver_maxiter": ("linear_solver_maxiter", 150, int),
    "ilu_relaxation": ("ilu_relaxation", 0.9, float),
    "linear_solver_verbosity": ("linear_solver_verbosity", 0, int),
    "linear_solver_ignoreconvergencefailure": ("linear_solver_ignoreconvergencefailure", False, bool),
    "require_full_sparsity_pattern": ("require_full_sparsity_pattern", False, bool),
    "newton_use_gmres": ("newton_use_gmres", False, bool),
    "linear_solver_restart": ("linear_solver_restart", 40, int),
}
# keys of FlowLinearSolverParameters this solver accepts only at their default value
_UNSUPPORTED = {"linear_solver_use_amg": False, "ilu_fillin_level": 0,
                "ilu_milu": "ILU"}


def make_params(param: Optional[dict] = None, **kw) -> L.Params:
    """ParameterGroup-style construction: param.getDefault(key, default) for every key."""
    lib = L.load()
    p = L.Params()
    lib.opmgpu_default_params(C.byref(p))
    merged = dict(param or {})
    merged.update(kw)
    for key, val in merged.items():
        if key in _PARAM_KEYS:
            fld, _, cast = _PARAM_KEYS[key]
            if cast is bool and isinstance(val, str):
                val = val.lower() in ("1", "true", "yes")
            setattr(p, fld, int(val) if cast is bool else cast(val))
        elif key == "max_half_steps":
            p.max_half_steps = int(val)
        elif key in _UNSUPPORTED:
            if str(val).lower() != str(_UNSUPPORTED[key]).lower():
                raise ValueError(f"{key}={val} is not supported by solver_approach=gpu "
                                 f"(only {key}={_UNSUPPORTED[key]})")
        # other keys belong to other components; ParameterGroup reports unused keys at exit
    return p


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def multicolour_order(rowptr, colidx):
    """The ordering rule of the multicolour ILU0 variant alone (host code of the library, no GPU):
    returns (ncolours, colour[N], n2p[N])."""
    lib = L.load()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    N = rowptr.size - 1
    nc = C.c_int(0)
    colour = np.empty(N, dtype=np.int32)
    n2p = np.empty(N, dtype=np.int32)
    rc = lib.opmgpu_multicolour_order(N, _ip(rowptr), _ip(colidx), C.byref(nc), _ip(colour), _ip(n2p))
    if rc != L.OK:
        raise ValueError("opmgpu_multicolour_order: bad pattern")
    return nc.value, colour, n2p


def line_order(nx, ny, nz):
    """n2p of the k-line ordering (red-black over the (i,j) columns, natural order along k); host only."""
    lib = L.load()
    n2p = np.empty(nx * ny * nz, dtype=np.int32)
    if lib.opmgpu_line_order(int(nx), int(ny), int(nz), _ip(n2p)) != L.OK:
        raise ValueError("opmgpu_line_order: bad dimensions")
    return n2p


class GpuLinearSolver:
    """One opmgpu handle (one GPU, one stream)."""

    def __init__(self, device: int = 0, experiments: bool = False):
        # experiments: bind libopmgpu_exp.so (A/B tools and tests of experimental kernel variants only)
        self.lib = L.load_experiments() if experiments else L.load()
        self.h = C.c_void_p()
        rc = self.lib.opmgpu_create(int(device), C.byref(self.h))
        if rc != L.OK:
            raise RuntimeError("opmgpu_create failed: " + self.lib.opmgpu_last_error(None).decode())
        self.N = 0
        self.nnzb = 0
        self.last = None

    @classmethod
    def multi(cls, devices):
        """One handle, several GPUs of this process (opmgpu_create_multi): the partition, the halo plan and
        the per-GPU worker threads live beneath the C ABI; set_pattern / solve_bcrs /
        solve_from_csc_blocks take and return GLOBAL arrays."""
        self = cls.__new__(cls)
        self.lib = L.load()
        self.h = C.c_void_p()
        dev = np.ascontiguousarray(list(devices), dtype=np.int32)
        rc = self.lib.opmgpu_create_multi(int(dev.size), _ip(dev), C.byref(self.h))
        if rc != L.OK:
            raise RuntimeError("opmgpu_create_multi failed: " + self.lib.opmgpu_last_error(None).decode())
        self.N = self.nnzb = 0
        self.last = None
        self.ngpus = int(dev.size)
        return self

    def multi_partition(self):
        """(axis, row_offsets) of a multi-GPU handle after its first solve; axis -1 = contiguous row blocks."""
        axis = C.c_int(0)
        offs = (C.c_longlong * (self.ngpus + 1))()
        self._check(self.lib.opmgpu_multi_partition(self.h, C.byref(axis), offs))
        return axis.value, list(offs)

    def close(self):
        if self.h:
            self.lib.opmgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------
    def error(self) -> str:
        return self.lib.opmgpu_last_error(self.h).decode()

    def _check(self, rc, allow=()):
        if rc == L.OK or rc in allow:
            return rc
        msg = self.error()
        if rc == L.NOT_CONVERGED:
            raise LinearSolverProblem("Convergence failure for linear solver.")
        if rc in (L.SINGULAR_BLOCK, L.BREAKDOWN):
            raise NumericalIssue(msg)
        if rc == L.BAD_PATTERN:
            raise ValueError(msg)
        if rc == L.BAD_ARGUMENT:
            raise ValueError(msg or "bad argument")
        raise RuntimeError(msg)

    def set_precision(self, single_precision: bool):
        """Select the instance: False = the reference's Impl<3,double>, True = its Impl<3,float>
        (LinearisedBlackoilResidual::singlePrecision, ...Interleaved.cpp:467-487).  The arrays that
        cross this wrapper stay float64 either way."""
        self._check(self.lib.opmgpu_set_precision(self.h, int(bool(single_precision))))

    def single_precision(self) -> bool:
        return bool(self.lib.opmgpu_get_precision(self.h))

    def use_torch_stream(self):
        import torch
        self._check(self.lib.opmgpu_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    # -- pattern / values ------------------------------------------------------------------
    def set_pattern(self, rowptr, colidx):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.N, self.nnzb = rowptr.size - 1, colidx.size
        self._check(self.lib.opmgpu_set_pattern_bcrs(self.h, self.N, self.nnzb, _ip(rowptr), _ip(colidx)))

    def set_pattern_operator_only(self, rowptr, colidx):
        """Pattern for spmv / spmv_dev only (no ILU0 analysis): SpMV micro-benchmarks at sizes where
        the factor records would not fit beside the matrix."""
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.N, self.nnzb = rowptr.size - 1, colidx.size
        self._check(self.lib.opmgpu_set_pattern_bcrs_operator_only(self.h, self.N, self.nnzb, _ip(rowptr), _ip(colidx)))

    # -- multicolour ILU0: a flagged variant, not the reference's preconditioner ----------------
    def set_ilu_ordering(self, multicolour):
        """Prepare the NEXT pattern for the natural (reference) ordering of the ILU0 (False / 0), the
        multicolour ordering (True / 1) or the k-line ordering ("lines" / 2; Cartesian stencils only).
        include/opm_gpu_solver.h: iteration counts of the variants are never parity."""
        o = L.ILU_MULTICOLOUR_LINES if multicolour == "lines" else int(multicolour)
        self._check(self.lib.opmgpu_set_ilu_ordering(self.h, o))

    def ilu_ordering(self) -> int:
        return int(self.lib.opmgpu_get_ilu_ordering(self.h))

    def ilu_permutation(self):
        """(ncolours, n2p) of the current multicolour pattern; n2p[i] = position of row i in P A P^T."""
        nc = C.c_int(0)
        n2p = np.empty(self.N, dtype=np.int32)
        self._check(self.lib.opmgpu_get_ilu_permutation(self.h, C.byref(nc), _ip(n2p)))
        return nc.value, n2p

    # -- block sizes other than 3 (np = 2: two-phase decks) ------------------------------------
    def set_block_size(self, np_: int):
        """Prepare the NEXT pattern for block size np_ (2..6); call before set_pattern."""
        self._check(self.lib.opmgpu_set_block_size(self.h, int(np_)))

    def spmv_np(self, np_, vals, x):
        vals = np.ascontiguousarray(vals, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.N * np_)
        self._check(self.lib.opmgpu_spmv_np(self.h, int(np_), _dp(vals), _dp(x), _dp(y)))
        return y.reshape(self.N, np_)

    def ilu0_np(self, np_, vals, w=0.9, d=None):
        """-> (lu[nnzb, np*np], v[N, np] or None, bad_row)"""
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        lu = np.empty((self.nnzb, np_ * np_))
        bad = C.c_int(-1)
        v = None
        if d is not None:
            d = np.ascontiguousarray(d, dtype=np.float64)
            v = np.empty(self.N * np_)
        rc = self.lib.opmgpu_ilu0_np(self.h, int(np_), _dp(vals), _dp(lu), float(w), _dp(d) if d is not None else None,
                                     _dp(v) if v is not None else None, C.byref(bad))
        if rc == L.SINGULAR_BLOCK:
            return lu, None, bad.value
        self._check(rc)
        return lu, (v.reshape(self.N, np_) if v is not None else None), -1

    def solve_bcrs_np(self, np_, vals, rhs, params: Optional[L.Params] = None, raise_on_failure=True, **kw):
        p = params if params is not None else make_params(**kw)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros(self.N * np_)
        res = L.Result()
        rc = self.lib.opmgpu_solve_bcrs_np(self.h, int(np_), _dp(vals), _dp(rhs), _dp(x), C.byref(p), C.byref(res))
        self.last = res.as_dict()
        self.last["status"] = rc
        self._check(rc, allow=() if raise_on_failure else (L.NOT_CONVERGED, L.SINGULAR_BLOCK, L.BREAKDOWN))
        return x.reshape(self.N, np_), self.last

    def solve_from_csc_blocks_np(self, N, np_, blocks, matbalscale, rhs_eqmajor, params: Optional[L.Params] = None,
                                 raise_on_failure=True, **kw):
        """...Interleaved.cpp:234-283 for np_ x np_ blocks; blocks[p1*np_+p2] = (colptr, rowidx, val)."""
        p = params if params is not None else make_params(**kw)
        arr = (L.Csc * (np_ * np_))()
        keep = []
        for q, (cp, ri, v) in enumerate(blocks):
            cp = np.ascontiguousarray(cp, dtype=np.int32)
            ri = np.ascontiguousarray(ri, dtype=np.int32)
            v = np.ascontiguousarray(v, dtype=np.float64)
            keep += [cp, ri, v]
            arr[q].colptr, arr[q].rowidx, arr[q].val = _ip(cp), _ip(ri), _dp(v)
        sc = np.ascontiguousarray(matbalscale, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs_eqmajor, dtype=np.float64)
        dx = np.zeros(np_ * N)
        res = L.Result()
        rc = self.lib.opmgpu_solve_from_csc_blocks_np(self.h, int(N), int(np_), arr, _dp(sc), _dp(rhs), _dp(dx),
                                                      C.byref(p), C.byref(res))
        self.last = res.as_dict()
        self.last["status"] = rc
        self.N = N
        self._check(rc, allow=() if raise_on_failure else (L.NOT_CONVERGED, L.SINGULAR_BLOCK, L.BREAKDOWN))
        return dx, self.last

    def set_values(self, vals):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        assert vals.size == self.nnzb * 9
        self._check(self.lib.opmgpu_set_values_bcrs3(self.h, _dp(vals)))

    def set_values_dev(self, vals_t):
        assert vals_t.is_cuda and vals_t.is_contiguous() and vals_t.numel() == self.nnzb * 9
        self._keep_vals = vals_t
        self._check(self.lib.opmgpu_set_values_bcrs3_dev(self.h, C.c_void_p(vals_t.data_ptr())))

    def num_levels(self):
        a, b = C.c_int(), C.c_int()
        self._check(self.lib.opmgpu_num_levels(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_profiling(self, on: bool = True):
        self._check(self.lib.opmgpu_set_profiling(self.h, int(on)))

    def profile(self):
        """-> {class: (ms, launches)} for ilu_apply, spmv, vector, factor since set_profiling."""
        ms = (C.c_double * 4)()
        cnt = (C.c_longlong * 4)()
        self._check(self.lib.opmgpu_get_profile(self.h, ms, cnt))
        names = ("ilu_apply", "spmv", "vector", "factor")
        return {n: (ms[i], int(cnt[i])) for i, n in enumerate(names)}

    def launch_count(self) -> int:
        return int(self.lib.opmgpu_launch_count(self.h))

    # -- kernels ---------------------------------------------------------------------------
    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.N * 3)
        self._check(self.lib.opmgpu_spmv(self.h, _dp(x), _dp(y)))
        return y.reshape(self.N, 3)

    def spmv_dev(self, x_t, y_t):
        self._check(self.lib.opmgpu_spmv_dev(self.h, C.c_void_p(x_t.data_ptr()), C.c_void_p(y_t.data_ptr())))

    def ilu0_factor(self):
        bad = C.c_int(-1)
        rc = self.lib.opmgpu_ilu0_factor(self.h, C.byref(bad))
        if rc == L.SINGULAR_BLOCK:
            return bad.value
        self._check(rc)
        return -1

    def ilu0_factors(self):
        lu = np.empty((self.nnzb, 9))
        self._check(self.lib.opmgpu_ilu0_get_factors(self.h, _dp(lu)))
        return lu

    def ilu0_apply(self, w, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        v = np.empty(self.N * 3)
        self._check(self.lib.opmgpu_ilu0_apply(self.h, float(w), _dp(d), _dp(v)))
        return v.reshape(self.N, 3)

    def ilu0_apply_dev(self, w, d_t, v_t):
        self._check(self.lib.opmgpu_ilu0_apply_dev(self.h, float(w), C.c_void_p(d_t.data_ptr()),
                                                   C.c_void_p(v_t.data_ptr())))

    def dot(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        y = np.ascontiguousarray(y, dtype=np.float64).ravel()
        out = C.c_double()
        self._check(self.lib.opmgpu_dot(self.h, _dp(x), _dp(y), x.size, C.byref(out)))
        return out.value

    def residual_history(self):
        n = C.c_int()
        buf = np.zeros(4096)
        self._check(self.lib.opmgpu_residual_history(self.h, _dp(buf), buf.size, C.byref(n)))
        return buf[:min(n.value, buf.size)].copy()

    # -- solves ----------------------------------------------------------------------------
    def solve_bcrs(self, vals, rhs, params: Optional[L.Params] = None, raise_on_failure=True, **kw):
        """ISTLSolver::solve(A, x, b) with host arrays. Returns x[N,3], result dict."""
        p = params if params is not None else make_params(**kw)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros(self.N * 3)
        res = L.Result()
        rc = self.lib.opmgpu_solve_bcrs3(self.h, _dp(vals), _dp(rhs), _dp(x), C.byref(p), C.byref(res))
        self.last = res.as_dict()
        self.last["status"] = rc
        self._check(rc, allow=() if raise_on_failure else (L.NOT_CONVERGED, L.SINGULAR_BLOCK, L.BREAKDOWN))
        return x.reshape(self.N, 3), self.last

    def solve_bcrs_dev(self, vals_t, rhs_t, x_t, params: Optional[L.Params] = None,
                       raise_on_failure=True, **kw):
        """Same with CUDA tensors (device-resident inputs).  vals_t=None: the values set by
        set_values / set_values_dev."""
        p = params if params is not None else make_params(**kw)
        res = L.Result()
        rc = self.lib.opmgpu_solve_bcrs3_dev(self.h, C.c_void_p(vals_t.data_ptr()) if vals_t is not None else None,
                                             C.c_void_p(rhs_t.data_ptr()),
                                             C.c_void_p(x_t.data_ptr()), C.byref(p), C.byref(res))
        self.last = res.as_dict()
        self.last["status"] = rc
        self._check(rc, allow=() if raise_on_failure else (L.NOT_CONVERGED, L.SINGULAR_BLOCK, L.BREAKDOWN))
        return self.last

    def solve_from_csc_blocks(self, N, blocks9, matbalscale, rhs_eqmajor, params: Optional[L.Params] = None,
                              raise_on_failure=True, out=None, **kw):
        """...Interleaved.cpp:234-283 in one call; blocks9[p1*3+p2] = (colptr, rowidx, val).
        out: optional float64 array of 3N entries that receives the increment (page-locked memory
        makes the device-to-host copy asynchronous and ~3x faster than into a fresh pageable array)."""
        p = params if params is not None else make_params(**kw)
        arr = (L.Csc * 9)()
        keep = []
        for q, (cp, ri, v) in enumerate(blocks9):
            cp = np.ascontiguousarray(cp, dtype=np.int32)
            ri = np.ascontiguousarray(ri, dtype=np.int32)
            v = np.ascontiguousarray(v, dtype=np.float64)
            keep += [cp, ri, v]
            arr[q].colptr, arr[q].rowidx, arr[q].val = _ip(cp), _ip(ri), _dp(v)
        sc = np.ascontiguousarray(matbalscale, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs_eqmajor, dtype=np.float64)
        if out is not None:
            dx = out
            assert dx.dtype == np.float64 and dx.size == 3 * N and dx.flags["C_CONTIGUOUS"]
        else:
            dx = np.zeros(3 * N)
        res = L.Result()
        rc = self.lib.opmgpu_solve_from_csc_blocks(self.h, int(N), arr, _dp(sc), _dp(rhs), _dp(dx),
                                                   C.byref(p), C.byref(res))
        self.last = res.as_dict()
        self.last["status"] = rc
        self.N = N
        self._check(rc, allow=() if raise_on_failure else (L.NOT_CONVERGED, L.SINGULAR_BLOCK, L.BREAKDOWN))
        return dx, self.last


# ---------------------------------------------------------------------------------------------
# Stand-ins for the reference's input types (no Eigen here): same members, same meaning.
# ---------------------------------------------------------------------------------------------
@dataclass
class ADB:
    """AutoDiffBlock<double>: value() and derivative()[block] (opm/autodiff/AutoDiffBlock.hpp:99,
    458-461).  Jacobian blocks are scipy.sparse CSC matrices (Eigen's default is column-major)."""
    value: np.ndarray
    jac: List[object] = field(default_factory=list)

    def size(self):
        return self.value.size


@dataclass
class LinearisedBlackoilResidual:
    """opm/autodiff/LinearisedBlackoilResidual.hpp:47-72."""
    material_balance_eq: List[ADB]
    well_flux_eq: Optional[ADB] = None
    well_eq: Optional[ADB] = None
    matbalscale: Sequence[float] = (1.1169, 1.0031, 0.0031)
    singlePrecision: bool = False


def eliminateVariable(eqs: List[ADB], n: int) -> List[ADB]:
    """Schur complement A - B D^-1 C removing variable/equation n, host side as in
    opm/autodiff/NewtonIterationUtilities.cpp:45-128 (sparse LU of D solved against I)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    num_eq = len(eqs)
    if num_eq != len(eqs[0].jac):
        raise ValueError("eliminateVariable() requires the same number of variables and equations.")
    if n >= num_eq:
        raise ValueError("Trying to eliminate variable from too small set of equations.")
    Jn = eqs[n].jac
    D = sp.csc_matrix(Jn[n])
    lu = spl.splu(D)
    Di = sp.csc_matrix(lu.solve(np.eye(D.shape[0])))
    Dibn = lu.solve(eqs[n].value)
    out = []
    for eq in range(num_eq):
        if eq == n:
            continue
        B = sp.csc_matrix(eqs[eq].jac[n])
        val = eqs[eq].value - B @ Dibn
        jacs = []
        for var in range(num_eq):
            if var == n:
                continue
            u = Di @ sp.csc_matrix(Jn[var])
            J = sp.csc_matrix(eqs[eq].jac[var]) + (B @ u) * -1.0
            J.sort_indices()
            jacs.append(sp.csc_matrix(J))
        out.append(ADB(val, jacs))
    return out


def recoverVariable(equation: ADB, partial_solution: np.ndarray, n: int) -> np.ndarray:
    """y = D^-1 (b - C x) spliced back at the eliminated offset,
    opm/autodiff/NewtonIterationUtilities.cpp:134-184."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    D = sp.csc_matrix(equation.jac[n])
    Cj = [sp.csc_matrix(j) for k, j in enumerate(equation.jac) if k != n]
    Cm = sp.hstack(Cj, format="csc")
    b = equation.value - Cm @ partial_solution
    y = spl.splu(D).solve(b)
    start = sum(sp.csc_matrix(equation.jac[i]).shape[1] for i in range(n))
    return np.concatenate([partial_solution[:start], y, partial_solution[start:]])


class NewtonIterationBlackoilGPU:
    """Drop-in for NewtonIterationBlackoilInterleaved behind NewtonIterationBlackoilInterface
    (solver_approach=gpu, opm/autodiff/FlowMain.hpp:806-830)."""

    def __init__(self, param: Optional[dict] = None, parallelInformation=None, device: int = 0):
        self.parameters_ = make_params(param)
        self.parallelInformation_ = parallelInformation      # empty boost::any: serial branches
        self.iterations_ = 0
        self._solver = GpuLinearSolver(device)
        # ilu_redblack (ISTLSolver.hpp:207-209): the ILU0 of a colour-sorted reordering -- here the library's
        # multicolour variant (greedy natural-order colouring, not the reference's Welsh-Powell / sphere
        # reordering of opm-simulators' GraphColoring.hpp: iteration counts match neither reference ordering)
        rb = (param or {}).get("ilu_redblack", False)
        if isinstance(rb, str):
            rb = rb.lower() in ("1", "true", "yes")
        if rb:
            self._solver.set_ilu_ordering(True)

    def iterations(self) -> int:
        return self.iterations_

    def parallelInformation(self):
        return self.parallelInformation_

    def computeNewtonIncrement(self, residual: LinearisedBlackoilResidual) -> np.ndarray:
        """...Interleaved.cpp:202-292 (np = 2..6; the double or the float instance as
        residual.singlePrecision asks, :467-487).  Returns dx ordered
        [p(N), sw(N), xvar(N), qs(nw*np), bhp(nw)]."""
        import scipy.sparse as sp
        npz = len(residual.material_balance_eq)
        if not 2 <= npz <= 6:
            raise NotImplementedError("NewtonIterationBlackoilGPU: np outside the reference's range 2..6")
        eqs = list(residual.material_balance_eq)
        has_wells = residual.well_flux_eq is not None and residual.well_flux_eq.size() > 0
        elim = []
        if has_wells:
            eqs += [residual.well_flux_eq, residual.well_eq]
            elim.append(eqs[npz])
            eqs = eliminateVariable(eqs, npz)            # well flux unknowns
            elim.append(eqs[npz])
            eqs = eliminateVariable(eqs, npz)            # bhp unknowns
        N = eqs[0].size()
        blocks = []
        for p1 in range(npz):
            for p2 in range(npz):
                J = sp.csc_matrix(eqs[p1].jac[p2])
                J.sort_indices()
                blocks.append((J.indptr, J.indices, J.data))
        rhs = np.concatenate([eqs[p].value for p in range(npz)])
        # the dispatcher of the reference: Impl<3,float> when the residual asks for it (GMRES exists
        # for the double instance only; that combination stays in double)
        self._solver.set_precision(bool(residual.singlePrecision) and not self.parameters_.newton_use_gmres)
        try:
            if npz == 3:
                dx, res = self._solver.solve_from_csc_blocks(N, blocks, residual.matbalscale, rhs,
                                                             params=self.parameters_)
            else:
                dx, res = self._solver.solve_from_csc_blocks_np(N, npz, blocks, residual.matbalscale[:npz], rhs,
                                                                params=self.parameters_)
        finally:
            if self._solver.last is not None:            # valid also on the exception path
                self.iterations_ = self._solver.last["iterations"]
        if has_wells:
            dx = recoverVariable(elim[1], dx, npz)
            dx = recoverVariable(elim[0], dx, npz)
        return dx
