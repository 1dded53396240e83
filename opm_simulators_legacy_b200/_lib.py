"""ctypes binding of libopmgpu.so (the C ABI declared in include/opm_gpu_solver.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C csrc`.  There is no
fallback of any kind: if the shared object is missing or no sm_100 device is usable, the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OPMGPU_LIB: alternative build of the same library (kernel experiments), else the in-tree one
LIB_PATH = os.environ.get("OPMGPU_LIB") or os.path.join(_HERE, "libopmgpu.so")
# the experiments build (make exp; -DOPMGPU_EXPERIMENTS): slower kernel variants kept for A/B
# measurements, the tuning switches of DESIGN.md section 10, tracing entry points.  Never the product.
EXP_LIB_PATH = os.path.join(_HERE, "libopmgpu_exp.so")
# test hooks (include/opm_gpu_solver_testhooks.h): in the shipping library / in the experiments build only
TEST_HOOKS = ["opmgpu_debug_analyse_only", "opmgpu_debug_host_program_apply", "opmgpu_debug_host_factor_program",
              "opmgpu_debug_partition", "opmgpu_debug_set_watchdog_word", "opmgpu_debug_host_mc_apply"]
EXP_HOOKS = ["opmgpu_debug_host_col_apply", "opmgpu_debug_trace_apply", "opmgpu_debug_gtrace_apply"]

# every symbol include/opm_gpu_solver.h declares (tests check the export list against this)
EXPORTS = [
    "opmgpu_default_params", "opmgpu_create", "opmgpu_destroy", "opmgpu_last_error",
    "opmgpu_nccl_unique_id", "opmgpu_create_distributed", "opmgpu_set_stream",
    "opmgpu_create_multi", "opmgpu_multi_partition",
    "opmgpu_set_pattern_bcrs", "opmgpu_set_pattern_bcrs_distributed",
    "opmgpu_solve_bcrs3", "opmgpu_solve_bcrs3_dev", "opmgpu_solve_from_csc_blocks",
    "opmgpu_set_values_bcrs3", "opmgpu_set_values_bcrs3_dev", "opmgpu_spmv", "opmgpu_spmv_dev",
    "opmgpu_ilu0_factor", "opmgpu_ilu0_get_factors", "opmgpu_ilu0_apply", "opmgpu_ilu0_apply_dev",
    "opmgpu_dot", "opmgpu_num_levels", "opmgpu_launch_count", "opmgpu_residual_history",
    "opmgpu_set_profiling", "opmgpu_get_profile", "opmgpu_set_precision", "opmgpu_get_precision",
    "opmgpu_set_pattern_bcrs_operator_only",
    "opmgpu_set_ilu_ordering", "opmgpu_get_ilu_ordering", "opmgpu_get_ilu_permutation", "opmgpu_multicolour_order", "opmgpu_line_order",
    "opmgpu_set_block_size", "opmgpu_solve_bcrs_np", "opmgpu_solve_from_csc_blocks_np", "opmgpu_spmv_np", "opmgpu_ilu0_np",
]

OK, NOT_CONVERGED, SINGULAR_BLOCK, BREAKDOWN, BAD_PATTERN, BAD_ARGUMENT = 0, 1, 2, 3, 4, 5
CUDA_ERROR, NCCL_ERROR = -1, -2
ILU_NATURAL, ILU_MULTICOLOUR, ILU_MULTICOLOUR_LINES = 0, 1, 2


class Params(C.Structure):
    _fields_ = [("linear_solver_reduction", C.c_double),
                ("linear_solver_maxiter", C.c_int),
                ("ilu_relaxation", C.c_double),
                ("linear_solver_verbosity", C.c_int),
                ("linear_solver_ignoreconvergencefailure", C.c_int),
                ("require_full_sparsity_pattern", C.c_int),
                ("max_half_steps", C.c_int),
                ("newton_use_gmres", C.c_int),
                ("linear_solver_restart", C.c_int)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("half_steps", C.c_int),
                ("bad_row", C.c_int), ("reduction", C.c_double), ("norm0", C.c_double),
                ("ms_analysis", C.c_double), ("ms_h2d", C.c_double), ("ms_interleave", C.c_double),
                ("ms_factor", C.c_double), ("ms_solve", C.c_double), ("ms_d2h", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Csc(C.Structure):
    _fields_ = [("colptr", C.POINTER(C.c_int)), ("rowidx", C.POINTER(C.c_int)),
                ("val", C.POINTER(C.c_double))]


_lib = None
_exp = None


def load_experiments():
    """Load libopmgpu_exp.so (same C ABI plus the experiment hooks); for A/B tools and the tests of
    the experimental kernel variants only."""
    global _exp
    if _exp is None:
        if not os.path.exists(EXP_LIB_PATH):
            raise RuntimeError(f"{EXP_LIB_PATH} is missing: make -C opm_simulators_legacy_b200/csrc exp")
        _exp = _bind(C.CDLL(EXP_LIB_PATH))
    return _exp


def load():
    """Load libopmgpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or make -C opm_simulators_legacy_b200/csrc). There is no CPU fallback.")
    _lib = _bind(C.CDLL(LIB_PATH))
    return _lib


def _bind(lib):
    H = C.c_void_p
    ip, dp, vp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p
    PP, RP = C.POINTER(Params), C.POINTER(Result)
    sig = {
        "opmgpu_default_params": (None, [PP]),
        "opmgpu_create": (C.c_int, [C.c_int, C.POINTER(H)]),
        "opmgpu_destroy": (C.c_int, [H]),
        "opmgpu_last_error": (C.c_char_p, [H]),
        "opmgpu_nccl_unique_id": (C.c_int, [vp]),
        "opmgpu_create_distributed": (C.c_int, [C.c_int, C.c_int, C.c_int, vp, C.POINTER(H)]),
        "opmgpu_create_multi": (C.c_int, [C.c_int, ip, C.POINTER(H)]),
        "opmgpu_multi_partition": (C.c_int, [H, ip, C.POINTER(C.c_longlong)]),
        "opmgpu_set_stream": (C.c_int, [H, vp]),
        "opmgpu_set_precision": (C.c_int, [H, C.c_int]),
        "opmgpu_get_precision": (C.c_int, [H]),
        "opmgpu_set_pattern_bcrs": (C.c_int, [H, C.c_int, C.c_int, ip, ip]),
        "opmgpu_set_pattern_bcrs_operator_only": (C.c_int, [H, C.c_int, C.c_int, ip, ip]),
        "opmgpu_set_block_size": (C.c_int, [H, C.c_int]),
        "opmgpu_set_ilu_ordering": (C.c_int, [H, C.c_int]),
        "opmgpu_get_ilu_ordering": (C.c_int, [H]),
        "opmgpu_get_ilu_permutation": (C.c_int, [H, ip, ip]),
        "opmgpu_multicolour_order": (C.c_int, [C.c_int, ip, ip, ip, ip, ip]),
        "opmgpu_line_order": (C.c_int, [C.c_int, C.c_int, C.c_int, ip]),
        "opmgpu_solve_bcrs_np": (C.c_int, [H, C.c_int, dp, dp, dp, PP, RP]),
        "opmgpu_solve_from_csc_blocks_np": (C.c_int, [H, C.c_int, C.c_int, C.POINTER(Csc), dp, dp, dp, PP, RP]),
        "opmgpu_spmv_np": (C.c_int, [H, C.c_int, dp, dp, dp]),
        "opmgpu_ilu0_np": (C.c_int, [H, C.c_int, dp, dp, C.c_double, dp, dp, ip]),
        "opmgpu_set_pattern_bcrs_distributed": (C.c_int, [H, C.c_int, C.c_int, ip, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
        "opmgpu_solve_bcrs3": (C.c_int, [H, dp, dp, dp, PP, RP]),
        "opmgpu_solve_bcrs3_dev": (C.c_int, [H, vp, vp, vp, PP, RP]),
        "opmgpu_solve_from_csc_blocks": (C.c_int, [H, C.c_int, C.POINTER(Csc), dp, dp, dp, PP, RP]),
        "opmgpu_set_values_bcrs3": (C.c_int, [H, dp]),
        "opmgpu_set_values_bcrs3_dev": (C.c_int, [H, vp]),
        "opmgpu_spmv": (C.c_int, [H, dp, dp]),
        "opmgpu_spmv_dev": (C.c_int, [H, vp, vp]),
        "opmgpu_ilu0_factor": (C.c_int, [H, ip]),
        "opmgpu_ilu0_get_factors": (C.c_int, [H, dp]),
        "opmgpu_ilu0_apply": (C.c_int, [H, C.c_double, dp, dp]),
        "opmgpu_ilu0_apply_dev": (C.c_int, [H, C.c_double, vp, vp]),
        "opmgpu_dot": (C.c_int, [H, dp, dp, C.c_int, dp]),
        "opmgpu_num_levels": (C.c_int, [H, ip, ip]),
        "opmgpu_launch_count": (C.c_longlong, [H]),
        "opmgpu_residual_history": (C.c_int, [H, dp, C.c_int, ip]),
        "opmgpu_set_profiling": (C.c_int, [H, C.c_int]),
        "opmgpu_get_profile": (C.c_int, [H, dp, C.POINTER(C.c_longlong)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    return lib
