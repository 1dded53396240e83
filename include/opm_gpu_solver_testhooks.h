/*
 * opm_gpu_solver_testhooks.h -- test hooks of libopmgpu.so.  NOT part of the drop-in boundary
 * (include/opm_gpu_solver.h is): a caller of the solver never needs these.  They let the test suite
 * check host-side analysis code without a GPU and inject a failure the recovery path must survive.
 *
 * Shipping library (libopmgpu.so): the six hooks of the first section.
 * Experiments build (libopmgpu_exp.so, `make -C opm_simulators_legacy_b200/csrc exp`, -DOPMGPU_EXPERIMENTS):
 * additionally the hooks of the second section, the slower kernel variants kept for A/B measurements and
 * the tuning switches of DESIGN.md section 10.
 */
#ifndef OPM_GPU_SOLVER_TESTHOOKS_H
#define OPM_GPU_SOLVER_TESTHOOKS_H

#include "opm_gpu_solver.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- shipping library ----------------------------------------------------------------------- */
/* Host only: wall time (ms) of the pattern analysis for P sweep CTAs (tools/analysis_time.py). */
double opmgpu_debug_analyse_only(int N, const int* rowptr, const int* colidx, int P);
/* Host only: the pipelined sweep programs (record streams, push slots, windows) interpreted
 * sequentially exactly as the kernels index them: v = w U^-1 L^-1 d from BCRS factors lu. */
int opmgpu_debug_host_program_apply(int N, const int* rowptr, const int* colidx, const double* lu, int P,
                                    double w, const double* d, double* v, int* info /*[8]*/);
/* Host only: the pipelined factorisation program interpreted sequentially; lu receives the factors. */
int opmgpu_debug_host_factor_program(int N, const int* rowptr, const int* colidx, const double* vals, int P,
                                     double* lu, int* bad_row, int* info /*[4]*/);
/* Host only: the halo plan of one rank of a row-partitioned pattern (multi-GPU set-up). */
int opmgpu_debug_partition(int N_local, const int* rowptr, const long long* colidx_global,
                           const long long* row_offsets, int world, int rank, int* colidx_full,
                           long long* ghost_global, int* recv_cnt, int* rowptr_diag, int* colidx_diag,
                           int* lu_src, int* nnzb_diag_out);
/* Host only: the multicolour ILU0 program (permutation, [ L | Dinv | U ] layout, update lists) interpreted
 * sequentially as the kernels index it: factors in the caller's slots (lu_out, may be NULL) and
 * v = w P^T U^-1 L^-1 P d; lines != 0: the k-line ordering (returns -2 when it refuses the pattern). */
int opmgpu_debug_host_mc_apply(int N, const int* rowptr, const int* colidx, const double* vals, int lines,
                               double w, const double* d, double* v, double* lu_out, int* info /*[2]*/);
/* Sets the device watchdog word as a sweep kernel does when a dependency is never delivered: the
 * next call that collects it must fail with OPMGPU_CUDA_ERROR, re-arm and leave the handle usable. */
int opmgpu_debug_set_watchdog_word(opmgpu_handle h, int code);

/* ---- experiments build only ------------------------------------------------------------------ */
#ifdef OPMGPU_EXPERIMENTS
/* Host only: the column-owned sweep program (colprog.cpp) interpreted sequentially. */
int opmgpu_debug_host_col_apply(int N, const int* rowptr, const int* colidx, int nx, int ny, int nz,
                                const double* lu, int P, double w, const double* d, double* v, int* info /*[8]*/);
/* clock64 / %globaltimer stamps of the pipelined sweeps of the next apply (tools/trace_sweep.py,
 * tools/gtrace_sweep.py, tools/gtrace_cluster.py). */
int opmgpu_debug_trace_apply(opmgpu_handle h, int cta, double w, const double* d_dev, double* v_dev, long long* out);
int opmgpu_debug_gtrace_apply(opmgpu_handle h, int steps, double w, const double* d_dev, double* v_dev,
                              long long* out, int* P_out);
#endif

#ifdef __cplusplus
}
#endif
#endif
