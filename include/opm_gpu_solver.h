/*
 * opm_gpu_solver.h -- C ABI of the B200-native Newton-step linear solver.
 *
 * Drop-in boundary: the body of
 *     NewtonIterationBlackoilInterleavedImpl<3,double>::computeNewtonIncrement
 *     (opm/autodiff/NewtonIterationBlackoilInterleaved.cpp:202-292)
 * minus the host-side well elimination/recovery, i.e. formInterleavedSystem (:110-194),
 * the rhs/solution interleave (:263-283) and ISTLSolver::solve (opm/autodiff/ISTLSolver.hpp:
 * 283-306 -> :124-189 ILU0 construction -> :250-274 BiCGSTAB -> :358-368 checkConvergence).
 * Precedent in the reference for a raw-array solver signature:
 * LinearSolverInterface::solve(size,nnz,ia,ja,sa,rhs,solution) at
 * opm/core/linalg/LinearSolverInterface.hpp:67-74 and struct CSRMatrix at
 * opm/core/linalg/sparse_sys.h:38-47.
 *
 * Conventions: every function returns an opmgpu_status value, 0 = ok; no exception crosses the
 * ABI; all pointers are HOST pointers unless the name ends in _dev; block size is 3
 * (np = 3, water/oil/gas); FP64 values, int32 indices.  A handle owns one GPU (one CUDA
 * stream); it is not re-entrant, matching the reference's single-threaded caller
 * (BlackoilModelBase_impl.hpp:289).  There is no CPU fallback: without a usable sm_100
 * device opmgpu_create fails.
 */
#ifndef OPM_GPU_SOLVER_H
#define OPM_GPU_SOLVER_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct opmgpu_solver* opmgpu_handle;

typedef enum {
    OPMGPU_OK             = 0,
    OPMGPU_NOT_CONVERGED  = 1,   /* -> Opm::LinearSolverProblem (ISTLSolver.hpp:364-367)            */
    OPMGPU_SINGULAR_BLOCK = 2,   /* ILU0 pivot block singular/missing -> Dune::MatrixBlockError class */
    OPMGPU_BREAKDOWN      = 3,   /* |rho|,|omega| <= 1e-80 or |h| < 1e-80 -> Dune::SolverAbort class   */
    OPMGPU_BAD_PATTERN    = 4,   /* Jacobian entry outside the pattern (dune throws in istlA[r][c])   */
    OPMGPU_BAD_ARGUMENT   = 5,
    OPMGPU_CUDA_ERROR     = -1,  /* -> std::runtime_error; text in opmgpu_last_error                  */
    OPMGPU_NCCL_ERROR     = -2
} opmgpu_status;

/* The keys FlowLinearSolverParameters reads from the ParameterGroup (member names visible at
 * ISTLSolver.hpp:142,204-208,255-262,364 and ...Interleaved.cpp:127); there are no hidden
 * defaults below the ABI -- opmgpu_default_params fills in upstream's 2019.04 values. */
typedef struct {
    double linear_solver_reduction;                /* 1e-2 */
    int    linear_solver_maxiter;                  /* 150  */
    double ilu_relaxation;                         /* 0.9  */
    int    linear_solver_verbosity;                /* 0    */
    int    linear_solver_ignoreconvergencefailure; /* 0    */
    int    require_full_sparsity_pattern;          /* 0    */
    int    max_half_steps;                         /* -1; >=0 stops after that many half iterations
                                                      (parity checks at equal half-step counts)      */
    int    newton_use_gmres;                       /* 0; 1: Dune::RestartedGMResSolver instead of
                                                      BiCGSTABSolver (ISTLSolver.hpp:257-265)        */
    int    linear_solver_restart;                  /* 40   */
} opmgpu_params;

/* Dune::InverseOperatorResult plus what the benchmark reports. */
typedef struct {
    int    iterations;        /* ceil(it), valid also on failure (BlackoilModelBase_impl.hpp:291,295) */
    int    converged;
    int    half_steps;
    int    bad_row;           /* OPMGPU_SINGULAR_BLOCK: block row of the failing pivot */
    double reduction;         /* |r| / |r0| at exit */
    double norm0;
    double ms_analysis;       /* pattern analysis (only non-zero when the pattern changed) */
    double ms_h2d;            /* host->device copies                                       */
    double ms_interleave;     /* K1 scatter of the CSC blocks into BCRS                    */
    double ms_factor;         /* K3 ILU0 factorisation                                     */
    double ms_solve;          /* BiCGStab loop                                             */
    double ms_d2h;
} opmgpu_result;

/* One scalar Jacobian block d(eq p1)/d(var p2) in Eigen's column-major compressed layout
 * (AutoDiffMatrix::getSparse, opm/autodiff/AutoDiffMatrix.hpp:635-648). */
typedef struct {
    const int*    colptr;     /* N+1  (outerIndexPtr) */
    const int*    rowidx;     /* nnz  (innerIndexPtr), ascending per column */
    const double* val;        /* nnz  (valuePtr) */
} opmgpu_csc;

void opmgpu_default_params(opmgpu_params* p);

/* Lifetime: created once where FlowMain::setupLinearSolver builds fis_solver_
 * (opm/autodiff/FlowMain.hpp:806-830) and kept for the whole run.
 *
 * GPU exclusivity: the ILU0 sweep and factorisation kernels are persistent CTAs that wait for each
 * other (push slots, distributed shared memory, cluster barriers), so every CTA of a launch must
 * be resident at once.  They are launched cooperatively (the driver refuses a launch that cannot
 * be co-resident) and the handle's stream runs one kernel at a time; do not share the device's
 * SMs with other work (other streams, MPS clients) while a solve is in flight.  All waits are
 * bounded: an undelivered dependency ends the solve with OPMGPU_CUDA_ERROR ("sweep watchdog"),
 * after which the handle is usable again.
 *
 * Borrowed device pointers: opmgpu_set_values_bcrs3_dev keeps the caller's pointer until the next
 * set_values / solve call -- the buffer must stay valid and unchanged that long (opmgpu_spmv*,
 * opmgpu_ilu0_factor and opmgpu_ilu0_get_factors read it).  opmgpu_solve_bcrs3_dev borrows its
 * pointers for the duration of the call only. */
int  opmgpu_create(int device, opmgpu_handle* out);
int  opmgpu_destroy(opmgpu_handle h);

/* Scalar type of the instance the handle runs.  The reference keeps two instances of its solver,
 * Impl<np,double> and Impl<np,float>, and picks one per call from
 * LinearisedBlackoilResidual::singlePrecision (NewtonIterationBlackoilInterleaved.cpp:467-487;
 * BlackoilModelBase_impl.hpp:284 sets it for time steps below 20 days).  single_precision != 0
 * selects the float instance: the interleaved matrix, the right-hand side and every vector are
 * rounded to float once (as the assignments at ...Interleaved.cpp:189, :266 do), SpMV, ILU0,
 * sweeps, vector updates, scalar products and the convergence test run in float.  The ABI still
 * exchanges doubles everywhere; device pointers passed to the *_dev entry points are copied and
 * rounded, never borrowed, in this mode.  Switching invalidates values and factors, not the
 * pattern analysis.  Restarted GMRES exists for the double instance only. */
int  opmgpu_set_precision(opmgpu_handle h, int single_precision);
int  opmgpu_get_precision(opmgpu_handle h);                 /* 1: float instance */
const char* opmgpu_last_error(opmgpu_handle h);          /* h may be NULL: creation errors */

/* Multi-GPU (one process per GPU): rank r of `world` owns the contiguous block rows
 * [row_offsets[r], row_offsets[r+1]) of the global system.  The operator is the global one
 * (x is exchanged over NVLink with ncclSend/ncclRecv before every SpMV, dot products are
 * ncclAllReduce'd); the preconditioner is block-Jacobi ILU0 (couplings to other ranks dropped),
 * the counterpart of the reference's per-subdomain ILU0 under MPI (ISTLSolver.hpp:218-235,
 * 286-298).  nccl_unique_id is the 128-byte ncclUniqueId rank 0 obtained from
 * opmgpu_nccl_unique_id and distributed by any means.  After set_pattern_bcrs_distributed the
 * solve / set_values / spmv entry points take the rank's LOCAL rows (values in the order of the
 * local pattern passed here) and the solves are collective. */
int  opmgpu_nccl_unique_id(void* id128);

/* Multi-GPU beneath this ABI: ONE host process, one worker thread per
 * GPU, the caller unaware of ranks -- the counterpart of the MPI branch inside ISTLSolver::solve
 * (ISTLSolver.hpp:283-306), which a caller of NewtonIterationBlackoilInterface does not see either.
 * The handle accepts opmgpu_set_pattern_bcrs, opmgpu_solve_bcrs3, opmgpu_solve_from_csc_blocks,
 * opmgpu_destroy and opmgpu_last_error with GLOBAL arrays; it partitions the rows itself (slabs along
 * the weakest-coupling grid axis for Cartesian patterns, contiguous row blocks otherwise; the halo
 * plan comes from the pattern), runs block-Jacobi ILU0 per GPU and the global operator with halo
 * exchange, and returns the increment in the caller's ordering.  opmgpu_multi_partition reports
 * the partition chosen at the first solve (axis 0/1/2 = i/j/k, -1 = contiguous row blocks). */
int  opmgpu_create_multi(int ngpus, const int* device_ids, opmgpu_handle* out);
int  opmgpu_multi_partition(opmgpu_handle h, int* axis, long long* row_offsets /* ngpus+1 */);
int  opmgpu_create_distributed(int device, int rank, int world, const void* nccl_unique_id,
                               opmgpu_handle* out);

/* Launch on this stream (a cudaStream_t) instead of the handle's own. */
int  opmgpu_set_stream(opmgpu_handle h, void* cuda_stream);

/* ---- the Newton-step solve ----------------------------------------------------------- */

/* Replaces the pattern half of formInterleavedSystem (...Interleaved.cpp:118-155) when the
 * caller already holds BCRS: rowptr[N+1], colidx[nnzb] ascending per row, diagonal present.
 * Runs the dependency analysis of the ILU0 sweeps; cached until the next call.
 * Distributed handles pass their LOCAL rows with GLOBAL column ids plus every rank's row
 * range. */
int  opmgpu_set_pattern_bcrs(opmgpu_handle h, int N, int nnzb, const int* rowptr, const int* colidx);
int  opmgpu_set_pattern_bcrs_distributed(opmgpu_handle h, int N_local, int nnzb_local,
                                         const int* rowptr, const long long* colidx_global,
                                         const long long* row_offsets /* [world+1] */);

/* Replaces ISTLSolver::solve(A,x,b) (ISTLSolver.hpp:283-306): vals[nnzb*9] row-major 3x3
 * blocks [eq][var], rhs/x cell-major [cell][3], x0 = 0 (...Interleaved.cpp:272-273). */
int  opmgpu_solve_bcrs3(opmgpu_handle h, const double* vals, const double* rhs, double* x,
                        const opmgpu_params* params, opmgpu_result* result);
/* Same with device-resident inputs/outputs (benchmarks; a caller that assembles on the GPU).
 * vals_dev == NULL: factorise and solve with the values already resident in the handle
 * (opmgpu_set_values_bcrs3[_dev]). */
int  opmgpu_solve_bcrs3_dev(opmgpu_handle h, const double* vals_dev, const double* rhs_dev,
                            double* x_dev, const opmgpu_params* params, opmgpu_result* result);

/* Replaces ...Interleaved.cpp:234-283 in one call: eq p1 of the nine CSC blocks is scaled by
 * matbalscale[p1] (:234-236), the pattern is the union of the pressure-derivative patterns
 * (:118-123; all nine with require_full_sparsity_pattern :127-134) and is re-analysed only
 * when it changed, values are scattered on the device (:178-193), rhs_eqmajor (3N, stride N,
 * unscaled equation values) is scaled and interleaved (:263-269), the system is solved and
 * dx_varmajor (3N, stride N) de-interleaved (:279-283). */
int  opmgpu_solve_from_csc_blocks(opmgpu_handle h, int N, const opmgpu_csc blocks[9],
                                  const double matbalscale[3], const double* rhs_eqmajor,
                                  double* dx_varmajor, const opmgpu_params* params,
                                  opmgpu_result* result);

/* ---- block sizes other than 3 ----------------------------------------------------------------
 * The reference's dispatcher instantiates Impl<np,Scalar> for np = 2..6
 * (NewtonIterationBlackoilInterleaved.cpp:467-487, .hpp:73; np = 2: two-phase decks, np >= 4: the
 * polymer / solvent extensions).  np = 3 runs the pipelined kernels; np = 2, 4, 5, 6 run
 * level-scheduled kernels with the same arithmetic (one launch per dependency level: correct and
 * bit-comparable with the oracle of that block size, not tuned); the diagonal blocks are inverted as
 * the reference's MatrixBlock does for that size (2, 3, 4: closed forms; 5, 6: dune's LU with
 * thresholded row pivoting).  Any other np answers OPMGPU_BAD_ARGUMENT.
 * opmgpu_set_block_size prepares the NEXT pattern for that block size (call it
 * before opmgpu_set_pattern_bcrs; opmgpu_solve_from_csc_blocks_np does both itself).  All arrays are
 * the np-sized analogues of the np = 3 entry points: vals[nnzb*np*np], rhs/x[N*np] cell-major,
 * blocks[np*np] with blocks[p1*np+p2] = d(eq p1)/d(var p2), rhs_eqmajor / dx_varmajor[np*N]. */
int  opmgpu_set_block_size(opmgpu_handle h, int np);
int  opmgpu_solve_bcrs_np(opmgpu_handle h, int np, const double* vals, const double* rhs, double* x,
                          const opmgpu_params* params, opmgpu_result* result);
int  opmgpu_solve_from_csc_blocks_np(opmgpu_handle h, int N, int np, const opmgpu_csc* blocks,
                                     const double* matbalscale, const double* rhs_eqmajor,
                                     double* dx_varmajor, const opmgpu_params* params,
                                     opmgpu_result* result);
/* kernel-level (parity tests): y = A x; ILU0 factors (lu_out, may be NULL) and v = w U^-1 L^-1 d
 * (d / v may be NULL) of the given values */
int  opmgpu_spmv_np(opmgpu_handle h, int np, const double* vals, const double* x, double* y);
int  opmgpu_ilu0_np(opmgpu_handle h, int np, const double* vals, double* lu_out, double w,
                    const double* d, double* v, int* bad_row);

/* ---- multicolour ILU0: a FLAGGED VARIANT, not the reference's preconditioner --------------------
 * The reference factorises in the natural cell order (Dune::bilu0_decomposition on istlA,
 * ISTLSolver.hpp:201-211), whose sweeps are a chain of nx+ny+nz-2 dependency levels.  With
 * OPMGPU_ILU_MULTICOLOUR the same ILU0 is built of the symmetric permutation P A P^T that sorts the
 * rows by colour (greedy colouring of the symmetrised block graph in natural order: red-black on a
 * 7-point stencil), so each sweep is one fully parallel pass per colour.  This is a different
 * preconditioner: iteration counts differ from the reference's and are reported separately, never
 * as parity.  What IS bit-comparable: factors and applies against the oracle run on P A P^T.
 * BiCGStab / GMRES, the operator and every vector stay in the caller's ordering.
 * opmgpu_set_ilu_ordering prepares the NEXT pattern (call it before opmgpu_set_pattern_bcrs or the
 * first opmgpu_solve_from_csc_blocks); plain single-GPU handles, 3x3 blocks, both precisions.
 *
 * OPMGPU_ILU_MULTICOLOUR_LINES ("k-lines"): for Cartesian stencil patterns in natural numbering the two
 * colours are taken over the (i,j) COLUMNS of the grid and the natural order is kept along k, i.e. the
 * vertical couplings -- the strong ones of a reservoir grid -- stay in the reference's order and only
 * the horizontal ones are re-ordered.  Each column is then a chain owned by one thread triple and each
 * sweep is two launches that walk the planes.  Any other pattern answers OPMGPU_BAD_ARGUMENT at
 * opmgpu_set_pattern_bcrs (checked row by row, not guessed).  Same rules: flagged, bit-comparable with
 * the oracle on P A P^T, never parity with the natural-order solve. */
enum { OPMGPU_ILU_NATURAL = 0, OPMGPU_ILU_MULTICOLOUR = 1, OPMGPU_ILU_MULTICOLOUR_LINES = 2 };
int  opmgpu_set_ilu_ordering(opmgpu_handle h, int ordering);
int  opmgpu_get_ilu_ordering(opmgpu_handle h);
/* Colours and permutation of the current pattern (n2p[i] = position of row i in P A P^T). */
int  opmgpu_get_ilu_permutation(opmgpu_handle h, int* ncolours, int* n2p);
/* The ordering rule alone, host only (no GPU, no handle); any output may be NULL. */
int  opmgpu_multicolour_order(int N, const int* rowptr, const int* colidx, int* ncolours, int* colour, int* n2p);
/* The k-line ordering of an nx x ny x nz grid in natural numbering, host only. */
int  opmgpu_line_order(int nx, int ny, int nz, int* n2p);

/* ---- kernel-level entry points (parity tests, micro-benchmarks) -------------------------- */

/* The operator alone: uploads the pattern for opmgpu_spmv* without the ILU0 analysis (the SpMV
 * sweep of BASELINE.json config 5 reaches sizes where matrix + factor records exceed the HBM).
 * Factorisation and solves answer OPMGPU_BAD_ARGUMENT until opmgpu_set_pattern_bcrs is called. */
int  opmgpu_set_pattern_bcrs_operator_only(opmgpu_handle h, int N, int nnzb, const int* rowptr, const int* colidx);

/* Upload BCRS values for the current pattern (device copy kept in the handle). */
int  opmgpu_set_values_bcrs3(opmgpu_handle h, const double* vals);
int  opmgpu_set_values_bcrs3_dev(opmgpu_handle h, const double* vals_dev);
/* y = A x  (Dune::MatrixAdapter::apply, ISTLSolver.hpp:303). */
int  opmgpu_spmv(opmgpu_handle h, const double* x, double* y);
int  opmgpu_spmv_dev(opmgpu_handle h, const double* x_dev, double* y_dev);
/* ILU0 of the current values (ParallelOverlappingILU0 ctor, ISTLSolver.hpp:201-211). */
int  opmgpu_ilu0_factor(opmgpu_handle h, int* bad_row);
/* Copy the factors back in the BCRS layout of the pattern, diagonal blocks inverted. */
int  opmgpu_ilu0_get_factors(opmgpu_handle h, double* lu);
/* v = w U^-1 L^-1 d  (ParallelOverlappingILU0::apply). */
int  opmgpu_ilu0_apply(opmgpu_handle h, double w, const double* d, double* v);
int  opmgpu_ilu0_apply_dev(opmgpu_handle h, double w, const double* d_dev, double* v_dev);
/* sum_i x_i y_i (Dune::SeqScalarProduct::dot), deterministic tree order. */
int  opmgpu_dot(opmgpu_handle h, const double* x, const double* y, int n, double* out);

/* Analysis facts for reports: ILU0 dependency levels, kernel launches since creation. */
int  opmgpu_num_levels(opmgpu_handle h, int* lower_levels, int* upper_levels);
long long opmgpu_launch_count(opmgpu_handle h);
/* Per-kernel-class device time of the solves since profiling was switched on, from CUDA events
 * recorded on the launching stream: [0] ILU0 applies, [1] SpMV (+fused dots), [2] BiCGStab
 * vector kernels, [3] ILU0 factorisation (the reference's own split is one number,
 * report.linear_solve_time, BlackoilModelBase_impl.hpp:290-294). */
int  opmgpu_set_profiling(opmgpu_handle h, int on);
int  opmgpu_get_profile(opmgpu_handle h, double ms[4], long long count[4]);
/* |r| after every half step of the last solve (verbosity / parity reports). */
int  opmgpu_residual_history(opmgpu_handle h, double* out, int cap, int* n);

#ifdef __cplusplus
}
#endif
#endif
