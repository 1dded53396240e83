/*
 * oracle.h -- CPU restatement of OPM legacy's interleaved Newton-step linear solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (opm_simulators_legacy_b200/,
 * include/) may include, link or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker
 * or as the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at the dune-istl level.  The arithmetic of this path
 * (Dune::BCRSMatrix::mv, Dune::bilu0_decomposition, Opm::ParallelOverlappingILU0::apply,
 * Opm::MatrixBlock 3x3 inverse, Dune::BiCGSTABSolver::apply, Dune::SeqScalarProduct) lives in
 * dune-istl (>= 2.4, branches up to 2.6; dune.module:12) and opm-simulators (2019.04-pre),
 * neither of which is under /root/reference nor buildable in this image, and no reference
 * test asserts a numerical result at this boundary (tests/test_linearsolver.cpp:109-121
 * computes `exact` and never compares it).  The oracle restates the published algorithms
 * of those versions and is pinned by (1) the reference's own fixtures turned into
 * known-answer tests, (2) an independent scipy implementation, (3) defining properties of
 * ILU0 -- see tests/test_oracle_pins.py.
 *
 * Floating-point contract of the restatement: every `y += a*x` / `y -= a*x` in the dune
 * loops is ONE fused multiply-add (what gcc -O3 -march=<any FMA target> emits for them on any FMA
 * host); everything else is a separately rounded IEEE operation.  The file is compiled
 * with -ffp-contract=off so the compiler adds or removes nothing.
 */
#ifndef OPM_B200_ORACLE_H
#define OPM_B200_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* One scalar Jacobian block d(eq p1)/d(var p2) as Eigen hands it over: column-major
 * compressed (AutoDiffMatrix::getSparse, opm/autodiff/AutoDiffMatrix.hpp:635-648). */
typedef struct {
    const int*    colptr;   /* N+1 column starts (outerIndexPtr)  */
    const int*    rowidx;   /* nnz row ids, ascending per column   */
    const double* val;      /* nnz values                          */
} oracle_csc;

/* Scalar type of the instance: double = the reference's Impl<3,double>, float (build with
 * -DORACLE_REAL=float -> liboracle_f32.so) = its Impl<3,float>, selected by
 * LinearisedBlackoilResidual::singlePrecision (NewtonIterationBlackoilInterleaved.cpp:467-487).
 * Matrix values, factors and vectors are `real`; the caller-side data (CSC Jacobian values,
 * scaling, equation-major right-hand side and increment) and the parameters stay double. */
#ifndef ORACLE_REAL
#define ORACLE_REAL double
#endif
typedef ORACLE_REAL real;

typedef struct {
    int    iterations;      /* ceil(it), dune convention                     */
    int    converged;       /* 1 when a stop test fired                      */
    int    half_steps;      /* number of half iterations actually performed  */
    int    status;          /* 0 ok, 1 not converged, 2 singular block, 3 breakdown */
    int    bad_row;         /* row of singular diagonal block (status 2)     */
    double reduction;       /* |r| / |r0| at exit                            */
    double norm0;           /* |r0|                                          */
} oracle_result;

/* Pattern of the interleaved system: union of the np pressure-derivative patterns
 * (blocks[p1*np+0]), optionally of all np*np blocks, converted to row-major with
 * ascending columns.  NewtonIterationBlackoilInterleaved.cpp:118-155.
 * rowptr has N+1 entries; *colidx_out is malloc'ed (free with oracle_free). */
int oracle_interleave_pattern(int N, int np, const oracle_csc* blocks, int require_full,
                              int* rowptr, int** colidx_out);

/* Scatter the np*np CSC blocks into row-major np x np BCRS blocks, eq-major inside the
 * block ([p1][p2]); eq p1 is multiplied by scale[p1] first.  Returns -(k+1) if entry k
 * falls outside the pattern (dune throws there).  ...Interleaved.cpp:178-193, :234-236. */
int oracle_interleave_values(int N, int np, const oracle_csc* blocks, const double* scale,
                             const int* rowptr, const int* colidx, real* vals);

/* y = A x, Dune::BCRSMatrix::mv: y=0 then umv per block, ascending columns. */
void oracle_spmv3(int N, const int* rowptr, const int* colidx, const real* vals,
                  const real* x, real* y);

/* In-place block ILU(0), natural order (Dune::bilu0_decomposition); diagonal blocks end
 * up INVERTED (Opm::MatrixBlock 3x3 cofactor inverse).  Returns 0, or 1+row on a missing /
 * singular diagonal block. */
int oracle_ilu0_factor3(int N, const int* rowptr, const int* colidx, real* lu);

/* v = w * U^-1 L^-1 d  (Opm::ParallelOverlappingILU0::apply, sequential case). */
void oracle_ilu0_apply3(int N, const int* rowptr, const int* colidx, const real* lu,
                        double w, const real* d, real* v);

/* Dune::BiCGSTABSolver::apply with SeqScalarProduct, x0 as given (caller passes zeros),
 * preconditioner = oracle_ilu0_apply3 on `lu` (pass lu=NULL for the identity).
 * b is overwritten by the final residual like dune does.  max_half_steps < 0 means no
 * extra limit; otherwise the loop also stops after that many half iterations (used to
 * compare iterates at equal half-step counts).  history (optional) receives |r| after
 * every half step, at most history_cap entries. */
void oracle_bicgstab3(int N, const int* rowptr, const int* colidx, const real* vals,
                      const real* lu, double w, real* b, real* x,
                      double reduction, int maxiter, int max_half_steps,
                      double* history, int history_cap, oracle_result* res);

/* The whole reference path a6..a11 for np = 3 (Impl<3,real>::computeNewtonIncrement,
 * ...Interleaved.cpp:234-283, with the wells already eliminated): scale, pattern, values,
 * interleave rhs, ILU0, BiCGStab, de-interleave.  rhs_eqmajor / dx_varmajor have 3N
 * entries with stride N. */
void oracle_solve_from_csc_blocks(int N, const oracle_csc* blocks9, const double* matbalscale,
                                  const double* rhs_eqmajor, double* dx_varmajor,
                                  double reduction, int maxiter, double relax,
                                  int require_full, oracle_result* res);

/* Convenience: factor + solve on a BCRS system (vals untouched). */
void oracle_solve_bcrs3(int N, const int* rowptr, const int* colidx, const real* vals,
                        const real* rhs_cellmajor, real* x_cellmajor,
                        double reduction, int maxiter, double relax, int max_half_steps,
                        oracle_result* res);

/* Dune::RestartedGMResSolver::apply (dune-istl 2.6: left preconditioned, modified Gram-Schmidt,
 * Givens rotations; ISTLSolver.hpp:257-265 with restart = linear_solver_restart).  iterations =
 * Arnoldi steps; history receives the preconditioned defect norm after every step. */
void oracle_gmres3(int N, const int* rowptr, const int* colidx, const real* vals,
                   const real* lu, double w, real* b, real* x,
                   double reduction, int maxiter, int restart,
                   double* history, int history_cap, oracle_result* res);
void oracle_solve_gmres_bcrs3(int N, const int* rowptr, const int* colidx, const real* vals,
                              const real* rhs_cellmajor, real* x_cellmajor,
                              double reduction, int maxiter, double relax, int restart,
                              oracle_result* res);

void oracle_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
