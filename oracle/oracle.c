/*
 * oracle.c -- CPU restatement of the reference's Newton-step linear solve (np = 3).
 * See oracle.h for the status of this file (test infrastructure, parity unpinned at the
 * dune-istl level) and for the floating-point contract.  Single-threaded like the
 * reference (sequential dune-istl; OpenMP is off by default, CMakeLists.txt:23).
 *
 * Each function names the reference call site it follows (paths relative to the
 * reference root) and, where the arithmetic is external, the upstream routine restated.
 */
#include "oracle.h"

#include <tgmath.h>     /* fma / sqrt / fabs follow the type of `real` */
#include <stdlib.h>
#include <string.h>

/* Block size np: 3 (three-phase black oil, the default build) or 2 (two-phase decks; the reference
 * instantiates Impl<np,Scalar> for np = 2..6, NewtonIterationBlackoilInterleaved.cpp:467-487).
 * -DORACLE_BS=2 / 4 / 5 / 6 builds liboracle_np<bs>*.so: the same functions with generic loops in
 * the same operation order (block umv / mmv: row outer, column inner) and the block inverse the
 * reference's MatrixBlock uses for that size (2x2: dune's closed form; 4x4: OPM's closed form;
 * 5, 6: dune's thresholded-pivoting LU). */
#ifndef ORACLE_BS
#define ORACLE_BS 3
#endif
#define BS ORACLE_BS              /* block size np */
#define BB (ORACLE_BS * ORACLE_BS)

void oracle_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------
 * a7  formInterleavedSystem -- pattern (opm/autodiff/NewtonIterationBlackoilInterleaved.cpp:110-155)
 * ---------------------------------------------------------------------------------- */
static int cmp_int(const void* a, const void* b)
{
    const int x = *(const int*)a, y = *(const int*)b;
    return (x > y) - (x < y);
}

int oracle_interleave_pattern(int N, int np, const oracle_csc* blocks, int require_full,
                              int* rowptr, int** colidx_out)
{
    /* Which scalar blocks contribute: d(eq)/d(pressure) always (:118-123), all of them
     * when require_full_sparsity_pattern (:127-134). */
    int nsel = 0;
    int sel[36];
    for (int p1 = 0; p1 < np; ++p1) sel[nsel++] = p1 * np + 0;
    if (require_full)
        for (int p1 = 0; p1 < np; ++p1)
            for (int p2 = 1; p2 < np; ++p2) sel[nsel++] = p1 * np + p2;

    int cap = 16;
    int* tmp = (int*)malloc(sizeof(int) * cap);
    /* pass 1: union per column, count per row.  pass 2: fill (columns visited ascending,
     * so every row receives ascending column ids -- the row-major conversion at :137). */
    int* colidx = NULL;
    int* fill = (int*)calloc((size_t)N + 1, sizeof(int));
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            rowptr[0] = 0;
            for (int r = 0; r < N; ++r) rowptr[r + 1] = rowptr[r] + fill[r];
            colidx = (int*)malloc(sizeof(int) * (size_t)(rowptr[N] > 0 ? rowptr[N] : 1));
            for (int r = 0; r < N; ++r) fill[r] = rowptr[r];
        }
        for (int c = 0; c < N; ++c) {
            int n = 0;
            for (int s = 0; s < nsel; ++s) {
                const oracle_csc* m = &blocks[sel[s]];
                for (int k = m->colptr[c]; k < m->colptr[c + 1]; ++k) {
                    if (n == cap) { cap *= 2; tmp = (int*)realloc(tmp, sizeof(int) * cap); }
                    tmp[n++] = m->rowidx[k];
                }
            }
            qsort(tmp, n, sizeof(int), cmp_int);
            int last = -1;
            for (int q = 0; q < n; ++q) {
                if (tmp[q] == last) continue;
                last = tmp[q];
                if (pass == 0) fill[last]++;
                else colidx[fill[last]++] = c;
            }
        }
    }
    free(tmp);
    free(fill);
    *colidx_out = colidx;
    return rowptr[N];
}

/* a6 + a7  scaling and value scatter (...Interleaved.cpp:234-236, :178-193).
 * AutoDiffBlock * scalar multiplies every Jacobian value once (AutoDiffBlock.hpp:617-626). */
int oracle_interleave_values(int N, int np, const oracle_csc* blocks, const double* scale,
                             const int* rowptr, const int* colidx, real* vals)
{
    memset(vals, 0, sizeof(real) * (size_t)rowptr[N] * np * np);
    for (int p1 = 0; p1 < np; ++p1) {
        for (int p2 = 0; p2 < np; ++p2) {
            const oracle_csc* s = &blocks[p1 * np + p2];
            for (int col = 0; col < N; ++col) {
                for (int k = s->colptr[col]; k < s->colptr[col + 1]; ++k) {
                    const int row = s->rowidx[k];
                    /* istlA[row][col]: binary search in the row */
                    int lo = rowptr[row], hi = rowptr[row + 1] - 1, pos = -1;
                    while (lo <= hi) {
                        const int mid = (lo + hi) >> 1;
                        if (colidx[mid] == col) { pos = mid; break; }
                        if (colidx[mid] < col) lo = mid + 1; else hi = mid - 1;
                    }
                    if (pos < 0) return -(k + 1);
                    vals[(size_t)pos * np * np + p1 * np + p2] = (real)(s->val[k] * scale[p1]);   /* rounded once when real = float, like the assignment at :189 */
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * a10  operator apply: Dune::MatrixAdapter::apply -> BCRSMatrix::mv -> block umv
 * (call site opm/autodiff/ISTLSolver.hpp:303)
 * ---------------------------------------------------------------------------------- */
void oracle_spmv3(int N, const int* rowptr, const int* colidx, const real* vals,
                  const real* x, real* y)
{
#if BS != 3
    for (int i = 0; i < N; ++i) {
        real yb[BS];
        for (int r = 0; r < BS; ++r) yb[r] = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const real* a = vals + (size_t)k * BB;
            const real* xj = x + (size_t)colidx[k] * BS;
            for (int r = 0; r < BS; ++r)
                for (int c = 0; c < BS; ++c) yb[r] = fma(a[r * BS + c], xj[c], yb[r]);
        }
        for (int r = 0; r < BS; ++r) y[(size_t)i * BS + r] = yb[r];
    }
    return;
#else
    for (int i = 0; i < N; ++i) {
        real y0 = 0.0, y1 = 0.0, y2 = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const real* a = vals + (size_t)k * BB;
            const real* xj = x + (size_t)colidx[k] * BS;
            /* umv: y[r] += a[r][c] * x[c], r outer, c inner */
            y0 = fma(a[0], xj[0], y0); y0 = fma(a[1], xj[1], y0); y0 = fma(a[2], xj[2], y0);
            y1 = fma(a[3], xj[0], y1); y1 = fma(a[4], xj[1], y1); y1 = fma(a[5], xj[2], y1);
            y2 = fma(a[6], xj[0], y2); y2 = fma(a[7], xj[1], y2); y2 = fma(a[8], xj[2], y2);
        }
        y[(size_t)i * BS + 0] = y0; y[(size_t)i * BS + 1] = y1; y[(size_t)i * BS + 2] = y2;
    }
#endif
}

/* ------------------------------------------------------------------------------------
 * a9  ILU0 factorisation: Opm::ParallelOverlappingILU0 ctor -> Dune::bilu0_decomposition
 * (call site opm/autodiff/ISTLSolver.hpp:201-211; 3x3 inverse is OPM's own, :192-194)
 * ---------------------------------------------------------------------------------- */
/* C = A * B with dune's DenseMatrix::{right,left}multiply accumulation: each entry starts
 * at 0 and adds k = 0,1,2 in turn. */
static void mat3_mul(const real* A, const real* B, real* C)
{
    for (int i = 0; i < BS; ++i)
        for (int j = 0; j < BS; ++j) {
            real s = 0.0;
            for (int k = 0; k < BS; ++k) s = fma(A[i * BS + k], B[k * BS + j], s);
            C[i * BS + j] = s;
        }
}

/* Opm::MatrixBlock<real,3,3>::invert -> ISTLUtility::invertMatrix (adjugate / det,
 * "code generated by maple").  Returns the determinant. */
static real mat3_invert(real* M)
{
#if BS == 2
    /* Dune::FMatrixHelp / DenseMatrix::invert for 2x2 (OPM's MatrixBlock forwards to it):
     * detinv = 1/(a00 a11 - a01 a10); a00 <-> a11 scaled, off-diagonals negated and scaled. */
    const real det = M[0] * M[3] - M[1] * M[2];
    const real detinv = (real)1.0 / det;
    const real temp = M[0];
    M[0] = M[3] * detinv;
    M[1] = -M[1] * detinv;
    M[2] = -M[2] * detinv;
    M[3] = temp * detinv;
    return det;
#elif BS == 4
    /* Opm's invertMatrix(FieldMatrix<K,4,4>&) (opm-simulators MatrixBlock.hpp / ISTLUtility; not in the
     * reference tree): the cofactor expansion published with Mesa's GLU (gluInvertMatrix), every entry
     * a left-to-right sum of six signed triple products (a*b)*c, det = first row . first column of the
     * adjugate, inverse = adjugate * (1/det).  inv4_terms[e] lists the products of adjugate entry e
     * (flat index 4*row+col); the first product carries inv4_lead[e], the signs then go - - + + - / + + - - +. */
    static const unsigned char inv4_terms[16][6][3] = {
        {{5,10,15},{5,11,14},{9,6,15},{9,7,14},{13,6,11},{13,7,10}},   /* 0 */
        {{1,10,15},{1,11,14},{9,2,15},{9,3,14},{13,2,11},{13,3,10}},   /* 1 */
        {{1,6,15},{1,7,14},{5,2,15},{5,3,14},{13,2,7},{13,3,6}},       /* 2 */
        {{1,6,11},{1,7,10},{5,2,11},{5,3,10},{9,2,7},{9,3,6}},         /* 3 */
        {{4,10,15},{4,11,14},{8,6,15},{8,7,14},{12,6,11},{12,7,10}},   /* 4 */
        {{0,10,15},{0,11,14},{8,2,15},{8,3,14},{12,2,11},{12,3,10}},   /* 5 */
        {{0,6,15},{0,7,14},{4,2,15},{4,3,14},{12,2,7},{12,3,6}},       /* 6 */
        {{0,6,11},{0,7,10},{4,2,11},{4,3,10},{8,2,7},{8,3,6}},         /* 7 */
        {{4,9,15},{4,11,13},{8,5,15},{8,7,13},{12,5,11},{12,7,9}},     /* 8 */
        {{0,9,15},{0,11,13},{8,1,15},{8,3,13},{12,1,11},{12,3,9}},     /* 9 */
        {{0,5,15},{0,7,13},{4,1,15},{4,3,13},{12,1,7},{12,3,5}},       /* 10 */
        {{0,5,11},{0,7,9},{4,1,11},{4,3,9},{8,1,7},{8,3,5}},           /* 11 */
        {{4,9,14},{4,10,13},{8,5,14},{8,6,13},{12,5,10},{12,6,9}},     /* 12 */
        {{0,9,14},{0,10,13},{8,1,14},{8,2,13},{12,1,10},{12,2,9}},     /* 13 */
        {{0,5,14},{0,6,13},{4,1,14},{4,2,13},{12,1,6},{12,2,5}},       /* 14 */
        {{0,5,10},{0,6,9},{4,1,10},{4,2,9},{8,1,6},{8,2,5}},           /* 15 */
    };
    /* leading sign of entry e: + where row+col is even */
    static const signed char sgn[6] = {+1, -1, -1, +1, +1, -1};
    real A[BB], inv[BB];
    memcpy(A, M, sizeof A);
    for (int e = 0; e < 16; ++e) {
        const int lead = (((e >> 2) + (e & 3)) & 1) ? -1 : +1;
        real acc = 0.0;
        for (int t = 0; t < 6; ++t) {
            const unsigned char* q = inv4_terms[e][t];
            const real prod = A[q[0]] * A[q[1]] * A[q[2]];
            if (t == 0) acc = (lead * sgn[0] > 0) ? prod : -prod;
            else if (lead * sgn[t] > 0) acc = acc + prod;
            else acc = acc - prod;
        }
        inv[e] = acc;
    }
    const real det = A[0] * inv[0] + A[1] * inv[4] + A[2] * inv[8] + A[3] * inv[12];
    const real inv_det = (real)1.0 / det;
    for (int e = 0; e < 16; ++e) M[e] = inv[e] * inv_det;
    return det;
#elif BS > 4
    /* Dune::DenseMatrix::invert, generic branch (dune-common 2.6 densematrix.hh; OPM's MatrixBlock
     * forwards to it for block sizes without a closed form): LU decomposition with row pivoting only
     * where |A_ii| < max(absolute_limit, |A|_inf * pivoting_limit) (FMatrixPrecision: 1e-80, 1e-8),
     * singular where the pivot is below max(absolute_limit, |A|_inf * singular_limit) (1e-14); then
     * L Y = I, U X = Y, and the recorded row swaps undone as column swaps, last first.
     * Returns the product of the pivots (sign of the swaps included), 0 where dune throws "matrix is singular". */
    real A[BS][BS], X[BS][BS];
    int pivot[BS];
    real norm = 0.0;
    for (int i = 0; i < BS; ++i) {
        real srow = 0.0;
        for (int j = 0; j < BS; ++j) { A[i][j] = M[i * BS + j]; srow += fabs(A[i][j]); }
        if (srow > norm) norm = srow;
        pivot[i] = i;
    }
    const real abslim = (real)1e-80;
    real pivthres = norm * (real)1e-8, singthres = norm * (real)1e-14;
    if (pivthres < abslim) pivthres = abslim;
    if (singthres < abslim) singthres = abslim;
    real det = 1.0;
    for (int i = 0; i < BS; ++i) {
        real pivmax = fabs(A[i][i]);
        if (pivmax < pivthres) {
            int imax = i;
            for (int k = i + 1; k < BS; ++k) {
                const real ab = fabs(A[k][i]);
                if (ab > pivmax) { pivmax = ab; imax = k; }
            }
            if (imax != i) {
                for (int j = 0; j < BS; ++j) { const real t = A[i][j]; A[i][j] = A[imax][j]; A[imax][j] = t; }
                pivot[i] = imax;
                det = -det;
            }
        }
        if (!(pivmax >= singthres)) return 0.0;      /* dune: FMatrixError "matrix is singular" (also catches NaN) */
        det *= A[i][i];
        for (int k = i + 1; k < BS; ++k) {
            const real factor = A[k][i] / A[i][i];
            A[k][i] = factor;
            for (int j = i + 1; j < BS; ++j) A[k][j] -= factor * A[i][j];
        }
    }
    for (int i = 0; i < BS; ++i)
        for (int j = 0; j < BS; ++j) X[i][j] = (i == j) ? 1.0 : 0.0;
    for (int i = 0; i < BS; ++i)
        for (int j = 0; j < i; ++j)
            for (int k = 0; k < BS; ++k) X[i][k] -= A[i][j] * X[j][k];
    for (int i = BS; i > 0;) {
        --i;
        for (int k = 0; k < BS; ++k) {
            for (int j = i + 1; j < BS; ++j) X[i][k] -= A[i][j] * X[j][k];
            X[i][k] /= A[i][i];
        }
    }
    for (int i = BS; i > 0;) {
        --i;
        if (i != pivot[i])
            for (int j = 0; j < BS; ++j) { const real t = X[j][pivot[i]]; X[j][pivot[i]] = X[j][i]; X[j][i] = t; }
    }
    for (int i = 0; i < BS; ++i)
        for (int j = 0; j < BS; ++j) M[i * BS + j] = X[i][j];
    return det;
#else
    real A[BB];
    memcpy(A, M, sizeof A);
    const real t4 = A[0] * A[4];
    const real t6 = A[0] * A[5];
    const real t8 = A[1] * A[3];
    const real t10 = A[2] * A[3];
    const real t12 = A[1] * A[6];
    const real t14 = A[2] * A[6];
    const real det = (t4 * A[8] - t6 * A[7] - t8 * A[8] + t10 * A[7] + t12 * A[5] - t14 * A[4]);
    const real t17 = 1.0 / det;
    M[0] = (A[4] * A[8] - A[5] * A[7]) * t17;
    M[1] = -(A[1] * A[8] - A[2] * A[7]) * t17;
    M[2] = (A[1] * A[5] - A[2] * A[4]) * t17;
    M[3] = -(A[3] * A[8] - A[5] * A[6]) * t17;
    M[4] = (A[0] * A[8] - t14) * t17;
    M[5] = -(t6 - t10) * t17;
    M[6] = (A[3] * A[7] - A[4] * A[6]) * t17;
    M[7] = -(A[0] * A[7] - t12) * t17;
    M[8] = (t4 - t8) * t17;
    return det;
#endif
}

static int find_diag(const int* rowptr, const int* colidx, int i)
{
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
        if (colidx[k] == i) return k;
    return -1;
}

int oracle_ilu0_factor3(int N, const int* rowptr, const int* colidx, real* lu)
{
    int* diag = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    for (int i = 0; i < N; ++i) {
        diag[i] = find_diag(rowptr, colidx, i);
        if (diag[i] < 0) { free(diag); return 1 + i; }     /* "diagonal entry missing" */
    }
    for (int i = 0; i < N; ++i) {
        const int iend = rowptr[i + 1];
        for (int ij = rowptr[i]; colidx[ij] < i; ++ij) {
            const int j = colidx[ij];
            real* Aij = lu + (size_t)ij * BB;
            /* L_ij = A_ij * inv(A_jj)   ((*ij).rightmultiply(*jj)) */
            real L[BB];
            mat3_mul(Aij, lu + (size_t)diag[j] * BB, L);
            memcpy(Aij, L, sizeof L);
            /* A_ik -= L_ij * A_jk for k > j present in both rows */
            int jk = diag[j] + 1, ik = ij + 1;
            const int jend = rowptr[j + 1];
            while (ik < iend && jk < jend) {
                if (colidx[ik] == colidx[jk]) {
                    real B[BB];
                    mat3_mul(L, lu + (size_t)jk * BB, B);        /* B.leftmultiply(*ij) */
                    real* Aik = lu + (size_t)ik * BB;
                    for (int q = 0; q < BB; ++q) Aik[q] -= B[q];
                    ++ik; ++jk;
                } else if (colidx[ik] < colidx[jk]) ++ik;
                else ++jk;
            }
        }
        const real det = mat3_invert(lu + (size_t)diag[i] * BB);
        if (!(det != 0.0) || !isfinite(det)) { free(diag); return 1 + i; }
    }
    free(diag);
    return 0;
}

/* a10  preconditioner apply: Opm::ParallelOverlappingILU0::apply, sequential case.
 * Lower sweep walks columns ascending; the upper factor is stored by convertToCRS in
 * reverse order, so the upper sweep walks columns DESCENDING; then v *= w. */
void oracle_ilu0_apply3(int N, const int* rowptr, const int* colidx, const real* lu,
                        double w, const real* d, real* v)
{
#if BS != 3
    for (int i = 0; i < N; ++i) {
        real rb[BS];
        for (int r = 0; r < BS; ++r) rb[r] = d[(size_t)i * BS + r];
        for (int k = rowptr[i]; k < rowptr[i + 1] && colidx[k] < i; ++k) {
            const real* a = lu + (size_t)k * BB;
            const real* vj = v + (size_t)colidx[k] * BS;
            for (int r = 0; r < BS; ++r)
                for (int c = 0; c < BS; ++c) rb[r] = fma(-a[r * BS + c], vj[c], rb[r]);      /* mmv */
        }
        for (int r = 0; r < BS; ++r) v[(size_t)i * BS + r] = rb[r];
    }
    for (int i = N - 1; i >= 0; --i) {
        real rb[BS], yb[BS];
        for (int r = 0; r < BS; ++r) rb[r] = v[(size_t)i * BS + r];
        int k = rowptr[i + 1] - 1;
        for (; colidx[k] > i; --k) {
            const real* a = lu + (size_t)k * BB;
            const real* vj = v + (size_t)colidx[k] * BS;
            for (int r = 0; r < BS; ++r)
                for (int c = 0; c < BS; ++c) rb[r] = fma(-a[r * BS + c], vj[c], rb[r]);
        }
        const real* di = lu + (size_t)k * BB;
        for (int r = 0; r < BS; ++r) {
            yb[r] = 0.0;
            for (int c = 0; c < BS; ++c) yb[r] = fma(di[r * BS + c], rb[c], yb[r]);           /* inv_[i].mv */
        }
        for (int r = 0; r < BS; ++r) v[(size_t)i * BS + r] = yb[r];
    }
    if (fabs(w - 1.0) > 1e-15)
        for (size_t q = 0; q < (size_t)N * BS; ++q) v[q] *= (real)w;
    return;
#else
    for (int i = 0; i < N; ++i) {
        real r0 = d[(size_t)i * BS], r1 = d[(size_t)i * BS + 1], r2 = d[(size_t)i * BS + 2];
        for (int k = rowptr[i]; k < rowptr[i + 1] && colidx[k] < i; ++k) {
            const real* a = lu + (size_t)k * BB;
            const real* vj = v + (size_t)colidx[k] * BS;
            /* mmv: y[r] -= a[r][c] * x[c] */
            r0 = fma(-a[0], vj[0], r0); r0 = fma(-a[1], vj[1], r0); r0 = fma(-a[2], vj[2], r0);
            r1 = fma(-a[3], vj[0], r1); r1 = fma(-a[4], vj[1], r1); r1 = fma(-a[5], vj[2], r1);
            r2 = fma(-a[6], vj[0], r2); r2 = fma(-a[7], vj[1], r2); r2 = fma(-a[8], vj[2], r2);
        }
        v[(size_t)i * BS] = r0; v[(size_t)i * BS + 1] = r1; v[(size_t)i * BS + 2] = r2;
    }
    for (int i = N - 1; i >= 0; --i) {
        real r0 = v[(size_t)i * BS], r1 = v[(size_t)i * BS + 1], r2 = v[(size_t)i * BS + 2];
        int k = rowptr[i + 1] - 1;
        for (; colidx[k] > i; --k) {
            const real* a = lu + (size_t)k * BB;
            const real* vj = v + (size_t)colidx[k] * BS;
            r0 = fma(-a[0], vj[0], r0); r0 = fma(-a[1], vj[1], r0); r0 = fma(-a[2], vj[2], r0);
            r1 = fma(-a[3], vj[0], r1); r1 = fma(-a[4], vj[1], r1); r1 = fma(-a[5], vj[2], r1);
            r2 = fma(-a[6], vj[0], r2); r2 = fma(-a[7], vj[1], r2); r2 = fma(-a[8], vj[2], r2);
        }
        const real* di = lu + (size_t)k * BB;               /* k is the diagonal now */
        /* inv_[i].mv(rhs, vBlock): y[r] = 0; y[r] += a[r][c]*x[c] */
        real y0 = 0.0, y1 = 0.0, y2 = 0.0;
        y0 = fma(di[0], r0, y0); y0 = fma(di[1], r1, y0); y0 = fma(di[2], r2, y0);
        y1 = fma(di[3], r0, y1); y1 = fma(di[4], r1, y1); y1 = fma(di[5], r2, y1);
        y2 = fma(di[6], r0, y2); y2 = fma(di[7], r1, y2); y2 = fma(di[8], r2, y2);
        v[(size_t)i * BS] = y0; v[(size_t)i * BS + 1] = y1; v[(size_t)i * BS + 2] = y2;
    }
    if (fabs(w - 1.0) > 1e-15)                                /* relaxation_ flag */
        for (size_t q = 0; q < (size_t)N * BS; ++q) v[q] *= (real)w;
#endif
}

/* ------------------------------------------------------------------------------------
 * a10  Dune::BiCGSTABSolver::apply  (call site opm/autodiff/ISTLSolver.hpp:267-272),
 * Dune::SeqScalarProduct dot / norm: per-block partial sums accumulated in block order.
 * ---------------------------------------------------------------------------------- */
static real vdot(int N, const real* x, const real* y)
{
    real sum = 0.0;
    for (int i = 0; i < N; ++i) {
        real s = 0.0;
        for (int c = 0; c < BS; ++c) s = fma(x[(size_t)i * BS + c], y[(size_t)i * BS + c], s);
        sum += s;
    }
    return sum;
}
static real vnorm(int N, const real* x) { return sqrt(vdot(N, x, x)); }

static void precond(int N, const int* rowptr, const int* colidx, const real* lu, double w,
                    const real* d, real* v)
{
    if (lu) oracle_ilu0_apply3(N, rowptr, colidx, lu, w, d, v);
    else memcpy(v, d, sizeof(real) * (size_t)N * BS);
}

void oracle_bicgstab3(int N, const int* rowptr, const int* colidx, const real* vals,
                      const real* lu, double w, real* b, real* x,
                      double reduction_, int maxiter, int max_half_steps,
                      double* history, int history_cap, oracle_result* res)
{
    const real EPSILON = (real)1e-80;                          /* dune: real_type EPSILON = 1e-80 (0 in float) */
    const real reduction = (real)reduction_;
    const size_t n = (size_t)N * BS;
    real* r = b;                                            /* X& r = b */
    real* p = (real*)calloc(n ? n : 1, sizeof(real));
    real* v = (real*)calloc(n ? n : 1, sizeof(real));
    real* t = (real*)calloc(n ? n : 1, sizeof(real));
    real* y = (real*)calloc(n ? n : 1, sizeof(real));
    real* rt = (real*)malloc(sizeof(real) * (n ? n : 1));
    real rho = 1.0, alpha = 1.0, omega = 1.0, rho_new, beta, h;
    real norm, norm_0;
    double it;
    int half = 0;

    memset(res, 0, sizeof *res);
    /* r = b - A x   (_op.applyscaleadd(-1,x,r)); t doubles as scratch */
    oracle_spmv3(N, rowptr, colidx, vals, x, t);
    for (size_t q = 0; q < n; ++q) r[q] -= t[q];
    memcpy(rt, r, sizeof(real) * n);
    norm = norm_0 = vnorm(N, r);
    res->norm0 = norm_0;

    if (norm < norm_0 * reduction || norm < 1e-30) {
        res->converged = 1; res->iterations = 0; res->reduction = 0.0;
        goto done;
    }

    for (it = 0.5; it < maxiter; it += 0.5) {
        if (max_half_steps >= 0 && half >= max_half_steps) break;
        rho_new = vdot(N, rt, r);
        if (fabs(rho) <= EPSILON || fabs(omega) <= EPSILON) { res->status = 3; break; }
        if (it < 1) {
            memcpy(p, r, sizeof(real) * n);
        } else {
            beta = (rho_new / rho) * (alpha / omega);
            for (size_t q = 0; q < n; ++q) {                  /* p.axpy(-omega,v); p*=beta; p+=r */
                real pq = fma(-omega, v[q], p[q]);
                pq *= beta;
                p[q] = pq + r[q];
            }
        }
        memset(y, 0, sizeof(real) * n);
        precond(N, rowptr, colidx, lu, w, p, y);
        oracle_spmv3(N, rowptr, colidx, vals, y, v);
        h = vdot(N, rt, v);
        if (fabs(h) < EPSILON) { res->status = 3; break; }
        alpha = rho_new / h;
        for (size_t q = 0; q < n; ++q) x[q] = fma(alpha, y[q], x[q]);
        for (size_t q = 0; q < n; ++q) r[q] = fma(-alpha, v[q], r[q]);
        norm = vnorm(N, r);
        if (history && half < history_cap) history[half] = norm;
        ++half;
        if (norm < norm_0 * reduction) { res->converged = 1; break; }
        it += 0.5;
        if (max_half_steps >= 0 && half >= max_half_steps) break;

        memset(y, 0, sizeof(real) * n);
        precond(N, rowptr, colidx, lu, w, r, y);
        oracle_spmv3(N, rowptr, colidx, vals, y, t);
        omega = vdot(N, t, r) / vdot(N, t, t);
        for (size_t q = 0; q < n; ++q) x[q] = fma(omega, y[q], x[q]);
        for (size_t q = 0; q < n; ++q) r[q] = fma(-omega, t[q], r[q]);
        rho = rho_new;
        norm = vnorm(N, r);
        if (history && half < history_cap) history[half] = norm;
        ++half;
        if (norm < norm_0 * reduction || norm < 1e-30) { res->converged = 1; break; }
    }
    if (it > maxiter) it = maxiter;                           /* it = min(maxit, it) */
    res->iterations = (int)ceil(it);
    res->reduction = norm / norm_0;
    if (!res->converged && res->status == 0) res->status = 1;

done:
    res->half_steps = half;
    free(p); free(v); free(t); free(y); free(rt);
}

/* ------------------------------------------------------------------------------------
 * f3  Dune::RestartedGMResSolver::apply  (call site opm/autodiff/ISTLSolver.hpp:257-265,
 * selected by newton_use_gmres, restart = linear_solver_restart), dune-istl 2.6: LEFT
 * preconditioned (the Krylov space is built on W^-1 A, the defect that is measured is the
 * preconditioned one), modified Gram-Schmidt, Givens rotations, x updated at the end of
 * every cycle.  Like the BiCGStab above this is a restatement of an external routine
 * ("parity unpinned" at the dune level; pinned by scipy and known-answer tests).
 * b is overwritten (defect of the last restart), x0 as given.
 * ---------------------------------------------------------------------------------- */
static void gen_rotation(real dx, real dy, real* cs, real* sn)
{
    const real ndx = fabs(dx), ndy = fabs(dy);
    if (ndy < 1e-15) { *cs = 1.0; *sn = 0.0; }
    else if (ndx < 1e-15) { *cs = 0.0; *sn = 1.0; }
    else if (ndy > ndx) {
        const real temp = ndx / ndy;
        *cs = 1.0 / sqrt(1.0 + temp * temp);
        *sn = *cs;
        *cs *= temp;
        *sn *= dx / ndx;
        *sn *= dy / ndy;
    } else {
        const real temp = ndy / ndx;
        *cs = 1.0 / sqrt(1.0 + temp * temp);
        *sn = *cs;
        *sn *= dy / dx;
    }
}
static void apply_rotation(real* dx, real* dy, real cs, real sn)
{
    const real temp = cs * (*dx) + sn * (*dy);
    *dy = -sn * (*dx) + cs * (*dy);
    *dx = temp;
}
/* b -= A x  (MatrixAdapter::applyscaleadd(-1,x,b) -> BCRSMatrix::usmv: per block y -= a*x) */
static void residual_update(int N, const int* rowptr, const int* colidx, const real* vals,
                            const real* x, real* b)
{
#if BS != 3
    for (int i = 0; i < N; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const real* a = vals + (size_t)k * BB;
            const real* xj = x + (size_t)colidx[k] * BS;
            for (int r = 0; r < BS; ++r)
                for (int c = 0; c < BS; ++c) b[(size_t)i * BS + r] = fma(-a[r * BS + c], xj[c], b[(size_t)i * BS + r]);
        }
    return;
#endif
    for (int i = 0; i < N; ++i) {
        real r0 = b[(size_t)i * BS], r1 = b[(size_t)i * BS + 1], r2 = b[(size_t)i * BS + 2];
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const real* a = vals + (size_t)k * BB;
            const real* xj = x + (size_t)colidx[k] * BS;
            r0 = fma(-a[0], xj[0], r0); r0 = fma(-a[1], xj[1], r0); r0 = fma(-a[2], xj[2], r0);
            r1 = fma(-a[3], xj[0], r1); r1 = fma(-a[4], xj[1], r1); r1 = fma(-a[5], xj[2], r1);
            r2 = fma(-a[6], xj[0], r2); r2 = fma(-a[7], xj[1], r2); r2 = fma(-a[8], xj[2], r2);
        }
        b[(size_t)i * BS] = r0; b[(size_t)i * BS + 1] = r1; b[(size_t)i * BS + 2] = r2;
    }
}

void oracle_gmres3(int N, const int* rowptr, const int* colidx, const real* vals,
                   const real* lu, double wrelax, real* b, real* x,
                   double reduction_, int maxiter, int restart,
                   double* history, int history_cap, oracle_result* res)
{
    const real reduction = (real)reduction_;
    const real EPSILON = 1e-80;
    const size_t n = (size_t)N * BS;
    const int m = restart;
    real* s = (real*)calloc((size_t)m + 1, sizeof(real));
    real* sn = (real*)calloc((size_t)m, sizeof(real));
    real* cs = (real*)calloc((size_t)m, sizeof(real));
    real* H = (real*)calloc((size_t)(m + 1) * m, sizeof(real));      /* H[k][i] = H[k*m + i] */
    real* V = (real*)calloc((size_t)(m + 1) * (n ? n : 1), sizeof(real));
    real* w = (real*)calloc(n ? n : 1, sizeof(real));
    real* b2 = (real*)malloc(sizeof(real) * (n ? n : 1));
    real* yv = (real*)calloc((size_t)m + 1, sizeof(real));
    real norm, norm_0;
    int j = 1, nh = 0;

    memset(res, 0, sizeof *res);
    memcpy(b2, b, sizeof(real) * n);
    residual_update(N, rowptr, colidx, vals, x, b);               /* b -= A x */
    precond(N, rowptr, colidx, lu, wrelax, b, V);                 /* v[0] = W^-1 b */
    norm_0 = vnorm(N, V);
    norm = norm_0;
    res->norm0 = norm_0;
    if (norm_0 < EPSILON) { res->converged = 1; res->iterations = 0; goto done; }

    while (j <= maxiter && !res->converged) {
        int i = 0;
        const real inv = 1.0 / norm;
        for (size_t q = 0; q < n; ++q) V[q] *= inv;               /* v[0] *= 1/norm */
        s[0] = norm;
        for (i = 1; i < m + 1; ++i) s[i] = 0.0;
        for (i = 0; i < m && j <= maxiter && !res->converged; ++i, ++j) {
            real* vi = V + (size_t)i * n;
            real* vn = V + (size_t)(i + 1) * n;
            oracle_spmv3(N, rowptr, colidx, vals, vi, vn);        /* _A.apply(v[i], v[i+1]) */
            precond(N, rowptr, colidx, lu, wrelax, vn, w);        /* _W.apply(w, v[i+1]) */
            for (int k = 0; k < i + 1; ++k) {
                const real* vk = V + (size_t)k * n;
                const real hki = vdot(N, vk, w);
                H[(size_t)k * m + i] = hki;
                for (size_t q = 0; q < n; ++q) w[q] = fma(-hki, vk[q], w[q]);     /* w.axpy(-H[k][i], v[k]) */
            }
            const real hn = vnorm(N, w);
            H[(size_t)(i + 1) * m + i] = hn;
            if (fabs(hn) < EPSILON) { res->status = 3; goto finish; }             /* breakdown */
            const real hinv = 1.0 / hn;
            for (size_t q = 0; q < n; ++q) vn[q] = w[q] * hinv;                   /* v[i+1] = w; v[i+1] *= 1/H */
            for (int k = 0; k < i; ++k)
                apply_rotation(&H[(size_t)k * m + i], &H[(size_t)(k + 1) * m + i], cs[k], sn[k]);
            gen_rotation(H[(size_t)i * m + i], H[(size_t)(i + 1) * m + i], &cs[i], &sn[i]);
            apply_rotation(&H[(size_t)i * m + i], &H[(size_t)(i + 1) * m + i], cs[i], sn[i]);
            apply_rotation(&s[i], &s[i + 1], cs[i], sn[i]);
            norm = fabs(s[i + 1]);
            if (history && nh < history_cap) history[nh] = norm;
            ++nh;
            if (norm < reduction * norm_0) res->converged = 1;
        }
        /* update(w, i, H, s, v): back substitution, x += sum y[a] v[a] accumulated in w */
        memset(w, 0, sizeof(real) * n);
        for (int a = 0; a < m + 1; ++a) yv[a] = s[a];
        for (int a = i - 1; a >= 0; --a) {
            real rhs = s[a];
            for (int c = a + 1; c < i; ++c) rhs -= H[(size_t)a * m + c] * yv[c];
            yv[a] = rhs / H[(size_t)a * m + a];
            const real* va = V + (size_t)a * n;
            for (size_t q = 0; q < n; ++q) w[q] = fma(yv[a], va[q], w[q]);
        }
        for (size_t q = 0; q < n; ++q) x[q] += w[q];
        if (!res->converged && j <= maxiter) {
            memcpy(b, b2, sizeof(real) * n);
            residual_update(N, rowptr, colidx, vals, x, b);
            precond(N, rowptr, colidx, lu, wrelax, b, V);
            norm = vnorm(N, V);
        }
    }
finish:
    res->iterations = j - 1;
    res->reduction = norm / norm_0;
    if (!res->converged && res->status == 0) res->status = 1;
done:
    res->half_steps = nh;
    free(s); free(sn); free(cs); free(H); free(V); free(w); free(b2); free(yv);
}

/* factor + GMRES on a BCRS system (vals untouched), x0 = 0 */
void oracle_solve_gmres_bcrs3(int N, const int* rowptr, const int* colidx, const real* vals,
                              const real* rhs_cellmajor, real* x_cellmajor,
                              double reduction, int maxiter, double relax, int restart,
                              oracle_result* res)
{
    const size_t nnzb = (size_t)rowptr[N];
    real* lu = (real*)malloc(sizeof(real) * (nnzb ? nnzb : 1) * BB);
    real* b = (real*)malloc(sizeof(real) * ((size_t)N * BS + 1));
    memcpy(lu, vals, sizeof(real) * nnzb * BB);
    memcpy(b, rhs_cellmajor, sizeof(real) * (size_t)N * BS);
    memset(x_cellmajor, 0, sizeof(real) * (size_t)N * BS);
    const int bad = oracle_ilu0_factor3(N, rowptr, colidx, lu);
    if (bad) { memset(res, 0, sizeof *res); res->status = 2; res->bad_row = bad - 1; }
    else oracle_gmres3(N, rowptr, colidx, vals, lu, relax, b, x_cellmajor, reduction, maxiter, restart, NULL, 0, res);
    free(lu); free(b);
}

void oracle_solve_bcrs3(int N, const int* rowptr, const int* colidx, const real* vals,
                        const real* rhs_cellmajor, real* x_cellmajor,
                        double reduction, int maxiter, double relax, int max_half_steps,
                        oracle_result* res)
{
    const size_t nnzb = (size_t)rowptr[N];
    real* lu = (real*)malloc(sizeof(real) * (nnzb ? nnzb : 1) * BB);
    real* b = (real*)malloc(sizeof(real) * ((size_t)N * BS + 1));
    memcpy(lu, vals, sizeof(real) * nnzb * BB);             /* ILU works on a copy of A */
    memcpy(b, rhs_cellmajor, sizeof(real) * (size_t)N * BS);
    memset(x_cellmajor, 0, sizeof(real) * (size_t)N * BS);  /* x = 0.0, ...Interleaved.cpp:272-273 */
    const int bad = oracle_ilu0_factor3(N, rowptr, colidx, lu);
    if (bad) {
        memset(res, 0, sizeof *res);
        res->status = 2; res->bad_row = bad - 1;
    } else {
        oracle_bicgstab3(N, rowptr, colidx, vals, lu, relax, b, x_cellmajor, reduction, maxiter,
                         max_half_steps, NULL, 0, res);
    }
    free(lu); free(b);
}

/* a6..a11  Impl<3,real>::computeNewtonIncrement without wells (...Interleaved.cpp:234-283) */
void oracle_solve_from_csc_blocks(int N, const oracle_csc* blocks9, const double* matbalscale,
                                  const double* rhs_eqmajor, double* dx_varmajor,
                                  double reduction, int maxiter, double relax,
                                  int require_full, oracle_result* res)
{
    int* rowptr = (int*)malloc(sizeof(int) * ((size_t)N + 1));
    int* colidx = NULL;
    const int nnzb = oracle_interleave_pattern(N, BS, blocks9, require_full, rowptr, &colidx);
    real* vals = (real*)malloc(sizeof(real) * (size_t)(nnzb > 0 ? nnzb : 1) * BB);
    real* b = (real*)malloc(sizeof(real) * ((size_t)N * BS + 1));
    real* x = (real*)malloc(sizeof(real) * ((size_t)N * BS + 1));
    const int bad = oracle_interleave_values(N, BS, blocks9, matbalscale, rowptr, colidx, vals);
    if (bad) {
        memset(res, 0, sizeof *res);
        res->status = 4;                                      /* entry outside pattern */
    } else {
        /* b is the concatenation of the SCALED equation values (:234-253), interleaved at :263-269 */
        for (int i = 0; i < N; ++i)
            for (int pp = 0; pp < BS; ++pp)
                b[(size_t)i * BS + pp] = (real)(rhs_eqmajor[(size_t)pp * N + i] * matbalscale[pp]);
        oracle_solve_bcrs3(N, rowptr, colidx, vals, b, x, reduction, maxiter, relax, -1, res);
        for (int i = 0; i < N; ++i)                           /* :279-283 */
            for (int pp = 0; pp < BS; ++pp)
                dx_varmajor[(size_t)pp * N + i] = x[(size_t)i * BS + pp];
    }
    free(rowptr); free(colidx); free(vals); free(b); free(x);
}
