/*
 * oracle_omp.c -- "Baseline B" of BASELINE.md / SURVEY.md section 8(d): the oracle's ILU0 +
 * BiCGStab with OpenMP on all host cores (level-scheduled factorisation and sweeps, row-parallel
 * SpMV, parallel dot products).
 *
 * TEST / BENCHMARK INFRASTRUCTURE ONLY (same rule as oracle.h).  This is NOT the reference: the
 * reference's solver is sequential (OpenMP is off by default, CMakeLists.txt:23, and never used
 * inside the solver).  It answers "what would the same algorithm do on every core of the GPU
 * box's host".  Per row it runs the oracle's arithmetic in the oracle's order, so factors, sweeps
 * and SpMV are bit-identical to oracle.c; the dot products are reduced in a different order, so
 * iterates agree to rounding, not bitwise.
 */
#include "oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#define BS 3
#define BB 9

typedef struct {
    int nlev;
    int* ptr;    /* [nlev+1] */
    int* rows;   /* [N] rows grouped by level, ascending inside a level */
} levels_t;

static void build_levels(int N, const int* rowptr, const int* colidx, int lower, levels_t* out)
{
    int* lvl = (int*)calloc((size_t)(N > 0 ? N : 1), sizeof(int));
    int nlev = 0;
    if (lower) {
        for (int i = 0; i < N; ++i) {
            int l = 0;
            for (int k = rowptr[i]; k < rowptr[i + 1] && colidx[k] < i; ++k) if (lvl[colidx[k]] + 1 > l) l = lvl[colidx[k]] + 1;
            lvl[i] = l;
            if (l + 1 > nlev) nlev = l + 1;
        }
    } else {
        for (int i = N - 1; i >= 0; --i) {
            int l = 0;
            for (int k = rowptr[i + 1] - 1; k >= rowptr[i] && colidx[k] > i; --k) if (lvl[colidx[k]] + 1 > l) l = lvl[colidx[k]] + 1;
            lvl[i] = l;
            if (l + 1 > nlev) nlev = l + 1;
        }
    }
    out->nlev = nlev;
    out->ptr = (int*)calloc((size_t)nlev + 2, sizeof(int));
    out->rows = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    for (int i = 0; i < N; ++i) out->ptr[lvl[i] + 1]++;
    for (int l = 0; l < nlev; ++l) out->ptr[l + 1] += out->ptr[l];
    int* fill = (int*)malloc(sizeof(int) * ((size_t)nlev + 1));
    memcpy(fill, out->ptr, sizeof(int) * ((size_t)nlev + 1));
    for (int i = 0; i < N; ++i) out->rows[fill[lvl[i]]++] = i;
    free(fill); free(lvl);
}
static void free_levels(levels_t* l) { free(l->ptr); free(l->rows); }

static void mat3_mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < BS; ++i)
        for (int j = 0; j < BS; ++j) {
            double s = 0.0;
            for (int k = 0; k < BS; ++k) s = fma(A[i * BS + k], B[k * BS + j], s);
            C[i * BS + j] = s;
        }
}
static double mat3_invert(double* M)
{
    double A[BB];
    memcpy(A, M, sizeof A);
    const double t4 = A[0] * A[4], t6 = A[0] * A[5], t8 = A[1] * A[3], t10 = A[2] * A[3], t12 = A[1] * A[6], t14 = A[2] * A[6];
    const double det = (t4 * A[8] - t6 * A[7] - t8 * A[8] + t10 * A[7] + t12 * A[5] - t14 * A[4]);
    const double t17 = 1.0 / det;
    M[0] = (A[4] * A[8] - A[5] * A[7]) * t17; M[1] = -(A[1] * A[8] - A[2] * A[7]) * t17; M[2] = (A[1] * A[5] - A[2] * A[4]) * t17;
    M[3] = -(A[3] * A[8] - A[5] * A[6]) * t17; M[4] = (A[0] * A[8] - t14) * t17; M[5] = -(t6 - t10) * t17;
    M[6] = (A[3] * A[7] - A[4] * A[6]) * t17; M[7] = -(A[0] * A[7] - t12) * t17; M[8] = (t4 - t8) * t17;
    return det;
}

static void omp_spmv3(int N, const int* rowptr, const int* colidx, const double* vals, const double* x, double* y)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const double* a = vals + (size_t)k * BB;
            const double* xj = x + (size_t)colidx[k] * BS;
            y0 = fma(a[0], xj[0], y0); y0 = fma(a[1], xj[1], y0); y0 = fma(a[2], xj[2], y0);
            y1 = fma(a[3], xj[0], y1); y1 = fma(a[4], xj[1], y1); y1 = fma(a[5], xj[2], y1);
            y2 = fma(a[6], xj[0], y2); y2 = fma(a[7], xj[1], y2); y2 = fma(a[8], xj[2], y2);
        }
        y[(size_t)i * BS] = y0; y[(size_t)i * BS + 1] = y1; y[(size_t)i * BS + 2] = y2;
    }
}

/* rows of one level only read rows of earlier levels: bilu0_decomposition row by row */
static int omp_ilu0_factor3(int N, const int* rowptr, const int* colidx, const int* diag, const levels_t* L, double* lu)
{
    int bad = 0;
    for (int l = 0; l < L->nlev; ++l) {
#pragma omp parallel for schedule(static)
        for (int q = L->ptr[l]; q < L->ptr[l + 1]; ++q) {
            const int i = L->rows[q], iend = rowptr[i + 1];
            for (int ij = rowptr[i]; colidx[ij] < i; ++ij) {
                const int j = colidx[ij];
                double* Aij = lu + (size_t)ij * BB;
                double Lb[BB];
                mat3_mul(Aij, lu + (size_t)diag[j] * BB, Lb);
                memcpy(Aij, Lb, sizeof Lb);
                int jk = diag[j] + 1, ik = ij + 1;
                const int jend = rowptr[j + 1];
                while (ik < iend && jk < jend) {
                    if (colidx[ik] == colidx[jk]) {
                        double B[BB];
                        mat3_mul(Lb, lu + (size_t)jk * BB, B);
                        double* Aik = lu + (size_t)ik * BB;
                        for (int t = 0; t < BB; ++t) Aik[t] -= B[t];
                        ++ik; ++jk;
                    } else if (colidx[ik] < colidx[jk]) ++ik;
                    else ++jk;
                }
            }
            const double det = mat3_invert(lu + (size_t)diag[i] * BB);
            if (!(det != 0.0) || !isfinite(det)) {
#pragma omp critical
                if (!bad || i + 1 < bad) bad = i + 1;
            }
        }
        if (bad) return bad;
    }
    (void)N;
    return 0;
}

static void omp_ilu0_apply3(int N, const int* rowptr, const int* colidx, const int* diag, const levels_t* LL,
                            const levels_t* LU, const double* lu, double w, const double* d, double* v)
{
    for (int l = 0; l < LL->nlev; ++l) {
#pragma omp parallel for schedule(static)
        for (int q = LL->ptr[l]; q < LL->ptr[l + 1]; ++q) {
            const int i = LL->rows[q];
            double r0 = d[(size_t)i * BS], r1 = d[(size_t)i * BS + 1], r2 = d[(size_t)i * BS + 2];
            for (int k = rowptr[i]; k < diag[i]; ++k) {
                const double* a = lu + (size_t)k * BB;
                const double* vj = v + (size_t)colidx[k] * BS;
                r0 = fma(-a[0], vj[0], r0); r0 = fma(-a[1], vj[1], r0); r0 = fma(-a[2], vj[2], r0);
                r1 = fma(-a[3], vj[0], r1); r1 = fma(-a[4], vj[1], r1); r1 = fma(-a[5], vj[2], r1);
                r2 = fma(-a[6], vj[0], r2); r2 = fma(-a[7], vj[1], r2); r2 = fma(-a[8], vj[2], r2);
            }
            v[(size_t)i * BS] = r0; v[(size_t)i * BS + 1] = r1; v[(size_t)i * BS + 2] = r2;
        }
    }
    for (int l = 0; l < LU->nlev; ++l) {
#pragma omp parallel for schedule(static)
        for (int q = LU->ptr[l]; q < LU->ptr[l + 1]; ++q) {
            const int i = LU->rows[q];
            double r0 = v[(size_t)i * BS], r1 = v[(size_t)i * BS + 1], r2 = v[(size_t)i * BS + 2];
            for (int k = rowptr[i + 1] - 1; k > diag[i]; --k) {
                const double* a = lu + (size_t)k * BB;
                const double* vj = v + (size_t)colidx[k] * BS;
                r0 = fma(-a[0], vj[0], r0); r0 = fma(-a[1], vj[1], r0); r0 = fma(-a[2], vj[2], r0);
                r1 = fma(-a[3], vj[0], r1); r1 = fma(-a[4], vj[1], r1); r1 = fma(-a[5], vj[2], r1);
                r2 = fma(-a[6], vj[0], r2); r2 = fma(-a[7], vj[1], r2); r2 = fma(-a[8], vj[2], r2);
            }
            const double* di = lu + (size_t)diag[i] * BB;
            double y0 = 0.0, y1 = 0.0, y2 = 0.0;
            y0 = fma(di[0], r0, y0); y0 = fma(di[1], r1, y0); y0 = fma(di[2], r2, y0);
            y1 = fma(di[3], r0, y1); y1 = fma(di[4], r1, y1); y1 = fma(di[5], r2, y1);
            y2 = fma(di[6], r0, y2); y2 = fma(di[7], r1, y2); y2 = fma(di[8], r2, y2);
            v[(size_t)i * BS] = y0; v[(size_t)i * BS + 1] = y1; v[(size_t)i * BS + 2] = y2;
        }
    }
    if (fabs(w - 1.0) > 1e-15) {
#pragma omp parallel for schedule(static)
        for (long long q = 0; q < (long long)N * BS; ++q) v[q] *= w;
    }
}

static double omp_dot(size_t n, const double* x, const double* y)
{
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (long long q = 0; q < (long long)n; ++q) s += x[q] * y[q];
    return s;
}

/* factor + Dune::BiCGSTABSolver::apply (x0 = 0) with nthreads OpenMP threads (<= 0: all);
 * ms_factor / ms_solve are wall-clock milliseconds of the two phases (levels excluded: they
 * belong to the pattern, like the GPU library's analysis). */
int oracle_omp_solve_bcrs3(int N, const int* rowptr, const int* colidx, const double* vals, const double* rhs,
                           double* x, double reduction, int maxiter, double relax, int nthreads,
                           oracle_result* res, double* ms_factor, double* ms_solve, int* threads_used)
{
    const double EPSILON = 1e-80;
    const size_t n = (size_t)N * BS, nnzb = (size_t)rowptr[N];
    if (nthreads > 0) omp_set_num_threads(nthreads);
    int nt = 1;
#pragma omp parallel
    {
#pragma omp single
        nt = omp_get_num_threads();
    }
    if (threads_used) *threads_used = nt;
    memset(res, 0, sizeof *res);
    int* diag = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    for (int i = 0; i < N; ++i) {
        diag[i] = -1;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) if (colidx[k] == i) diag[i] = k;
        if (diag[i] < 0) { res->status = 2; res->bad_row = i; free(diag); return 2; }
    }
    levels_t LL, LU;
    build_levels(N, rowptr, colidx, 1, &LL);
    build_levels(N, rowptr, colidx, 0, &LU);
    double* lu = (double*)malloc(sizeof(double) * (nnzb ? nnzb : 1) * BB);
    double *r = (double*)malloc(sizeof(double) * (n + 1)), *rt = (double*)malloc(sizeof(double) * (n + 1));
    double *p = (double*)calloc(n + 1, sizeof(double)), *v = (double*)calloc(n + 1, sizeof(double));
    double *t = (double*)calloc(n + 1, sizeof(double)), *y = (double*)calloc(n + 1, sizeof(double));
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)(nnzb * BB); ++q) lu[q] = vals[q];
    const int bad = omp_ilu0_factor3(N, rowptr, colidx, diag, &LL, lu);
    double t1 = omp_get_wtime();
    if (ms_factor) *ms_factor = (t1 - t0) * 1e3;
    if (bad) { res->status = 2; res->bad_row = bad - 1; goto done; }
    {
        memset(x, 0, sizeof(double) * n);
        memcpy(r, rhs, sizeof(double) * n);
        memcpy(rt, r, sizeof(double) * n);
        double rho = 1.0, alpha = 1.0, omega = 1.0, rho_new, beta, h, norm, norm_0, it;
        norm = norm_0 = sqrt(omp_dot(n, r, r));
        res->norm0 = norm_0;
        if (norm < norm_0 * reduction || norm < 1e-30) { res->converged = 1; goto timed; }
        for (it = 0.5; it < maxiter; it += 0.5) {
            rho_new = omp_dot(n, rt, r);
            if (fabs(rho) <= EPSILON || fabs(omega) <= EPSILON) { res->status = 3; break; }
            if (it < 1) memcpy(p, r, sizeof(double) * n);
            else {
                beta = (rho_new / rho) * (alpha / omega);
#pragma omp parallel for schedule(static)
                for (long long q = 0; q < (long long)n; ++q) { double pq = fma(-omega, v[q], p[q]); pq *= beta; p[q] = pq + r[q]; }
            }
            omp_ilu0_apply3(N, rowptr, colidx, diag, &LL, &LU, lu, relax, p, y);
            omp_spmv3(N, rowptr, colidx, vals, y, v);
            h = omp_dot(n, rt, v);
            if (fabs(h) < EPSILON) { res->status = 3; break; }
            alpha = rho_new / h;
#pragma omp parallel for schedule(static)
            for (long long q = 0; q < (long long)n; ++q) { x[q] = fma(alpha, y[q], x[q]); r[q] = fma(-alpha, v[q], r[q]); }
            norm = sqrt(omp_dot(n, r, r));
            res->half_steps++;
            if (norm < norm_0 * reduction) { res->converged = 1; break; }
            it += 0.5;
            omp_ilu0_apply3(N, rowptr, colidx, diag, &LL, &LU, lu, relax, r, y);
            omp_spmv3(N, rowptr, colidx, vals, y, t);
            omega = omp_dot(n, t, r) / omp_dot(n, t, t);
#pragma omp parallel for schedule(static)
            for (long long q = 0; q < (long long)n; ++q) { x[q] = fma(omega, y[q], x[q]); r[q] = fma(-omega, t[q], r[q]); }
            rho = rho_new;
            norm = sqrt(omp_dot(n, r, r));
            res->half_steps++;
            if (norm < norm_0 * reduction || norm < 1e-30) { res->converged = 1; break; }
        }
        if (it > maxiter) it = maxiter;
        res->iterations = (int)ceil(it);
        res->reduction = norm / norm_0;
        if (!res->converged && res->status == 0) res->status = 1;
    }
timed:
    if (ms_solve) *ms_solve = (omp_get_wtime() - t1) * 1e3;
done:
    free(lu); free(r); free(rt); free(p); free(v); free(t); free(y); free(diag);
    free_levels(&LL); free_levels(&LU);
    return res->status;
}
