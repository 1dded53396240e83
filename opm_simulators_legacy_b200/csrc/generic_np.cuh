// Block sizes other than 3: the reference instantiates Impl<np,Scalar> for np = 2..6
// (opm/autodiff/NewtonIterationBlackoilInterleaved.cpp:467-487; np = 2 for two-phase decks, 4 and more
// for the polymer / solvent extensions).  The pipelined kernels of this library are written for 3x3
// blocks; np = 2, 4, 5, 6 run on the level-scheduled kernels below -- the same arithmetic in the same order (block umv / mmv: row outer, column inner;
// dune's 2x2 inverse), one launch per dependency level, matrix values / factors / vectors stored as T.
// Parity: bit-identical SpMV, factors and apply against the oracle built with -DORACLE_BS=2.
#pragma once
#include "kernels.cuh"

namespace opmgpu {

template <int NP, class T>
__device__ __forceinline__ void npmat_mul(const T* A, const T* B, T* C)
{
#pragma unroll
    for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < NP; ++k) s = fma(A[i * NP + k], B[k * NP + j], s);
            C[i * NP + j] = s;
        }
}
// adjugate of a 4x4 block: products of entry e = 4*row+col (OPM's invertMatrix(FieldMatrix<K,4,4>&),
// the cofactor expansion published with Mesa's GLU); signs: lead, then - - + + - relative to it
__device__ __constant__ unsigned char kInv4Terms[16][6][3] = {
    {{5,10,15},{5,11,14},{9,6,15},{9,7,14},{13,6,11},{13,7,10}}, {{1,10,15},{1,11,14},{9,2,15},{9,3,14},{13,2,11},{13,3,10}},
    {{1,6,15},{1,7,14},{5,2,15},{5,3,14},{13,2,7},{13,3,6}},     {{1,6,11},{1,7,10},{5,2,11},{5,3,10},{9,2,7},{9,3,6}},
    {{4,10,15},{4,11,14},{8,6,15},{8,7,14},{12,6,11},{12,7,10}}, {{0,10,15},{0,11,14},{8,2,15},{8,3,14},{12,2,11},{12,3,10}},
    {{0,6,15},{0,7,14},{4,2,15},{4,3,14},{12,2,7},{12,3,6}},     {{0,6,11},{0,7,10},{4,2,11},{4,3,10},{8,2,7},{8,3,6}},
    {{4,9,15},{4,11,13},{8,5,15},{8,7,13},{12,5,11},{12,7,9}},   {{0,9,15},{0,11,13},{8,1,15},{8,3,13},{12,1,11},{12,3,9}},
    {{0,5,15},{0,7,13},{4,1,15},{4,3,13},{12,1,7},{12,3,5}},     {{0,5,11},{0,7,9},{4,1,11},{4,3,9},{8,1,7},{8,3,5}},
    {{4,9,14},{4,10,13},{8,5,14},{8,6,13},{12,5,10},{12,6,9}},   {{0,9,14},{0,10,13},{8,1,14},{8,2,13},{12,1,10},{12,2,9}},
    {{0,5,14},{0,6,13},{4,1,14},{4,2,13},{12,1,6},{12,2,5}},     {{0,5,10},{0,6,9},{4,1,10},{4,2,9},{8,1,6},{8,2,5}},
};

// in-place inverse, returns the determinant (np >= 5: the product of the pivots, 0 where dune throws).
// NP = 2: Dune::DenseMatrix::invert's 2x2 branch; NP = 3: OPM's MatrixBlock (mat3_invert); NP = 4: OPM's
// closed form; NP = 5, 6: Dune::DenseMatrix::invert's LU branch (dune-common 2.6: rows are swapped only
// where the pivot is below max(1e-80, |A|_inf * 1e-8), singular below max(1e-80, |A|_inf * 1e-14))
template <int NP, class T>
__device__ __forceinline__ T npmat_invert(T* M)
{
    static_assert(NP >= 2 && NP <= 6, "block sizes 2..6 are built");
    if (NP == 3) return mat3_invert(M);
    if (NP == 4) {
        T A[16], inv[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) A[e] = M[e];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const bool neg_lead = (((e >> 2) + (e & 3)) & 1) != 0;
            T acc = T(0);
#pragma unroll
            for (int t = 0; t < 6; ++t) {
                const T prod = A[kInv4Terms[e][t][0]] * A[kInv4Terms[e][t][1]] * A[kInv4Terms[e][t][2]];
                const bool plus = (t == 0 || t == 3 || t == 4) != neg_lead;
                if (t == 0) acc = plus ? prod : -prod;
                else acc = plus ? acc + prod : acc - prod;
            }
            inv[e] = acc;
        }
        const T det = A[0] * inv[0] + A[1] * inv[4] + A[2] * inv[8] + A[3] * inv[12];
        const T inv_det = T(1) / det;
#pragma unroll
        for (int e = 0; e < 16; ++e) M[e] = inv[e] * inv_det;
        return det;
    }
    if (NP > 4) {
        T A[NP][NP], X[NP][NP];
        int pivot[NP];
        T norm = T(0);
        for (int i = 0; i < NP; ++i) {
            T srow = T(0);
            for (int j = 0; j < NP; ++j) { A[i][j] = M[i * NP + j]; srow += fabs(A[i][j]); }
            if (srow > norm) norm = srow;
            pivot[i] = i;
        }
        const T abslim = sizeof(T) == 8 ? (T)1e-80 : T(0);          // (float)1e-80 == 0
        T pivthres = norm * (T)1e-8, singthres = norm * (T)1e-14;
        if (pivthres < abslim) pivthres = abslim;
        if (singthres < abslim) singthres = abslim;
        T det = T(1);
        for (int i = 0; i < NP; ++i) {
            T pivmax = fabs(A[i][i]);
            if (pivmax < pivthres) {
                int imax = i;
                for (int k = i + 1; k < NP; ++k) {
                    const T ab = fabs(A[k][i]);
                    if (ab > pivmax) { pivmax = ab; imax = k; }
                }
                if (imax != i) {
                    for (int j = 0; j < NP; ++j) { const T t = A[i][j]; A[i][j] = A[imax][j]; A[imax][j] = t; }
                    pivot[i] = imax;
                    det = -det;
                }
            }
            if (!(pivmax >= singthres)) return T(0);
            det *= A[i][i];
            for (int k = i + 1; k < NP; ++k) {
                const T factor = A[k][i] / A[i][i];
                A[k][i] = factor;
                for (int j = i + 1; j < NP; ++j) A[k][j] -= factor * A[i][j];
            }
        }
        for (int i = 0; i < NP; ++i)
            for (int j = 0; j < NP; ++j) X[i][j] = (i == j) ? T(1) : T(0);
        for (int i = 0; i < NP; ++i)
            for (int j = 0; j < i; ++j)
                for (int k = 0; k < NP; ++k) X[i][k] -= A[i][j] * X[j][k];
        for (int i = NP; i > 0;) {
            --i;
            for (int k = 0; k < NP; ++k) {
                for (int j = i + 1; j < NP; ++j) X[i][k] -= A[i][j] * X[j][k];
                X[i][k] /= A[i][i];
            }
        }
        for (int i = NP; i > 0;) {
            --i;
            if (i != pivot[i])
                for (int j = 0; j < NP; ++j) { const T t = X[j][pivot[i]]; X[j][pivot[i]] = X[j][i]; X[j][i] = t; }
        }
        for (int i = 0; i < NP; ++i)
            for (int j = 0; j < NP; ++j) M[i * NP + j] = X[i][j];
        return det;
    }
    const T det = M[0] * M[3] - M[1] * M[2];
    const T detinv = T(1) / det;
    const T temp = M[0];
    M[0] = M[3] * detinv;
    M[1] = -M[1] * detinv;
    M[2] = -M[2] * detinv;
    M[3] = temp * detinv;
    return det;
}

// y = A x: one thread per (block row, component), blocks in ascending column order
template <int NP, class T>
__global__ void __launch_bounds__(256)
np_spmv_kernel(int N, const int* __restrict__ rowptr, const int* __restrict__ colidx, const T* __restrict__ vals,
               const T* __restrict__ x, T* __restrict__ y)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)N * NP) return;
    const int row = (int)(gid / NP), r = (int)(gid - (size_t)row * NP);
    T acc = T(0);
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
        const T* a = vals + (size_t)k * NP * NP + r * NP;
        const T* xj = x + (size_t)colidx[k] * NP;
#pragma unroll
        for (int c = 0; c < NP; ++c) acc = fma(a[c], xj[c], acc);
    }
    y[gid] = acc;
}

// Dune::bilu0_decomposition, one dependency level per launch, one thread per row of the level
template <int NP, class T>
__global__ void __launch_bounds__(128)
np_factor_level_kernel(const int* __restrict__ lvl_rows, int begin, int end, const int* __restrict__ rowptr,
                       const int* __restrict__ colidx, const int* __restrict__ diag, T* lu, int* bad_row)
{
    constexpr int BB = NP * NP;
    const int q = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= end) return;
    const int i = lvl_rows[q];
    const int iend = rowptr[i + 1], idiag = diag[i];
    for (int ij = rowptr[i]; ij < idiag; ++ij) {
        const int j = colidx[ij];
        T Aij[BB], Dj[BB], L[BB];
        const int jd = diag[j];
#pragma unroll
        for (int t = 0; t < BB; ++t) { Aij[t] = lu[(size_t)ij * BB + t]; Dj[t] = lu[(size_t)jd * BB + t]; }
        npmat_mul<NP, T>(Aij, Dj, L);                                // L_ij = A_ij * inv(A_jj)
#pragma unroll
        for (int t = 0; t < BB; ++t) lu[(size_t)ij * BB + t] = L[t];
        int jk = jd + 1, ik = ij + 1;
        const int jend = rowptr[j + 1];
        while (ik < iend && jk < jend) {
            const int ci = colidx[ik], cj = colidx[jk];
            if (ci == cj) {
                T Ajk[BB], B[BB];
#pragma unroll
                for (int t = 0; t < BB; ++t) Ajk[t] = lu[(size_t)jk * BB + t];
                npmat_mul<NP, T>(L, Ajk, B);                         // A_ik -= L_ij * A_jk
#pragma unroll
                for (int t = 0; t < BB; ++t) lu[(size_t)ik * BB + t] -= B[t];
                ++ik; ++jk;
            } else if (ci < cj) ++ik;
            else ++jk;
        }
    }
    T D[BB];
#pragma unroll
    for (int t = 0; t < BB; ++t) D[t] = lu[(size_t)idiag * BB + t];
    const T det = npmat_invert<NP, T>(D);
#pragma unroll
    for (int t = 0; t < BB; ++t) lu[(size_t)idiag * BB + t] = D[t];
    if (!(det != T(0)) || isinf(det) || isnan(det)) atomicMin(bad_row, i);
}

// Opm::ParallelOverlappingILU0::apply, one dependency level per launch, one thread per row.
// LOWER: work = L^-1 d (reads d, earlier rows of work).  UPPER: work <- D^-1 (work - U work) walking the
// columns descending, out = w * work (when scale) or work.
template <int NP, class T, bool LOWER>
__global__ void __launch_bounds__(128)
np_sweep_level_kernel(const int* __restrict__ lvl_rows, int begin, int end, const int* __restrict__ rowptr,
                      const int* __restrict__ colidx, const int* __restrict__ diag, const T* __restrict__ lu,
                      const T* __restrict__ d, T* work, T* out, T w, int scale)
{
    constexpr int BB = NP * NP;
    const int q = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= end) return;
    const int i = lvl_rows[q];
    T rb[NP];
    if (LOWER) {
#pragma unroll
        for (int r = 0; r < NP; ++r) rb[r] = d[(size_t)i * NP + r];
        for (int k = rowptr[i]; k < diag[i]; ++k) {
            const T* a = lu + (size_t)k * BB;
            const T* vj = work + (size_t)colidx[k] * NP;
#pragma unroll
            for (int r = 0; r < NP; ++r)
#pragma unroll
                for (int c = 0; c < NP; ++c) rb[r] = fma(-a[r * NP + c], vj[c], rb[r]);
        }
#pragma unroll
        for (int r = 0; r < NP; ++r) work[(size_t)i * NP + r] = rb[r];
    } else {
#pragma unroll
        for (int r = 0; r < NP; ++r) rb[r] = work[(size_t)i * NP + r];
        for (int k = rowptr[i + 1] - 1; k > diag[i]; --k) {
            const T* a = lu + (size_t)k * BB;
            const T* vj = work + (size_t)colidx[k] * NP;
#pragma unroll
            for (int r = 0; r < NP; ++r)
#pragma unroll
                for (int c = 0; c < NP; ++c) rb[r] = fma(-a[r * NP + c], vj[c], rb[r]);
        }
        const T* di = lu + (size_t)diag[i] * BB;
        T yb[NP];
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            yb[r] = T(0);
#pragma unroll
            for (int c = 0; c < NP; ++c) yb[r] = fma(di[r * NP + c], rb[c], yb[r]);
        }
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            work[(size_t)i * NP + r] = yb[r];
            out[(size_t)i * NP + r] = scale ? yb[r] * w : yb[r];
        }
    }
}

// CSC front end for np x np scalar blocks (formInterleavedSystem, ...Interleaved.cpp:110-194)
__global__ void __launch_bounds__(256)
np_build_gather_map_kernel(int N, int q, int bb, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                           long long base, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                           long long* __restrict__ map, int* bad)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    for (int k = colptr[c]; k < colptr[c + 1]; ++k) {
        const int row = rowidx[k];
        int lo = rowptr[row], hi = rowptr[row + 1] - 1, pos = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int cm = colidx[mid];
            if (cm == c) { pos = mid; break; }
            if (cm < c) lo = mid + 1; else hi = mid - 1;
        }
        if (pos < 0) { atomicExch(bad, 1); continue; }
        map[(size_t)pos * bb + q] = base + k;
    }
}
struct NpScale { double s[6]; };
template <class T>
__global__ void __launch_bounds__(256)
np_interleave_gather_kernel(size_t nvals, int np, const long long* __restrict__ map, const double* __restrict__ cscval,
                            NpScale sc, T* __restrict__ vals)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nvals) return;
    const long long m = map[e];
    const int p1 = (int)((e % (size_t)(np * np)) / np);
    vals[e] = (T)(m >= 0 ? cscval[m] * sc.s[p1] : 0.0);
}
template <class T>
__global__ void __launch_bounds__(256)
np_interleave_rhs_kernel(int N, int np, const double* __restrict__ b_eqmajor, NpScale sc, T* __restrict__ b_cellmajor)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * np) return;
    const size_t i = e / np;
    const int p = (int)(e - i * np);
    b_cellmajor[e] = (T)(b_eqmajor[(size_t)p * N + i] * sc.s[p]);
}
template <class T>
__global__ void __launch_bounds__(256)
np_deinterleave_x_kernel(int N, int np, const T* __restrict__ x_cellmajor, double* __restrict__ dx_varmajor)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * np) return;
    const size_t p = e / N, i = e - p * N;
    dx_varmajor[e] = (double)x_cellmajor[i * np + p];
}

}  // namespace opmgpu
