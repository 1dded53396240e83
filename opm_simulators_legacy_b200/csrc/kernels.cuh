// Device code of the B200-native Newton-step linear solver (sm_100a).
//
// All kernels are HBM-bound (SpMV: 18 flop per 76 B in FP64); none uses tensor cores --
// 3x3 blocks are not a dense contraction (BASELINE.json north_star).  Every kernel that does
// arithmetic is a template on the scalar type T: double is the reference's Impl<3,double>, float its
// Impl<3,float> (NewtonIterationBlackoilInterleaved.cpp:478-480, selected by
// LinearisedBlackoilResidual::singlePrecision).  In the float instance the matrix values and all
// vectors are stored as float; the ILU0 factors and the sweep records keep their 8-byte containers
// (enc / dec below), the scalar block holds floats widened exactly, so the layouts the host analysis
// produces are the same for both instances.  Arithmetic order inside
// a block row follows the reference's dune-istl loops so that SpMV, the ILU0 factors and the
// ILU0 sweeps are bit-identical to the CPU oracle: every `y +-= a*x` is one fma(), nothing
// else is contracted (the file is compiled with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace opmgpu {

// 8-byte containers of the factor arrays, sweep / factorisation records, windows, push slots and
// program-ordered right-hand sides.  The double instance stores the value itself.  The float
// instance stores the float's bit pattern in the UPPER word (lower word 0): reading it back is a
// register move, not a conversion instruction on the sweeps' critical path, and the all-ones
// "empty" pattern of the self-validating slots stays a NaN no arithmetic produces.
template <class T> __device__ __forceinline__ T dec(double c);
template <> __device__ __forceinline__ double dec<double>(double c) { return c; }
template <> __device__ __forceinline__ float dec<float>(double c) { return __int_as_float(__double2hiint(c)); }
__device__ __forceinline__ double enc(double v) { return v; }
__device__ __forceinline__ double enc(float v) { return __hiloint2double(__float_as_int(v), 0); }

constexpr int kBS = 3;
constexpr int kBB = 9;
constexpr int kExtBitDev = 0x40000000;

// ---- device scalar block of the BiCGStab recurrences (lives in HBM, read by every kernel)
enum ScalarSlot {
    S_RHO_OLD = 0, S_ALPHA = 1, S_OMEGA = 2, S_H = 3, S_TR = 4, S_TT = 5,
    S_NRM2 = 6,
    S_ERRW = 7,       // != 0: the sweep watchdog word was set when the half-step ended (summed over ranks with S_NRM2)
    S_RHO_NEW = 8, S_DOT = 9, S_COUNT = 16
};

// Host mailbox (page-locked, device-mapped): the kernel that finishes a BiCGStab half-step
// writes the scalars the host's convergence logic needs, the sweep watchdog word and then a
// sequence number straight into host memory, so the host polls a cache line instead of paying
// a copy + stream synchronisation per half-step.  hS == nullptr: disabled (partitioned runs
// reduce the scalars with NCCL after the kernel).
struct HostBox {
    double* hS;                        // [S_COUNT]
    int* herr;
    unsigned long long* hseq;
    const int* derr;                   // device watchdog word of the sweeps
    unsigned long long seq;
};
__device__ __forceinline__ void hostbox_publish(const HostBox& hb)
{
    if (!hb.hS) return;
    *hb.herr = *reinterpret_cast<const volatile int*>(hb.derr);
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(hb.hseq) = hb.seq;
}
// Partitioned runs: the scalars are summed over the ranks (ncclAllReduce) after the kernel that
// produced them, so the mailbox is written by this one-warp kernel that follows the reduction.
// S[S_ERRW] then holds the number of ranks whose sweep watchdog fired: every rank sees the same
// verdict and leaves the solve with the same status.
__global__ void publish_scalars_kernel(const double* __restrict__ S, HostBox hb)
{
    if (threadIdx.x < S_COUNT) hb.hS[threadIdx.x] = S[threadIdx.x];
    __syncwarp();
    if (threadIdx.x == 0) {
        *hb.herr = S[S_ERRW] != 0.0 ? 9 : 0;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(hb.hseq) = hb.seq;
    }
}

// ------------------------------------------------------------------------------------------
// Deterministic grid reduction: fixed-shape shuffle tree per block, per-block partials in
// HBM, the last block to arrive (ticket) folds the partials in a fixed order.  No floating
// point atomics, so results are reproducible run to run for a given launch shape.
// ------------------------------------------------------------------------------------------
struct ReduceWs {
    double*   partials;      // [kMaxRed][kMaxBlocks]
    unsigned* ticket;        // wraps to 0 by atomicInc
};
constexpr int kMaxRedBlocks = 2048;

template <class T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int NRED, class T>
__device__ __forceinline__ void block_sum(T (&v)[NRED], T* smem /*[NRED*32]*/)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int r = 0; r < NRED; ++r) {
        v[r] = warp_sum(v[r]);
        if (lane == 0) smem[r * 32 + wid] = v[r];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int r = 0; r < NRED; ++r) {
            T t = (lane < nw) ? smem[r * 32 + lane] : T(0);
            v[r] = warp_sum(t);
        }
    }
    __syncthreads();
}

// Finaliser hook: called by thread 0 of the last block with the NRED totals.  (The per-block
// partials travel through 8-byte slots whatever T is.)
template <int NRED, class T, class Fin>
__device__ __forceinline__ void grid_reduce(T (&v)[NRED], ReduceWs ws, Fin fin)
{
    __shared__ T red_smem[NRED * 32];
    __shared__ bool is_last;
    block_sum<NRED, T>(v, red_smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NRED; ++r) ws.partials[r * kMaxRedBlocks + blockIdx.x] = (double)v[r];
        __threadfence();
        const unsigned t = atomicInc(ws.ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        T acc[NRED];
#pragma unroll
        for (int r = 0; r < NRED; ++r) {
            acc[r] = T(0);
            for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
                acc[r] += (T)__ldcg(&ws.partials[r * kMaxRedBlocks + b]);
        }
        block_sum<NRED, T>(acc, red_smem);
        if (threadIdx.x == 0) fin(acc);
    }
}

// ------------------------------------------------------------------------------------------
// K2  BCRS 3x3 SpMV  y = A x   (Dune::MatrixAdapter::apply -> BCRSMatrix::mv -> umv;
// call site opm/autodiff/ISTLSolver.hpp:303).  One thread per (block row, component),
// blocks visited in ascending column order.  MODE 0: plain.  MODE 1: also S[S_H] = w1.y.
// MODE 2: also S[S_TR] = y.w1, S[S_TT] = y.y.  (fused dot epilogues of BiCGStab)
// ------------------------------------------------------------------------------------------
template <int MODE, class T>
__global__ void __launch_bounds__(256)
spmv3_kernel(int N, const int* __restrict__ rowptr, const int* __restrict__ colidx,
             const T* __restrict__ vals, const T* __restrict__ x,
             T* __restrict__ y, const T* __restrict__ w1, double* S, ReduceWs ws)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int row = (int)(gid / 3), c = (int)(gid - 3LL * row);
    T acc = T(0);
    if (row < N) {
        const int kb = __ldg(rowptr + row), ke = __ldg(rowptr + row + 1);
        for (int k = kb; k < ke; ++k) {
            const T* a = vals + (size_t)k * kBB + c * kBS;
            const T* xj = x + (size_t)__ldg(colidx + k) * kBS;
            acc = fma(__ldcs(a + 0), xj[0], acc);
            acc = fma(__ldcs(a + 1), xj[1], acc);
            acc = fma(__ldcs(a + 2), xj[2], acc);
        }
        y[gid] = acc;
    }
    if (MODE == 1) {
        T v[1] = { row < N ? w1[gid] * acc : T(0) };
        grid_reduce<1, T>(v, ws, [=](T (&t)[1]) { S[S_H] = t[0]; });
    } else if (MODE == 2) {
        T v[2] = { row < N ? acc * w1[gid] : T(0), row < N ? acc * acc : T(0) };
        grid_reduce<2, T>(v, ws, [=](T (&t)[2]) { S[S_TR] = t[0]; S[S_TT] = t[1]; });
    }
}

// ------------------------------------------------------------------------------------------
// K5  fused BiCGStab vector kernels (Dune::BiCGSTABSolver::apply; elementwise order as there:
// p.axpy(-omega,v); p*=beta; p+=r   /   x.axpy(alpha,y); r.axpy(-alpha,v)).
// ------------------------------------------------------------------------------------------
// S[S_NRM2] = S[S_RHO_NEW] = r.r ; recurrences reset (rho = alpha = omega = 1)
template <class T>
__global__ void __launch_bounds__(256)
bicg_init_kernel(size_t n, const T* __restrict__ r, double* S, ReduceWs ws, HostBox hb)
{
    T v[1] = {T(0)};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v[0] = fma(r[i], r[i], v[0]);
    grid_reduce<1, T>(v, ws, [=](T (&t)[1]) {
        S[S_NRM2] = t[0]; S[S_RHO_NEW] = t[0];
        S[S_RHO_OLD] = 1.0; S[S_ALPHA] = 1.0; S[S_OMEGA] = 1.0;
        S[S_ERRW] = *reinterpret_cast<const volatile int*>(hb.derr) != 0 ? 1.0 : 0.0;
        if (hb.hS) { hb.hS[S_NRM2] = t[0]; hostbox_publish(hb); }
    });
}

// p = r + beta (p - omega v),  beta = (rho_new/rho)(alpha/omega)
// lpos != nullptr: p is also written in the lower sweep's program order (row r -> position
// lpos[r]), which is what the next preconditioner application streams as its right-hand side
template <class T>
__global__ void __launch_bounds__(256)
bicg_update_p_kernel(size_t n, T* __restrict__ p, const T* __restrict__ r,
                     const T* __restrict__ v, const double* __restrict__ S,
                     const int* __restrict__ lpos, double* __restrict__ pperm)
{
    const T omega = (T)S[S_OMEGA];
    const T beta = ((T)S[S_RHO_NEW] / (T)S[S_RHO_OLD]) * ((T)S[S_ALPHA] / omega);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T pq = fma(-omega, v[i], p[i]);
        pq *= beta;
        const T pn = pq + r[i];
        p[i] = pn;
        if (lpos) { const size_t row = i / 3; pperm[(size_t)lpos[row] * 3 + (i - row * 3)] = enc(pn); }
    }
}

// alpha = rho_new / h ; x += alpha y ; r -= alpha v ; S[S_NRM2] = r.r
template <class T>
__global__ void __launch_bounds__(256)
bicg_update1_kernel(size_t n, T* __restrict__ x, T* __restrict__ r,
                    const T* __restrict__ y, const T* __restrict__ v, double* S, ReduceWs ws, HostBox hb,
                    const int* __restrict__ lpos, double* __restrict__ rperm)
{
    const T hdot = (T)S[S_H];
    const T alpha = (T)S[S_RHO_NEW] / hdot;
    T s[1] = {T(0)};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, y[i], x[i]);
        const T ri = fma(-alpha, v[i], r[i]);
        r[i] = ri;
        if (lpos) { const size_t row = i / 3; rperm[(size_t)lpos[row] * 3 + (i - row * 3)] = enc(ri); }     // see bicg_update_p_kernel
        s[0] = fma(ri, ri, s[0]);
    }
    grid_reduce<1, T>(s, ws, [=](T (&t)[1]) {
        S[S_NRM2] = t[0]; S[S_ALPHA] = alpha;
        S[S_ERRW] = *reinterpret_cast<const volatile int*>(hb.derr) != 0 ? 1.0 : 0.0;
        if (hb.hS) { hb.hS[S_NRM2] = t[0]; hb.hS[S_H] = hdot; hostbox_publish(hb); }
    });
}

// omega = (t.r)/(t.t) ; x += omega y ; r -= omega t ; S[S_NRM2] = r.r ; rho <- rho_new ; rho_new = rt.r
template <class T>
__global__ void __launch_bounds__(256)
bicg_update2_kernel(size_t n, T* __restrict__ x, T* __restrict__ r,
                    const T* __restrict__ y, const T* __restrict__ t,
                    const T* __restrict__ rt, double* S, ReduceWs ws, HostBox hb)
{
    const T omega = (T)S[S_TR] / (T)S[S_TT];
    T s[2] = {T(0), T(0)};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        x[i] = fma(omega, y[i], x[i]);
        const T ri = fma(-omega, t[i], r[i]);
        r[i] = ri;
        s[0] = fma(ri, ri, s[0]);
        s[1] = fma(rt[i], ri, s[1]);
    }
    grid_reduce<2, T>(s, ws, [=](T (&u)[2]) {
        const double rho_old = S[S_RHO_NEW];
        S[S_OMEGA] = omega;
        S[S_RHO_OLD] = rho_old;
        S[S_NRM2] = u[0];
        S[S_RHO_NEW] = u[1];
        S[S_ERRW] = *reinterpret_cast<const volatile int*>(hb.derr) != 0 ? 1.0 : 0.0;
        if (hb.hS) { hb.hS[S_OMEGA] = omega; hb.hS[S_RHO_OLD] = rho_old; hb.hS[S_NRM2] = u[0]; hostbox_publish(hb); }
    });
}

template <class T>
__global__ void __launch_bounds__(256)
dot_kernel(size_t n, const T* __restrict__ a, const T* __restrict__ b, double* S, ReduceWs ws)
{
    T v[1] = {T(0)};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v[0] = fma(a[i], b[i], v[0]);
    grid_reduce<1, T>(v, ws, [=](T (&t)[1]) { S[S_DOT] = t[0]; });
}

// ------------------------------------------------------------------------------------------
// 3x3 helpers mirroring dune's DenseMatrix::{left,right}multiply and OPM's MatrixBlock inverse
// ------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void mat3_mul(const T* A, const T* B, T* C)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < 3; ++k) s = fma(A[i * 3 + k], B[k * 3 + j], s);
            C[i * 3 + j] = s;
        }
}

// adjugate / determinant, same operation order as Opm::MatrixBlock<T,3,3>::invert
template <class T>
__device__ __forceinline__ T mat3_invert(T* M)
{
    T A[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) A[q] = M[q];
    const T t4 = A[0] * A[4], t6 = A[0] * A[5], t8 = A[1] * A[3];
    const T t10 = A[2] * A[3], t12 = A[1] * A[6], t14 = A[2] * A[6];
    const T det = (t4 * A[8] - t6 * A[7] - t8 * A[8] + t10 * A[7] + t12 * A[5] - t14 * A[4]);
    const T t17 = T(1) / det;
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
}

// ------------------------------------------------------------------------------------------
// K3  block ILU0 factorisation, natural order, one dependency level per launch
// (Opm::ParallelOverlappingILU0 ctor -> Dune::bilu0_decomposition; ISTLSolver.hpp:201-211).
// One thread per row of the level; rows of a level are independent by construction.
// ------------------------------------------------------------------------------------------
// (lu: 8-byte containers; AT = arithmetic type, every stored value is an AT widened exactly)
template <class AT>
__global__ void __launch_bounds__(128)
ilu0_factor_level_kernel(const int* __restrict__ lvl_rows, int begin, int end,
                         const int* __restrict__ rowptr, const int* __restrict__ colidx,
                         const int* __restrict__ diag, double* lu, int* bad_row)
{
    const int q = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= end) return;
    const int i = lvl_rows[q];
    const int iend = rowptr[i + 1], idiag = diag[i];
    for (int ij = rowptr[i]; ij < idiag; ++ij) {
        const int j = colidx[ij];
        AT Aij[9], Dj[9], L[9];
        const int jd = diag[j];
#pragma unroll
        for (int t = 0; t < 9; ++t) { Aij[t] = dec<AT>(lu[(size_t)ij * 9 + t]); Dj[t] = dec<AT>(lu[(size_t)jd * 9 + t]); }
        mat3_mul(Aij, Dj, L);                                 // L_ij = A_ij * inv(A_jj)
#pragma unroll
        for (int t = 0; t < 9; ++t) lu[(size_t)ij * 9 + t] = enc(L[t]);
        int jk = jd + 1, ik = ij + 1;
        const int jend = rowptr[j + 1];
        while (ik < iend && jk < jend) {
            const int ci = colidx[ik], cj = colidx[jk];
            if (ci == cj) {
                AT Ajk[9], B[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) Ajk[t] = dec<AT>(lu[(size_t)jk * 9 + t]);
                mat3_mul(L, Ajk, B);                          // B = L_ij * A_jk
#pragma unroll
                for (int t = 0; t < 9; ++t) lu[(size_t)ik * 9 + t] = enc(dec<AT>(lu[(size_t)ik * 9 + t]) - B[t]);
                ++ik; ++jk;
            } else if (ci < cj) ++ik;
            else ++jk;
        }
    }
    AT D[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) D[t] = dec<AT>(lu[(size_t)idiag * 9 + t]);
    const AT det = mat3_invert(D);
#pragma unroll
    for (int t = 0; t < 9; ++t) lu[(size_t)idiag * 9 + t] = enc(D[t]);
    if (!(det != AT(0)) || isinf(det) || isnan(det)) atomicMin(bad_row, i);
}

// ------------------------------------------------------------------------------------------
// K3 (fast path)  block ILU0 factorisation as a pipelined wavefront of persistent CTAs: the
// rows of a CTA's tile are processed level by level (one thread per row, __syncthreads
// between levels); a row of another tile is awaited through its per-row flag.  The update
// lists (which A_ik a given L_ij touches) are precomputed on the host, so every address is
// known up front.  Arithmetic order is Dune::bilu0_decomposition's (bit parity).
// ------------------------------------------------------------------------------------------
struct FactorDev {
    const int* cta_step_ptr;
    const int* step_row_ptr;
    const int4* frow;         // {row, diag slot, first entry, count | kFactorSimpleDev}
    const int4* fent;         // two per entry: {ij, jd, dep | ext, npairs} {jk0, ik0, pair_ptr, -}
    const int* pair_jk;
    const int* pair_ik;
    const unsigned char* needs_flag;   // some consumer in another CTA waits on the row's flag
    const int* fpush_ptr;              // slots the row's inverted pivot is pushed to
    const int* fpush_slot;
    double* fslots;                    // [n][9], all-ones = empty
};
constexpr int kFactorSimpleDev = 1 << 30;

// self-validating 9-double slot written by another CTA (each double is one atomic 8-byte store)
template <class AT>
__device__ __forceinline__ bool factor_poll_slot(const double* slot, AT (&d)[9], int* err)
{
    const long long* s = reinterpret_cast<const long long*>(slot);
    unsigned spins = 0;
    for (;;) {
        long long v[9];
        bool ok = true;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v[t]) : "l"(s + t) : "memory");
            ok = ok && v[t] != -1;
        }
        if (ok) {
#pragma unroll
            for (int t = 0; t < 9; ++t) d[t] = dec<AT>(__longlong_as_double(v[t]));
            return true;
        }
        if (++spins > (1u << 22)) { atomicExch(err, 6); return false; }
        if ((spins & 255u) == 0 && *(volatile int*)err) return false;
    }
}

__device__ __forceinline__ int ld_acquire_gpu_f(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <class AT>
__device__ __forceinline__ void load9(AT (&d)[9], const double* p, bool bypass_l1)
{
#pragma unroll
    for (int t = 0; t < 9; ++t) d[t] = dec<AT>(bypass_l1 ? __ldcg(p + t) : p[t]);
}
__device__ __forceinline__ bool factor_wait_row(const int* flags, int j, int epoch, int* err)
{
    unsigned spins = 0;
    while (ld_acquire_gpu_f(flags + j) != epoch) {
        if (++spins > (1u << 22)) { atomicExch(err, 5); return false; }
        if ((spins & 1023u) == 0 && *(volatile int*)err) return false;
    }
    return true;
}

// row c of C = A * B where the thread holds row c of A (dune's accumulation order per entry)
template <class AT>
__device__ __forceinline__ void mat3_row_mul(const AT (&Arow)[3], const AT (&B)[9], AT (&Crow)[3])
{
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        AT s = AT(0);
#pragma unroll
        for (int k = 0; k < 3; ++k) s = fma(Arow[k], B[k * 3 + j], s);
        Crow[j] = s;
    }
}

// Three threads per block row (thread c owns row c of every 3x3 block of the row; ten rows per
// warp, lanes 30/31 idle), one level of the CTA's tile per iteration.
constexpr int kFactorThreads = 256;
constexpr int kFactorRowsPerPass = (kFactorThreads / 32) * 10;

template <class AT>
__global__ void __launch_bounds__(kFactorThreads)
ilu0_factor_tile_kernel(FactorDev pg, double* lu, int* flags, int epoch, int* bad_row, int* err)
{
    const int s_begin = pg.cta_step_ptr[blockIdx.x], s_end = pg.cta_step_ptr[blockIdx.x + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = lane / 3, c = lane - rl * 3;
    const bool lane_on = lane < 30;
    const int base = rl * 3;
    for (int s = s_begin; s < s_end; ++s) {
        const int q0 = pg.step_row_ptr[s], q1 = pg.step_row_ptr[s + 1];
        // pull the coming levels towards L2 while this one computes: row records two levels
        // ahead, entry records and the row's own blocks (lower part + diagonal are contiguous in
        // BCRS) one level ahead
        if (s + 2 < s_end) {
            const int qn = pg.step_row_ptr[s + 2] + threadIdx.x;
            if (qn < pg.step_row_ptr[s + 3]) asm volatile("prefetch.global.L2 [%0];" ::"l"(pg.frow + qn));
        }
        if (s + 1 < s_end) {
            const int qn = q1 + threadIdx.x;
            if (qn < pg.step_row_ptr[s + 2]) {
                const int4 fn = pg.frow[qn];
                const int nn = fn.w & ~kFactorSimpleDev;
                const char* e = reinterpret_cast<const char*>(pg.fent + 2 * fn.z);
                for (int o = 0; o < nn * 32; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(e + o));
                const char* blk = reinterpret_cast<const char*>(lu + (size_t)(fn.y - nn) * 9);
                for (int o = 0; o < (nn + 1) * 72 + 127; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + o));
            }
        }
        for (int qb = q0; qb < q1; qb += kFactorRowsPerPass) {              // warp-uniform trip count
            const int q = qb + warp * 10 + rl;
            const bool on = lane_on && q < q1;
            int4 fr = make_int4(0, 0, 0, 0);
            if (on) fr = pg.frow[q];
            const int i = fr.x, idiag = fr.y, e0 = fr.z, n = fr.w & ~kFactorSimpleDev;
            const bool simple = on && (fr.w & kFactorSimpleDev);
            AT Drow[3] = {AT(0), AT(0), AT(0)};
            int rearm[3] = {-1, -1, -1};
            if (simple) {
                // stencil rows: every address is known now, issue all loads before any arithmetic
                int4 ea[3];
                int jk0[3];
                AT Arow[3][3], Dj[3][9], Ajk[3][9];
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k < n) {
                        ea[k] = pg.fent[2 * (e0 + k)];
                        const int4 eb = pg.fent[2 * (e0 + k) + 1];
                        jk0[k] = eb.x; rearm[k] = eb.w;
                    }
#pragma unroll
                for (int t = 0; t < 3; ++t) Drow[t] = dec<AT>(lu[(size_t)idiag * 9 + c * 3 + t]);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k < n) {
#pragma unroll
                        for (int t = 0; t < 3; ++t) Arow[k][t] = dec<AT>(lu[(size_t)ea[k].x * 9 + c * 3 + t]);
                    }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k < n) {
                        const bool ext = (ea[k].z & kExtBitDev) != 0;
                        // the coupling block A_ji of a simple row j is never modified: no wait
                        if (ea[k].w && rearm[k] >= 0) load9(Ajk[k], lu + (size_t)jk0[k] * 9, true);
                        if (ext && rearm[k] < 0) factor_wait_row(flags, ea[k].z & ~kExtBitDev, epoch, err);
                        if (!(ext && rearm[k] >= 0)) {
                            load9(Dj[k], lu + (size_t)ea[k].y * 9, ext);
                            if (ea[k].w) load9(Ajk[k], lu + (size_t)jk0[k] * 9, ext);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k)                                  // pushed pivots last: they arrive latest
                    if (k < n && rearm[k] >= 0) factor_poll_slot(pg.fslots + (size_t)rearm[k] * 9, Dj[k], err);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k < n) {
                        AT Lrow[3], Brow[3];
                        mat3_row_mul(Arow[k], Dj[k], Lrow);                 // L_ij = A_ij * inv(A_jj)
#pragma unroll
                        for (int t = 0; t < 3; ++t) lu[(size_t)ea[k].x * 9 + c * 3 + t] = enc(Lrow[t]);
                        if (ea[k].w) {
                            mat3_row_mul(Lrow, Ajk[k], Brow);               // A_ii -= L_ij * A_ji
#pragma unroll
                            for (int t = 0; t < 3; ++t) Drow[t] -= Brow[t];
                        }
                    }
                }
            } else if (on && c == 0) {
                // general rows (more than three lower blocks, fill outside the diagonal): one thread
                AT D[9];
                load9(D, lu + (size_t)idiag * 9, false);
                for (int b = e0; b < e0 + n; ++b) {
                    const int4 ea = pg.fent[2 * b], eb = pg.fent[2 * b + 1];
                    const int ij = ea.x;
                    const bool ext = (ea.z & kExtBitDev) != 0;
                    if (ext) factor_wait_row(flags, ea.z & ~kExtBitDev, epoch, err);
                    AT Aij[9], Dj[9], L[9];
                    load9(Aij, lu + (size_t)ij * 9, false);
                    load9(Dj, lu + (size_t)ea.y * 9, ext);
                    mat3_mul(Aij, Dj, L);
#pragma unroll
                    for (int t = 0; t < 9; ++t) lu[(size_t)ij * 9 + t] = enc(L[t]);
                    for (int pp = eb.z; pp < eb.z + ea.w; ++pp) {
                        const int jk = pg.pair_jk[pp], ik = pg.pair_ik[pp];
                        AT Ajk[9], B[9];
                        load9(Ajk, lu + (size_t)jk * 9, ext);
                        mat3_mul(L, Ajk, B);
                        if (ik == idiag) {
#pragma unroll
                            for (int t = 0; t < 9; ++t) D[t] -= B[t];
                        } else {
#pragma unroll
                            for (int t = 0; t < 9; ++t) lu[(size_t)ik * 9 + t] = enc(dec<AT>(lu[(size_t)ik * 9 + t]) - B[t]);
                        }
                    }
                }
                const AT det = mat3_invert(D);
#pragma unroll
                for (int t = 0; t < 9; ++t) lu[(size_t)idiag * 9 + t] = enc(D[t]);
                if (!(det != AT(0)) || isinf(det) || isnan(det)) atomicMin(bad_row, i);
            }
            // simple rows: gather the whole pivot block inside the warp, invert, store own row
            AT D[9];
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int t = 0; t < 3; ++t) D[m * 3 + t] = __shfl_sync(0xffffffffu, Drow[t], base + m);
            if (simple) {
                const AT det = mat3_invert(D);
                const double o0 = enc(c == 0 ? D[0] : (c == 1 ? D[3] : D[6]));
                const double o1 = enc(c == 0 ? D[1] : (c == 1 ? D[4] : D[7]));
                const double o2 = enc(c == 0 ? D[2] : (c == 1 ? D[5] : D[8]));
                double* dst = lu + (size_t)idiag * 9 + c * 3;
                dst[0] = o0; dst[1] = o1; dst[2] = o2;
                for (int t = pg.fpush_ptr[q]; t < pg.fpush_ptr[q + 1]; ++t) {      // push the pivot to other CTAs
                    double* sl = pg.fslots + (size_t)pg.fpush_slot[t] * 9 + c * 3;
                    __stcg(sl, o0); __stcg(sl + 1, o1); __stcg(sl + 2, o2);
                }
                if (c == 0 && (!(det != AT(0)) || isinf(det) || isnan(det))) atomicMin(bad_row, i);
            }
            __syncwarp();
            if (simple && c == 0) {                                          // re-arm the slots this row consumed
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (rearm[k] >= 0) {
                        long long* sl = reinterpret_cast<long long*>(pg.fslots + (size_t)rearm[k] * 9);
#pragma unroll
                        for (int t = 0; t < 9; ++t) __stcg(sl + t, -1LL);
                    }
            }
            if (on && c == 0 && pg.needs_flag[q]) {
                __threadfence();
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + i), "r"(epoch) : "memory");
            }
        }
        __syncthreads();
    }
}

// factors (BCRS order) -> the sweep programs' streaming layout
__global__ void __launch_bounds__(256)
repack_blocks_kernel(size_t nblk, const int* __restrict__ psrc, const double* __restrict__ lu,
                     double* __restrict__ pval)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one double each
    if (e >= nblk * 9) return;
    const size_t b = e / 9;
    pval[e] = lu[(size_t)psrc[b] * 9 + (e - b * 9)];
}
// factors (BCRS order) -> step records of the pipelined sweeps.  Element [c][e] of source block
// b goes to double index dst8[b] + c*stride[b] + e.
__global__ void __launch_bounds__(256)
repack_pipe_kernel(size_t n, const int* __restrict__ src, const unsigned* __restrict__ dst8,
                   const int* __restrict__ stride, const double* __restrict__ lu, double* __restrict__ bufd)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 9) return;
    const size_t b = t / 9;
    const int q = (int)(t - b * 9), c = q / 3, e = q - c * 3;
    const int st = stride[b];
    const size_t dst = (size_t)dst8[b] + (size_t)c * st + e;
    bufd[dst] = lu[(size_t)src[b] * 9 + q];
}
__global__ void __launch_bounds__(256)
repack_dinv_kernel(int N, const int* __restrict__ prow, const int* __restrict__ diag,
                   const double* __restrict__ lu, double* __restrict__ pdinv)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * 9) return;
    const size_t q = e / 9;
    pdinv[e] = lu[(size_t)diag[prow[q]] * 9 + (e - q * 9)];
}

// ------------------------------------------------------------------------------------------
// K4  ILU0 apply  v = w U^-1 L^-1 d   (Opm::ParallelOverlappingILU0::apply).
// Persistent cooperative grid: CTA c runs its program (analysis.hpp) step by step; rows of
// a step are independent; dependencies owned by the same CTA are ordered by __syncthreads,
// dependencies owned by another CTA by a per-row flag carrying the sweep's epoch
// (point-to-point, no grid-wide barrier -> the CTAs form a pipelined wavefront).
// ------------------------------------------------------------------------------------------
struct SweepDev {
    const int* cta_step_ptr;
    const int* step_row_ptr;
    const int* prow;
    const int* pblk_ptr;
    const int* pcol;
    const unsigned char* publish;
    const double* pval;      // [nblk*9] program order
    const double* pdinv;     // [N*9]    program-row order (upper only)
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr unsigned kSpinLimit = 1u << 22;

template <bool LOWER, class T>
__global__ void __launch_bounds__(256)
ilu0_sweep_kernel(SweepDev pg, const T* __restrict__ rhs, T* work, T* out,
                  double w_, int scale, int* flags, int epoch, int* err)
{
    const T w = (T)w_;
    // LOWER: rhs = d, work = yL (written), out unused.
    // UPPER: rhs = yL, work = vU (unscaled, written, read by dependants), out = w * vU.
    const int s_begin = pg.cta_step_ptr[blockIdx.x], s_end = pg.cta_step_ptr[blockIdx.x + 1];
    for (int s = s_begin; s < s_end; ++s) {
        const int q0 = pg.step_row_ptr[s], q1 = pg.step_row_ptr[s + 1];
        for (int q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
            const int row = pg.prow[q];
            T r0 = rhs[(size_t)row * 3], r1 = rhs[(size_t)row * 3 + 1], r2 = rhs[(size_t)row * 3 + 2];
            const int b0 = pg.pblk_ptr[q], b1 = pg.pblk_ptr[q + 1];
            for (int b = b0; b < b1; ++b) {
                const int c = pg.pcol[b];
                const int j = c & ~kExtBitDev;
                T y0, y1, y2;
                if (c & kExtBitDev) {
                    unsigned spins = 0;
                    while (ld_acquire_gpu(flags + j) != epoch) {
                        if (++spins > kSpinLimit) { atomicExch(err, 1); break; }
                        if ((spins & 1023u) == 0 && *(volatile int*)err) break;
                    }
                    y0 = __ldcg(work + (size_t)j * 3); y1 = __ldcg(work + (size_t)j * 3 + 1); y2 = __ldcg(work + (size_t)j * 3 + 2);
                } else {
                    y0 = work[(size_t)j * 3]; y1 = work[(size_t)j * 3 + 1]; y2 = work[(size_t)j * 3 + 2];
                }
                const double* a = pg.pval + (size_t)b * 9;
                r0 = fma(-dec<T>(a[0]), y0, r0); r0 = fma(-dec<T>(a[1]), y1, r0); r0 = fma(-dec<T>(a[2]), y2, r0);
                r1 = fma(-dec<T>(a[3]), y0, r1); r1 = fma(-dec<T>(a[4]), y1, r1); r1 = fma(-dec<T>(a[5]), y2, r1);
                r2 = fma(-dec<T>(a[6]), y0, r2); r2 = fma(-dec<T>(a[7]), y1, r2); r2 = fma(-dec<T>(a[8]), y2, r2);
            }
            if (!LOWER) {
                const double* di = pg.pdinv + (size_t)q * 9;
                T v0 = T(0), v1 = T(0), v2 = T(0);
                v0 = fma(dec<T>(di[0]), r0, v0); v0 = fma(dec<T>(di[1]), r1, v0); v0 = fma(dec<T>(di[2]), r2, v0);
                v1 = fma(dec<T>(di[3]), r0, v1); v1 = fma(dec<T>(di[4]), r1, v1); v1 = fma(dec<T>(di[5]), r2, v1);
                v2 = fma(dec<T>(di[6]), r0, v2); v2 = fma(dec<T>(di[7]), r1, v2); v2 = fma(dec<T>(di[8]), r2, v2);
                r0 = v0; r1 = v1; r2 = v2;
                if (scale) { v0 *= w; v1 *= w; v2 *= w; }
                out[(size_t)row * 3] = v0; out[(size_t)row * 3 + 1] = v1; out[(size_t)row * 3 + 2] = v2;
            }
            work[(size_t)row * 3] = r0; work[(size_t)row * 3 + 1] = r1; work[(size_t)row * 3 + 2] = r2;
            if (pg.publish[q]) {
                __threadfence();
                st_release_gpu(flags + row, epoch);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K1  interleave: nine CSC scalar blocks -> row-major 3x3 BCRS blocks
// (formInterleavedSystem value scatter, ...Interleaved.cpp:178-193, with the matbalscale
// row scaling of :234-236 folded in).  The gather map is built once per pattern on the device.
// ------------------------------------------------------------------------------------------
// one thread per (scalar block q, column c): locate every CSC entry in the BCRS pattern
__global__ void __launch_bounds__(256)
build_gather_map_kernel(int N, int q, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                        long long base, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                        long long* __restrict__ map9, int* bad)
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
        map9[(size_t)pos * 9 + q] = base + k;
    }
}

// one thread per output double: vals[slot][p1][p2] = cscval[map] * scale[p1]  (0 if absent)
// (T = float: the scaled value is rounded once, as the assignment to the reference's float matrix does)
template <class T>
__global__ void __launch_bounds__(256)
interleave_gather_kernel(size_t nvals, const long long* __restrict__ map9,
                         const double* __restrict__ cscval, double s0, double s1, double s2,
                         T* __restrict__ vals)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nvals) return;
    const long long m = map9[e];
    const int p1 = (int)((e % 9) / 3);
    const double sc = p1 == 0 ? s0 : (p1 == 1 ? s1 : s2);
    vals[e] = (T)(m >= 0 ? cscval[m] * sc : 0.0);
}

// a8  rhs interleave (+ scaling) and solution de-interleave (...Interleaved.cpp:263-269, 279-283)
template <class T>
__global__ void __launch_bounds__(256)
interleave_rhs_kernel(int N, const double* __restrict__ b_eqmajor, double s0, double s1, double s2,
                      T* __restrict__ b_cellmajor)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * 3) return;
    const size_t i = e / 3;
    const int p = (int)(e - i * 3);
    const double sc = p == 0 ? s0 : (p == 1 ? s1 : s2);
    b_cellmajor[e] = (T)(b_eqmajor[(size_t)p * N + i] * sc);
}
template <class T>
__global__ void __launch_bounds__(256)
deinterleave_x_kernel(int N, const T* __restrict__ x_cellmajor, double* __restrict__ dx_varmajor)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * 3) return;
    const size_t p = e / N, i = e - p * N;
    dx_varmajor[e] = (double)x_cellmajor[i * 3 + p];
}

// values -> containers and back (opmgpu_ilu0_get_factors hands the caller plain doubles)
template <class T>
__global__ void __launch_bounds__(256)
encode_kernel(size_t n, const T* __restrict__ in, double* __restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = enc(in[i]);
}
template <class T>
__global__ void __launch_bounds__(256)
decode_kernel(size_t n, const double* __restrict__ in, double* __restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)dec<T>(in[i]);
}

// plain conversions between the caller's doubles and the instance's scalar type
template <class TI, class TO>
__global__ void __launch_bounds__(256)
convert_kernel(size_t n, const TI* __restrict__ in, TO* __restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (TO)in[i];
}

}  // namespace opmgpu
