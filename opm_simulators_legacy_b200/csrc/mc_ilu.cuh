// Multicolour block ILU0 (flagged variant, mcorder.hpp): the reference's ILU0 of P A P^T.
//
// Storage: one "unified" factor array of 3x3 blocks of T, [ L | Dinv | U ], rows in permuted order
// (sorted by colour).  L holds the strictly lower blocks of a permuted row in ascending permuted
// column order, U the strictly upper ones in DESCENDING column order -- the order
// Opm::ParallelOverlappingILU0::apply visits them -- so both sweeps stream their operand linearly.
// Rows of one colour have no blocks among themselves: a colour is one fully parallel, HBM-bound pass
// over its block rows, i.e. an SpMV with an epilogue.  The sweep kernel below therefore has the
// structure of spmv3_tma_kernel (persistent CTAs, TMA-fed shared-memory ring, three consumer groups).
//
// Arithmetic and its order are the oracle's (bit parity with the oracle run on the permuted system):
//   factorisation  Dune::bilu0_decomposition (oracle.c oracle_ilu0_factor3)
//   lower          rb = d_i;  rb -= L_ij y_j  (ascending j)
//   upper          rb = y_i;  rb -= U_ij x_j  (descending j);  x_i = Dinv_i rb;  out = w x_i
// Algorithmic bytes per apply: 76 per off-diagonal block + 72 (Dinv) + 5*24 (vectors) per row.
#pragma once
#include "kernels.cuh"
#include "spmv_tma.cuh"

namespace opmgpu {

// A (natural BCRS) -> unified array: one thread per scalar
template <class T>
__global__ void __launch_bounds__(256)
mc_gather_values_kernel(size_t nscal, const int* __restrict__ psrc, const int* __restrict__ ppos,
                        const T* __restrict__ vals, T* __restrict__ uni)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nscal) return;
    const size_t kp = e / 9;
    const int t = (int)(e - kp * 9);
    uni[(size_t)ppos[kp] * 9 + t] = vals[(size_t)psrc[kp] * 9 + t];
}
// unified array -> BCRS factor array in the caller's (natural) slot order, as doubles
template <class T>
__global__ void __launch_bounds__(256)
mc_scatter_factors_kernel(size_t nscal, const int* __restrict__ psrc, const int* __restrict__ ppos,
                          const T* __restrict__ uni, double* __restrict__ lu)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nscal) return;
    const size_t kp = e / 9;
    const int t = (int)(e - kp * 9);
    lu[(size_t)psrc[kp] * 9 + t] = (double)uni[(size_t)ppos[kp] * 9 + t];
}

// Dune::bilu0_decomposition on the permuted pattern, one dependency level per launch, one thread
// per row; entry b of the permuted pattern lives at block ppos[b] of the unified array
template <class T>
__global__ void __launch_bounds__(128)
mc_factor_level_kernel(const int* __restrict__ lvl_rows, int begin, int end, const int* __restrict__ prowptr,
                       const int* __restrict__ pcol, const int* __restrict__ pdiag, const int* __restrict__ ppos,
                       T* uni, int* bad_row)
{
    const int s = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= end) return;
    const int i = lvl_rows[s];
    const int iend = prowptr[i + 1], idiag = pdiag[i];
    for (int ij = prowptr[i]; ij < idiag; ++ij) {
        const int j = pcol[ij];
        T Aij[9], Dj[9], L[9];
        const int jd = pdiag[j];
        T* pij = uni + (size_t)ppos[ij] * 9;
        const T* pjd = uni + (size_t)ppos[jd] * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) { Aij[t] = pij[t]; Dj[t] = pjd[t]; }
        mat3_mul(Aij, Dj, L);                                 // L_ij = A_ij * inv(A_jj)
#pragma unroll
        for (int t = 0; t < 9; ++t) pij[t] = L[t];
        int jk = jd + 1, ik = ij + 1;
        const int jend = prowptr[j + 1];
        while (ik < iend && jk < jend) {
            const int ci = pcol[ik], cj = pcol[jk];
            if (ci == cj) {
                T Ajk[9], B[9];
                const T* pjk = uni + (size_t)ppos[jk] * 9;
                T* pik = uni + (size_t)ppos[ik] * 9;
#pragma unroll
                for (int t = 0; t < 9; ++t) Ajk[t] = pjk[t];
                mat3_mul(L, Ajk, B);                          // A_ik -= L_ij * A_jk
#pragma unroll
                for (int t = 0; t < 9; ++t) pik[t] = pik[t] - B[t];
                ++ik; ++jk;
            } else if (ci < cj) ++ik;
            else ++jk;
        }
    }
    T D[9];
    T* pd = uni + (size_t)ppos[idiag] * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) D[t] = pd[t];
    const T det = mat3_invert(D);
#pragma unroll
    for (int t = 0; t < 9; ++t) pd[t] = D[t];
    if (!(det != T(0)) || isinf(det) || isnan(det)) atomicMin(bad_row, i);
}

// rows of the first colour have no lower blocks: y = P d
template <class T>
__global__ void __launch_bounds__(256)
mc_copy_rows_kernel(int row0, int row1, const int* __restrict__ p2n, const T* __restrict__ d, T* __restrict__ W)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)(row1 - row0) * 3) return;
    const size_t q = row0 + e / 3;
    const int c = (int)(e % 3);
    W[q * 3 + c] = d[(size_t)p2n[q] * 3 + c];
}

struct McSweepArgs {
    int N, nnz;                  // rows of the system, blocks of this operand (L or U)
    const int* rowptr;           // [N+1] of the operand
    const int* colidx;           // permuted columns
    const void* vals;            // blocks of T
    int row0, row1;              // permuted rows of the colour
    const int* p2n;
    const void* d;               // right-hand side in natural order (lower kinds)
    void* W;                     // work vector in permuted order, updated in place
    const void* dinv;            // [N] inverted pivots in permuted order
    void* out;                   // result in natural order (kinds with the Dinv epilogue)
    double w;
    int scale;
};

constexpr size_t kMcSmemBytes = kSpmvSmemBytes + 2 * kSpmvGroups * kSpmvComputeWarps * 32 * sizeof(double);

// KIND 0: lower sweep of one colour.  KIND 1: lower sweep of the LAST colour (no upper blocks) with
// the upper epilogue.  KIND 2: upper sweep of one colour.
template <int KIND, class T, bool ROWT, int ROWS>
__global__ void __launch_bounds__(kSpmvThreads, 1)
mc_sweep_tma_kernel(McSweepArgs a)
{
    constexpr int kAl = sizeof(T) == 8 ? 2 : 4;
    constexpr int kCap = SpmvTile<ROWS>::cap, kValB = SpmvTile<ROWS>::val_bytes, kColB = SpmvTile<ROWS>::col_bytes;
    constexpr int kGroupThreads = kSpmvComputeWarps * 32;
    static_assert(ROWS == kSpmvRows || (ROWT && sizeof(T) == 4), "128-row tiles: float, one thread per row");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* empty = full + kSpmvStages;
    unsigned char* stages = smem_raw + 128;
    T* scratch = reinterpret_cast<T*>(smem_raw + kSpmvSmemBytes);      // [2][groups][kGroupThreads]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, nnzb = a.nnz;
    const int* __restrict__ rowptr = a.rowptr;
    const int* __restrict__ colidx = a.colidx;
    const T* __restrict__ vals = static_cast<const T*>(a.vals);
    const T* __restrict__ dinv = static_cast<const T*>(a.dinv);
    const int* __restrict__ p2n = a.p2n;
    const T* __restrict__ drhs = static_cast<const T*>(a.d);
    T* W = static_cast<T*>(a.W);
    T* out = static_cast<T*>(a.out);
    const T wrel = (T)a.w;
    const int t_begin = a.row0 / ROWS, t_end = (a.row1 + ROWS - 1) / ROWS;

    if (tid == 0) {
        for (int i = 0; i < kSpmvStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kSpmvComputeWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        int it = 0;
        int t = t_begin + blockIdx.x;
        int b0n = 0, b1n = 0;
        if (t < t_end) { b0n = rowptr[t * ROWS]; b1n = rowptr[min(N, t * ROWS + ROWS)]; }
        for (; t < t_end; t += gridDim.x, ++it) {
            const int st = it % kSpmvStages, k = it / kSpmvStages;
            const int r0 = t * ROWS, r1 = min(N, r0 + ROWS);
            const int b0 = b0n, b1 = b1n;
            const int tn = t + (int)gridDim.x;
            if (tn < t_end) { b0n = rowptr[tn * ROWS]; b1n = rowptr[min(N, tn * ROWS + ROWS)]; }
            const int b0a = b0 & ~(kAl - 1);
            const int b1a = (b1 + kAl - 1) & ~(kAl - 1);
            const bool direct = (b1a > nnzb) || (b1a - b0a > kCap);
            if (k > 0) { while (!mbar_try_wait(&empty[st], (unsigned)((k - 1) & 1))) {} }
            unsigned char* stage = stages + (size_t)st * kSpmvStageBytes;
            const unsigned pbytes = (unsigned)(((r1 - r0 + 1) * 4 + 15) & ~15);
            const int c0a = b0 & ~3;
            const unsigned cbytes = (direct || b1 == b0) ? 0u : (unsigned)((((b1 - c0a) * 4) + 15) & ~15);
            const unsigned vbytes = (direct || b1 == b0) ? 0u : (unsigned)((b1a - b0a) * 9 * (int)sizeof(T));
            if (lane == 0) mbar_arrive_expect_tx(&full[st], pbytes + cbytes + vbytes);
            __syncwarp();
            void* dst = lane == 0 ? (void*)stage : (lane == 1 ? (void*)(stage + kValB) : (void*)(stage + kValB + kColB));
            const void* src = lane == 0 ? (const void*)(vals + (size_t)b0a * 9) : (lane == 1 ? (const void*)(colidx + c0a) : (const void*)(rowptr + r0));
            const unsigned bytes = lane == 0 ? vbytes : (lane == 1 ? cbytes : pbytes);
            if (lane < 3 && bytes > 0) tma_bulk_g2s(dst, src, bytes, &full[st]);
        }
    } else {
        const int g = (warp - 1) / kSpmvComputeWarps;
        const int ct = tid - 32 - g * kGroupThreads;
        const int rl = ROWT ? ct : ct / 3, c = ROWT ? 0 : ct - rl * 3;
        int it = 0, mine = 0;
        for (int t = t_begin + blockIdx.x; t < t_end; t += gridDim.x, ++it) {
            if (it % kSpmvGroups != g) continue;
            const int st = it % kSpmvStages, k = it / kSpmvStages;
            const int r0 = t * ROWS, r1 = min(N, r0 + ROWS);
            const int r = r0 + rl;
            const bool active = rl < ROWS && r < r1 && r >= a.row0 && r < a.row1;
            // operands that do not depend on the tile's blocks: issued before the wait
            T init0 = T(0), init1 = T(0), init2 = T(0);
            T di[ROWT ? 9 : 3];
            int rn = 0;
            if (active) {
                rn = p2n[r];
                if (ROWT) {
                    const T* s = KIND == 2 ? W + (size_t)r * 3 : drhs + (size_t)rn * 3;
                    init0 = s[0]; init1 = s[1]; init2 = s[2];
                    if (KIND != 0) {
#pragma unroll
                        for (int q = 0; q < 9; ++q) di[q] = dinv[(size_t)r * 9 + q];
                    }
                } else {
                    init0 = KIND == 2 ? W[(size_t)r * 3 + c] : drhs[(size_t)rn * 3 + c];
                    if (KIND != 0) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) di[q] = dinv[(size_t)r * 9 + c * 3 + q];
                    }
                }
            }
            while (!mbar_try_wait(&full[st], (unsigned)(k & 1))) {}
            const unsigned char* stage = stages + (size_t)st * kSpmvStageBytes;
            const int* rp = reinterpret_cast<const int*>(stage + kValB + kColB);
            const int b0 = rp[0], b1 = rp[r1 - r0];
            const int b0a = b0 & ~(kAl - 1), b1a = (b1 + kAl - 1) & ~(kAl - 1), c0a = b0 & ~3;
            const bool direct = (b1a > nnzb) || (b1a - b0a > kCap);
            if (ROWT) {
                if (active) {
                    const int kb = rp[rl], ke = rp[rl + 1];
                    T a0 = init0, a1 = init1, a2 = init2;
                    if (!direct) {
                        const T* vs = reinterpret_cast<const T*>(stage);
                        const int* cs = reinterpret_cast<const int*>(stage + kValB);
                        for (int kk = kb; kk < ke; kk += 8) {
                            T xv[8][3];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                if (kk + u < ke) {
                                    const T* xj = W + (size_t)cs[kk + u - c0a] * 3;
                                    xv[u][0] = xj[0]; xv[u][1] = xj[1]; xv[u][2] = xj[2];
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                if (kk + u < ke) {
                                    const T* m = vs + (size_t)(kk + u - b0a) * 9;
                                    a0 = fma(-m[0], xv[u][0], a0); a0 = fma(-m[1], xv[u][1], a0); a0 = fma(-m[2], xv[u][2], a0);
                                    a1 = fma(-m[3], xv[u][0], a1); a1 = fma(-m[4], xv[u][1], a1); a1 = fma(-m[5], xv[u][2], a1);
                                    a2 = fma(-m[6], xv[u][0], a2); a2 = fma(-m[7], xv[u][1], a2); a2 = fma(-m[8], xv[u][2], a2);
                                }
                            }
                        }
                    } else {
                        for (int kk = kb; kk < ke; ++kk) {
                            const T* m = vals + (size_t)kk * 9;
                            const T* xj = W + (size_t)colidx[kk] * 3;
                            a0 = fma(-m[0], xj[0], a0); a0 = fma(-m[1], xj[1], a0); a0 = fma(-m[2], xj[2], a0);
                            a1 = fma(-m[3], xj[0], a1); a1 = fma(-m[4], xj[1], a1); a1 = fma(-m[5], xj[2], a1);
                            a2 = fma(-m[6], xj[0], a2); a2 = fma(-m[7], xj[1], a2); a2 = fma(-m[8], xj[2], a2);
                        }
                    }
                    const size_t o = (size_t)r * 3;
                    if (KIND == 0) {
                        W[o] = a0; W[o + 1] = a1; W[o + 2] = a2;
                    } else {
                        T v0 = T(0), v1 = T(0), v2 = T(0);
                        v0 = fma(di[0], a0, v0); v0 = fma(di[1], a1, v0); v0 = fma(di[2], a2, v0);
                        v1 = fma(di[3], a0, v1); v1 = fma(di[4], a1, v1); v1 = fma(di[5], a2, v1);
                        v2 = fma(di[6], a0, v2); v2 = fma(di[7], a1, v2); v2 = fma(di[8], a2, v2);
                        W[o] = v0; W[o + 1] = v1; W[o + 2] = v2;
                        const size_t on = (size_t)rn * 3;
                        out[on] = a.scale ? v0 * wrel : v0; out[on + 1] = a.scale ? v1 * wrel : v1; out[on + 2] = a.scale ? v2 * wrel : v2;
                    }
                }
            } else {
                T acc = init0;
                if (active) {
                    const int kb = rp[rl], ke = rp[rl + 1];
                    if (!direct) {
                        const T* vs = reinterpret_cast<const T*>(stage) + c * 3;
                        const int* cs = reinterpret_cast<const int*>(stage + kValB);
                        for (int kk = kb; kk < ke; kk += 8) {
                            T xv[8][3];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                if (kk + u < ke) {
                                    const T* xj = W + (size_t)cs[kk + u - c0a] * 3;
                                    xv[u][0] = xj[0]; xv[u][1] = xj[1]; xv[u][2] = xj[2];
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                if (kk + u < ke) {
                                    const T* m = vs + (size_t)(kk + u - b0a) * 9;
                                    acc = fma(-m[0], xv[u][0], acc);
                                    acc = fma(-m[1], xv[u][1], acc);
                                    acc = fma(-m[2], xv[u][2], acc);
                                }
                            }
                        }
                    } else {
                        for (int kk = kb; kk < ke; ++kk) {
                            const T* m = vals + (size_t)kk * 9 + c * 3;
                            const T* xj = W + (size_t)colidx[kk] * 3;
                            acc = fma(-m[0], xj[0], acc);
                            acc = fma(-m[1], xj[1], acc);
                            acc = fma(-m[2], xj[2], acc);
                        }
                    }
                }
                if (KIND == 0) {
                    if (active) W[(size_t)r * 3 + c] = acc;
                } else {
                    // the three components of a row meet through shared memory (a row's threads may
                    // sit in two warps); two buffers, one named barrier per tile
                    T* sc = scratch + ((size_t)(mine & 1) * kSpmvGroups + g) * kGroupThreads;
                    sc[ct] = acc;
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kGroupThreads) : "memory");
                    if (active) {
                        const T r0v = sc[rl * 3], r1v = sc[rl * 3 + 1], r2v = sc[rl * 3 + 2];
                        T v = T(0);
                        v = fma(di[0], r0v, v); v = fma(di[1], r1v, v); v = fma(di[2], r2v, v);
                        W[(size_t)r * 3 + c] = v;
                        out[(size_t)rn * 3 + c] = a.scale ? v * wrel : v;
                    }
                }
            }
            ++mine;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
    }
}

}  // namespace opmgpu
