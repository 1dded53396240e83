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

// Dune::bilu0_decomposition of one row of the permuted pattern.  Lower entry l of the row (block l of the
// unified array) eliminates with the inverted pivot of row j = Lcol[l] (block offD + j); the blocks it
// updates were found on the host (pair lists: the same pairs, in the same order, as the reference's walk
// over both rows).
template <class T>
__device__ __forceinline__ void mc_factor_row(int i, const int* __restrict__ Lrowptr, const int* __restrict__ Lcol,
                                              const int* __restrict__ pair_ptr, const int* __restrict__ pair_jk,
                                              const int* __restrict__ pair_ik, long long offD, T* uni, int* bad_row)
{
    const int lend = Lrowptr[i + 1];
    for (int l = Lrowptr[i]; l < lend; ++l) {
        const int j = Lcol[l];
        T Aij[9], Dj[9], L[9];
        T* pij = uni + (size_t)l * 9;
        const T* pjd = uni + (size_t)(offD + j) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) { Aij[t] = pij[t]; Dj[t] = pjd[t]; }
        mat3_mul(Aij, Dj, L);                                 // L_ij = A_ij * inv(A_jj)
#pragma unroll
        for (int t = 0; t < 9; ++t) pij[t] = L[t];
        const int pend = pair_ptr[l + 1];
        for (int p = pair_ptr[l]; p < pend; ++p) {
            T Ajk[9], B[9];
            const T* pjk = uni + (size_t)pair_jk[p] * 9;
            T* pik = uni + (size_t)pair_ik[p] * 9;
#pragma unroll
            for (int t = 0; t < 9; ++t) Ajk[t] = pjk[t];
            mat3_mul(L, Ajk, B);                              // A_ik -= L_ij * A_jk
#pragma unroll
            for (int t = 0; t < 9; ++t) pik[t] = pik[t] - B[t];
        }
    }
    T D[9];
    T* pd = uni + (size_t)(offD + i) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) D[t] = pd[t];
    const T det = mat3_invert(D);
#pragma unroll
    for (int t = 0; t < 9; ++t) pd[t] = D[t];
    if (!(det != T(0)) || isinf(det) || isnan(det)) atomicMin(bad_row, i);
}

// one dependency level per launch, one thread per row of the level
template <class T>
__global__ void __launch_bounds__(128)
mc_factor_level_kernel(const int* __restrict__ lvl_rows, int begin, int end, const int* __restrict__ Lrowptr,
                       const int* __restrict__ Lcol, const int* __restrict__ pair_ptr, const int* __restrict__ pair_jk,
                       const int* __restrict__ pair_ik, long long offD, T* uni, int* bad_row)
{
    const int s = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= end) return;
    mc_factor_row<T>(lvl_rows[s], Lrowptr, Lcol, pair_ptr, pair_jk, pair_ik, offD, uni, bad_row);
}

// k-line ordering: one launch per colour, one thread per column walking the planes (a column's rows depend
// on each other and on rows of the colour factorised by the previous launch)
template <class T>
__global__ void __launch_bounds__(32)
mc_factor_lines_kernel(int base, int ncols, int nz, const int* __restrict__ Lrowptr, const int* __restrict__ Lcol,
                       const int* __restrict__ pair_ptr, const int* __restrict__ pair_jk, const int* __restrict__ pair_ik,
                       long long offD, T* uni, int* bad_row)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    for (int k = 0; k < nz; ++k)
        mc_factor_row<T>(base + k * ncols + col, Lrowptr, Lcol, pair_ptr, pair_jk, pair_ik, offD, uni, bad_row);
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


// ------------------------------------------------------------------------------------------------
// k-line variant (OPMGPU_ILU_MULTICOLOUR_LINES): red-black over the (i,j) COLUMNS of a Cartesian grid,
// natural order along k inside a column.  The strong vertical couplings keep their natural order (one
// half step more than the reference's ordering at 1M cells, where point red-black more than doubles
// the count: tools/ordering_study.py), and a column is a chain that belongs to ONE thread triple: in
// the reference's visiting order the chain's own previous row is the LAST block of a row, so only its
// three fused multiply-adds per component (and Dinv in the upper sweep) wait for the previous step.
// No result ever travels between CTAs inside a launch: a CTA owns RB columns of one colour and walks
// the planes, its rows of a plane being one contiguous tile of the permuted operand (TMA-streamed
// several planes ahead); the other colour's results were written by the previous launch.
// Per step: block rows from shared memory, neighbours' results gathered from L2, chain values kept
// in registers; the three components of a row meet through shared memory once per step.
// ------------------------------------------------------------------------------------------------
struct McLineArgs {
    int nnz;                     // blocks of the operand (L or U)
    const int* rowptr;           // operand row pointers (permuted rows)
    const int* colidx;           // permuted columns (L ascending, U descending)
    const void* vals;
    int base, ncols, nz, rb;     // the colour's first permuted row, its columns, planes, columns per CTA
    const int* p2n;
    const void* d;               // right-hand side, natural order (lower sweep)
    void* W;                     // work vector, permuted order, updated in place
    const void* dinv;
    void* out;                   // result, natural order (upper sweep)
    double w;
    int scale;
    int obase, oncols, nx, plane;      // the other colour's first row and columns, grid width, cells per plane (L2 prefetch ranges)
    long long* trace;                  // experiments build: clock64 stamps of CTA 0, first consumer warp, [nz][8]
};

// L2 prefetch of a byte range (rounded outwards to 16 bytes, clipped to [lo, hi))
__device__ __forceinline__ void line_prefetch_l2(const void* base, long long lo_byte, long long hi_byte, long long limit_byte)
{
    lo_byte = lo_byte < 0 ? 0 : (lo_byte & ~15ll);
    hi_byte = (hi_byte + 15) & ~15ll;
    if (hi_byte > (limit_byte & ~15ll)) hi_byte = limit_byte & ~15ll;
    if (hi_byte <= lo_byte) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(static_cast<const char*>(base) + lo_byte), "r"((unsigned)(hi_byte - lo_byte)) : "memory");
}

constexpr int kLineGroupThreads = kSpmvComputeWarps * 32;          // 192 = 64 rows x 3 components
constexpr int kLineThreads = 32 + kLineGroupThreads;
constexpr int kLineStages = 5;
constexpr int kLineAhead = 4;             // planes the per-row operands are fetched ahead
constexpr int kLineStageBytes = kSpmvValBytes + kSpmvColBytes;
// barriers | tile headers | stages | scratch [4][threads] | value ring [ahead][threads][4] | index ring [ahead][threads][2] | p2n ring [2*ahead][threads]
constexpr size_t kLineSmemBytes = 256 + (size_t)kLineStages * kLineStageBytes + 4 * kLineGroupThreads * sizeof(double) +
                                  (size_t)kLineAhead * kLineGroupThreads * (4 * sizeof(double) + 2 * sizeof(int)) +
                                  (size_t)2 * kLineAhead * kLineGroupThreads * sizeof(int);
static_assert(kLineSmemBytes <= 227 * 1024, "line sweep: shared memory of one CTA");

template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst_smem, const void* src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "n"(BYTES) : "memory");
}

template <bool UPPER, class T>
__global__ void __launch_bounds__(kLineThreads, 1)
mc_line_sweep_kernel(McLineArgs a)
{
    constexpr int kAl = sizeof(T) == 8 ? 2 : 4;
    constexpr int kCap = kSpmvCapBlocks;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* empty = full + kLineStages;
    int* hdr = reinterpret_cast<int*>(smem_raw + 128);          // [stages][2]: block range of the tile in the stage
    unsigned char* stages = smem_raw + 256;
    T* scratch = reinterpret_cast<T*>(smem_raw + 256 + (size_t)kLineStages * kLineStageBytes);      // [4][kLineGroupThreads]
    T* ring_v = reinterpret_cast<T*>(smem_raw + 256 + (size_t)kLineStages * kLineStageBytes + 4 * kLineGroupThreads * sizeof(double));
    int* ring_i = reinterpret_cast<int*>(smem_raw + 256 + (size_t)kLineStages * kLineStageBytes + 4 * kLineGroupThreads * sizeof(double) +
                                         (size_t)kLineAhead * kLineGroupThreads * 4 * sizeof(double));
    int* ring_rn = ring_i + (size_t)kLineAhead * kLineGroupThreads * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nnzb = a.nnz;
    const int* __restrict__ rowptr = a.rowptr;
    const int* __restrict__ colidx = a.colidx;
    const T* __restrict__ vals = static_cast<const T*>(a.vals);
    const T* __restrict__ dinv = static_cast<const T*>(a.dinv);
    const int* __restrict__ p2n = a.p2n;
    const T* __restrict__ drhs = static_cast<const T*>(a.d);
    T* W = static_cast<T*>(a.W);
    T* out = static_cast<T*>(a.out);
    const T wrel = (T)a.w;
    const int col0 = blockIdx.x * a.rb;
    const int cnt = min(a.rb, a.ncols - col0);            // rows of a tile
    const int nz = a.nz;

    if (tid == 0) {
        for (int i = 0; i < kLineStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kSpmvComputeWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // producer: the blocks and column indices of the tile of every plane, several planes ahead.  The
        // tile's block range comes from the row pointers: lane l fetches the range of plane s + l once per
        // 32 planes (a load per plane would serialise this loop on its own latency)
        int rb0 = 0, rb1 = 0;
        for (int s = 0; s < nz; ++s) {
            const int st = s % kLineStages, ph = s / kLineStages;
            if ((s & 31) == 0) {
                const int sl = s + lane;
                if (sl < nz) {
                    const int ql = a.base + (UPPER ? nz - 1 - sl : sl) * a.ncols + col0;
                    rb0 = rowptr[ql]; rb1 = rowptr[ql + cnt];
                }
            }
            const int b0 = __shfl_sync(0xffffffffu, rb0, s & 31), b1 = __shfl_sync(0xffffffffu, rb1, s & 31);
            const int b0a = b0 & ~(kAl - 1);
            const int b1a = (b1 + kAl - 1) & ~(kAl - 1);
            const bool direct = (b1a > nnzb) || (b1a - b0a > kCap);
            if (ph > 0) { while (!mbar_try_wait(&empty[st], (unsigned)((ph - 1) & 1))) {} }
            unsigned char* stage = stages + (size_t)st * kLineStageBytes;
            const int c0a = b0 & ~3;
            const unsigned cbytes = (direct || b1 == b0) ? 0u : (unsigned)((((b1 - c0a) * 4) + 15) & ~15);
            const unsigned vbytes = (direct || b1 == b0) ? 0u : (unsigned)((b1a - b0a) * 9 * (int)sizeof(T));
            if (lane == 0) {
                hdr[st * 2] = b0; hdr[st * 2 + 1] = b1;          // (the arrive below releases them to the consumers)
                if (cbytes + vbytes == 0) mbar_arrive(&full[st]);
                else mbar_arrive_expect_tx(&full[st], cbytes + vbytes);
            }
            __syncwarp();
            void* dst = lane == 0 ? (void*)stage : (void*)(stage + kSpmvValBytes);
            const void* src = lane == 0 ? (const void*)(vals + (size_t)b0a * 9) : (const void*)(colidx + c0a);
            const unsigned bytes = lane == 0 ? vbytes : cbytes;
            if (lane < 2 && bytes > 0) tma_bulk_g2s(dst, src, bytes, &full[st]);
            // what the consumers load themselves (their rows' pointers, permutation, pivots, right-hand side
            // and the other colour's results around their columns) is pulled into L2 at the same lead
            {
                const int k = UPPER ? nz - 1 - s : s;
                const long long q0 = a.base + (long long)k * a.ncols + col0;
                const long long ntot = (long long)a.plane * nz;
                if (lane == 2) line_prefetch_l2(rowptr, q0 * 4, (q0 + cnt + 1) * 4, (ntot + 1) * 4);
                if (lane == 3) line_prefetch_l2(p2n, q0 * 4, (q0 + cnt) * 4, ntot * 4);
                if (lane == 4 && UPPER) line_prefetch_l2(dinv, q0 * 9 * (long long)sizeof(T), (q0 + cnt) * 9 * (long long)sizeof(T), ntot * 9 * (long long)sizeof(T));
                if (lane == 5 && UPPER) line_prefetch_l2(W, q0 * 3 * (long long)sizeof(T), (q0 + cnt) * 3 * (long long)sizeof(T), ntot * 3 * (long long)sizeof(T));
                if (lane == 5 && !UPPER) {          // right-hand side: the cells of my columns in plane k (natural numbering)
                    const long long c_lo = (long long)k * a.plane + 2ll * col0 - 2, c_hi = (long long)k * a.plane + 2ll * (col0 + cnt) + 2;
                    line_prefetch_l2(drhs, c_lo * 3 * (long long)sizeof(T), c_hi * 3 * (long long)sizeof(T), ntot * 3 * (long long)sizeof(T));
                }
                if (lane == 6 && a.oncols > 0 && b1 - b0 > cnt) {      // rows with horizontal neighbours: the other colour's plane-k results near my columns
                    const long long o0 = a.obase + (long long)k * a.oncols + col0 - (a.nx / 2 + 2), o1 = a.obase + (long long)k * a.oncols + col0 + cnt + (a.nx / 2 + 2);
                    const long long lo = a.obase + (long long)k * a.oncols, hi = lo + a.oncols;
                    line_prefetch_l2(W, (o0 < lo ? lo : o0) * 3 * (long long)sizeof(T), (o1 > hi ? hi : o1) * 3 * (long long)sizeof(T), ntot * 3 * (long long)sizeof(T));
                }
            }
        }
    } else {
        // Consumers.  What a row needs besides its blocks (block range, natural row, right-hand side or
        // lower result, its row of Dinv) is fetched kLineAhead planes ahead with cp.async into a small
        // per-thread ring in shared memory (the natural row two rings ahead, because the right-hand side
        // is gathered through it): a load issued at plane s and consumed at plane s + 1 costs a full
        // memory latency (~1.1 us measured) per plane.
        const int ct = tid - 32;
        const int rl = ct / 3, c = ct - rl * 3;
        const bool active = rl < cnt;
        T y0 = T(0), y1 = T(0), y2 = T(0);          // the chain: my row's result of the previous plane
        int buf = 0;
        auto row_of = [&](int s) { return a.base + (UPPER ? nz - 1 - s : s) * a.ncols + col0 + rl; };
        auto issue = [&](int s) {                   // loads of plane s (p2n of plane s + kLineAhead); one group
            if (active) {
                if (s < nz) {
                    const int q = row_of(s);
                    int* ri = ring_i + ((size_t)(s % kLineAhead) * kLineGroupThreads + ct) * 2;
                    T* rv = ring_v + ((size_t)(s % kLineAhead) * kLineGroupThreads + ct) * 4;
                    cp_async<4>(ri, rowptr + q); cp_async<4>(ri + 1, rowptr + q + 1);
                    if (UPPER) {
                        cp_async<sizeof(T)>(rv, W + (size_t)q * 3 + c);
                        cp_async<sizeof(T)>(rv + 1, dinv + (size_t)q * 9 + c * 3);
                        cp_async<sizeof(T)>(rv + 2, dinv + (size_t)q * 9 + c * 3 + 1);
                        cp_async<sizeof(T)>(rv + 3, dinv + (size_t)q * 9 + c * 3 + 2);
                    } else {
                        const int rn4 = ring_rn[(size_t)(s % (2 * kLineAhead)) * kLineGroupThreads + ct];
                        cp_async<sizeof(T)>(rv, drhs + (size_t)rn4 * 3 + c);
                    }
                }
                if (s + kLineAhead < nz) cp_async<4>(ring_rn + (size_t)((s + kLineAhead) % (2 * kLineAhead)) * kLineGroupThreads + ct, p2n + row_of(s + kLineAhead));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // prologue: natural rows of the first kLineAhead planes, then the planes themselves
        if (active)
            for (int s = 0; s < kLineAhead && s < nz; ++s) cp_async<4>(ring_rn + (size_t)s * kLineGroupThreads + ct, p2n + row_of(s));
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        for (int s = 0; s < kLineAhead; ++s) issue(s);
        for (int s = 0; s < nz; ++s) {
            const int st = s % kLineStages, ph = s / kLineStages;
            asm volatile("cp.async.wait_group %0;" ::"n"(kLineAhead - 1) : "memory");      // the group of plane s has landed
            const int q = row_of(s);
            int kb = 0, ke = 0, rn = 0;
            T init = T(0), di0 = T(0), di1 = T(0), di2 = T(0);
            if (active) {
                const int* ri = ring_i + ((size_t)(s % kLineAhead) * kLineGroupThreads + ct) * 2;
                const T* rv = ring_v + ((size_t)(s % kLineAhead) * kLineGroupThreads + ct) * 4;
                kb = ri[0]; ke = ri[1]; init = rv[0];
                if (UPPER) { di0 = rv[1]; di1 = rv[2]; di2 = rv[3]; }
                rn = ring_rn[(size_t)(s % (2 * kLineAhead)) * kLineGroupThreads + ct];
            }
            issue(s + kLineAhead);
            const bool tr = a.trace && blockIdx.x == 0 && ct == 0;
            if (tr) a.trace[s * 8 + 0] = clock64();
            const int qchain = s == 0 ? -1 : (UPPER ? q + a.ncols : q - a.ncols);      // my column one plane back in the walk
            while (!mbar_try_wait(&full[st], (unsigned)(ph & 1))) {}
            const unsigned char* stage = stages + (size_t)st * kLineStageBytes;
            const int b0 = hdr[st * 2], b1 = hdr[st * 2 + 1];          // the tile's block range, left by the producer
            const int b0a = b0 & ~(kAl - 1), b1a = (b1 + kAl - 1) & ~(kAl - 1), c0a = b0 & ~3;
            const bool direct = (b1a > nnzb) || (b1a - b0a > kCap);
            if (tr) a.trace[s * 8 + 1] = clock64();
            T acc = init;
            if (tr) a.trace[s * 8 + 2] = pipe_clock_after((int)__double_as_longlong((double)acc));
            // every block but the chain's is gathered from the work vector (written by earlier launches); the
            // chain's block comes last (mcorder.cpp checks it) and uses the values this thread kept.
            // (Pointers into the stage must stay shared-memory pointers for the compiler: two copies of the loop.)
            if (active) {
                if (!direct) {
                    const T* vs = reinterpret_cast<const T*>(stage) + c * 3;
                    const int* cs = reinterpret_cast<const int*>(stage + kSpmvValBytes);
                    for (int kk = kb; kk < ke; kk += 8) {
                        T xv[8][3];
                        int cj[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (kk + u < ke) {
                                cj[u] = cs[kk + u - c0a];
                                if (cj[u] != qchain) {
                                    const T* xj = W + (size_t)cj[u] * 3;
                                    xv[u][0] = xj[0]; xv[u][1] = xj[1]; xv[u][2] = xj[2];
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (kk + u < ke) {
                                const T* m = vs + (size_t)(kk + u - b0a) * 9;
                                const bool chain = cj[u] == qchain;
                                const T x0 = chain ? y0 : xv[u][0], x1 = chain ? y1 : xv[u][1], x2 = chain ? y2 : xv[u][2];
                                acc = fma(-m[0], x0, acc);
                                acc = fma(-m[1], x1, acc);
                                acc = fma(-m[2], x2, acc);
                            }
                        }
                    }
                } else {
                    for (int kk = kb; kk < ke; ++kk) {
                        const T* m = vals + (size_t)kk * 9 + c * 3;
                        const int cjk = colidx[kk];
                        const bool chain = cjk == qchain;
                        const T* xj = W + (size_t)(chain ? 0 : cjk) * 3;
                        const T x0 = chain ? y0 : xj[0], x1 = chain ? y1 : xj[1], x2 = chain ? y2 : xj[2];
                        acc = fma(-m[0], x0, acc);
                        acc = fma(-m[1], x1, acc);
                        acc = fma(-m[2], x2, acc);
                    }
                }
            }
            if (tr) a.trace[s * 8 + 3] = pipe_clock_after((int)__double_as_longlong((double)acc));
            // the three components of a row meet (a row's threads may sit in two warps)
            T* sc = scratch + (size_t)buf * kLineGroupThreads;
            sc[ct] = acc;
            asm volatile("bar.sync 1, %0;" ::"n"(kLineGroupThreads) : "memory");
            if (tr) a.trace[s * 8 + 4] = clock64();
            T mine = acc;
            if (active) {
                const T r0v = sc[rl * 3], r1v = sc[rl * 3 + 1], r2v = sc[rl * 3 + 2];
                if (!UPPER) {
                    y0 = r0v; y1 = r1v; y2 = r2v;
                } else {
                    T v = T(0);
                    v = fma(di0, r0v, v); v = fma(di1, r1v, v); v = fma(di2, r2v, v);
                    mine = v;
                }
            }
            if (UPPER) {          // the next plane needs all three components of x: a second meeting
                T* sx = scratch + (size_t)(2 + buf) * kLineGroupThreads;
                sx[ct] = mine;
                asm volatile("bar.sync 1, %0;" ::"n"(kLineGroupThreads) : "memory");
                if (active) { y0 = sx[rl * 3]; y1 = sx[rl * 3 + 1]; y2 = sx[rl * 3 + 2]; }
            }
            buf ^= 1;
            if (active) {
                W[(size_t)q * 3 + c] = mine;
                if (UPPER) out[(size_t)rn * 3 + c] = a.scale ? mine * wrel : mine;
            }
            if (tr) a.trace[s * 8 + 5] = pipe_clock_after((int)__double_as_longlong((double)y0));
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
            if (tr) a.trace[s * 8 + 6] = clock64();
            if (tr) a.trace[s * 8 + 7] = clock64();
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
}

}  // namespace opmgpu
