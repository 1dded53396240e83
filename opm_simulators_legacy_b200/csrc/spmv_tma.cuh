// K2 (fast path)  BCRS 3x3 SpMV  y = A x  as a persistent TMA-pipelined stream (sm_100a).
//
// (Dune::MatrixAdapter::apply -> BCRSMatrix::mv -> umv; call site opm/autodiff/ISTLSolver.hpp:303)
//
// The values of consecutive block rows are one contiguous byte range, so a CTA walks tiles of
// 64 rows: one elected lane fetches the tile's values, column indices and row pointers with
// bulk async copies (cp.async.bulk + mbarrier complete_tx) into a 6-stage shared-memory ring;
// three groups of 192 compute threads, one thread per (row, component), read the blocks from shared memory in
// ascending column order (the reference's order: bit parity) and gather x through L1/L2
// (x is 24 MB at 1M cells, L2 resident).  The BiCGStab dot products that follow the SpMV are
// fused as an epilogue: MODE 1: S[S_H] = w1.y; MODE 2: S[S_TR] = y.w1, S[S_TT] = y.y, reduced
// deterministically over the (at most 148*k) CTAs.
// Algorithmic bytes: 76 per block + 52 per row.
#pragma once
#include "kernels.cuh"
#include "sweep_pipe.cuh"      // mbarrier / bulk-copy helpers

namespace opmgpu {

constexpr int kSpmvRows = 64;                   // rows per tile
constexpr int kSpmvCapBlocks = 480;             // blocks a stage can hold (>= 64*7 + alignment slack)
constexpr int kSpmvStages = 6;                  // a multiple of kSpmvGroups: a stage always serves the same
                                                // group, so that group sees every phase of its mbarriers
constexpr int kSpmvComputeWarps = 6;            // 192 threads = 64 rows x 3 components
constexpr int kSpmvGroups = 3;                  // consumer groups working on different tiles at once
constexpr int kSpmvThreads = 32 * (1 + kSpmvGroups * kSpmvComputeWarps);
constexpr int kSpmvValBytes = kSpmvCapBlocks * 72;
constexpr int kSpmvColBytes = ((kSpmvCapBlocks + 4) * 4 + 15) / 16 * 16;
constexpr int kSpmvPtrBytes = ((kSpmvRows + 1 + 3) * 4 + 15) / 16 * 16;
constexpr int kSpmvStageBytes = kSpmvValBytes + kSpmvColBytes + kSpmvPtrBytes;
constexpr size_t kSpmvSmemBytes = 128 + (size_t)kSpmvStages * kSpmvStageBytes;
static_assert(kSpmvStages % 3 == 0, "stages must be a multiple of the consumer groups");
// Float instance: 36-byte blocks, so a stage of the same size holds a tile of 128 rows (one thread
// per row) -- half as many tiles, i.e. half the per-tile cost of the producer warp.
constexpr int kSpmvRowsF32 = 128;
constexpr int kSpmvCapBlocksF32 = kSpmvRowsF32 * 7 + 8;
constexpr int kSpmvValBytesF32 = kSpmvCapBlocksF32 * 36;
constexpr int kSpmvColBytesF32 = ((kSpmvCapBlocksF32 + 4) * 4 + 15) / 16 * 16;
constexpr int kSpmvPtrBytesF32 = ((kSpmvRowsF32 + 1 + 3) * 4 + 15) / 16 * 16;
static_assert(kSpmvValBytesF32 % 16 == 0 && kSpmvValBytesF32 + kSpmvColBytesF32 + kSpmvPtrBytesF32 <= kSpmvStageBytes, "float tiles fit the stage");
template <int ROWS> struct SpmvTile {
    static constexpr int cap = ROWS == kSpmvRows ? kSpmvCapBlocks : kSpmvCapBlocksF32;
    static constexpr int val_bytes = ROWS == kSpmvRows ? kSpmvValBytes : kSpmvValBytesF32;
    static constexpr int col_bytes = ROWS == kSpmvRows ? kSpmvColBytes : kSpmvColBytesF32;
};

// rowptr / colidx must be readable up to 16 bytes past their end (the solver's own buffers are
// padded); vals is never read past its end (a misaligned last tile takes the direct path).
// T = float (the reference's Impl<3,float>): 36-byte blocks, so a tile's value range starts at a
// multiple of four blocks (16-byte alignment of the bulk copy) instead of two.
// row_skip != nullptr (row-partitioned systems): bit r of row_skip[t] marks row r of tile t as a
// boundary row -- it references a ghost column whose value may still be in flight (the halo
// exchange runs beside this kernel), so it is neither stored nor counted in the dot products here;
// spmv3_rows_kernel computes those rows when the ghosts have arrived.
// ROWT: one thread per block ROW (the first 64 threads of a group; three accumulators) instead of one
// per (row, component): a third of the x gathers, which is what bounds the float instance once the
// matrix bytes are halved.  Same operation order per component either way.
template <int MODE, class T, bool ROWT = false, int ROWS = kSpmvRows>
__global__ void __launch_bounds__(kSpmvThreads, 1)
spmv3_tma_kernel(int N, int nnzb, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                 const T* __restrict__ vals, const T* __restrict__ x, T* __restrict__ y,
                 const T* __restrict__ w1, double* S, ReduceWs ws, const unsigned long long* __restrict__ row_skip = nullptr)
{
    constexpr int kAl = sizeof(T) == 8 ? 2 : 4;          // blocks per 16-byte aligned unit (144 bytes)
    constexpr int kCap = SpmvTile<ROWS>::cap, kValB = SpmvTile<ROWS>::val_bytes, kColB = SpmvTile<ROWS>::col_bytes;
    static_assert(ROWS == kSpmvRows || (ROWT && sizeof(T) == 4), "128-row tiles: float, one thread per row");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* empty = full + kSpmvStages;
    unsigned char* stages = smem_raw + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (N + ROWS - 1) / ROWS;

    if (tid == 0) {
        for (int i = 0; i < kSpmvStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kSpmvComputeWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    T d0 = T(0), d1 = T(0);                     // dot partials of this thread
    if (warp == 0) {
        // Producer warp.  The per-tile cost of this loop bounds the whole kernel once the tiles are
        // small (a tile is 32 KB of values in double, 16 KB in float): the block range of the NEXT tile
        // is fetched while this one is issued, and the tile's three bulk copies (values, column
        // indices, row pointers) leave as ONE warp instruction, lanes 0..2 with their own operands.
        int it = 0;
        int t = blockIdx.x;
        int b0n = 0, b1n = 0;
        if (t < ntiles) { b0n = rowptr[t * ROWS]; b1n = rowptr[min(N, t * ROWS + ROWS)]; }
        for (; t < ntiles; t += gridDim.x, ++it) {
            const int st = it % kSpmvStages, k = it / kSpmvStages;
            const int r0 = t * ROWS, r1 = min(N, r0 + ROWS);
            const int b0 = b0n, b1 = b1n;
            const int tn = t + (int)gridDim.x;
            if (tn < ntiles) { b0n = rowptr[tn * ROWS]; b1n = rowptr[min(N, tn * ROWS + ROWS)]; }
            const int b0a = b0 & ~(kAl - 1);                          // 16-byte aligned start
            int b1a = (b1 + kAl - 1) & ~(kAl - 1);
            const bool direct = (b1a > nnzb) || (b1a - b0a > kCap);   // odd tail / oversized tile
            if (k > 0) { while (!mbar_try_wait(&empty[st], (unsigned)((k - 1) & 1))) {} }
            unsigned char* stage = stages + (size_t)st * kSpmvStageBytes;
            // the row pointers always travel; values and columns unless the tile is direct
            const unsigned pbytes = (unsigned)(((r1 - r0 + 1) * 4 + 15) & ~15);
            const int c0a = b0 & ~3;
            const unsigned cbytes = direct ? 0u : (unsigned)((((b1 - c0a) * 4) + 15) & ~15);
            const unsigned vbytes = direct ? 0u : (unsigned)((b1a - b0a) * 9 * (int)sizeof(T));
            if (lane == 0) mbar_arrive_expect_tx(&full[st], pbytes + cbytes + vbytes);
            __syncwarp();
            void* dst = lane == 0 ? (void*)stage : (lane == 1 ? (void*)(stage + kValB) : (void*)(stage + kValB + kColB));
            const void* src = lane == 0 ? (const void*)(vals + (size_t)b0a * 9) : (lane == 1 ? (const void*)(colidx + c0a) : (const void*)(rowptr + r0));
            const unsigned bytes = lane == 0 ? vbytes : (lane == 1 ? cbytes : pbytes);
            if (lane < 3 && bytes > 0) tma_bulk_g2s(dst, src, bytes, &full[st]);
        }
    } else {
        // consumer group g takes every kSpmvGroups-th tile of this CTA, so several tiles' x
        // gathers (L2 latency) are in flight per SM
        const int g = (warp - 1) / kSpmvComputeWarps;
        const int ct = tid - 32 - g * kSpmvComputeWarps * 32;
        const int rl = ROWT ? ct : ct / 3, c = ROWT ? 0 : ct - rl * 3;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            if (it % kSpmvGroups != g) continue;
            const int st = it % kSpmvStages, k = it / kSpmvStages;
            const int r0 = t * ROWS, r1 = min(N, r0 + ROWS);
            while (!mbar_try_wait(&full[st], (unsigned)(k & 1))) {}
            const unsigned char* stage = stages + (size_t)st * kSpmvStageBytes;
            const int* rp = reinterpret_cast<const int*>(stage + kValB + kColB);
            const int b0 = rp[0], b1 = rp[r1 - r0];
            const int b0a = b0 & ~(kAl - 1), b1a = (b1 + kAl - 1) & ~(kAl - 1), c0a = b0 & ~3;
            const bool direct = (b1a > nnzb) || (b1a - b0a > kCap);
            const int r = r0 + rl;
            const bool skip = row_skip && rl < ROWS && ((row_skip[t] >> rl) & 1ull);
            if (ROWT && rl < ROWS && r < r1 && !skip) {
                const int kb = rp[rl], ke = rp[rl + 1];
                T a0 = T(0), a1 = T(0), a2 = T(0);
                if (!direct) {
                    const T* vs = reinterpret_cast<const T*>(stage);
                    const int* cs = reinterpret_cast<const int*>(stage + kValB);
                    for (int kk = kb; kk < ke; kk += 8) {
                        T xv[8][3];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (kk + u < ke) {
                                const T* xj = x + (size_t)cs[kk + u - c0a] * 3;
                                xv[u][0] = xj[0]; xv[u][1] = xj[1]; xv[u][2] = xj[2];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (kk + u < ke) {
                                const T* a = vs + (size_t)(kk + u - b0a) * 9;
                                a0 = fma(a[0], xv[u][0], a0); a0 = fma(a[1], xv[u][1], a0); a0 = fma(a[2], xv[u][2], a0);
                                a1 = fma(a[3], xv[u][0], a1); a1 = fma(a[4], xv[u][1], a1); a1 = fma(a[5], xv[u][2], a1);
                                a2 = fma(a[6], xv[u][0], a2); a2 = fma(a[7], xv[u][1], a2); a2 = fma(a[8], xv[u][2], a2);
                            }
                        }
                    }
                } else {
                    for (int kk = kb; kk < ke; ++kk) {
                        const T* a = vals + (size_t)kk * 9;
                        const T* xj = x + (size_t)colidx[kk] * 3;
                        a0 = fma(a[0], xj[0], a0); a0 = fma(a[1], xj[1], a0); a0 = fma(a[2], xj[2], a0);
                        a1 = fma(a[3], xj[0], a1); a1 = fma(a[4], xj[1], a1); a1 = fma(a[5], xj[2], a1);
                        a2 = fma(a[6], xj[0], a2); a2 = fma(a[7], xj[1], a2); a2 = fma(a[8], xj[2], a2);
                    }
                }
                const size_t o = (size_t)r * 3;
                y[o] = a0; y[o + 1] = a1; y[o + 2] = a2;
                if (MODE == 1) { d0 = fma(w1[o], a0, d0); d0 = fma(w1[o + 1], a1, d0); d0 = fma(w1[o + 2], a2, d0); }
                if (MODE == 2) {
                    d0 = fma(a0, w1[o], d0); d1 = fma(a0, a0, d1);
                    d0 = fma(a1, w1[o + 1], d0); d1 = fma(a1, a1, d1);
                    d0 = fma(a2, w1[o + 2], d0); d1 = fma(a2, a2, d1);
                }
            } else if (!ROWT && r < r1 && !skip) {
                const int kb = rp[rl], ke = rp[rl + 1];
                T acc = T(0);
                if (!direct) {
                    const T* vs = reinterpret_cast<const T*>(stage) + c * 3;
                    const int* cs = reinterpret_cast<const int*>(stage + kValB);
                    // all x gathers of up to eight blocks are issued before the FMA chain
                    for (int kk = kb; kk < ke; kk += 8) {
                        T xv[8][3];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (kk + u < ke) {
                                const T* xj = x + (size_t)cs[kk + u - c0a] * 3;
                                xv[u][0] = xj[0]; xv[u][1] = xj[1]; xv[u][2] = xj[2];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (kk + u < ke) {
                                const T* a = vs + (size_t)(kk + u - b0a) * 9;
                                acc = fma(a[0], xv[u][0], acc);
                                acc = fma(a[1], xv[u][1], acc);
                                acc = fma(a[2], xv[u][2], acc);
                            }
                        }
                    }
                } else {
                    for (int kk = kb; kk < ke; ++kk) {
                        const T* a = vals + (size_t)kk * 9 + c * 3;
                        const T* xj = x + (size_t)colidx[kk] * 3;
                        acc = fma(a[0], xj[0], acc);
                        acc = fma(a[1], xj[1], acc);
                        acc = fma(a[2], xj[2], acc);
                    }
                }
                const size_t o = (size_t)r * 3 + c;
                y[o] = acc;
                if (MODE == 1) d0 = fma(w1[o], acc, d0);
                if (MODE == 2) { d0 = fma(acc, w1[o], d0); d1 = fma(acc, acc, d1); }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
    }
    if (MODE == 1) {
        T v[1] = {d0};
        grid_reduce<1, T>(v, ws, [=](T (&u)[1]) { S[S_H] = u[0]; });
    } else if (MODE == 2) {
        T v[2] = {d0, d1};
        grid_reduce<2, T>(v, ws, [=](T (&u)[2]) { S[S_TR] = u[0]; S[S_TT] = u[1]; });
    }
}

// The boundary rows of a row-partitioned system (those spmv3_tma_kernel skipped): one thread per
// (listed row, component), blocks in ascending column order as everywhere; the dot products of
// MODE 1 / 2 are ADDED to what the first kernel left in the scalar slots.
template <int MODE, class T>
__global__ void __launch_bounds__(256)
spmv3_rows_kernel(int nrows, const int* __restrict__ rows, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                  const T* __restrict__ vals, const T* __restrict__ x, T* __restrict__ y,
                  const T* __restrict__ w1, double* S, ReduceWs ws)
{
    T d0 = T(0), d1 = T(0);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)nrows * 3; e += (size_t)gridDim.x * blockDim.x) {
        const int row = rows[e / 3], c = (int)(e % 3);
        T acc = T(0);
        for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
            const T* a = vals + (size_t)k * 9 + c * 3;
            const T* xj = x + (size_t)colidx[k] * 3;
            acc = fma(a[0], xj[0], acc);
            acc = fma(a[1], xj[1], acc);
            acc = fma(a[2], xj[2], acc);
        }
        const size_t o = (size_t)row * 3 + c;
        y[o] = acc;
        if (MODE == 1) d0 = fma(w1[o], acc, d0);
        if (MODE == 2) { d0 = fma(acc, w1[o], d0); d1 = fma(acc, acc, d1); }
    }
    if (MODE == 1) {
        T v[1] = {d0};
        grid_reduce<1, T>(v, ws, [=](T (&u)[1]) { S[S_H] = (double)((T)S[S_H] + u[0]); });
    } else if (MODE == 2) {
        T v[2] = {d0, d1};
        grid_reduce<2, T>(v, ws, [=](T (&u)[2]) { S[S_TR] = (double)((T)S[S_TR] + u[0]); S[S_TT] = (double)((T)S[S_TT] + u[1]); });
    }
}

}  // namespace opmgpu
