// K3 (fast path)  Pipelined-wavefront block ILU0 factorisation for sm_100a.
//
//   Opm::ParallelOverlappingILU0 ctor -> Dune::bilu0_decomposition, natural order
//   (call site opm/autodiff/ISTLSolver.hpp:201-211); 3x3 inverse by OPM's adjugate formula.
//
// On a stencil-like pattern every lower block (i,j) only updates the diagonal of row i:
//     L_ij = A_ij * inv(D_j),   D_i = A_ii - sum_j L_ij * A_ji,   then D_i is inverted in place,
// so the only thing that travels along the dependency graph is the inverted pivot block (9
// doubles) -- the same graph as the lower triangular sweep.  The kernel therefore reuses the
// sweep's machinery (sweep_pipe.cuh): persistent CTAs own (i,j) column tiles and walk them level
// by level; the A blocks a row needs (A_ii, A_ij, A_ji) arrive as a linear stream of step records
// through a TMA-fed shared-memory ring (packed by pack_factor_records_kernel); pivots of the
// CTA's own recent rows live in a shared-memory window; pivots of other tiles are pushed into
// self-validating slots in L2 and staged by a helper warp; two groups of compute warps
// ping-pong the steps on named barriers.  One thread owns one block row; the arithmetic is the
// reference's, operation for operation (bit parity with the oracle).
//
// Output: only the inverted pivots, in program order (FactorPipeProgram::fpos), written by bulk
// async stores (TMA) straight from the shared-memory window -- scattered 8-byte stores of L and
// D into the BCRS array cost more LSU time than the whole chain (measured: 920 us with them,
// 380 us without, 1M cells).  L_ij = A_ij * inv(D_j) is recomputed where it is consumed
// (repack_pipe2_kernel: same three fused multiply-adds per element, so still bit-identical).
// Store protocol: the rows of step s sit in consecutive window entries; after the hand-over
// barrier of step s+1 the elected thread of the OTHER group issues the bulk store of step s
// (writers fence the async proxy before arriving) and, before its next barrier, waits until
// the store has read shared memory.  Entries are only overwritten >= 3 steps later
// (kFWindow >= 3 * kLeanStepRows + the two steps in between), i.e. after that wait.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "analysis.hpp"
#include "sweep_pipe.cuh"

namespace opmgpu {

constexpr int kFGroups = 2;
constexpr int kFComputeWarps = 3;
constexpr int kFThreads = 32 * (2 + kFGroups * kFComputeWarps);
static_assert(kFComputeWarps * 32 == kLeanStepRows, "one pass of a group covers a step");
constexpr int kFDepBytes = (kFWindow + kFRing) * kFEntry * 8;
static_assert(kFWindow >= 3 * kLeanStepRows, "window entries must survive until their bulk store has been read");
constexpr int kFMaxStages = 8;

struct FactorPipeDev {
    const unsigned char* buf;
    const int* cta_step_ptr;
    const unsigned* step_off16;
    const unsigned* step_bytes;
    const long long* cta_ext_base;
    const int* cta_row_base;     // program position of the CTA's first row
    double* ext;                 // push slots, 9 doubles each, all-ones when empty
    double* fout;                // inverted pivots in program order, kFEntry doubles per row
    int stage_bytes;
    int nstages;
    int dbg;
};

__host__ __device__ inline size_t factor_pipe_smem_bytes(int nstages, int stage_bytes)
{
    return 512 + (size_t)kFDepBytes + (size_t)nstages * (size_t)stage_bytes;
}

// A blocks (BCRS) -> step records: one thread per block row (3 values).  TA = type the blocks are
// stored in; CONT: they are containers already (the rank's diagonal block of a partitioned system)
template <class TA, bool CONT = false>
__global__ void __launch_bounds__(256)
pack_factor_records_kernel(size_t nval, const int* __restrict__ src, const unsigned* __restrict__ dst8,
                           const TA* __restrict__ vals, double* __restrict__ bufd)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nval * 3) return;
    const size_t b = t / 3;
    const int c = (int)(t - b * 3);
    const TA* a = vals + (size_t)src[b] * 9 + c * 3;
    double* d = bufd + (size_t)dst8[b] + c * 3;
    const TA a0 = a[0], a1 = a[1], a2 = a[2];
    if (CONT) { d[0] = (double)a0; d[1] = (double)a1; d[2] = (double)a2; }       // (TA = double: a plain copy)
    else { d[0] = enc(a0); d[1] = enc(a1); d[2] = enc(a2); }
}

template <class AT>
__device__ __forceinline__ AT factor_invert3(AT (&M)[9])
{
    AT A[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) A[q] = M[q];
    const AT t4 = A[0] * A[4], t6 = A[0] * A[5], t8 = A[1] * A[3];
    const AT t10 = A[2] * A[3], t12 = A[1] * A[6], t14 = A[2] * A[6];
    const AT det = (t4 * A[8] - t6 * A[7] - t8 * A[8] + t10 * A[7] + t12 * A[5] - t14 * A[4]);
    const AT t17 = AT(1) / det;
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

// pivots of one step: window entries [q0 % kFWindow, ...) -> fout[(row_base + q0) ...], one or two
// bulk async stores (the window wraps), committed as one group of the calling thread
__device__ __forceinline__ void factor_store_step(const FactorPipeDev& pg, const double* dep, int row_base, int q0, int n)
{
    if (n <= 0) return;
    const int e0 = q0 % kFWindow;
    const int n0 = min(n, kFWindow - e0);
    double* dst = pg.fout + (size_t)(row_base + q0) * kFEntry;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(dep + (size_t)e0 * kFEntry)), "r"((unsigned)(n0 * kFEntry * 8)) : "memory");
    if (n0 < n)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(dst + (size_t)n0 * kFEntry), "r"(smem_u32(dep)), "r"((unsigned)((n - n0) * kFEntry * 8)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// AT: arithmetic type (double / float); records, window, push slots and fout are 8-byte containers
template <class AT>
__global__ void __launch_bounds__(kFThreads, 1)
ilu0_factor_pipe_kernel(FactorPipeDev pg, int* bad_row, int* err)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeCtl* ctl = reinterpret_cast<PipeCtl*>(smem_raw);
    double* dep = reinterpret_cast<double*>(smem_raw + 512);      // window | pushed ring (9-double entries)
    unsigned char* stages = smem_raw + 512 + kFDepBytes;
    const int S = pg.nstages;
    const size_t stage_stride = (size_t)pg.stage_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = pg.cta_step_ptr[blockIdx.x];
    const int nsteps = pg.cta_step_ptr[blockIdx.x + 1] - s0;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], kFComputeWarps); }
        ctl->ext_consumed = 0; ctl->ext_ready = 0; ctl->abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (nsteps == 0) return;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (one elected lane)
        if (lane == 0) {
            for (int i = 0; i < nsteps; ++i) {
                const int st = i % S, k = i / S;
                const unsigned off16 = pg.step_off16[s0 + i], bytes = pg.step_bytes[s0 + i];
                if (k > 0 && !pipe_wait(&ctl->empty[st], (unsigned)((k - 1) & 1), ctl, err)) break;
                mbar_arrive_expect_tx(&ctl->full[st], bytes);
                tma_bulk_g2s(stages + st * stage_stride, pg.buf + (size_t)off16 * 16, bytes, &ctl->full[st]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ pushed-pivot helper
        // polls the next kFPoll slots in consumption order and delivers the valid prefix into the
        // shared-memory ring, then publishes ext_ready and re-arms the slots
        const long long base = pg.cta_ext_base[blockIdx.x];
        const int total = (int)(pg.cta_ext_base[blockIdx.x + 1] - base);
        const long long* slots = reinterpret_cast<const long long*>(pg.ext) + (size_t)base * 9;
        double* ring = dep + (size_t)kFWindow * kFEntry;
        unsigned spins = 0;
        int e = 0;
        while (e < total) {
            const int limit = min(total, ctl->ext_consumed + kFRing);
            long long a[2][9];
            unsigned m[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int idx = e + u * 32 + lane;
                bool ok = idx < limit;
                if (ok) {
                    const long long* sl = slots + (size_t)idx * 9;
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(a[u][t]) : "l"(sl + t) : "memory");
#pragma unroll
                    for (int t = 0; t < 9; ++t) ok = ok && a[u][t] != -1;
                }
                m[u] = __ballot_sync(0xffffffffu, ok);
            }
            int n = 0;
            if (m[0] != 0xffffffffu) n = __ffs(~m[0]) - 1;
            else if (m[1] != 0xffffffffu) n = 32 + __ffs(~m[1]) - 1;
            else n = 64;
            if (n > 0) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int idx = e + u * 32 + lane;
                    if (u * 32 + lane < n) {
                        double* dst = ring + (size_t)(idx & (kFRing - 1)) * kFEntry;
#pragma unroll
                        for (int t = 0; t < 9; ++t) dst[t] = __longlong_as_double(a[u][t]);
                    }
                }
                __syncwarp();                        // ring data before ext_ready
                e += n;
                if (lane == 0) ctl->ext_ready = e;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int idx = e - n + u * 32 + lane;
                    if (u * 32 + lane < n) {
                        long long* sl = const_cast<long long*>(slots) + (size_t)idx * 9;
#pragma unroll
                        for (int t = 0; t < 9; ++t) __stcg(sl + t, -1LL);
                    }
                }
                spins = 0;
            } else {
                ++spins;
                const int ab = __shfl_sync(0xffffffffu, (int)ctl->abort_flag, 0);
                if (ab) break;
                if (spins > kPipeSpinLimit) { if (lane == 0) { ctl->abort_flag = 1; atomicExch(err, 7); } break; }
            }
        }
    } else {
        // ------------------------------------------------ compute warps (two groups ping-pong)
        constexpr int G = kFGroups;
        constexpr int NPP = 2 * kFComputeWarps * 32;
        const int g = (warp - 2) / kFComputeWarps;
        const int cw = (warp - 2) - g * kFComputeWarps;
        const int r = cw * 32 + lane;
        const bool elected = cw == 0 && lane == 0;
        const int bar_prev = 1 + (g + G - 1) % G, bar_own = 1 + g;
        bool dead = false;
        int st = g % S;
        unsigned par = (unsigned)((g / S) & 1);
        int ext_prev_end = -1;
        const int row_base = pg.cta_row_base[blockIdx.x];
        int prev_q0 = 0, prev_n = 0;          // step s-1 (the other group's): its rows' window entries
        for (int s = g; s < nsteps; s += G) {
            const unsigned char* rec = stages + st * stage_stride;
            // ---- everything that does not depend on earlier rows: into registers, stage back
            int n = 0, ext_end = 0, ext_cnt = 0;
            int4 ri0 = make_int4(0, 0, 0, 0), ri1 = make_int4(0, -1, -1, 0), ri2 = make_int4(-1, -1, -1, 0);
            AT v[kFRowVals];
            bool on = false;
            if (!dead) {
                if (pipe_wait(&ctl->full[st], par, ctl, err)) {
                    const int4 h0 = *reinterpret_cast<const int4*>(rec);
                    n = h0.x; ext_end = h0.z; ext_cnt = h0.w;
                    prev_n = reinterpret_cast<const int*>(rec)[4];      // rows of step s-1: [qbase - prev_n, qbase)
                    prev_q0 = h0.y - prev_n;
                    on = r < n;
                    if (on) {
                        const int4* rip = reinterpret_cast<const int4*>(rec + 32) + 3 * r;
                        ri0 = rip[0]; ri1 = rip[1]; ri2 = rip[2];
                        const double* vp = reinterpret_cast<const double*>(rec + 32 + (size_t)n * (kFRowInts * 4)) + (size_t)r * kFRowVals;
#pragma unroll
                        for (int q = 0; q < kFRowVals; ++q) v[q] = dec<AT>(vp[q]);
                    }
                } else dead = true;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->empty[st]);
            // the bulk store this thread issued two steps ago has read its window entries
            if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            // ext_ready only grows: the usual answer is fetched while the group waits for its turn
            const bool ext_ok = dead || ext_cnt <= 0 || ctl->ext_ready >= ext_end;
            if (s > 0) asm volatile("bar.sync %0, %1;" ::"r"(bar_prev), "n"(NPP) : "memory");     // step s-1 done
            if (!ext_ok) {
                if (!pipe_wait_ext(ctl, ext_end, err)) { dead = true; on = false; }
                asm volatile("" ::: "memory");
            }
            AT D[9];
            AT det = AT(1);
            if (on) {
                const int mask = ri1.x;
#pragma unroll
                for (int t = 0; t < 9; ++t) D[t] = v[t];
#pragma unroll
                for (int kb = 0; kb < kFastBlocks; ++kb) {
                    if (mask & (1 << kb)) {
                        const int de = kb == 0 ? ri0.y : (kb == 1 ? ri0.z : ri0.w);
                        const double* dj = dep + (size_t)de * kFEntry;
                        AT Dj[9], L[9];
#pragma unroll
                        for (int t = 0; t < 9; ++t) Dj[t] = dec<AT>(dj[t]);
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                AT sacc = AT(0);
#pragma unroll
                                for (int k = 0; k < 3; ++k) sacc = fma(v[9 + kb * 18 + c * 3 + k], Dj[k * 3 + j], sacc);
                                L[c * 3 + j] = sacc;
                            }
                        if (mask & (1 << (4 + kb))) {
#pragma unroll
                            for (int c = 0; c < 3; ++c)
#pragma unroll
                                for (int j = 0; j < 3; ++j) {
                                    AT sacc = AT(0);
#pragma unroll
                                    for (int k = 0; k < 3; ++k) sacc = fma(L[c * 3 + k], v[18 + kb * 18 + k * 3 + j], sacc);
                                    D[c * 3 + j] -= sacc;
                                }
                        }
                    }
                }
                det = factor_invert3(D);
                double* w = dep + (size_t)ri1.w * kFEntry;
#pragma unroll
                for (int t = 0; t < 9; ++t) w[t] = enc(D[t]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // window -> bulk store
            }
            asm volatile("bar.arrive %0, %1;" ::"r"(bar_own), "n"(NPP) : "memory");                // step s done
            // off the in-tile critical path: pivots other CTAs wait for, pivots of step s-1 -> HBM
            if (on) {
                if (ri1.y >= 0) {
                    double* sl = pg.ext + (size_t)ri1.y * 9;
#pragma unroll
                    for (int t = 0; t < 9; ++t) push_f64(sl + t, enc(D[t]));
                }
                if (ri1.z >= 0) {
                    double* sl = pg.ext + (size_t)ri1.z * 9;
#pragma unroll
                    for (int t = 0; t < 9; ++t) push_f64(sl + t, enc(D[t]));
                }
            }
            if (elected && s > 0) factor_store_step(pg, dep, row_base, prev_q0, prev_n);
            // every warp of this group passed the bar.sync of this step, i.e. is done with its
            // step s-G: release the pushed-pivot ring entries of that step
            if (elected && ext_prev_end >= 0) ctl->ext_consumed = ext_prev_end;
            if (on && (!(det != AT(0)) || isinf(det) || isnan(det))) atomicMin(bad_row, ri0.x);
            ext_prev_end = ext_end;
            st += G;
            if (st >= S) { st -= S; par ^= 1u; }
        }
        // consume the last hand-over addressed to this group so no barrier is left half-arrived,
        // and store the last step's pivots
        if (nsteps > 0 && (nsteps % G) == g) {
            asm volatile("bar.sync %0, %1;" ::"r"(bar_prev), "n"(NPP) : "memory");
            if (elected) {
                const int* lh = reinterpret_cast<const int*>(pg.buf + (size_t)pg.step_off16[s0 + nsteps - 1] * 16);
                factor_store_step(pg, dep, row_base, lh[1], lh[0]);
            }
        }
                // shared memory must stay alive until the bulk stores have read it
        if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// factors -> step records of the pipelined sweeps, from A and the program-ordered pivots.  One
// thread per (record block b, block row c): a lower block becomes row c of L_ij = A_ij *
// inv(D_j) (the factorisation's own three fused multiply-adds per element), a diagonal block
// becomes row c of inv(D_j), an upper block is A's.  Element [c][e] goes to dst8[b] + c*stride[b] + e.
// PART 0: everything, 1: only the blocks copied from A, 2: only the pivots
// TA: type A is stored in (CONT: as containers), AT: arithmetic type of the product.
template <bool LOWER, int PART, class TA, class AT, bool CONT = false>
__global__ void __launch_bounds__(256)
repack_pipe2_kernel(size_t nval, const int* __restrict__ src, const unsigned* __restrict__ dst8,
                    const int* __restrict__ stride, const int* __restrict__ colidx, const int* __restrict__ diag,
                    const int* __restrict__ fpos, const TA* __restrict__ A, const double* __restrict__ fout,
                    double* __restrict__ bufd)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nval * 3; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t / 3;
        const int c = (int)(t - b * 3);
        const int k = src[b];
        const int j = colidx[k];
        double o[3];
        const bool pivot = k == diag[j];
        if (PART == 1 && pivot) continue;
        if (PART == 2 && !pivot) continue;
        if (pivot) {
            const double* d = fout + (size_t)fpos[j] * kFEntry + c * 3;
            o[0] = d[0]; o[1] = d[1]; o[2] = d[2];
        } else {
            const TA* a = A + (size_t)k * 9 + c * 3;
            if (LOWER) {
                const double* d = fout + (size_t)fpos[j] * kFEntry;
                const AT a0 = CONT ? dec<AT>((double)a[0]) : (AT)a[0], a1 = CONT ? dec<AT>((double)a[1]) : (AT)a[1],
                         a2 = CONT ? dec<AT>((double)a[2]) : (AT)a[2];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    AT sacc = AT(0);
                    sacc = fma(a0, dec<AT>(d[e]), sacc); sacc = fma(a1, dec<AT>(d[3 + e]), sacc); sacc = fma(a2, dec<AT>(d[6 + e]), sacc);
                    o[e] = enc(sacc);
                }
            } else if (CONT) { o[0] = (double)a[0]; o[1] = (double)a[1]; o[2] = (double)a[2]; }
            else { o[0] = enc((AT)a[0]); o[1] = enc((AT)a[1]); o[2] = enc((AT)a[2]); }
        }
        double* dst = bufd + (size_t)dst8[b] + (size_t)c * stride[b];
        dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    }
}

// BCRS factor array on demand (opmgpu_ilu0_get_factors): in place on a copy of A.  One thread
// per (row, block row c).
template <class AT>
__global__ void __launch_bounds__(256)
materialise_lu_kernel(int N, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                      const int* __restrict__ fpos, const double* __restrict__ fout, double* lu)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)N * 3) return;
    const int i = (int)(t / 3), c = (int)(t - (size_t)i * 3);
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        if (j > i) break;
        double* a = lu + (size_t)k * 9 + c * 3;
        const double* d = fout + (size_t)fpos[j] * kFEntry;
        if (j == i) { a[0] = d[c * 3]; a[1] = d[c * 3 + 1]; a[2] = d[c * 3 + 2]; }
        else {
            const AT a0 = dec<AT>(a[0]), a1 = dec<AT>(a[1]), a2 = dec<AT>(a[2]);
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                AT sacc = AT(0);
                sacc = fma(a0, dec<AT>(d[e]), sacc); sacc = fma(a1, dec<AT>(d[3 + e]), sacc); sacc = fma(a2, dec<AT>(d[6 + e]), sacc);
                a[e] = enc(sacc);
            }
        }
    }
}

}  // namespace opmgpu
