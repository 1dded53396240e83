// K4 (fast path)  Pipelined-wavefront ILU0 sweeps for sm_100a.
//
//   v = w U^-1 L^-1 d      (Opm::ParallelOverlappingILU0::apply; natural-order block ILU0,
//                           call site opm/autodiff/ISTLSolver.hpp:201-211)
//
// The exact natural-order sweeps have nx+ny+nz-2 dependency levels on a Cartesian grid
// (298 at 100^3).  A grid-wide barrier per level costs more than the whole sweep's HBM time,
// so the sweep is a pipelined wavefront of persistent CTAs instead:
//
//   * every CTA owns an (i,j) column tile of cells (analysis.cpp) and walks its rows level by
//     level ("steps"); dependencies inside the tile are served from a shared-memory window of
//     recent results, ordered by a CTA-local named barrier;
//   * a result another CTA needs is PUSHED by its producer into that consumer's slot in HBM/L2.
//     Slots are self-validating (all-ones = empty, each double is one atomic 8-byte store), so
//     neither side needs a flag or a memory fence; a helper warp polls the CTA's slots in
//     consumption order, stages them into a shared-memory ring and re-arms them;
//   * the CTA's part of the factors is a linear byte stream of step records laid out per
//     consuming thread (analysis.hpp), and the right-hand side arrives in the same program
//     order; one elected producer lane fetches both with bulk async copies (TMA,
//     cp.async.bulk + mbarrier complete_tx) into a ring of stages several steps ahead.
//
// A lone warp retires one dependent instruction every ~5-7 cycles, so the step time is the
// length of the dependent instruction stream between two step barriers.  The compute warps
// therefore keep on that path only: dependency loads (shared memory) -> 9-FMA chain in the
// reference's order (bit parity) -> window store -> barrier; everything that does not depend
// on the previous step is loaded into registers one step ahead.
//
// Two groups of compute warps alternate steps (ping-pong): while one group runs the chain of
// step s, the other has already pulled the static data of step s+1 into registers and waits
// on a named barrier; the global stores of step s are issued after the hand-over.
//
// Warp roles: warp 0 = TMA producer, warp 1 = pushed-result helper, warps 2..9 and 10..17 =
// the two compute groups (one thread per block row and component).  Every wait is bounded;
// on expiry the kernel raises *err and all roles drain (barrier hand-shakes keep running).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "analysis.hpp"

namespace opmgpu {

constexpr int kPipeComputeWarps = 8;
constexpr int kPipeThreads = 32 * (2 + 2 * kPipeComputeWarps);     // two ping-pong compute groups
constexpr int kPipeRowsPerPass = kPipeComputeWarps * 10;
constexpr int kPipeMaxStages = 8;
constexpr unsigned kPipeSpinLimit = 1u << 21;
constexpr int kPipeDepBytes = ((kDepZeroSlot + 1) * 24 + 15) / 16 * 16;

struct PipeDev {
    const unsigned char* buf;
    const int* cta_step_ptr;
    const unsigned* step_off16;
    const unsigned* step_bytes;
    const unsigned* step_rhs_row;
    const unsigned* step_rhs_bytes;
    const long long* cta_ext_base;
    double* ext;                 // push slots, 3 doubles each, all-ones when empty
    int stage_bytes;             // record capacity of a stage (multiple of 16)
    int rhs_bytes;               // rhs area of a stage (multiple of 16)
    int nstages;
    long long* trace;            // optional (debug): per-step clock64 stamps of CTA trace_cta
    int trace_cta;
};

struct PipeCtl {
    unsigned long long full[kPipeMaxStages];    // record + rhs landed (TMA complete_tx)
    unsigned long long empty[kPipeMaxStages];   // stage consumed
    volatile int ext_consumed;
    volatile int ext_ready;
    volatile int abort_flag;
};

__host__ __device__ inline size_t pipe_smem_bytes(int nstages, int stage_bytes, int rhs_bytes)
{
    return 512 + (size_t)kPipeDepBytes + (size_t)nstages * ((size_t)stage_bytes + rhs_bytes);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// named barrier over the compute warps that also ORs a predicate: a uniform abort decision
__device__ __forceinline__ bool bar_or(int nthreads, bool pred)
{
    unsigned r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "bar.red.or.pred q, 1, %2, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r) : "r"((unsigned)pred), "r"(nthreads) : "memory");
    return r != 0;
}

__device__ __forceinline__ bool pipe_wait(unsigned long long* bar, unsigned parity, PipeCtl* ctl, int* err)
{
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (ctl->abort_flag) return false;
        if (++spins > kPipeSpinLimit) { ctl->abort_flag = 1; atomicExch(err, 2); return false; }
    }
    return true;
}

__device__ __forceinline__ bool pipe_wait_ext(PipeCtl* ctl, int ext_end, int* err)
{
    unsigned spins = 0;
    while (ctl->ext_ready < ext_end) {
        if (ctl->abort_flag) return false;
        if (++spins > kPipeSpinLimit * 8u) { ctl->abort_flag = 1; atomicExch(err, 4); return false; }
    }
    return true;
}

// ---- compute-warp helpers --------------------------------------------------------------------
// Everything about a thread's (row r, component c; slot j = 3r + c) in a step that does not
// depend on earlier results; loaded from the landed stage one step ahead.
template <bool UPPER>
struct StepPre {
    int n, qbase, ext_end, ext_cnt;
    int4 ri0, ri1;              // rowinfo, dep0, dep1, dep2 | upos, push0, push1, -
    bool on;
    double rhs;
    double cf[9];
    double dv[3];

    // stage = rhs area followed by the record
    __device__ __forceinline__ void load(const unsigned char* stage, int rhs_bytes, int r, int j, bool lane_on)
    {
        const unsigned char* rec = stage + rhs_bytes;
        const int4 h0 = *reinterpret_cast<const int4*>(rec);
        n = h0.x; qbase = h0.y; ext_end = h0.z; ext_cnt = h0.w;
        on = lane_on && r < n;
        if (on) {
            const double* cfp = reinterpret_cast<const double*>(rec + 32) + j * 9;
#pragma unroll
            for (int q = 0; q < 9; ++q) cf[q] = cfp[q];
            const int T = 3 * n;
            size_t off = 32 + (size_t)T * 72;
            if (UPPER) {
                const double* dp = reinterpret_cast<const double*>(rec + off) + j * 3;
                dv[0] = dp[0]; dv[1] = dp[1]; dv[2] = dp[2];
                off += (size_t)T * 24;
            }
            const int4* rip = reinterpret_cast<const int4*>(rec + ((off + 15) & ~(size_t)15)) + 2 * r;
            ri0 = rip[0]; ri1 = rip[1];
            rhs = reinterpret_cast<const double*>(stage)[j];
        }
    }
};

__device__ __forceinline__ const double* dep_ptr(int code, const double* dep, const double* work)
{
    return code >= 0 ? dep + code : work + (size_t)(code & kDepValueMask) * 3;
}

// blocks beyond the three held in registers (rows with many couplings, e.g. well cells)
__device__ __noinline__ double sweep_tail_blocks(const unsigned char* rec, int r, int c, const double* dep,
                                                 const double* work, double acc)
{
    const int* hdr = reinterpret_cast<const int*>(rec);
    const int n = hdr[0];
    const int* lists = reinterpret_cast<const int*>(rec + (size_t)hdr[5] * 8);
    const int* tail_end = lists;
    const int* tail_dep = lists + 2 * n;
    const double* tail_vals = reinterpret_cast<const double*>(rec + (size_t)hdr[6] * 8);
    for (int t = r ? tail_end[r - 1] : 0; t < tail_end[r]; ++t) {
        const double* yp = dep_ptr(tail_dep[t], dep, work);
        const double* ap = tail_vals + (size_t)t * 9 + c * 3;
        acc = fma(-ap[0], yp[0], acc);
        acc = fma(-ap[1], yp[1], acc);
        acc = fma(-ap[2], yp[2], acc);
    }
    return acc;
}
// pushes beyond the two held in registers
__device__ __noinline__ void sweep_extra_pushes(const unsigned char* rec, int r, int c, double* ext, double acc)
{
    const int* hdr = reinterpret_cast<const int*>(rec);
    const int n = hdr[0];
    const int* lists = reinterpret_cast<const int*>(rec + (size_t)hdr[5] * 8);
    const int* xpush_end = lists + n;
    const int* xpush_slot = lists + 2 * n + hdr[4];
    for (int t = r ? xpush_end[r - 1] : 0; t < xpush_end[r]; ++t) __stcg(ext + (size_t)xpush_slot[t] * 3 + c, acc);
}

// one (row, component), critical part: dependency loads, FMA chain, Dinv (upper), window store
template <bool UPPER, bool LEAN>
__device__ __forceinline__ double sweep_row_chain(const StepPre<UPPER>& p, const unsigned char* rec, int r, int c, int rl,
                                                  double* dep, const double* work)
{
    double acc = 0.0;
    if (p.on) {
        acc = p.rhs;
        double y[9];
        if (LEAN || (p.ri0.y | p.ri0.z | p.ri0.w) >= 0) {    // all three in shared memory (the common case)
            const double* y0 = dep + p.ri0.y; const double* y1 = dep + p.ri0.z; const double* y2 = dep + p.ri0.w;
            y[0] = y0[0]; y[1] = y0[1]; y[2] = y0[2];
            y[3] = y1[0]; y[4] = y1[1]; y[5] = y1[2];
            y[6] = y2[0]; y[7] = y2[1]; y[8] = y2[2];
        } else {
            const double* y0 = dep_ptr(p.ri0.y, dep, work); const double* y1 = dep_ptr(p.ri0.z, dep, work);
            const double* y2 = dep_ptr(p.ri0.w, dep, work);
            y[0] = y0[0]; y[1] = y0[1]; y[2] = y0[2];
            y[3] = y1[0]; y[4] = y1[1]; y[5] = y1[2];
            y[6] = y2[0]; y[7] = y2[1]; y[8] = y2[2];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) acc = fma(-p.cf[q], y[q], acc);
        if (!LEAN && (p.ri0.x & kRowSlow)) acc = sweep_tail_blocks(rec, r, c, dep, work, acc);
    }
    if (UPPER) {
        // v = Dinv * r needs the whole row vector: exchange inside the warp
        const int base = rl * 3;
        const double r0 = __shfl_sync(0xffffffffu, acc, base);
        const double r1 = __shfl_sync(0xffffffffu, acc, base + 1);
        const double r2 = __shfl_sync(0xffffffffu, acc, base + 2);
        if (p.on) {
            double v = 0.0;
            v = fma(p.dv[0], r0, v); v = fma(p.dv[1], r1, v); v = fma(p.dv[2], r2, v);
            acc = v;
        }
    }
    if (p.on) dep[p.ri1.w + c] = acc;
    return acc;
}
// ... and the part nobody inside the CTA waits for: results to HBM, pushes to other CTAs
template <bool UPPER, bool LEAN>
__device__ __forceinline__ void sweep_row_stores(const StepPre<UPPER>& p, const unsigned char* rec, int r, int c, double acc,
                                                 double* work, double* hand_off, double* out, double* ext, double w, int scale)
{
    if (p.on) {
        const int row = p.ri0.x & kRowMask;
        if (p.ri1.y >= 0) __stcg(ext + (size_t)p.ri1.y * 3 + c, acc);
        if (p.ri1.z >= 0) __stcg(ext + (size_t)p.ri1.z * 3 + c, acc);
        if (!LEAN && (p.ri0.x & kRowSlow)) sweep_extra_pushes(rec, r, c, ext, acc);
        if (UPPER) out[(size_t)row * 3 + c] = scale ? acc * w : acc;
        else hand_off[(size_t)p.ri1.x * 3 + c] = acc;
        if (!LEAN && (p.ri0.x & kRowWriteGlobal)) work[(size_t)row * 3 + c] = acc;
    }
}

// rows [80, n) of a step wider than one pass over the compute warps
template <bool UPPER>
__device__ __noinline__ void sweep_extra_rows(const unsigned char* stage, int rhs_bytes, int n, int r_first, int c, int rl,
                                              bool lane_on, double* dep, double* work, double* hand_off,
                                              double* out, double* ext, double w, int scale)
{
    for (int rbase = kPipeRowsPerPass; rbase < n; rbase += kPipeRowsPerPass) {
        const int r = rbase + r_first;
        StepPre<UPPER> p;
        p.load(stage, rhs_bytes, r, 3 * r + c, lane_on);
        const double acc = sweep_row_chain<UPPER, false>(p, stage + rhs_bytes, r, c, rl, dep, work);
        sweep_row_stores<UPPER, false>(p, stage + rhs_bytes, r, c, acc, work, hand_off, out, ext, w, scale);
    }
}

// LEAN: the program has no slow rows, no own-result reads from HBM and no step wider than one
// pass (every Cartesian stencil case): those paths are compiled out.
template <bool UPPER, bool LEAN>
__global__ void __launch_bounds__(kPipeThreads, 1)
ilu0_sweep_pipe_kernel(PipeDev pg, const double* __restrict__ rhs_perm, double* work, double* hand_off,
                       double* out, double w, int scale, int* err)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeCtl* ctl = reinterpret_cast<PipeCtl*>(smem_raw);
    double* dep = reinterpret_cast<double*>(smem_raw + 512);      // window | pushed ring | zero entry
    unsigned char* stages = smem_raw + 512 + kPipeDepBytes;
    const int S = pg.nstages;
    const size_t stage_stride = (size_t)pg.stage_bytes + pg.rhs_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = pg.cta_step_ptr[blockIdx.x];
    const int nsteps = pg.cta_step_ptr[blockIdx.x + 1] - s0;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
        ctl->ext_consumed = 0; ctl->ext_ready = 0; ctl->abort_flag = 0;
        dep[kDepZeroSlot * 3] = 0.0; dep[kDepZeroSlot * 3 + 1] = 0.0; dep[kDepZeroSlot * 3 + 2] = 0.0;
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
                const unsigned rrow = pg.step_rhs_row[s0 + i], rbytes = pg.step_rhs_bytes[s0 + i];
                if (k > 0 && !pipe_wait(&ctl->empty[st], (unsigned)((k - 1) & 1), ctl, err)) break;
                unsigned char* stage = stages + st * stage_stride;
                mbar_arrive_expect_tx(&ctl->full[st], bytes + rbytes);
                tma_bulk_g2s(stage + pg.rhs_bytes, pg.buf + (size_t)off16 * 16, bytes, &ctl->full[st]);
                tma_bulk_g2s(stage, rhs_perm + (size_t)rrow * 3, rbytes, &ctl->full[st]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ pushed-result helper
        // Polls up to 96 slots per round trip, in consumption order; delivers the valid prefix.
        const long long base = pg.cta_ext_base[blockIdx.x];
        const int total = (int)(pg.cta_ext_base[blockIdx.x + 1] - base);
        double* ring = dep + kWindowRows * 3;
        int e = 0;
        unsigned spins = 0;
        long long polls = 0;
        const bool htr = pg.trace && blockIdx.x == pg.trace_cta && lane == 0;
        if (htr) pg.trace[509 * 16 + 8] = clock64();
        while (e < total) {
            ++polls;
            const int limit = min(total, ctl->ext_consumed + kExtRing);
            long long a[3][3];
            unsigned m[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int idx = e + u * 32 + lane;
                a[u][0] = a[u][1] = a[u][2] = -1;
                if (idx < limit) {
                    const volatile long long* sl = reinterpret_cast<const volatile long long*>(pg.ext) + (size_t)(base + idx) * 3;
                    a[u][0] = sl[0]; a[u][1] = sl[1]; a[u][2] = sl[2];
                }
            }
#pragma unroll
            for (int u = 0; u < 3; ++u)
                m[u] = __ballot_sync(0xffffffffu, a[u][0] != -1 && a[u][1] != -1 && a[u][2] != -1);
            int n = 0;
            if (m[0] != 0xffffffffu) n = __ffs(~m[0]) - 1;
            else if (m[1] != 0xffffffffu) n = 32 + __ffs(~m[1]) - 1;
            else if (m[2] != 0xffffffffu) n = 64 + __ffs(~m[2]) - 1;
            else n = 96;
            if (n > 0) {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int idx = e + u * 32 + lane;
                    if (u * 32 + lane < n) {
                        double* dst = ring + (idx & (kExtRing - 1)) * 3;
                        dst[0] = __longlong_as_double(a[u][0]); dst[1] = __longlong_as_double(a[u][1]); dst[2] = __longlong_as_double(a[u][2]);
                        long long* sl = reinterpret_cast<long long*>(pg.ext) + (size_t)(base + idx) * 3;
                        __stcg(sl + 0, -1LL); __stcg(sl + 1, -1LL); __stcg(sl + 2, -1LL);      // re-arm
                    }
                }
                __syncwarp();
                __threadfence_block();
                if (htr && e == 0) { pg.trace[509 * 16 + 9] = clock64(); pg.trace[509 * 16 + 10] = polls; pg.trace[509 * 16 + 11] = n; }
                e += n;
                if (lane == 0) ctl->ext_ready = e;
                spins = 0;
            } else {
                ++spins;                                              // n is warp-uniform, so is spins
                const int ab = __shfl_sync(0xffffffffu, (int)ctl->abort_flag, 0);
                if (ab) break;
                if (spins > kPipeSpinLimit) { if (lane == 0) { ctl->abort_flag = 1; atomicExch(err, 3); } break; }
            }
        }
    } else {
        // ------------------------------------------------ compute warps (two ping-pong groups)
        // lane = 3 * (row % 10) + c, ten rows per warp (lanes 30, 31 idle) so the three
        // components of a row sit in one warp (shuffles in the upper sweep).  Group g owns the
        // steps s = g, g+2, ...; it signals "step s done" on named barrier 1+g (bar.arrive)
        // and waits for "step s-1 done" on barrier 2-g (bar.sync).
        constexpr int NPP = 2 * kPipeComputeWarps * 32;
        const int g = (warp - 2) / kPipeComputeWarps;
        const int cw = (warp - 2) - g * kPipeComputeWarps;
        const int rl = lane / 3, c = lane - rl * 3;
        const bool lane_on = lane < 30;
        const int r_first = cw * 10 + rl;
        const int j_first = 3 * r_first + c;
        const bool elected = cw == 0 && lane == 0;
        bool dead = false;
        int st = g % S;
        unsigned par = (unsigned)((g / S) & 1);
        int st_prev = -1, ext_prev_end = 0;
        for (int s = g; s < nsteps; s += 2) {
            const bool tr = pg.trace && blockIdx.x == pg.trace_cta && elected && s < 512;
            const unsigned char* stage = stages + st * stage_stride;
            StepPre<UPPER> p;
            p.n = 0; p.qbase = 0; p.ext_end = 0; p.ext_cnt = 0; p.on = false;
            if (!dead) {
                if (pipe_wait(&ctl->full[st], par, ctl, err)) p.load(stage, pg.rhs_bytes, r_first, j_first, lane_on);
                else dead = true;
            }
            if (tr) pg.trace[s * 16 + 0] = clock64();
            if (s > 0) asm volatile("bar.sync %0, %1;" ::"r"(2 - g), "n"(NPP) : "memory");     // step s-1 done
            if (tr) pg.trace[s * 16 + 1] = clock64() + (ctl->abort_flag == 12345);
            // pushed inputs of this step staged by the helper warp?  (ring data is written
            // before ext_ready, and shared-memory accesses of a thread are not reordered)
            if (!dead && p.ext_cnt > 0) {
                if (ctl->ext_ready < p.ext_end && !pipe_wait_ext(ctl, p.ext_end, err)) { dead = true; p.on = false; }
                asm volatile("" ::: "memory");
            }
            const double acc = sweep_row_chain<UPPER, LEAN>(p, stage + pg.rhs_bytes, r_first, c, rl, dep, work);
            if (!LEAN && p.n > kPipeRowsPerPass)
                sweep_extra_rows<UPPER>(stage, pg.rhs_bytes, p.n, r_first, c, rl, lane_on, dep, work, hand_off, out, pg.ext, w, scale);
            if (tr) { pg.trace[s * 16 + 2] = clock64(); pg.trace[s * 16 + 4] = p.n; }
            asm volatile("bar.arrive %0, %1;" ::"r"(1 + g), "n"(NPP) : "memory");             // step s done
            // every warp of this group passed the bar.sync of this step, i.e. is done with its
            // step s-2: release that stage and its pushed-result ring entries
            if (elected && st_prev >= 0) {
                ctl->ext_consumed = ext_prev_end;
                mbar_arrive(&ctl->empty[st_prev]);
            }
            sweep_row_stores<UPPER, LEAN>(p, stage + pg.rhs_bytes, r_first, c, acc, work, hand_off, out, pg.ext, w, scale);
            if (tr) pg.trace[s * 16 + 3] = clock64();
            st_prev = st; ext_prev_end = p.ext_end;
            st += 2;
            if (st >= S) { st -= S; par ^= 1u; }
        }
        // consume the other group's last hand-over so no barrier is left half-arrived
        if (((nsteps - 1) & 1) != g) asm volatile("bar.sync %0, %1;" ::"r"(2 - g), "n"(NPP) : "memory");
    }
}

// natural order -> program order (right-hand side of the lower sweep); perm_row < 0 is padding
__global__ void __launch_bounds__(256)
permute_rows_kernel(size_t nperm, const int* __restrict__ perm_row, const double* __restrict__ x,
                    double* __restrict__ xp)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nperm * 3) return;
    const size_t q = e / 3;
    const int row = perm_row[q];
    xp[e] = row >= 0 ? x[(size_t)row * 3 + (e - q * 3)] : 0.0;
}

}  // namespace opmgpu
