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
// Warp roles: warp 0 = TMA producer, warps 1..kPipeHelpers = pushed-result helpers, then kPipeGroups compute
// groups of kPipeComputeWarps warps (one thread per block row).  Every wait is bounded;
// on expiry the kernel raises *err and all roles drain (barrier hand-shakes keep running).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstddef>

#include "analysis.hpp"

namespace opmgpu {

#ifndef OPMGPU_PIPE_GROUPS
#define OPMGPU_PIPE_GROUPS 3
#endif
constexpr int kPipeGroups = OPMGPU_PIPE_GROUPS;                    // compute groups taking turns
constexpr int kPipeComputeWarps = 3;                               // warps per group, one thread per block row
constexpr int kPipeHelpers = 1;                                    // warps polling for pushed results
#ifndef OPMGPU_POLL_PER_LANE
#define OPMGPU_POLL_PER_LANE 1
#endif
#ifndef OPMGPU_SPIN_HANDOVER
#define OPMGPU_SPIN_HANDOVER 0
#endif
// 1: the groups hand the turn over through a shared-memory counter the waiting warps poll,
// instead of named barriers (experiment)
constexpr bool kPipeSpinHandover = OPMGPU_SPIN_HANDOVER != 0;
constexpr int kPipePollPerLane = OPMGPU_POLL_PER_LANE;             // slots a helper lane examines per poll (a short poll = a short crossing)
constexpr int kPipeThreads = 32 * (1 + kPipeHelpers + kPipeGroups * kPipeComputeWarps);
constexpr int kPipeRowsPerPass = kPipeComputeWarps * 32;
static_assert(kPipeRowsPerPass == kLeanStepRows, "analysis.hpp sizes lean steps for one pass");
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
    long long* gtrace;           // optional (debug): %globaltimer stamps of every CTA, [cta][gtrace_steps][8], then helper deliveries [cta][512][4]
    int gtrace_steps;
    int dbg;                     // experiments: bit 0 = skip the result stores to HBM
    int cluster_size;            // CTAs per thread-block cluster (1: none)
    int cx_bytes;                // intra-cluster result entries of a CTA (analysis.hpp: kCxBase), bytes
};
__device__ __forceinline__ long long pipe_gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// timer read that cannot issue before `dep` is available (stamps after barriers / loads / chains)
__device__ __forceinline__ long long pipe_gtime_after(int dep) { long long t; asm volatile("mov.u64 %0, %globaltimer; // %1" : "=l"(t) : "r"(dep)); return t; }
__device__ __forceinline__ long long pipe_clock_after(int dep) { long long t; asm volatile("mov.u64 %0, %clock64; // %1" : "=l"(t) : "r"(dep)); return t; }

struct PipeCtl {
    unsigned long long full[kPipeMaxStages];    // record + rhs landed (TMA complete_tx)
    unsigned long long empty[kPipeMaxStages];   // stage consumed
    volatile int ext_consumed;
    volatile int ext_ready;
    volatile int abort_flag;
    volatile int steps_done;                    // OPMGPU_SPIN_HANDOVER: compute warps that finished a step, all steps
};

// dependency array (window | pushed ring | zero entry) followed by the intra-cluster entries
__host__ __device__ inline size_t pipe_dep_bytes(int cx_bytes)
{
    return cx_bytes > 0 ? ((size_t)kCxBase * 24 + (size_t)cx_bytes + 127) / 128 * 128 : (size_t)kPipeDepBytes;
}
__host__ __device__ inline size_t pipe_smem_bytes(int nstages, int stage_bytes, int rhs_bytes, int cx_bytes = 0)
{
    return 512 + pipe_dep_bytes(cx_bytes) + (size_t)nstages * ((size_t)stage_bytes + rhs_bytes);
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
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s32_volatile(uint32_t a) { int v; asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

#ifndef OPMGPU_PUSH_MODE
#define OPMGPU_PUSH_MODE 0
#endif
__device__ __forceinline__ void push_f64(double* p, double v)
{
#if OPMGPU_PUSH_MODE == 0
    __stcg(p, v);
#elif OPMGPU_PUSH_MODE == 1
    atomicExch(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__double_as_longlong(v));
#else
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#endif
}

// Everything about a thread's block row r of a step that does not depend on earlier results;
// loaded from the landed stage into registers one step ahead.  One thread owns the whole row
// (all three components), so the chain needs no shuffles and its nine dependency loads are
// issued once per row, not once per component.
// T: arithmetic type of the sweep (double, or float for the reference's Impl<3,float>); records,
// window and push slots hold 8-byte containers either way (kernels.cuh: enc / dec).
template <bool UPPER, class T>
struct StepPre {
    int n, qbase, ext_end, ext_cnt;
    int4 ri0, ri1;              // rowinfo, dep0, dep1, dep2 | upos, push0, push1, own window slot
    uint32_t a0, a1, a2, aw;    // shared addresses of the three dependencies and of the own window slot
    bool on, has_cx;            // has_cx: some dependency arrives through distributed shared memory
    T rhs[3];
    T cf[27];                   // cf[c*9 + k*3 + e]
    T dv[9];                    // dv[c*3 + e] (upper)

    // stage = rhs area followed by the record
    __device__ __forceinline__ void load(const unsigned char* stage, int rhs_bytes, int r, bool lane_on, uint32_t dep_s = 0)
    {
        a0 = a1 = a2 = aw = dep_s; has_cx = false;
        const unsigned char* rec = stage + rhs_bytes;
        const int4 h0 = *reinterpret_cast<const int4*>(rec);
        n = h0.x; qbase = h0.y; ext_end = h0.z; ext_cnt = h0.w;
        on = lane_on && r < n;
        if (on) {
            const double* cfp = reinterpret_cast<const double*>(rec + 32) + r * 27;
#pragma unroll
            for (int q = 0; q < 27; ++q) cf[q] = dec<T>(cfp[q]);
            const int T3 = 3 * n;
            size_t off = 32 + (size_t)T3 * 72;
            if (UPPER) {
                const double* dp = reinterpret_cast<const double*>(rec + off) + r * 9;
#pragma unroll
                for (int q = 0; q < 9; ++q) dv[q] = dec<T>(dp[q]);
                off += (size_t)T3 * 24;
            }
            const int4* rip = reinterpret_cast<const int4*>(rec + ((off + 15) & ~(size_t)15)) + 2 * r;
            ri0 = rip[0]; ri1 = rip[1];
            // (off the critical path: the step's addresses are ready before its turn comes)
            a0 = dep_s + 8u * (uint32_t)ri0.y; a1 = dep_s + 8u * (uint32_t)ri0.z; a2 = dep_s + 8u * (uint32_t)ri0.w;
            aw = dep_s + 8u * (uint32_t)ri1.w;
            has_cx = max(ri0.y, max(ri0.z, ri0.w)) >= kCxBase * 3;
            const double* rp = reinterpret_cast<const double*>(stage) + 3 * r;
            rhs[0] = dec<T>(rp[0]); rhs[1] = dec<T>(rp[1]); rhs[2] = dec<T>(rp[2]);
        }
    }
};

__device__ __forceinline__ const double* dep_ptr(int code, const double* dep, const double* work)
{
    return code >= 0 ? dep + code : work + (size_t)(code & kDepValueMask) * 3;
}

// blocks beyond the three held in registers (rows with many couplings, e.g. well cells)
template <class T>
__device__ __noinline__ void sweep_tail_blocks(const unsigned char* rec, int r, const double* dep,
                                               const double* work, T (&acc)[3])
{
    const int* hdr = reinterpret_cast<const int*>(rec);
    const int n = hdr[0];
    const int* lists = reinterpret_cast<const int*>(rec + (size_t)hdr[5] * 8);
    const int* tail_end = lists;
    const int* tail_dep = lists + 2 * n;
    const double* tail_vals = reinterpret_cast<const double*>(rec + (size_t)hdr[6] * 8);
    for (int t = r ? tail_end[r - 1] : 0; t < tail_end[r]; ++t) {
        const double* yp = dep_ptr(tail_dep[t], dep, work);
        const T y0 = dec<T>(yp[0]), y1 = dec<T>(yp[1]), y2 = dec<T>(yp[2]);
        const double* ap = tail_vals + (size_t)t * 9;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            acc[c] = fma(-dec<T>(ap[c * 3 + 0]), y0, acc[c]);
            acc[c] = fma(-dec<T>(ap[c * 3 + 1]), y1, acc[c]);
            acc[c] = fma(-dec<T>(ap[c * 3 + 2]), y2, acc[c]);
        }
    }
}
// pushes beyond the two held in registers
template <class T>
__device__ __noinline__ void sweep_extra_pushes(const unsigned char* rec, int r, double* ext, const T (&acc)[3])
{
    const int* hdr = reinterpret_cast<const int*>(rec);
    const int n = hdr[0];
    const int* lists = reinterpret_cast<const int*>(rec + (size_t)hdr[5] * 8);
    const int* xpush_end = lists + n;
    const int* xpush_slot = lists + 2 * n + hdr[4];
    for (int t = r ? xpush_end[r - 1] : 0; t < xpush_end[r]; ++t) {
        double* sl = ext + (size_t)xpush_slot[t] * 3;
        __stcg(sl, enc(acc[0])); __stcg(sl + 1, enc(acc[1])); __stcg(sl + 2, enc(acc[2]));
    }
}

// one block row, critical part: dependency loads, three 9-FMA chains in the reference's order
// (bit parity), Dinv (upper), window store, pushes to other CTAs
// an entry written by another CTA of the cluster (self-validating: all-ones = not there yet)
__device__ __forceinline__ bool cx_wait(uint32_t a, double& y0, double& y1, double& y2, PipeCtl* ctl, int* err)
{
    unsigned spins = 0;
    while (__double_as_longlong(y0) == -1LL || __double_as_longlong(y1) == -1LL || __double_as_longlong(y2) == -1LL) {
        if (++spins > kPipeSpinLimit * 8u || ctl->abort_flag) { ctl->abort_flag = 1; atomicExch(err, 8); return false; }
        y0 = lds_f64(a); y1 = lds_f64(a + 8); y2 = lds_f64(a + 16);
    }
    return true;
}

#ifndef OPMGPU_CX_UNIFORM
#define OPMGPU_CX_UNIFORM 0
#endif
// 1 (experiment): results delivered through distributed shared memory are awaited by the whole
// warp (vote) before the dependency loads, not by the lanes that need them afterwards.  A lone
// spinning lane costs ~1000 cycles per hand-over against 200-300 for a warp-uniform poll
// (tools/ubench/xwarp*.cu), but in the sweeps the entries are usually there already and the extra
// vote + nine 4-byte loads per step cost more: 100^3 apply 282 us (0) against 293 us (1), round 2.
constexpr bool kCxUniformWait = OPMGPU_CX_UNIFORM != 0;

__device__ __forceinline__ bool cx_entry_there(uint32_t a)
{
    int h0, h1, h2;
    asm volatile("ld.volatile.shared.s32 %0, [%1+4];" : "=r"(h0) : "r"(a) : "memory");
    asm volatile("ld.volatile.shared.s32 %0, [%1+12];" : "=r"(h1) : "r"(a) : "memory");
    asm volatile("ld.volatile.shared.s32 %0, [%1+20];" : "=r"(h2) : "r"(a) : "memory");
    return (h0 != -1) & (h1 != -1) & (h2 != -1);      // a result never has an all-ones upper half (that is a NaN no arithmetic produces)
}

template <bool UPPER, bool LEAN, bool CX = false, class T = double>
__device__ __forceinline__ void sweep_row_chain(const StepPre<UPPER, T>& p, const unsigned char* rec, int r,
                                                double* dep, uint32_t dep_s, const double* work, double* ext, T (&acc)[3],
                                                PipeCtl* ctl = nullptr, int* err = nullptr)
{
    if (CX && kCxUniformWait) {
        // (all lanes of the warp are here: the turn barrier and the pushed-input check are warp-uniform)
        const bool mine = p.on && p.has_cx;
        if (__any_sync(0xffffffffu, mine)) {
            unsigned spins = 0;
            for (;;) {
                bool missing = false;
                if (mine) {
                    if (p.ri0.y >= kCxBase * 3) missing |= !cx_entry_there(p.a0);
                    if (p.ri0.z >= kCxBase * 3) missing |= !cx_entry_there(p.a1);
                    if (p.ri0.w >= kCxBase * 3) missing |= !cx_entry_there(p.a2);
                }
                if (!__any_sync(0xffffffffu, missing)) break;
                if (++spins > kPipeSpinLimit * 8u || ctl->abort_flag) { ctl->abort_flag = 1; atomicExch(err, 8); break; }
            }
        }
    }
    if (!p.on) { acc[0] = acc[1] = acc[2] = T(0); }
    if (p.on) {
        double y[9];        // containers as loaded (validity of cluster-delivered entries is a bit pattern)
        if (LEAN || (p.ri0.y | p.ri0.z | p.ri0.w) >= 0) {    // all three in shared memory (the common case)
            const uint32_t a0 = LEAN ? p.a0 : dep_s + 8u * (uint32_t)p.ri0.y, a1 = LEAN ? p.a1 : dep_s + 8u * (uint32_t)p.ri0.z,
                           a2 = LEAN ? p.a2 : dep_s + 8u * (uint32_t)p.ri0.w;
            y[0] = lds_f64(a0); y[1] = lds_f64(a0 + 8); y[2] = lds_f64(a0 + 16);
            y[3] = lds_f64(a1); y[4] = lds_f64(a1 + 8); y[5] = lds_f64(a1 + 16);
            y[6] = lds_f64(a2); y[7] = lds_f64(a2 + 8); y[8] = lds_f64(a2 + 16);
            if (CX && !kCxUniformWait && p.has_cx) {      // delivered by a CTA of the cluster?
                if (p.ri0.y >= kCxBase * 3) cx_wait(a0, y[0], y[1], y[2], ctl, err);
                if (p.ri0.z >= kCxBase * 3) cx_wait(a1, y[3], y[4], y[5], ctl, err);
                if (p.ri0.w >= kCxBase * 3) cx_wait(a2, y[6], y[7], y[8], ctl, err);
            }
        } else {
            const double* y0 = dep_ptr(p.ri0.y, dep, work); const double* y1 = dep_ptr(p.ri0.z, dep, work);
            const double* y2 = dep_ptr(p.ri0.w, dep, work);
            y[0] = y0[0]; y[1] = y0[1]; y[2] = y0[2];
            y[3] = y1[0]; y[4] = y1[1]; y[5] = y1[2];
            y[6] = y2[0]; y[7] = y2[1]; y[8] = y2[2];
        }
        acc[0] = p.rhs[0]; acc[1] = p.rhs[1]; acc[2] = p.rhs[2];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const T yq = dec<T>(y[q]);
            acc[0] = fma(-p.cf[q], yq, acc[0]);
            acc[1] = fma(-p.cf[9 + q], yq, acc[1]);
            acc[2] = fma(-p.cf[18 + q], yq, acc[2]);
        }
        if (!LEAN && (p.ri0.x & kRowSlow)) sweep_tail_blocks<T>(rec, r, dep, work, acc);
        if (UPPER) {
            T v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T t = T(0);
                t = fma(p.dv[c * 3 + 0], acc[0], t); t = fma(p.dv[c * 3 + 1], acc[1], t); t = fma(p.dv[c * 3 + 2], acc[2], t);
                v[c] = t;
            }
            acc[0] = v[0]; acc[1] = v[1]; acc[2] = v[2];
        }
        const uint32_t w = LEAN ? p.aw : dep_s + 8u * (uint32_t)p.ri1.w;
        sts_f64(w, enc(acc[0])); sts_f64(w + 8, enc(acc[1])); sts_f64(w + 16, enc(acc[2]));
    }
}
// results other CTAs wait for.  Issued right after the hand-over to the next group: the
// in-tile hand-over is on the critical path of every step, a tile crossing only once per tile.
// result -> entry `id & 0xfffff` of CTA `(id >> 20) & 15` of this cluster (distributed shared memory)
template <class T>
__device__ __forceinline__ void push_dsmem(uint32_t cx_s, int id, const T (&acc)[3])
{
    const uint32_t la = cx_s + 24u * (uint32_t)(id & 0xfffff);
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"((id >> 20) & 0xf));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(enc(acc[0])) : "memory");
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra + 8), "d"(enc(acc[1])) : "memory");
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra + 16), "d"(enc(acc[2])) : "memory");
}

template <bool UPPER, bool LEAN, bool CX = false, class T = double>
__device__ __forceinline__ void sweep_row_pushes(const StepPre<UPPER, T>& p, const unsigned char* rec, int r,
                                                 double* ext, const T (&acc)[3], uint32_t cx_s = 0)
{
    if (p.on) {
        if (CX && p.ri1.y >= 0 && (p.ri1.y & kPushDsmem)) push_dsmem<T>(cx_s, p.ri1.y, acc);
        else if (p.ri1.y >= 0) { double* sl = ext + (size_t)p.ri1.y * 3; push_f64(sl, enc(acc[0])); push_f64(sl + 1, enc(acc[1])); push_f64(sl + 2, enc(acc[2])); }
        if (CX && p.ri1.z >= 0 && (p.ri1.z & kPushDsmem)) push_dsmem<T>(cx_s, p.ri1.z, acc);
        else if (p.ri1.z >= 0) { double* sl = ext + (size_t)p.ri1.z * 3; push_f64(sl, enc(acc[0])); push_f64(sl + 1, enc(acc[1])); push_f64(sl + 2, enc(acc[2])); }
        if (!LEAN && (p.ri0.x & kRowSlow)) sweep_extra_pushes<T>(rec, r, ext, acc);
    }
}
// ... and the part nobody waits for: results to HBM
template <bool UPPER, bool LEAN, class T>
__device__ __forceinline__ void sweep_row_stores(const StepPre<UPPER, T>& p, const T (&acc)[3],
                                                 double* work, double* hand_off, T* out, T w, int scale)
{
    if (p.on) {
        const int row = p.ri0.x & kRowMask;
        if (UPPER) {
            T* o = out + (size_t)row * 3;
            o[0] = scale ? acc[0] * w : acc[0]; o[1] = scale ? acc[1] * w : acc[1]; o[2] = scale ? acc[2] * w : acc[2];
        } else {
            double* o = hand_off + (size_t)p.ri1.x * 3;
            o[0] = enc(acc[0]); o[1] = enc(acc[1]); o[2] = enc(acc[2]);
        }
        if (!LEAN && (p.ri0.x & kRowWriteGlobal)) {
            double* o = work + (size_t)row * 3;
            o[0] = enc(acc[0]); o[1] = enc(acc[1]); o[2] = enc(acc[2]);
        }
    }
}

// rows [kPipeRowsPerPass, n) of a step wider than one pass over the compute warps
template <bool UPPER, class T>
__device__ __noinline__ void sweep_extra_rows(const unsigned char* stage, int rhs_bytes, int n, int r_first,
                                              double* dep, uint32_t dep_s, double* work, double* hand_off,
                                              T* out, double* ext, T w, int scale)
{
    for (int rbase = kPipeRowsPerPass; rbase < n; rbase += kPipeRowsPerPass) {
        const int r = rbase + r_first;
        StepPre<UPPER, T> p;
        p.load(stage, rhs_bytes, r, true);
        T acc[3];
        sweep_row_chain<UPPER, false, false, T>(p, stage + rhs_bytes, r, dep, dep_s, work, ext, acc);
        sweep_row_pushes<UPPER, false, false, T>(p, stage + rhs_bytes, r, ext, acc);
        sweep_row_stores<UPPER, false, T>(p, acc, work, hand_off, out, w, scale);
    }
}

// LEAN: the program has no slow rows, no own-result reads from HBM and no step wider than one
// pass (every Cartesian stencil case): those paths are compiled out.
// CX: launched in thread-block clusters; results for CTAs of the same cluster are stored straight
// into their shared memory (a tile crossing then costs a shared-memory round trip, not an L2 poll).
// TRACE: the per-step time stamps of the debug tools (tools/trace_sweep.py, gtrace_sweep.py) are
// compiled in; the production variants carry none of that code on the critical path.
template <bool UPPER, bool LEAN, bool CX = false, bool TRACE = false, class T = double>
__global__ void __launch_bounds__(kPipeThreads, 1)
ilu0_sweep_pipe_kernel(PipeDev pg, const double* __restrict__ rhs_perm, double* work, double* hand_off,
                       T* out, double w_, int scale, int* err)
{
    const T w = (T)w_;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeCtl* ctl = reinterpret_cast<PipeCtl*>(smem_raw);
    double* dep = reinterpret_cast<double*>(smem_raw + 512);      // window | pushed ring | zero entry
    unsigned char* stages = smem_raw + 512 + pipe_dep_bytes(CX ? pg.cx_bytes : 0);
    const int S = pg.nstages;
    const size_t stage_stride = (size_t)pg.stage_bytes + pg.rhs_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = pg.cta_step_ptr[blockIdx.x];
    const int nsteps = pg.cta_step_ptr[blockIdx.x + 1] - s0;
    long long* gtr = TRACE && pg.gtrace ? pg.gtrace + (size_t)blockIdx.x * pg.gtrace_steps * 8 : nullptr;
    long long* hlog = TRACE && pg.gtrace ? pg.gtrace + (size_t)gridDim.x * pg.gtrace_steps * 8 + (size_t)blockIdx.x * 2048 : nullptr;
    if (TRACE && gtr && tid == 0) gtr[(pg.gtrace_steps - 1) * 8 + 0] = pipe_gtime();

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], kPipeComputeWarps); }
        ctl->ext_consumed = 0; ctl->ext_ready = 0; ctl->abort_flag = 0; ctl->steps_done = 0;
        dep[kDepZeroSlot * 3] = 0.0; dep[kDepZeroSlot * 3 + 1] = 0.0; dep[kDepZeroSlot * 3 + 2] = 0.0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CX) {
        // intra-cluster entries start out empty; nobody may write into them before that
        long long* cx = reinterpret_cast<long long*>(dep + (size_t)kCxBase * 3);
        for (int i = tid; i < pg.cx_bytes / 8; i += kPipeThreads) cx[i] = -1LL;
        __syncthreads();
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    __syncthreads();
    if (!CX && nsteps == 0) return;

    if (nsteps == 0) {
        // (cluster launch) nothing to do, but stay until the cluster is done
    } else if (warp == 0) {
        // ------------------------------------------------ TMA producer
        // (the warp runs the loop uniformly: the descriptors of the NEXT step are fetched while this
        // one is issued, and the step's two bulk copies -- record and right-hand side -- leave as one
        // warp instruction, lanes 0 and 1 with their own operands)
        unsigned off16n = 0, bytesn = 0, rrown = 0, rbytesn = 0;
        if (nsteps > 0) { off16n = pg.step_off16[s0]; bytesn = pg.step_bytes[s0]; rrown = pg.step_rhs_row[s0]; rbytesn = pg.step_rhs_bytes[s0]; }
        for (int i = 0; i < nsteps; ++i) {
            const int st = i % S, k = i / S;
            const unsigned off16 = off16n, bytes = bytesn, rrow = rrown, rbytes = rbytesn;
            if (i + 1 < nsteps) {
                off16n = pg.step_off16[s0 + i + 1]; bytesn = pg.step_bytes[s0 + i + 1];
                rrown = pg.step_rhs_row[s0 + i + 1]; rbytesn = pg.step_rhs_bytes[s0 + i + 1];
            }
            if (k > 0 && !pipe_wait(&ctl->empty[st], (unsigned)((k - 1) & 1), ctl, err)) break;
            unsigned char* stage = stages + st * stage_stride;
            if (lane == 0) mbar_arrive_expect_tx(&ctl->full[st], bytes + rbytes);
            __syncwarp();
            void* dst = lane == 0 ? (void*)(stage + pg.rhs_bytes) : (void*)stage;
            const void* src = lane == 0 ? (const void*)(pg.buf + (size_t)off16 * 16) : (const void*)(rhs_perm + (size_t)rrow * 3);
            const unsigned nb = lane == 0 ? bytes : rbytes;
            if (lane < 2 && nb > 0) tma_bulk_g2s(dst, src, nb, &ctl->full[st]);
        }
    } else if (warp <= kPipeHelpers) {
        // ------------------------------------------------ pushed-result helpers
        // A poll reads the next 96 slots in consumption order and delivers the valid prefix:
        // values into the shared-memory ring, then ext_ready, then the slots are re-armed.  One
        // poll is a full L2 round trip (1300+ cycles while the record stream is running), so
        // kPipeHelpers warps poll the same window out of phase.  They do not coordinate beyond
        // ext_ready (atomicMax): two helpers may deliver the same entry, which writes the same
        // value to the same ring position twice.
        const int hid = warp - 1;
        const long long base = pg.cta_ext_base[blockIdx.x];
        const int total = (int)(pg.cta_ext_base[blockIdx.x + 1] - base);
        const long long* slots = reinterpret_cast<const long long*>(pg.ext) + (size_t)base * 3;
        double* ring = dep + kWindowRows * 3;
        int* ext_ready = const_cast<int*>(&ctl->ext_ready);
        unsigned spins = 0;
        long long polls = 0;
        int ndeliv = 0;
        // The slots were last touched a whole sweep ago and may have left L2: keep the lines
        // ahead of the poll window warm (helper 0).
        constexpr int kPfAhead = 768;
        int pf = 0;                            // slots [0, pf) have been prefetched
        int rearm_from = 0, rearm_n = 0;       // delivered, not yet re-armed
        if (hid > 0) __nanosleep(300 * hid);
        for (;;) {
            const int e = ctl->ext_ready;
            if (e >= total) break;
            ++polls;
            if (hid == 0 && pf < min(total, e + kPfAhead)) {               // one 128-byte line per lane
                const int want = min(total, e + kPfAhead);
                const char* p0 = reinterpret_cast<const char*>(slots + (size_t)pf * 3);
                const char* p1 = reinterpret_cast<const char*>(slots + (size_t)want * 3);
                const char* line = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p0) & ~(uintptr_t)127) + (size_t)lane * 128;
                if (line < p1) asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
                pf = min(want, pf + (32 * 128) / 24);
            }
            const int limit = min(total, ctl->ext_consumed + kExtRing);
            long long a[kPipePollPerLane][3];
            unsigned m[kPipePollPerLane];
#pragma unroll
            for (int u = 0; u < kPipePollPerLane; ++u) {
                const int idx = e + u * 32 + lane;
                a[u][0] = a[u][1] = a[u][2] = -1;
                if (idx < limit) {
                    const long long* sl = slots + (size_t)idx * 3;
                    asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(a[u][0]) : "l"(sl) : "memory");
                    asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(a[u][1]) : "l"(sl + 1) : "memory");
                    asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(a[u][2]) : "l"(sl + 2) : "memory");
                }
            }
            // re-arm the slots of the previous delivery while this poll's loads are in flight
            // (nobody waits for these stores; the producer writes them again one sweep later)
            if (rearm_n > 0) {
#pragma unroll
                for (int u = 0; u < kPipePollPerLane; ++u) {
                    const int idx = rearm_from + u * 32 + lane;
                    if (u * 32 + lane < rearm_n) {
                        long long* sl = const_cast<long long*>(slots) + (size_t)idx * 3;
                        __stcg(sl + 0, -1LL); __stcg(sl + 1, -1LL); __stcg(sl + 2, -1LL);
                    }
                }
                rearm_n = 0;
            }
#pragma unroll
            for (int u = 0; u < kPipePollPerLane; ++u)
                m[u] = __ballot_sync(0xffffffffu, a[u][0] != -1 && a[u][1] != -1 && a[u][2] != -1);
            int n = 32 * kPipePollPerLane;
#pragma unroll
            for (int u = kPipePollPerLane - 1; u >= 0; --u)
                if (m[u] != 0xffffffffu) n = 32 * u + __ffs(~m[u]) - 1;
            // another helper may have delivered part of the prefix meanwhile
            const int e_now = kPipeHelpers > 1 ? __shfl_sync(0xffffffffu, (int)ctl->ext_ready, 0) : e;
            if (e + n > e_now) {
#pragma unroll
                for (int u = 0; u < kPipePollPerLane; ++u) {
                    const int idx = e + u * 32 + lane;
                    if (idx >= e_now && u * 32 + lane < n) {
                        double* dst = ring + (idx & (kExtRing - 1)) * 3;
                        dst[0] = __longlong_as_double(a[u][0]); dst[1] = __longlong_as_double(a[u][1]); dst[2] = __longlong_as_double(a[u][2]);
                    }
                }
                // ring data before ext_ready: shared-memory stores of one warp are performed in order
                __syncwarp();
                if (lane == 0) { if (kPipeHelpers > 1) atomicMax(ext_ready, e + n); else *reinterpret_cast<volatile int*>(ext_ready) = e + n; }
                rearm_from = e_now; rearm_n = e + n - e_now;
                if (TRACE && gtr && lane == 0 && hid == 0) {
                    if (e_now == 0) { gtr[(pg.gtrace_steps - 1) * 8 + 1] = pipe_gtime(); gtr[(pg.gtrace_steps - 1) * 8 + 2] = polls; }
                    if (ndeliv < 512) {
                        hlog[4 * ndeliv] = pipe_gtime(); hlog[4 * ndeliv + 1] = ((long long)polls << 32) | (unsigned)(e + n);
                        hlog[4 * ndeliv + 2] = 0; hlog[4 * ndeliv + 3] = e + n - e_now;
                        ++ndeliv;
                    }
                }
                spins = 0;
            } else {
                ++spins;                                              // n is warp-uniform, so is spins
                const int ab = __shfl_sync(0xffffffffu, (int)ctl->abort_flag, 0);
                if (ab) break;
                if (spins > kPipeSpinLimit) { if (lane == 0) { ctl->abort_flag = 1; atomicExch(err, 3); } break; }
            }
        }
        if (rearm_n > 0) {
#pragma unroll
            for (int u = 0; u < kPipePollPerLane; ++u) {
                const int idx = rearm_from + u * 32 + lane;
                if (u * 32 + lane < rearm_n) {
                    long long* sl = const_cast<long long*>(slots) + (size_t)idx * 3;
                    __stcg(sl + 0, -1LL); __stcg(sl + 1, -1LL); __stcg(sl + 2, -1LL);
                }
            }
        }
    } else {
        // ------------------------------------------------ compute warps (kPipeGroups groups in turn)
        // lane = row of the step (kPipeRowsPerPass rows per pass).  Group g owns the steps
        // s = g, g+G, ...; it signals "step s done" on named barrier 1+g (bar.arrive) and waits
        // for "step s-1 done" on the barrier of the group before it (bar.sync).
        constexpr int G = kPipeGroups;
        constexpr int NPP = 2 * kPipeComputeWarps * 32;
        const int g = (warp - 1 - kPipeHelpers) / kPipeComputeWarps;
        const int cw = (warp - 1 - kPipeHelpers) - g * kPipeComputeWarps;
        const int r_first = cw * 32 + lane;
        const bool elected = cw == 0 && lane == 0;
        const int bar_prev = 1 + (g + G - 1) % G, bar_own = 1 + g;
        // 32-bit shared addresses held in registers (opaque to the compiler, which otherwise
        // rebuilds them from SR_CgaCtaId on the critical path of every step)
        uint32_t dep_s = smem_u32(dep), ctl_s = smem_u32(ctl);
        asm volatile("" : "+r"(dep_s), "+r"(ctl_s));
        bool dead = false;
        int st = g % S;
        unsigned par = (unsigned)((g / S) & 1);
        int st_prev = -1, ext_prev_end = 0;
        for (int s = g; s < nsteps; s += G) {
            const bool tr = TRACE && pg.trace && blockIdx.x == pg.trace_cta && elected && s < 512;
            const unsigned char* stage = stages + st * stage_stride;
            StepPre<UPPER, T> p;
            p.n = 0; p.qbase = 0; p.ext_end = 0; p.ext_cnt = 0; p.on = false; p.has_cx = false; p.a0 = p.a1 = p.a2 = p.aw = dep_s;
            if (!dead) {
                if (pipe_wait(&ctl->full[st], par, ctl, err)) p.load(stage, pg.rhs_bytes, r_first, true, dep_s);
                else dead = true;
            }
            if (LEAN) {             // everything the step needs is in registers now: hand the stage back
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->empty[st]);
            }
            if (tr) pg.trace[s * 16 + 0] = clock64();
            const bool gt = TRACE && gtr && elected && s < pg.gtrace_steps - 1;
            if (gt) gtr[s * 8 + 0] = pipe_gtime();
            // pushed inputs of this step staged by the helper warp?  ext_ready only grows, so the
            // usual answer (yes) is fetched while the group still waits for its turn.  (Ring data
            // is written before ext_ready, and shared-memory accesses of a thread are not reordered.)
            const bool ext_ok = dead || p.ext_cnt <= 0 || lds_s32_volatile(ctl_s + (uint32_t)offsetof(PipeCtl, ext_ready)) >= p.ext_end;
            if (kPipeSpinHandover) {                                                             // step s-1 done
                const int need = kPipeComputeWarps * s;
                while (lds_s32_volatile(ctl_s + (uint32_t)offsetof(PipeCtl, steps_done)) < need) { }
            } else if (s > 0) asm volatile("bar.sync %0, %1;" ::"r"(bar_prev), "n"(NPP) : "memory");
            if (tr) pg.trace[s * 16 + 1] = pipe_clock_after(ctl->abort_flag);
            if (gt) gtr[s * 8 + 1] = pipe_gtime_after(ctl->abort_flag);
            if (!ext_ok) {
                if (!pipe_wait_ext(ctl, p.ext_end, err)) { dead = true; p.on = false; }
                asm volatile("" ::: "memory");
            }
            if (gt) { gtr[s * 8 + 2] = pipe_gtime_after(ctl->ext_ready); gtr[s * 8 + 4] = p.ext_end; gtr[s * 8 + 5] = p.n; }
            if (tr) pg.trace[s * 16 + 5] = pipe_clock_after(ctl->ext_ready);
            T acc[3];
            sweep_row_chain<UPPER, LEAN, CX, T>(p, stage + pg.rhs_bytes, r_first, dep, dep_s, work, pg.ext, acc, ctl, err);
            if (!LEAN && p.n > kPipeRowsPerPass)
                sweep_extra_rows<UPPER, T>(stage, pg.rhs_bytes, p.n, r_first, dep, dep_s, work, hand_off, out, pg.ext, w, scale);
            if (tr) { pg.trace[s * 16 + 2] = pipe_clock_after(__double2hiint(enc(acc[0]))); pg.trace[s * 16 + 4] = p.n; }
            if (gt) gtr[s * 8 + 3] = pipe_gtime_after(__double2hiint(enc(acc[0])));
            long long rb = 0;           // trace: read the first pushed value back through L2
            const bool gt_rb = gt && p.on && p.ri1.y >= 0;
            if (gt_rb) asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(rb) : "l"(pg.ext + (size_t)p.ri1.y * 3) : "memory");
            if (kPipeSpinHandover) {                                                               // step s done
                __syncwarp();
                if (lane == 0) asm volatile("red.shared.add.s32 [%0], 1;" ::"r"(ctl_s + (uint32_t)offsetof(PipeCtl, steps_done)) : "memory");
            } else asm volatile("bar.arrive %0, %1;" ::"r"(bar_own), "n"(NPP) : "memory");
            sweep_row_pushes<UPPER, LEAN, CX, T>(p, stage + pg.rhs_bytes, r_first, pg.ext, acc, dep_s + 24u * (uint32_t)kCxBase);
            if (!LEAN) {            // tail lists of slow rows are read from the stage during the chain
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->empty[st]);
            }
            // every warp of this group passed the bar.sync of this step, i.e. is done with its
            // step s-G: release the pushed-result ring entries of that step
            if (elected && st_prev >= 0) ctl->ext_consumed = ext_prev_end;
            if (!(pg.dbg & 1)) sweep_row_stores<UPPER, LEAN, T>(p, acc, work, hand_off, out, w, scale);
            if (gt_rb) gtr[s * 8 + 6] = pipe_gtime_after((int)rb);
            if (tr) pg.trace[s * 16 + 3] = clock64();
            st_prev = st; ext_prev_end = p.ext_end;
            st += G;
            if (st >= S) { st -= S; par ^= 1u; }
        }
        // consume the last hand-over addressed to this group so no barrier is left half-arrived
        if (!kPipeSpinHandover && nsteps > 0 && (nsteps % G) == g) asm volatile("bar.sync %0, %1;" ::"r"(bar_prev), "n"(NPP) : "memory");
    }
    if (CX) {
        // other CTAs of the cluster may still be writing into this CTA's shared memory
        asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
    }
}

// natural order -> program order (right-hand side of the lower sweep); perm_row < 0 is padding
template <class T>
__global__ void __launch_bounds__(256)
permute_rows_kernel(size_t nperm, const int* __restrict__ perm_row, const T* __restrict__ x,
                    double* __restrict__ xp)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nperm * 3) return;
    const size_t q = e / 3;
    const int row = perm_row[q];
    xp[e] = row >= 0 ? enc(x[(size_t)row * 3 + (e - q * 3)]) : 0.0;
}

}  // namespace opmgpu
