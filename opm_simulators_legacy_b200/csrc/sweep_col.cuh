// K4 (Cartesian fast path)  Column-owned wavefront ILU0 sweeps for sm_100a.
//
//   v = w U^-1 L^-1 d      (Opm::ParallelOverlappingILU0::apply; natural-order block ILU0,
//                           call site opm/autodiff/ISTLSolver.hpp:201-211)
//
// See colprog.hpp for the program.  One lane owns one (i,j) column and walks it along k; per
// step the only dependent work is: two neighbour results by shuffle (or a shared-memory ring
// entry at a patch edge) -> three 9-FMA chains in the reference's order (bit parity) -> [upper:
// inverted diagonal] -> next step.  There is no CTA-wide barrier on that path: warps of a tile
// run as a dataflow pipeline, coupled only through self-validating ring entries.
//
// Warp roles: warps 0..W-1 = compute (one patch each), warp W = record producer (lane w feeds
// compute warp w's stage ring with bulk async copies, cp.async.bulk + mbarrier complete_tx; issuing
// a bulk copy costs ~100 instructions of uniform-register set-up, far too many for the compute
// warps), warp W+1 = helper that stages results pushed by other tiles (L2 slots) into the rings.
// Every spin loop is warp-uniform (a lone spinning lane beside lanes parked at a
// reconvergence point costs ~1000 cycles per hand-over, tools/ubench/xwarp.cu) and bounded; on
// expiry the kernel raises *err and all roles drain.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "colprog.hpp"
#include "sweep_pipe.cuh"

namespace opmgpu {

constexpr int kColThreads = 32 * (kColMaxWarps + 2);
constexpr unsigned kColSpinLimit = 1u << 22;
constexpr int kColHelperBatch = 8;                  // L2 slots a helper lane examines per poll
static_assert(kColHelperBatch <= kColRing, "a batch must fit the ring");
static_assert((kColRing & (kColRing - 1)) == 0, "ring index is k & (R-1)");

struct ColDev {
    ColGeom g;
    const double* rec;           // record stream of this sweep
    const int* cta_tile_ptr;
    const int* cta_tiles;        // tiles of a CTA in processing order
    double* ext;                 // push slots [tile][edge][nz][3], all-ones when empty
    int nstages;
    int pf_ahead;                // records the producer keeps ahead in L2 (0: none)
    int producer_warp, helper_warp;   // warp indices of the two service roles (compute warps are 0..W-1)
    long long* prof;             // debug (OPMGPU_COL_PROF): per CTA and warp {cycles waiting for records, cycles in the ring loop, cycles total, steps}
    long long* trace;            // debug (OPMGPU_COL_TRACE=cta): clock64 stamps of warp 0 of that CTA, [T][8]
    int trace_cta;
};

__device__ __forceinline__ bool col_valid3(double a, double b, double c)
{
    return __double_as_longlong(a) != -1LL && __double_as_longlong(b) != -1LL && __double_as_longlong(c) != -1LL;
}
__device__ __forceinline__ bool col_empty3(double a, double b, double c)
{
    return __double_as_longlong(a) == -1LL && __double_as_longlong(b) == -1LL && __double_as_longlong(c) == -1LL;
}
__device__ __forceinline__ double col_lds(uint32_t a) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void col_sts(uint32_t a, double v) { asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

__device__ __forceinline__ void col_prefetch_l2(const void* p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool col_warp_has_patch(const ColGeom& g, int tile, int warp)
{
    const int pa = (tile % g.nta) * g.ta + warp % g.ta, pb = (tile / g.nta) * g.tb + warp / g.ta;
    return pa < g.npa && pb < g.npb;
}

// predicated stores: one instruction each, no branch (the step loop must stay straight-line code)
__device__ __forceinline__ void col_sts_if(bool p, uint32_t a, double v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.f64 [%0], %1;\n\t}" ::"r"(a), "d"(v), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void col_sts2_if(bool p, uint32_t a, double v0, double v1)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.shared.v2.f64 [%0], {%1, %2};\n\t}" ::"r"(a), "d"(v0), "d"(v1), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ int col_lds_hi_vol(uint32_t a) { int v; asm volatile("ld.volatile.shared.s32 %0, [%1+4];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int col_lds_hi(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1+4];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void col_stg_if(bool p, double* a, double v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f64 [%0], %1;\n\t}" ::"l"(a), "d"(v), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void col_stcg_if(bool p, double* a, double v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.cg.f64 [%0], %1;\n\t}" ::"l"(a), "d"(v), "r"((unsigned)p) : "memory");
}
// record loads: plain shared-memory loads, pinned (asm volatile) after the step's shuffles
__device__ __forceinline__ double col_lds_var(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ bool col_try_wait(uint32_t bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// record of one lane-step, in registers
template <bool UPPER>
struct ColRegs {
    double rhs[3];
    double cf[UPPER ? kColNCU : kColNCL];
};
__device__ __forceinline__ bool col_test_wait(uint32_t bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void col_sts32_if(bool p, uint32_t a, int v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.volatile.shared.s32 [%0], %1;\n\t}" ::"r"(a), "r"(v), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void col_lds2(uint32_t a, double& x, double& y)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
#define COL_LOAD_REC(R_, stage_lane_s, stage_rhs_s)                                                       \
    do {                                                                                                  \
        (R_).rhs[0] = col_lds_var(stage_rhs_s); (R_).rhs[1] = col_lds_var((stage_rhs_s) + 8u); (R_).rhs[2] = col_lds_var((stage_rhs_s) + 16u); \
        _Pragma("unroll") for (int q_ = 0; q_ < (UPPER ? kColNCU : kColNCL); q_ += 2) col_lds2((stage_lane_s) + 256u * q_, (R_).cf[q_], (R_).cf[q_ + 1]); \
    } while (0)

constexpr int kColFixedSmem = 2048;     // barriers | abort flag | always-valid zero ring | always-empty ring (see col_smem_fixed)

// PROF: cycle counters / per-step stamps of the debug tools are compiled in
template <bool UPPER, bool PROF>
__global__ void __launch_bounds__(kColThreads, 1)
ilu0_sweep_col_kernel(ColDev pg, const double* __restrict__ rhs_perm, double* __restrict__ hand_off,
                      double* __restrict__ out, double w, int scale, int* err)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const ColGeom& g = pg.g;
    constexpr int NC = UPPER ? kColNCU : kColNCL;
    constexpr int R = kColRing;
    constexpr unsigned FULLM = 0xffffffffu;
    const int W = g.W, S = pg.nstages;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw);             // [W][S]
    volatile int* abort_flag = reinterpret_cast<volatile int*>(smem_raw + 768);
    volatile int* prog = reinterpret_cast<volatile int*>(smem_raw + 800);                   // [W] records used up by each compute warp
    double* zero_ring = reinterpret_cast<double*>(smem_raw + 1024);                         // [R][4], all 0: "no neighbour"
    double* none_ring = reinterpret_cast<double*>(smem_raw + 1536);                         // [R][4], all-ones: "no consumer"
    double* ring_i = reinterpret_cast<double*>(smem_raw + kColFixedSmem);                   // [W][ph][R][4]
    double* ring_j = ring_i + (size_t)W * g.ph * R * 4;                                     // [W][pw][R][4]
    unsigned char* stages = reinterpret_cast<unsigned char*>(ring_j + (size_t)W * g.pw * R * 4);
    constexpr unsigned stage_bytes = 768 + NC * 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q_begin = pg.cta_tile_ptr[blockIdx.x], q_end = pg.cta_tile_ptr[blockIdx.x + 1];
    const double EMPTY = __longlong_as_double(-1LL);

    if (tid == 0) {
        for (int i = 0; i < W * S; ++i) mbar_init(&full[i], 1);
        *abort_flag = 0;
        for (int i = 0; i < W; ++i) prog[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < R * 4; i += blockDim.x) { zero_ring[i] = 0.0; none_ring[i] = EMPTY; }
    for (int i = tid; i < W * (g.pw + g.ph) * R * 4; i += blockDim.x) ring_i[i] = EMPTY;
    __syncthreads();

    if (warp == pg.producer_warp) {
        // ------------------------------------------------ record producer
        // Record g of compute warp w may be issued into its stage once record g-S has been used up
        // (prog[w] > g-S).  The whole warp walks the compute warps in turn with warp-uniform state
        // (the bulk-copy instructions take their operands from uniform registers: per-lane state
        // costs ~100 instructions of set-up per copy), one elected lane issues.  The progress
        // counters are polled at a leisurely pace: a compute warp has S-1 records in flight or landed.
        int q_[kColMaxWarps], t_[kColMaxWarps], st_[kColMaxWarps], issued_[kColMaxWarps];
        bool done_[kColMaxWarps];
#pragma unroll
        for (int w_ = 0; w_ < kColMaxWarps; ++w_) {
            q_[w_] = q_begin; t_[w_] = 0; st_[w_] = 0; issued_[w_] = 0; done_[w_] = w_ >= W;
            while (!done_[w_] && q_[w_] < q_end && !col_warp_has_patch(g, pg.cta_tiles[q_[w_]], w_)) ++q_[w_];
            if (!done_[w_] && q_[w_] >= q_end) done_[w_] = true;
        }
        unsigned idle = 0;
        for (;;) {
            bool all_done = true, progressed = false;
#pragma unroll
            for (int w_ = 0; w_ < kColMaxWarps; ++w_) {
                if (done_[w_]) continue;
                all_done = false;
                const int used = prog[w_];
                // (several records per visit when the compute warp ran ahead)
                while (!done_[w_] && issued_[w_] < used + S) {
                    const size_t rec_no = ((size_t)pg.cta_tiles[q_[w_]] * W + w_) * g.T + t_[w_];
                    unsigned char* stage = stages + ((size_t)w_ * S + st_[w_]) * stage_bytes;
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full[w_ * S + st_[w_]], stage_bytes);
                        tma_bulk_g2s(stage + 768, pg.rec + rec_no * NC * 32, NC * 256, &full[w_ * S + st_[w_]]);
                        tma_bulk_g2s(stage, rhs_perm + rec_no * 96, 768, &full[w_ * S + st_[w_]]);
                    }
                    progressed = true;
                    ++issued_[w_];
                    if (++st_[w_] == S) st_[w_] = 0;
                    if (++t_[w_] == g.T) {
                        t_[w_] = 0; ++q_[w_];
                        while (q_[w_] < q_end && !col_warp_has_patch(g, pg.cta_tiles[q_[w_]], w_)) ++q_[w_];
                        if (q_[w_] >= q_end) done_[w_] = true;
                    }
                }
            }
            if (all_done) break;
            if (progressed) idle = 0;
            else {
                __nanosleep(40);
                if (*abort_flag) break;
                if (++idle > kColSpinLimit) { if (lane == 0) { *abort_flag = 1; atomicExch(err, 11); } break; }
            }
        }
    } else if (warp == pg.helper_warp) {
        // ------------------------------------------------ helper: L2 push slots of other tiles -> rings
        for (int q = q_begin; q < q_end; ++q) {
            const int tile = pg.cta_tiles[q];
            const int tA = tile % g.nta, tB = tile / g.nta;
            const bool up_i = UPPER ? tA < g.nta - 1 : tA > 0, up_j = UPPER ? tB < g.ntb - 1 : tB > 0;
            int kn[2] = {0, 0};
            uint32_t ring_a[2] = {0, 0};
            const long long* slots[2] = {nullptr, nullptr};
            bool ex[2] = {false, false};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = lane + 32 * u;
                if (e >= g.NE) continue;
                if (e < g.tb * g.ph) {
                    const int wb = e / g.ph, lj = e % g.ph;
                    const int pb = tB * g.tb + wb, j = pb * g.ph + lj;
                    const int wc = (UPPER ? g.ta - 1 : 0) + g.ta * wb;
                    ex[u] = up_i && pb < g.npb && j < g.ny;
                    ring_a[u] = smem_u32(ring_i + ((size_t)(wc * g.ph + lj) * R) * 4);
                } else {
                    const int ii = e - g.tb * g.ph, wa = ii / g.pw, li = ii % g.pw;
                    const int pa = tA * g.ta + wa, i = pa * g.pw + li;
                    const int wc = wa + g.ta * (UPPER ? g.tb - 1 : 0);
                    ex[u] = up_j && pa < g.npa && i < g.nx;
                    ring_a[u] = smem_u32(ring_j + ((size_t)(wc * g.pw + li) * R) * 4);
                }
                slots[u] = reinterpret_cast<const long long*>(pg.ext) + ((size_t)tile * g.NE + e) * g.nz * 3;
            }
            unsigned idle = 0;
            bool dead = false;
            while (__any_sync(FULLM, (ex[0] && kn[0] < g.nz) || (ex[1] && kn[1] < g.nz))) {
                bool progressed = false;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (!(ex[u] && kn[u] < g.nz)) continue;
                    long long a[kColHelperBatch][3];
                    const int nb = min(kColHelperBatch, g.nz - kn[u]);
#pragma unroll
                    for (int b = 0; b < kColHelperBatch; ++b) {
                        a[b][0] = a[b][1] = a[b][2] = -1;
                        if (b < nb) {
                            const long long* sl = slots[u] + (size_t)(kn[u] + b) * 3;
                            asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(a[b][0]) : "l"(sl) : "memory");
                            asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(a[b][1]) : "l"(sl + 1) : "memory");
                            asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(a[b][2]) : "l"(sl + 2) : "memory");
                        }
                    }
                    int n = 0;
#pragma unroll
                    for (int b = 0; b < kColHelperBatch; ++b) {
                        if (n != b) continue;                               // valid prefix only
                        if (a[b][0] == -1 || a[b][1] == -1 || a[b][2] == -1) continue;
                        const uint32_t ra = ring_a[u] + (uint32_t)((kn[u] + b) & (R - 1)) * 32;
                        if (!col_empty3(col_lds(ra), col_lds(ra + 8), col_lds(ra + 16))) continue;      // consumer not there yet
                        col_sts(ra, __longlong_as_double(a[b][0])); col_sts(ra + 8, __longlong_as_double(a[b][1]));
                        col_sts(ra + 16, __longlong_as_double(a[b][2]));
                        n = b + 1;
                    }
                    // re-arm the delivered slots (their producer writes them again one sweep later)
#pragma unroll
                    for (int b = 0; b < kColHelperBatch; ++b)
                        if (b < n) {
                            long long* sl = const_cast<long long*>(slots[u]) + (size_t)(kn[u] + b) * 3;
                            __stcg(sl, -1LL); __stcg(sl + 1, -1LL); __stcg(sl + 2, -1LL);
                        }
                    kn[u] += n;
                    progressed = progressed || n > 0;
                }
                if (__any_sync(FULLM, progressed)) idle = 0;
                else {
                    if (*abort_flag) { dead = true; break; }
                    if (++idle > kColSpinLimit) { if (lane == 0) { *abort_flag = 1; atomicExch(err, 13); } dead = true; break; }
                }
            }
            if (dead) break;
        }
    } else if (warp < W) {
        // ------------------------------------------------ compute warps
        const int li = lane % g.pw, lj = lane / g.pw;
        const bool lane_ok = lane < g.pw * g.ph;
        const int s = UPPER ? -1 : 1;
        const int wa = warp % g.ta, wb = warp / g.ta;
        const int dloc = col_lane_delay(g, UPPER, li, lj);
        const long long plane = (long long)g.nx * g.ny;
        // neighbours inside the patch (shuffle sources) and at its edges (ring entries)
        const bool iu_in = lane_ok && li - s >= 0 && li - s < g.pw, ju_in = lane_ok && lj - s >= 0 && lj - s < g.ph;
        const int src_i = iu_in ? lane - s : lane, src_j = ju_in ? lane - s * g.pw : lane;
        const uint32_t zero_s = smem_u32(zero_ring), none_s = smem_u32(none_ring);
        const uint32_t my_ring_i = smem_u32(ring_i + ((size_t)(warp * g.ph + lj) * R) * 4);
        const uint32_t my_ring_j = smem_u32(ring_j + ((size_t)(warp * g.pw + li) * R) * 4);
        const bool dn_i_tile = wa + s >= 0 && wa + s < g.ta, dn_j_tile = wb + s >= 0 && wb + s < g.tb;
        const uint32_t dn_ring_i = smem_u32(ring_i + ((size_t)((warp + s) * g.ph + lj) * R) * 4);
        const uint32_t dn_ring_j = smem_u32(ring_j + ((size_t)((warp + s * g.ta) * g.pw + li) * R) * 4);
        const uint32_t full_s = smem_u32(&full[warp * S]);
        const uint32_t stage0_s = smem_u32(stages + (size_t)warp * S * stage_bytes);
        int st = 0;                 // stage and parity of the next record to take
        unsigned par = 0;
        // records this warp has used up (their stages may be refilled): read by the producer warp
        const uint32_t prog_s = smem_u32(const_cast<int*>(prog) + warp);
        int nused = 0;
        long long prof_rec = 0, prof_ring = 0, prof_all = 0, nsteps_done = 0;
        bool dead = false;
        for (int q = q_begin; q < q_end; ++q) {
            const int tile = pg.cta_tiles[q];
            if (!col_warp_has_patch(g, tile, warp)) continue;
            const int tA = tile % g.nta, tB = tile / g.nta;
            const int i = (tA * g.ta + wa) * g.pw + li, j = (tB * g.tb + wb) * g.ph + lj;
            const bool col_ok = lane_ok && i < g.nx && j < g.ny;
            const int iu = i - s, ju = j - s, id = i + s, jd = j + s;
            const bool has_iu = col_ok && iu >= 0 && iu < g.nx, has_ju = col_ok && ju >= 0 && ju < g.ny;
            const bool has_id = col_ok && id >= 0 && id < g.nx, has_jd = col_ok && jd >= 0 && jd < g.ny;
            const bool iu_sh = has_iu && iu_in, ju_sh = has_ju && ju_in;
            const bool iu_ring = has_iu && !iu_in, ju_ring = has_ju && !ju_in;
            const bool id_out = has_id && !(li + s >= 0 && li + s < g.pw), jd_out = has_jd && !(lj + s >= 0 && lj + s < g.ph);
            const bool id_ring = id_out && dn_i_tile, jd_ring = jd_out && dn_j_tile;
            const bool id_l2 = id_out && !dn_i_tile, jd_l2 = jd_out && !dn_j_tile;
            // every lane reads two input entries and two output entries per step; lanes without a
            // ring on a side read the always-valid zero ring / the always-empty ring instead
            const uint32_t in_i = iu_ring ? my_ring_i : zero_s, in_j = ju_ring ? my_ring_j : zero_s;
            const uint32_t out_i = id_ring ? dn_ring_i : none_s, out_j = jd_ring ? dn_ring_j : none_s;
            const size_t ls_base = ((size_t)tile * W + warp) * g.T;
            // running pointers of the lane's cell (step 0 may lie before the column's first cell: only used when active)
            double* slot_i = pg.ext + ((size_t)(tile + s) * g.NE + (wb * g.ph + lj)) * g.nz * 3 - (ptrdiff_t)dloc * 3;
            double* slot_j = pg.ext + ((size_t)(tile + s * g.nta) * g.NE + (g.tb * g.ph + wa * g.pw + li)) * g.nz * 3 - (ptrdiff_t)dloc * 3;
            double* res_p = UPPER ? out + ((ptrdiff_t)i + (ptrdiff_t)g.nx * j + (ptrdiff_t)plane * (g.nz - 1 + dloc)) * 3
                                  : hand_off + ((ls_base + (size_t)(g.T - 1)) * 32 + lane) * 3;
            const ptrdiff_t res_step = UPPER ? -(ptrdiff_t)plane * 3 : -96;
            const bool warp_l2_out = __any_sync(FULLM, id_l2 || jd_l2);        // warp-uniform: any lane pushes to another tile
            int kl = -dloc;                      // cell of the current step along the column (sweep-local)
            uint32_t ro = (uint32_t)(kl & (R - 1)) * 32;
            double y0 = 0.0, y1 = 0.0, y2 = 0.0;

            // Ring entries of step t: two inputs (values) and two outputs (must be empty again),
            // fetched one step ahead, so on the usual path (producer ahead, consumer keeping up) the
            // step itself only votes on `ok`.  An input is there when none of its three words is
            // empty (upper half 0xffffffff); an output entry is free again when its word 2 is empty
            // (its consumer re-arms words 0,1 before word 2, and a thread's shared-memory stores are
            // performed in order).  vol: polling loop (the loads must not be hoisted out of it).
            struct RingPre { double ri0, ri1, ri2, rj0, rj1, rj2; bool ok; };
            auto ring_load = [&](bool vol, int kl_, uint32_t ro_, RingPre& r) {
                const bool active = col_ok && (unsigned)kl_ < (unsigned)g.nz;
                int oi, oj;
                if (vol) {
                    r.ri0 = col_lds(in_i + ro_); r.ri1 = col_lds(in_i + ro_ + 8); r.ri2 = col_lds(in_i + ro_ + 16);
                    r.rj0 = col_lds(in_j + ro_); r.rj1 = col_lds(in_j + ro_ + 8); r.rj2 = col_lds(in_j + ro_ + 16);
                    oi = __double2hiint(col_lds(out_i + ro_ + 16)); oj = __double2hiint(col_lds(out_j + ro_ + 16));
                } else {
                    col_lds2(in_i + ro_, r.ri0, r.ri1); r.ri2 = col_lds_var(in_i + ro_ + 16);
                    col_lds2(in_j + ro_, r.rj0, r.rj1); r.rj2 = col_lds_var(in_j + ro_ + 16);
                    oi = col_lds_hi(out_i + ro_ + 16); oj = col_lds_hi(out_j + ro_ + 16);
                }
                r.ok = !active | ((__double2hiint(r.ri0) != -1) & (__double2hiint(r.ri1) != -1) & (__double2hiint(r.ri2) != -1) &
                                  (__double2hiint(r.rj0) != -1) & (__double2hiint(r.rj1) != -1) & (__double2hiint(r.rj2) != -1) & ((oi & oj) == -1));
            };
            auto wait_record = [&]() {           // slow path of the record wait
                const uint32_t bar = full_s + 8u * st;
                unsigned spins = 0;
                while (!dead && !col_try_wait(bar, par)) {
                    if (*abort_flag) dead = true;
                    else if (++spins > kColSpinLimit) { *abort_flag = 1; atomicExch(err, 11); dead = true; }
                }
            };
            // one step: `cur` / `pre` hold the record and ring entries of step t, `nxt` / `pre_n`
            // receive those of step t+1
            auto step = [&](ColRegs<UPPER>& cur, ColRegs<UPPER>& nxt, RingPre& pre, RingPre& pre_n, int t) {
                const long long c0 = PROF ? clock64() : 0;
                if (!__all_sync(FULLM, pre.ok)) {
                    // Waiting must stay cheap: several warps of an SM wait most of the time, and a poll
                    // that reads whole entries every few cycles saturates the shared-memory pipe the
                    // working warps load their records through.  So: one 4-byte load per entry (upper
                    // half of word 2, written last) with exponential back-off, the values afterwards.
                    unsigned spins = 0, ns = 16;
                    const bool active = col_ok && (unsigned)kl < (unsigned)g.nz;
                    do {
                        bool ready;
                        do {
                            __nanosleep(ns);
                            ns = min(ns * 2u, 512u);
                            const int a = col_lds_hi_vol(in_i + ro + 16), b = col_lds_hi_vol(in_j + ro + 16);
                            const int c = col_lds_hi_vol(out_i + ro + 16), d = col_lds_hi_vol(out_j + ro + 16);
                            ready = !active | ((a != -1) & (b != -1) & ((c & d) == -1));
                            if (dead || *abort_flag || ++spins > kColSpinLimit) {
                                if (!dead && !*abort_flag) { *abort_flag = 1; atomicExch(err, 12); }
                                dead = true; ready = true;
                            }
                        } while (!__all_sync(FULLM, ready));
                        ring_load(true, kl, ro, pre);
                        if (dead) pre.ok = true;
                    } while (!__all_sync(FULLM, pre.ok));
                }
                const bool more = t + 1 < g.T;
                if (more && !col_test_wait(full_s + 8u * st, par)) wait_record();
                const long long c1 = PROF ? clock64() : 0;
                // ---- straight-line from here: loads of step t+1, shuffles, chain, stores
                const bool active = col_ok && (unsigned)kl < (unsigned)g.nz;
                const uint32_t ro_n = (ro + 32u) & (uint32_t)(R * 32 - 1);
                const double sj0 = __shfl_sync(FULLM, y0, src_j), sj1 = __shfl_sync(FULLM, y1, src_j), sj2 = __shfl_sync(FULLM, y2, src_j);
                const double si0 = __shfl_sync(FULLM, y0, src_i), si1 = __shfl_sync(FULLM, y1, src_i), si2 = __shfl_sync(FULLM, y2, src_i);
                ring_load(false, kl + 1, ro_n, pre_n);
                // (after a tile's last step this reads a stage that is being refilled: never used)
                const uint32_t stage_s = stage0_s + (uint32_t)st * stage_bytes;
                COL_LOAD_REC(nxt, stage_s + 768u + 16u * lane, stage_s + 24u * lane);
                // consumed: re-arm (words 0,1 before word 2)
                col_sts2_if(active && iu_ring, in_i + ro, EMPTY, EMPTY); col_sts_if(active && iu_ring, in_i + ro + 16, EMPTY);
                col_sts2_if(active && ju_ring, in_j + ro, EMPTY, EMPTY); col_sts_if(active && ju_ring, in_j + ro + 16, EMPTY);
                const double i0 = iu_sh ? si0 : pre.ri0, i1 = iu_sh ? si1 : pre.ri1, i2 = iu_sh ? si2 : pre.ri2;
                const double j0 = ju_sh ? sj0 : pre.rj0, j1 = ju_sh ? sj1 : pre.rj1, j2 = ju_sh ? sj2 : pre.rj2;
                double a0 = cur.rhs[0], a1 = cur.rhs[1], a2 = cur.rhs[2];
                a0 = fma(-cur.cf[0], y0, a0); a1 = fma(-cur.cf[9], y0, a1); a2 = fma(-cur.cf[18], y0, a2);
                a0 = fma(-cur.cf[1], y1, a0); a1 = fma(-cur.cf[10], y1, a1); a2 = fma(-cur.cf[19], y1, a2);
                a0 = fma(-cur.cf[2], y2, a0); a1 = fma(-cur.cf[11], y2, a1); a2 = fma(-cur.cf[20], y2, a2);
                a0 = fma(-cur.cf[3], j0, a0); a1 = fma(-cur.cf[12], j0, a1); a2 = fma(-cur.cf[21], j0, a2);
                a0 = fma(-cur.cf[4], j1, a0); a1 = fma(-cur.cf[13], j1, a1); a2 = fma(-cur.cf[22], j1, a2);
                a0 = fma(-cur.cf[5], j2, a0); a1 = fma(-cur.cf[14], j2, a1); a2 = fma(-cur.cf[23], j2, a2);
                a0 = fma(-cur.cf[6], i0, a0); a1 = fma(-cur.cf[15], i0, a1); a2 = fma(-cur.cf[24], i0, a2);
                a0 = fma(-cur.cf[7], i1, a0); a1 = fma(-cur.cf[16], i1, a1); a2 = fma(-cur.cf[25], i1, a2);
                a0 = fma(-cur.cf[8], i2, a0); a1 = fma(-cur.cf[17], i2, a1); a2 = fma(-cur.cf[26], i2, a2);
                if (UPPER) {
                    double v0 = fma(cur.cf[27], a0, 0.0), v1 = fma(cur.cf[30], a0, 0.0), v2 = fma(cur.cf[33], a0, 0.0);
                    v0 = fma(cur.cf[28], a1, v0); v1 = fma(cur.cf[31], a1, v1); v2 = fma(cur.cf[34], a1, v2);
                    v0 = fma(cur.cf[29], a2, v0); v1 = fma(cur.cf[32], a2, v1); v2 = fma(cur.cf[35], a2, v2);
                    a0 = v0; a1 = v1; a2 = v2;
                }
                // results the neighbouring warps wait for first (words 0,1 before word 2)
                col_sts2_if(active && id_ring, out_i + ro, a0, a1); col_sts_if(active && id_ring, out_i + ro + 16, a2);
                col_sts2_if(active && jd_ring, out_j + ro, a0, a1); col_sts_if(active && jd_ring, out_j + ro + 16, a2);
                y0 = active ? a0 : y0; y1 = active ? a1 : y1; y2 = active ? a2 : y2;
                const long long c2 = PROF ? pipe_clock_after(__double2hiint(a0)) : 0;
                if (more && ++st == S) { st = 0; par ^= 1u; }
                // this step's record has been used up (its loads completed before the chain read the
                // registers): the producer warp may refill its stage
                ++nused;
                col_sts32_if(lane == 0, prog_s, nused);
                // nobody on this SM waits for the rest
                if (warp_l2_out) {
                    double* sl = slot_i + (ptrdiff_t)t * 3;
                    col_stcg_if(active && id_l2, sl, a0); col_stcg_if(active && id_l2, sl + 1, a1); col_stcg_if(active && id_l2, sl + 2, a2);
                    sl = slot_j + (ptrdiff_t)t * 3;
                    col_stcg_if(active && jd_l2, sl, a0); col_stcg_if(active && jd_l2, sl + 1, a1); col_stcg_if(active && jd_l2, sl + 2, a2);
                }
                if (UPPER) { col_stg_if(active, res_p, scale ? a0 * w : a0); col_stg_if(active, res_p + 1, scale ? a1 * w : a1); col_stg_if(active, res_p + 2, scale ? a2 * w : a2); }
                else { col_stg_if(active, res_p, a0); col_stg_if(active, res_p + 1, a1); col_stg_if(active, res_p + 2, a2); }
                res_p += res_step;
                ++kl; ro = ro_n;
                if (PROF) {
                    const long long c3 = clock64();
                    prof_rec += c1 - c0; prof_ring += c2 - c1; prof_all += c3 - c0; ++nsteps_done;
                    if (pg.trace && (int)blockIdx.x == pg.trace_cta && warp == 0 && lane == 0) {
                        long long* tr = pg.trace + (size_t)t * 8;
                        tr[0] = c0; tr[1] = c1; tr[2] = c2; tr[3] = c3;
                    }
                }
            };

            ColRegs<UPPER> ra, rb;
            RingPre pa, pb;
            {
                if (!col_try_wait(full_s + 8u * st, par)) wait_record();
                const uint32_t stage_s = stage0_s + (uint32_t)st * stage_bytes;
                COL_LOAD_REC(ra, stage_s + 768u + 16u * lane, stage_s + 24u * lane);
                if (++st == S) { st = 0; par ^= 1u; }
                ring_load(true, kl, ro, pa);
            }
            for (int t = 0; t < g.T; t += 2) {
                step(ra, rb, pa, pb, t);
                if (t + 1 < g.T) step(rb, ra, pb, pa, t + 1);
            }
        }
        if (PROF && pg.prof && lane == 0) {
            long long* pr = pg.prof + ((size_t)blockIdx.x * 8 + warp) * 4;
            pr[0] = prof_rec; pr[1] = prof_ring; pr[2] = prof_all; pr[3] = nsteps_done;
        }
    }
}

// factors -> record streams of the column program, from A and the program-ordered pivots of the
// pipelined factorisation (factor_pipe.cuh).  One thread per (entry b, block row c): a lower
// block becomes row c of L_ij = A_ij * inv(D_j) (the factorisation's own three fused
// multiply-adds per element), the diagonal becomes row c of inv(D_i), an upper block is A's.
template <bool LOWER>
__global__ void __launch_bounds__(256)
repack_col_kernel(size_t nval, const int* __restrict__ src, const unsigned long long* __restrict__ dst,
                  const int* __restrict__ colidx, const int* __restrict__ fpos, const double* __restrict__ A,
                  const double* __restrict__ fout, double* __restrict__ rec)
{
    constexpr int NC = LOWER ? kColNCL : kColNCU;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nval * 3; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t / 3;
        const int c = (int)(t - b * 3);
        const int k = src[b];
        const unsigned long long d = dst[b];
        const int kb = (int)(d & 3);
        const size_t ls = (size_t)(d >> 2);
        double o[3];
        if (kb == 3) {
            const double* p = fout + (size_t)fpos[colidx[k]] * kFEntry + c * 3;
            o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        } else {
            const double* a = A + (size_t)k * 9 + c * 3;
            if (LOWER) {
                const double* p = fout + (size_t)fpos[colidx[k]] * kFEntry;
                const double a0 = a[0], a1 = a[1], a2 = a[2];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    double sacc = 0.0;
                    sacc = fma(a0, p[e], sacc); sacc = fma(a1, p[3 + e], sacc); sacc = fma(a2, p[6 + e], sacc);
                    o[e] = sacc;
                }
            } else { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
        }
        const int q0 = kb == 3 ? 27 + c * 3 : c * 9 + kb * 3;
        rec[col_rec_index(ls, NC, q0)] = o[0]; rec[col_rec_index(ls, NC, q0 + 1)] = o[1]; rec[col_rec_index(ls, NC, q0 + 2)] = o[2];
    }
}

// the same from a BCRS factor array (factorisation by the tile kernel)
template <bool LOWER>
__global__ void __launch_bounds__(256)
repack_col_from_lu_kernel(size_t nval, const int* __restrict__ src, const unsigned long long* __restrict__ dst,
                          const double* __restrict__ lu, double* __restrict__ rec)
{
    constexpr int NC = LOWER ? kColNCL : kColNCU;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nval * 3; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t / 3;
        const int c = (int)(t - b * 3);
        const unsigned long long d = dst[b];
        const int kb = (int)(d & 3);
        const double* a = lu + (size_t)src[b] * 9 + c * 3;
        const int q0 = kb == 3 ? 27 + c * 3 : c * 9 + kb * 3;
        const size_t ls = (size_t)(d >> 2);
        rec[col_rec_index(ls, NC, q0)] = a[0]; rec[col_rec_index(ls, NC, q0 + 1)] = a[1]; rec[col_rec_index(ls, NC, q0 + 2)] = a[2];
    }
}

}  // namespace opmgpu
