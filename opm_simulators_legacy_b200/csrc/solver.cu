// C ABI (include/opm_gpu_solver.h) and host driver of the B200-native Newton-step linear
// solver.  Replaces, behind NewtonIterationBlackoilInterface::computeNewtonIncrement, the
// reference's formInterleavedSystem + ISTLSolver::solve
// (opm/autodiff/NewtonIterationBlackoilInterleaved.cpp:110-194, :234-283;
//  opm/autodiff/ISTLSolver.hpp:124-189, 201-211, 250-274, 283-306, 358-368).
// There is no CPU fallback anywhere in this file.
#include "../../include/opm_gpu_solver.h"
#include "analysis.hpp"
#include "kernels.cuh"
#include "sweep_pipe.cuh"
#include "factor_pipe.cuh"
#include "colprog.hpp"
#include "spmv_tma.cuh"
#include "gmres.cuh"
#include "generic_np.cuh"
#include "mcorder.hpp"
#include "mc_ilu.cuh"
#include "multi.hpp"

#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>      // types only: the library is resolved at run time with dlopen

#include <algorithm>
#include <chrono>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <future>
#include <string>
#include <thread>
#include <vector>

// Experiments -- slower kernel variants kept for A/B measurements (column-owned sweeps, level-by-level
// factorisation, ...), tuning switches and the tracing entry points -- exist only in builds with
// -DOPMGPU_EXPERIMENTS (`make exp` -> libopmgpu_exp.so, loaded with OPMGPU_LIB).  The shipping
// library reads none of those switches and has one path per pattern class.
// (exp_env, analysis.hpp: getenv in such builds, nullptr otherwise.)
#ifdef OPMGPU_EXPERIMENTS
#include "sweep_col.cuh"
#endif

using namespace opmgpu;

namespace {

std::string g_create_error;

// NCCL is only needed by distributed handles and is resolved lazily (inside a PyTorch
// process this picks up the libnccl.so.2 torch already loaded).
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err)
    {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
        bool ok = true;
        auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) ok = false; return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!ok) { err = "NCCL library lacks a required symbol"; lib = nullptr; }
        return ok;
    }
} g_nccl;

template <class T>
struct DevArr {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct ProgramDevMem {
    DevArr<int> cta_step_ptr, step_row_ptr, prow, pblk_ptr, pcol, psrc;
    DevArr<unsigned char> publish;
    DevArr<int> pair_ptr, pair_jk, pair_ik, frow, fent, fpush_ptr, fpush_slot;
    DevArr<unsigned char> needs_flag;
    DevArr<double> pval, pdinv, fslots;
    size_t n_fslots = 0;
    size_t nblk = 0;
    int P = 0;
    void release()
    {
        cta_step_ptr.release(); step_row_ptr.release(); prow.release(); pblk_ptr.release();
        pcol.release(); psrc.release(); publish.release(); pval.release(); pdinv.release();
        pair_ptr.release(); pair_jk.release(); pair_ik.release(); frow.release(); fent.release();
        fpush_ptr.release(); fpush_slot.release(); needs_flag.release(); fslots.release();
    }
};

struct PipeDevMem {
    DevArr<unsigned char> buf;
    DevArr<int> cta_step_ptr, val_src, val_stride, perm_row, pos_of_row;
    DevArr<unsigned> step_off16, step_bytes, step_rhs_row, step_rhs_bytes, val_dst8;
    DevArr<long long> cta_ext_base;
    DevArr<double> ext, rhs_perm;
    size_t nval = 0, next = 0, nperm = 0;
    int P = 0, nstages = 0, stage_bytes = 0, rhs_bytes = 0, cluster_size = 1, cx_bytes = 0;
    bool lean = false;
    size_t smem = 0;
    void release()
    {
        buf.release(); cta_step_ptr.release(); val_src.release(); val_stride.release(); perm_row.release(); pos_of_row.release();
        step_off16.release(); step_bytes.release(); step_rhs_row.release(); step_rhs_bytes.release();
        val_dst8.release(); cta_ext_base.release(); ext.release(); rhs_perm.release();
    }
};

// column-owned sweep program (exact Cartesian stencils, colprog.hpp)
struct ColDevMem {
    bool valid = false;
    ColGeom g = {};
    int P = 0, nstagesL = 0, nstagesU = 0;
    size_t smemL = 0, smemU = 0, nperm = 0, next = 0, nvalL = 0, nvalU = 0;
    DevArr<double> recL, recU, extL, extU, rhsL, rhsU;
    DevArr<int> tile_ptr, tilesL, tilesU, valL_src, valU_src, perm_row, pos_of_row;
    DevArr<unsigned long long> valL_dst, valU_dst;
    void release()
    {
        recL.release(); recU.release(); extL.release(); extU.release(); rhsL.release(); rhsU.release();
        tile_ptr.release(); tilesL.release(); tilesU.release(); valL_src.release(); valU_src.release();
        perm_row.release(); pos_of_row.release(); valL_dst.release(); valU_dst.release();
        valid = false;
    }
};

// multicolour ILU0 variant (mcorder.hpp, mc_ilu.cuh)
struct McDevMem {
    bool valid = false;
    int ncolours = 0;
    bool lines = false;          // k-line ordering (mc_line_sweep_kernel): per colour the columns, planes, columns per CTA
    int nz = 0, nx = 0, ncols[2] = {0, 0}, base[2] = {0, 0}, rb[2] = {4, 4};
    long long nnzL = 0, nnzU = 0, offD = 0, offU = 0, total_blocks = 0;
    std::vector<int> colour_ptr, lvl_ptr, p2n_host, n2p_host;
    DevArr<int> p2n, psrc, ppos, Lrowptr, Lcol, Urowptr, Ucol, lvl_rows, pair_ptr, pair_jk, pair_ik;
    DevArr<double> uni;          // [ L | Dinv | U ] blocks of T (allocated for doubles)
    void release()
    {
        valid = false;
        p2n.release(); psrc.release(); ppos.release(); pair_ptr.release(); pair_jk.release(); pair_ik.release();
        Lrowptr.release(); Lcol.release(); Urowptr.release(); Ucol.release(); lvl_rows.release(); uni.release();
    }
};

struct FactorPipeDevMem {
    DevArr<unsigned char> buf;
    DevArr<int> cta_step_ptr, val_src, cta_row_base, fpos;
    DevArr<unsigned> step_off16, step_bytes, val_dst8;
    DevArr<long long> cta_ext_base;
    DevArr<double> ext, fout;
    size_t nval = 0, next = 0;
    int P = 0, nstages = 0, stage_bytes = 0;
    size_t smem = 0;
    bool valid = false;
    void release()
    {
        buf.release(); cta_step_ptr.release(); val_src.release(); step_off16.release(); step_bytes.release();
        val_dst8.release(); cta_ext_base.release(); ext.release(); fout.release(); cta_row_base.release(); fpos.release();
        valid = false;
    }
};

}  // namespace

// Halo exchange through peer memory: a rank's boundary rows are stored straight into its neighbours'
// ghost rows over NVLink (one small kernel), followed by an epoch flag; no send/recv pairing.
constexpr int kMaxPeers = 8;
struct PeerPushDesc {
    void* y[kMaxPeers];                       // peer's SpMV input vector (its d_y), mapped here
    unsigned long long* flags[kMaxPeers];     // peer's incoming-epoch array
    long long dst_row0[kMaxPeers];            // first ghost row (in the peer's numbering) my rows go to
    int send_off[kMaxPeers + 1];              // my send list, grouped by peer
    int world, rank;
};
// Scalar all-reduce through peer memory: every rank stores its partial sums into every other
// rank's inbox (double-buffered by the exchange's parity), then its epoch flag; each rank adds the
// W contributions in rank order, so all ranks hold bit-identical sums.
constexpr int kScalSlots = 16;                // = S_COUNT (kernels.cuh)
struct PeerScalDesc {
    double* inbox[kMaxPeers];                 // rank p's inbox [2][kMaxPeers][kScalSlots] (own entry: local pointer)
    unsigned long long* flags[kMaxPeers];     // rank p's scalar-epoch array [kMaxPeers]
    int world, rank;
};

struct opmgpu_solver {
    opmgpu::MultiSolver* multi = nullptr;      // opmgpu_create_multi: this handle only fronts one worker handle per GPU
    int device = 0;
    int sm_count = 0;
    int sweep_ctas = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    std::string err;
    long long launches = 0;

    // pattern
    bool have_pattern = false, have_values = false, have_factors = false;
    int N = 0, nnzb = 0, nlevL = 0, nlevU = 0;
    PatternAnalysis an;
    DevArr<int> d_rowptr, d_colidx, d_diag, d_lvl_rows;
    ProgramDevMem progL, progU;
    PipeDevMem pipeL, pipeU;
    FactorPipeDevMem pipeF;
    ColDevMem col;                 // column-owned sweeps (exact Cartesian stencils); replaces pipeL / pipeU when valid
    bool use_col = false, allow_col = false;     // OPMGPU_COL=1: column-owned sweeps on exact Cartesian stencils (experimental, slower so far: profiles/r02_column_sweeps.md)
    bool factor_tile = false;      // OPMGPU_FACTOR_TILE=1: keep the flag-synchronised tile kernel
    bool lu_lazy = false;          // the pipelined factorisation left only pivots: d_lu is built on demand
    ClusterCaps caps;              // co-resident CTAs of the cluster variants of the sweep kernels
    int cluster_size = 1;          // thread-block cluster size the sweeps are launched with
    int N_for_upload = 0;          // rows of the pattern being uploaded (set_pattern)
    bool fuse_permute = true;      // OPMGPU_FUSE_PERMUTE=0: separate permute kernel before every lower sweep
    bool use_pipe = false, force_simple = false, factor_by_levels = false, spmv_tma = true;
    int trace_cta = -1;
    DevArr<long long> d_trace;
    std::future<bool> pattern_check;   // CSC front end: full index compare running beside the upload
    int gtrace_steps = 0;          // > 0: all-CTA %globaltimer trace of the next apply (debug)
    DevArr<long long> d_gtrace;
    int max_smem_optin = 0;

    // values and factors
    // f32: the handle runs the reference's float instance (Impl<3,float>, ...Interleaved.cpp:478-480):
    // matrix values and vectors are float arrays (the vector buffers below are then used as float
    // arrays), factors / sweep records / scalar block keep their 8-byte containers.  The C ABI
    // exchanges doubles either way; d_stage holds them on their way in and out.
    bool f32 = false;
    // Block size.  np_req (opmgpu_set_block_size) is what the NEXT pattern is prepared for: np = 2 also
    // gets the level sets of the upper triangle (generic_np.cuh runs one launch per level); np is the
    // block size of the call in flight (3 except inside the *_np entry points).
    int np_req = 3, np = 3;
    std::vector<int> lvlU_ptr;
    DevArr<int> d_lvlU_rows;
    DevArr<long long> d_map_np;
    std::vector<std::vector<int>> npcsc_colptr, npcsc_rowidx;      // CSC front end cache of the np path
    std::vector<long long> npcsc_base;
    // ILU0 ordering: 0 natural (the reference's), 1 multicolour (flagged variant).  ilu_order_req is what
    // the NEXT pattern is prepared for (opmgpu_set_ilu_ordering), mc.valid what the current one runs.
    int ilu_order_req = 0;
    McDevMem mc;
    bool operator_only = false;    // opmgpu_set_pattern_bcrs_operator_only: no ILU0 programs, SpMV entry points only
    DevArr<double> d_vals_own, d_lu, d_stage;
    DevArr<float> d_vals32;
    const void* d_vals = nullptr;

    // vectors (3N each)
    DevArr<double> d_x, d_r, d_rt, d_p, d_v, d_t, d_y, d_yL, d_vU, d_tmp, d_tmp2;
    DevArr<double> d_S, d_partials;
    DevArr<double> d_gmres_V, d_gmres_H;      // restarted GMRES (newton_use_gmres): m+1 basis vectors, one Hessenberg column
    DevArr<unsigned> d_ticket;
    DevArr<int> d_flags, d_err, d_bad;
    int epoch = 0;
    double* h_S = nullptr;       // pinned
    int* h_flags2 = nullptr;     // pinned: [0] sweep watchdog, [1] bad row / bad pattern
    unsigned long long* h_seq = nullptr;   // pinned: host mailbox sequence number (kernels.cuh: HostBox)
    unsigned long long seq = 0;
    bool use_hostbox = true;     // OPMGPU_HOSTBOX=0: copy + stream synchronisation per half-step instead
    cudaEvent_t ev[8] = {};

    // distributed (one process per GPU): row partition, halo plan, NCCL communicator
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int n_ghost = 0, nnzb_full = 0;
    DevArr<int> d_rowptr_full, d_colidx_full, d_lu_src, d_send_rows;
    // halo exchange beside the SpMV: rows that reference no ghost column are multiplied while the
    // exchange is in flight on halo_stream, the boundary rows (d_bnd_rows; one bit per row in
    // d_row_skip, a word per SpMV tile) when it has arrived
    DevArr<int> d_bnd_rows;
    DevArr<unsigned long long> d_row_skip;
    int n_bnd_rows = 0;
    cudaStream_t halo_stream = nullptr;
    cudaEvent_t ev_x_ready = nullptr, ev_halo_done = nullptr;
    struct PeerHalo {
        bool ready = false;
        std::vector<void*> opened;                 // cudaIpcOpenMemHandle mappings to close
        DevArr<unsigned long long> flags_in;       // [kMaxPeers] epoch of the last exchange each peer completed into my ghosts
        DevArr<unsigned> ticket;
        PeerPushDesc desc;
        unsigned recv_mask = 0;
        unsigned long long epoch = 0;
        // scalar all-reduce
        DevArr<double> inbox;                      // [2][kMaxPeers][kScalSlots]
        DevArr<unsigned long long> sflags;         // [kMaxPeers]
        PeerScalDesc sdesc;
        unsigned long long sepoch = 0;
    } peer;
    bool use_peer_halo = true;     // OPMGPU_PEER_HALO=0: ncclSend/ncclRecv for every exchange
    bool overlap_halo = false;     // OPMGPU_HALO_OVERLAP=1: exchange beside the SpMV (measured: no gain over exchange-then-SpMV with NCCL send/recv, profiles/r02_halo_overlap.md)
    DevArr<double> d_sendbuf;
    std::vector<int> send_cnt, send_off, recv_cnt, recv_off;
    int n_send = 0;
    int nccl_fail(ncclResult_t r, const char* what)
    {
        err = std::string("NCCL error at ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        return OPMGPU_NCCL_ERROR;
    }

    // CSC front end cache
    std::vector<std::vector<int>> csc_colptr, csc_rowidx;
    std::vector<long long> csc_base;
    long long csc_total = 0;
    bool csc_full_pattern = false;
    DevArr<long long> d_map9;
    DevArr<double> d_cscval, d_rhs_stage;

    std::vector<double> history;

    // optional per-kernel-class timing (CUDA events on the launching stream)
    bool profile = false;
    std::vector<cudaEvent_t> prof_pool;
    std::vector<int> prof_kind;             // kind of each recorded [begin,end) pair
    size_t prof_used = 0;
    double prof_ms[4] = {0, 0, 0, 0};       // 0 precond apply, 1 spmv(+dots), 2 vector kernels, 3 factor
    long long prof_cnt[4] = {0, 0, 0, 0};
    void prof_begin(int kind)
    {
        if (!profile) return;
        if (prof_used + 2 > prof_pool.size()) {
            const size_t old = prof_pool.size();
            prof_pool.resize(old + 256);
            for (size_t i = old; i < prof_pool.size(); ++i) cudaEventCreate(&prof_pool[i]);
        }
        prof_kind.push_back(kind);
        cudaEventRecord(prof_pool[prof_used++], stream);
    }
    void prof_end() { if (profile) cudaEventRecord(prof_pool[prof_used++], stream); }
    void prof_collect()
    {
        if (!profile) return;
        cudaStreamSynchronize(stream);
        for (size_t i = 0; i < prof_kind.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, prof_pool[2 * i], prof_pool[2 * i + 1]);
            prof_ms[prof_kind[i]] += ms;
            prof_cnt[prof_kind[i]]++;
        }
        prof_kind.clear();
        prof_used = 0;
    }

    int fail(cudaError_t e, const char* what)
    {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
        err = buf;
        return OPMGPU_CUDA_ERROR;
    }
    int bad(const char* what) { err = what; return OPMGPU_BAD_ARGUMENT; }
    ReduceWs ws() { return ReduceWs{d_partials.p, d_ticket.p}; }
};

#define NK(call)                                                          \
    do {                                                                  \
        ncclResult_t r__ = (call);                                        \
        if (r__ != ncclSuccess) return h->nccl_fail(r__, #call);          \
    } while (0)

// block sizes of the level-scheduled path (generic_np.cuh): NPV is the compile-time block size inside CALL
#define NP_DISPATCH(np, CALL)                                             \
    switch (np) {                                                         \
    case 2: { constexpr int NPV = 2; CALL; } break;                       \
    case 4: { constexpr int NPV = 4; CALL; } break;                       \
    case 5: { constexpr int NPV = 5; CALL; } break;                       \
    case 6: { constexpr int NPV = 6; CALL; } break;                       \
    default: return h->bad("block size not built");                       \
    }

#define CK(call)                                                          \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return h->fail(e__, #call);               \
    } while (0)

namespace {

#ifndef OPMGPU_VEC_BLOCKS_PER_SM
#define OPMGPU_VEC_BLOCKS_PER_SM 8
#endif
constexpr int kVecBlocks = 148 * OPMGPU_VEC_BLOCKS_PER_SM;      // fixed launch shape of the vector kernels (determinism)
static_assert(kVecBlocks <= kMaxRedBlocks, "per-block partials of the grid reduction");

template <class T>
int upload(opmgpu_handle h, DevArr<T>& d, const std::vector<T>& v)
{
    CK(d.ensure(v.size()));
    if (!v.empty()) CK(cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int upload_program(opmgpu_handle h, const SweepProgram& p, ProgramDevMem& d, bool upper, bool with_values = true)
{
    int rc;
    if (!upper) {
        if ((rc = upload(h, d.pair_ptr, p.pair_ptr))) return rc;
        if ((rc = upload(h, d.pair_jk, p.pair_jk))) return rc;
        if ((rc = upload(h, d.pair_ik, p.pair_ik))) return rc;
        if ((rc = upload(h, d.frow, p.frow))) return rc;
        if ((rc = upload(h, d.fent, p.fent))) return rc;
        if ((rc = upload(h, d.fpush_ptr, p.fpush_ptr))) return rc;
        if ((rc = upload(h, d.fpush_slot, p.fpush_slot))) return rc;
        if ((rc = upload(h, d.needs_flag, p.needs_flag))) return rc;
        d.n_fslots = (size_t)p.n_fslots;
        CK(d.fslots.ensure(std::max<size_t>(d.n_fslots, 1) * 9));
        CK(cudaMemsetAsync(d.fslots.p, 0xff, std::max<size_t>(d.n_fslots, 1) * 9 * sizeof(double), h->stream));
    }
    if ((rc = upload(h, d.cta_step_ptr, p.cta_step_ptr))) return rc;
    if ((rc = upload(h, d.step_row_ptr, p.step_row_ptr))) return rc;
    if ((rc = upload(h, d.prow, p.prow))) return rc;
    if ((rc = upload(h, d.pblk_ptr, p.pblk_ptr))) return rc;
    if ((rc = upload(h, d.pcol, p.pcol))) return rc;
    if ((rc = upload(h, d.psrc, p.psrc))) return rc;
    if ((rc = upload(h, d.publish, p.publish))) return rc;
    d.nblk = p.pcol.size();
    d.P = p.P;
    if (with_values) {
        CK(d.pval.ensure(d.nblk * 9));
        if (upper) CK(d.pdinv.ensure((size_t)h->N * 9));
    }
    return 0;
}

// integer parts of the step records (compact host stream, analysis.hpp) -> their places in the
// zero-filled record buffer; one block per step, 16-byte copies
__global__ void __launch_bounds__(128)
expand_records_kernel(int nsteps, const uint4* __restrict__ ibuf, const unsigned* __restrict__ ioff16,
                      const unsigned* __restrict__ ilen, const unsigned* __restrict__ roff,
                      const unsigned* __restrict__ off16, uint4* __restrict__ buf)
{
    for (int s = blockIdx.x; s < nsteps; s += gridDim.x) {
        const uint4* src = ibuf + ioff16[s];
        uint4* rec = buf + off16[s];
        const unsigned n16 = ilen[s] / 16, r16 = roff[s] / 16;
        for (unsigned q = threadIdx.x; q < n16; q += blockDim.x) rec[q < 2 ? q : r16 + (q - 2)] = src[q];
    }
}

int expand_records(opmgpu_handle h, size_t total_bytes, const std::vector<unsigned char>& ibuf, const std::vector<unsigned>& ioff16,
                   const std::vector<unsigned>& ilen, const std::vector<unsigned>& roff, const unsigned* d_off16,
                   DevArr<unsigned char>& d_buf)
{
    int rc;
    DevArr<unsigned char> d_ibuf;
    DevArr<unsigned> d_ioff, d_ilen, d_roff;
    CK(d_buf.ensure(total_bytes));
    CK(cudaMemsetAsync(d_buf.p, 0, total_bytes, h->stream));
    if ((rc = upload(h, d_ibuf, ibuf))) return rc;
    if ((rc = upload(h, d_ioff, ioff16))) return rc;
    if ((rc = upload(h, d_ilen, ilen))) return rc;
    if ((rc = upload(h, d_roff, roff))) return rc;
    const int nsteps = (int)ioff16.size();
    if (nsteps > 0) {
        expand_records_kernel<<<std::min(nsteps, h->sm_count * 16), 128, 0, h->stream>>>(
            nsteps, (const uint4*)d_ibuf.p, d_ioff.p, d_ilen.p, d_roff.p, d_off16, (uint4*)d_buf.p);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));       // the temporaries go away
    d_ibuf.release(); d_ioff.release(); d_ilen.release(); d_roff.release();
    return 0;
}

int upload_pipe(opmgpu_handle h, const PipeProgram& p, PipeDevMem& d)
{
    int rc;
    if ((rc = upload(h, d.cta_step_ptr, p.cta_step_ptr))) return rc;
    if ((rc = upload(h, d.step_off16, p.step_off16))) return rc;
    if ((rc = expand_records(h, p.total_bytes, p.ibuf, p.step_ioff16, p.step_ilen, p.step_roff, d.step_off16.p, d.buf))) return rc;
    if ((rc = upload(h, d.step_bytes, p.step_bytes))) return rc;
    if ((rc = upload(h, d.step_rhs_row, p.step_rhs_row))) return rc;
    if ((rc = upload(h, d.step_rhs_bytes, p.step_rhs_bytes))) return rc;
    if ((rc = upload(h, d.cta_ext_base, p.cta_ext_base))) return rc;
    if ((rc = upload(h, d.val_src, p.val_src))) return rc;
    if ((rc = upload(h, d.val_dst8, p.val_dst8))) return rc;
    if ((rc = upload(h, d.val_stride, p.val_stride))) return rc;
    if ((rc = upload(h, d.perm_row, p.perm_row))) return rc;
    {
        // natural row -> position in this program's order (vector kernels write the next
        // right-hand side of the lower sweep in place, bicg_update_p_kernel)
        std::vector<int> pos((size_t)h->N_for_upload, 0);
        for (size_t q = 0; q < p.perm_row.size(); ++q)
            if (p.perm_row[q] >= 0) pos[p.perm_row[q]] = (int)q;
        if ((rc = upload(h, d.pos_of_row, pos))) return rc;
        CK(cudaStreamSynchronize(h->stream));          // pos is a local
    }
    d.nval = p.val_src.size(); d.next = (size_t)p.total_ext; d.P = p.P; d.nperm = (size_t)p.nperm; d.lean = p.lean;
    CK(d.ext.ensure(std::max<size_t>(d.next, 1) * 3));
    CK(cudaMemsetAsync(d.ext.p, 0xff, std::max<size_t>(d.next, 1) * 3 * sizeof(double), h->stream));
    CK(d.rhs_perm.ensure(std::max<size_t>(d.nperm, 2) * 3));
    CK(cudaMemsetAsync(d.rhs_perm.p, 0, std::max<size_t>(d.nperm, 2) * 3 * sizeof(double), h->stream));
    // ring geometry: as many stages as fit next to the dependency array
    d.stage_bytes = (p.max_step_bytes + 15) / 16 * 16;
    d.rhs_bytes = (p.max_step_rows * 24 + 15) / 16 * 16;
    d.cluster_size = p.cluster_size;
    d.cx_bytes = p.cluster_size > 1 ? (std::max(p.max_cx, 1) * 24 + 15) / 16 * 16 : 0;
    const size_t fixed = pipe_smem_bytes(0, 0, 0, d.cx_bytes);
    const size_t per_stage = (size_t)d.stage_bytes + d.rhs_bytes;
    int S = (size_t)h->max_smem_optin > fixed ? (int)(((size_t)h->max_smem_optin - fixed) / per_stage) : 0;
    d.nstages = std::min(S, kPipeMaxStages);
    if (const char* e = exp_env("OPMGPU_PIPE_STAGES")) d.nstages = std::max(kPipeGroups, std::min(d.nstages, std::atoi(e)));
    d.smem = pipe_smem_bytes(d.nstages, d.stage_bytes, d.rhs_bytes, d.cx_bytes);
    return 0;
}

int upload_factor_pipe(opmgpu_handle h, const FactorPipeProgram& p, FactorPipeDevMem& d)
{
    int rc;
    d.valid = false;
    if (!p.valid) return 0;
    d.stage_bytes = (p.max_step_bytes + 15) / 16 * 16;
    const size_t fixed = factor_pipe_smem_bytes(0, 0);
    if ((size_t)h->max_smem_optin < fixed + 3 * (size_t)d.stage_bytes) return 0;
    d.nstages = std::min((int)(((size_t)h->max_smem_optin - fixed) / (size_t)d.stage_bytes), kFMaxStages);
    d.smem = factor_pipe_smem_bytes(d.nstages, d.stage_bytes);
    if ((rc = upload(h, d.cta_step_ptr, p.cta_step_ptr))) return rc;
    if ((rc = upload(h, d.step_off16, p.step_off16))) return rc;
    if ((rc = expand_records(h, p.total_bytes, p.ibuf, p.step_ioff16, p.step_ilen, p.step_roff, d.step_off16.p, d.buf))) return rc;
    if ((rc = upload(h, d.step_bytes, p.step_bytes))) return rc;
    if ((rc = upload(h, d.cta_ext_base, p.cta_ext_base))) return rc;
    if ((rc = upload(h, d.val_src, p.val_src))) return rc;
    if ((rc = upload(h, d.val_dst8, p.val_dst8))) return rc;
    if ((rc = upload(h, d.cta_row_base, p.cta_row_base))) return rc;
    if ((rc = upload(h, d.fpos, p.fpos))) return rc;
    CK(d.fout.ensure((p.fpos.size() + 2) * kFEntry));
    d.nval = p.val_src.size(); d.next = (size_t)p.total_ext; d.P = p.P;
    CK(d.ext.ensure(std::max<size_t>(d.next, 1) * 9));
    CK(cudaMemsetAsync(d.ext.p, 0xff, std::max<size_t>(d.next, 1) * 9 * sizeof(double), h->stream));
    d.valid = true;
    return 0;
}

#ifdef OPMGPU_EXPERIMENTS
int upload_col(opmgpu_handle h, const ColProgram& p, ColDevMem& d)
{
    int rc;
    d.valid = false;
    d.g = p.g; d.P = p.P; d.nperm = (size_t)p.nperm; d.next = (size_t)p.next;
    d.nvalL = p.valL_src.size(); d.nvalU = p.valU_src.size();
    d.nstagesL = col_stage_count(p.g, false, (size_t)h->max_smem_optin);
    d.nstagesU = col_stage_count(p.g, true, (size_t)h->max_smem_optin);
    if (const char* e = exp_env("OPMGPU_COL_STAGES")) {
        d.nstagesL = std::max(2, std::min(d.nstagesL, std::atoi(e))); d.nstagesU = std::max(2, std::min(d.nstagesU, std::atoi(e)));
    }
    if (d.nstagesL < 2 || d.nstagesU < 2) return 0;
    d.smemL = col_smem_fixed(p.g) + (size_t)p.g.W * d.nstagesL * col_stage_bytes(false);
    d.smemU = col_smem_fixed(p.g) + (size_t)p.g.W * d.nstagesU * col_stage_bytes(true);
    if ((rc = upload(h, d.tile_ptr, p.cta_tile_ptr))) return rc;
    if ((rc = upload(h, d.tilesL, p.cta_tilesL))) return rc;
    if ((rc = upload(h, d.tilesU, p.cta_tilesU))) return rc;
    if ((rc = upload(h, d.valL_src, p.valL_src))) return rc;
    if ((rc = upload(h, d.valU_src, p.valU_src))) return rc;
    if ((rc = upload(h, d.valL_dst, p.valL_dst))) return rc;
    if ((rc = upload(h, d.valU_dst, p.valU_dst))) return rc;
    if ((rc = upload(h, d.perm_row, p.perm_rowL))) return rc;
    {
        // natural row -> lane-step of the lower sweep (the vector kernels write its right-hand side in place)
        std::vector<int> pos((size_t)h->N_for_upload, 0);
        for (size_t q = 0; q < p.perm_rowL.size(); ++q)
            if (p.perm_rowL[q] >= 0) pos[p.perm_rowL[q]] = (int)q;
        if ((rc = upload(h, d.pos_of_row, pos))) return rc;
        CK(cudaStreamSynchronize(h->stream));          // pos is a local
    }
    // records: zero wherever a block (or a whole lane-step) does not exist; values arrive per factorisation
    CK(d.recL.ensure(d.nperm * kColNCL)); CK(d.recU.ensure(d.nperm * kColNCU));
    CK(cudaMemsetAsync(d.recL.p, 0, d.nperm * kColNCL * sizeof(double), h->stream));
    CK(cudaMemsetAsync(d.recU.p, 0, d.nperm * kColNCU * sizeof(double), h->stream));
    CK(d.rhsL.ensure(d.nperm * 3)); CK(d.rhsU.ensure(d.nperm * 3));
    CK(cudaMemsetAsync(d.rhsL.p, 0, d.nperm * 3 * sizeof(double), h->stream));
    CK(cudaMemsetAsync(d.rhsU.p, 0, d.nperm * 3 * sizeof(double), h->stream));
    CK(d.extL.ensure(std::max<size_t>(d.next, 1) * 3)); CK(d.extU.ensure(std::max<size_t>(d.next, 1) * 3));
    CK(cudaMemsetAsync(d.extL.p, 0xff, std::max<size_t>(d.next, 1) * 3 * sizeof(double), h->stream));
    CK(cudaMemsetAsync(d.extU.p, 0xff, std::max<size_t>(d.next, 1) * 3 * sizeof(double), h->stream));
    d.valid = true;
    return 0;
}
#endif

PipeDev pipe_dev(const PipeDevMem& d)
{
    PipeDev p;
    p.buf = d.buf.p; p.cta_step_ptr = d.cta_step_ptr.p; p.step_off16 = d.step_off16.p; p.step_bytes = d.step_bytes.p;
    p.step_rhs_row = d.step_rhs_row.p; p.step_rhs_bytes = d.step_rhs_bytes.p;
    p.cta_ext_base = d.cta_ext_base.p; p.ext = d.ext.p;
    p.stage_bytes = d.stage_bytes; p.rhs_bytes = d.rhs_bytes; p.nstages = d.nstages;
    p.trace = nullptr; p.trace_cta = -1; p.gtrace = nullptr; p.gtrace_steps = 0;
    p.dbg = exp_env("OPMGPU_SDEBUG") ? atoi(exp_env("OPMGPU_SDEBUG")) : 0;
    p.cluster_size = d.cluster_size; p.cx_bytes = d.cx_bytes;
    return p;
}

SweepDev sweep_dev(const ProgramDevMem& d)
{
    SweepDev s;
    s.cta_step_ptr = d.cta_step_ptr.p; s.step_row_ptr = d.step_row_ptr.p; s.prow = d.prow.p;
    s.pblk_ptr = d.pblk_ptr.p; s.pcol = d.pcol.p; s.publish = d.publish.p;
    s.pval = d.pval.p; s.pdinv = d.pdinv.p;
    return s;
}

template <class T> inline T* vec(DevArr<double>& a) { return reinterpret_cast<T*>(a.p); }
template <class T> inline ncclDataType_t nccl_type() { return sizeof(T) == 4 ? ncclFloat : ncclDouble; }

// double <-> instance scalar type, device to device
template <class TI, class TO>
int convert(opmgpu_handle h, size_t n, const TI* in, TO* out)
{
    if (!n) return 0;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 16);
    convert_kernel<TI, TO><<<grid, 256, 0, h->stream>>>(n, in, out);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <class T>
int encode(opmgpu_handle h, size_t n, const T* in, double* out)
{
    if (!n) return 0;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 16);
    encode_kernel<T><<<grid, 256, 0, h->stream>>>(n, in, out);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int ensure_vectors(opmgpu_handle h)
{
    const size_t n = (size_t)h->N * std::max(3, h->np_req);          // np = 4..6: longer vectors
    CK(h->d_x.ensure(n)); CK(h->d_r.ensure(n)); CK(h->d_rt.ensure(n)); CK(h->d_p.ensure(n));
    CK(h->d_v.ensure(n)); CK(h->d_t.ensure(n)); CK(h->d_y.ensure(n + (size_t)h->n_ghost * 3)); CK(h->d_yL.ensure(n));
    CK(h->d_vU.ensure(n)); CK(h->d_tmp.ensure(n + (size_t)h->n_ghost * 3)); CK(h->d_tmp2.ensure(n));
    return 0;
}

// Multicolour variant: colours, permutation, permuted pattern and the operand layouts of the sweeps;
// none of the natural-order sweep programs is built.
int set_pattern_mc(opmgpu_handle h, int N, int nnzb, const int* rowptr, const int* colidx)
{
    if (h->np_req != 3) return h->bad("the multicolour ILU0 variant exists for 3x3 blocks");
    for (int i = 0; i < N; ++i) {
        bool have_diag = false;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) have_diag |= colidx[k] == i;
        if (!have_diag) {
            h->err = "diagonal entry missing in block row " + std::to_string(i);
            return OPMGPU_SINGULAR_BLOCK;
        }
    }
    McProgram m;
    const bool lines = h->ilu_order_req == OPMGPU_ILU_MULTICOLOUR_LINES;
    if (!build_mc_program(N, rowptr, colidx, m, lines))
        return h->bad("the k-line ordering needs a Cartesian stencil pattern in natural numbering whose only same-colour "
                      "couplings are the vertical ones (use OPMGPU_ILU_MULTICOLOUR for general patterns)");
    h->use_pipe = false; h->use_col = false; h->cluster_size = 1;
    h->pipeL.release(); h->pipeU.release(); h->pipeF.release(); h->progL.release(); h->progU.release(); h->col.release();
    h->d_lu.release();                  // built on demand by opmgpu_ilu0_get_factors
    h->lvlU_ptr.clear();
    h->N = N; h->nnzb = nnzb; h->n_ghost = 0;
    h->N_for_upload = N;
    CK(h->d_rowptr.ensure((size_t)N + 1 + 8));
    CK(h->d_colidx.ensure((size_t)nnzb + 8));
    CK(cudaMemcpyAsync(h->d_rowptr.p, rowptr, sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_colidx.p, colidx, sizeof(int) * (size_t)nnzb, cudaMemcpyHostToDevice, h->stream));
    McDevMem& d = h->mc;
    d.valid = false;
    // the sweeps' bulk copies round sizes up to 16 bytes: 8 ints of slack behind the index arrays
    auto up = [&](DevArr<int>& dst, const std::vector<int>& v) -> int {
        CK(dst.ensure(v.size() + 8));
        if (!v.empty()) CK(cudaMemcpyAsync(dst.p, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice, h->stream));
        return 0;
    };
    int rc;
    if ((rc = up(d.p2n, m.ord.p2n)) || (rc = up(d.pair_ptr, m.pair_ptr)) || (rc = up(d.pair_jk, m.pair_jk)) || (rc = up(d.pair_ik, m.pair_ik)) ||
        (rc = up(d.psrc, m.psrc)) || (rc = up(d.ppos, m.ppos)) || (rc = up(d.Lrowptr, m.Lrowptr)) || (rc = up(d.Lcol, m.Lcol)) ||
        (rc = up(d.Urowptr, m.Urowptr)) || (rc = up(d.Ucol, m.Ucol)) || (rc = up(d.lvl_rows, m.lvl_rows))) return rc;
    CK(d.uni.ensure((size_t)m.total_blocks * 9));
    CK(cudaMemsetAsync(d.uni.p, 0, sizeof(double) * (size_t)m.total_blocks * 9, h->stream));
    if ((rc = ensure_vectors(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    d.ncolours = m.ord.ncolours;
    d.nnzL = m.Lrowptr[N]; d.nnzU = m.Urowptr[N]; d.offD = m.offD; d.offU = m.offU; d.total_blocks = m.total_blocks;
    d.colour_ptr = m.ord.colour_ptr; d.lvl_ptr = m.lvl_ptr;
    d.p2n_host = std::move(m.ord.p2n); d.n2p_host = std::move(m.ord.n2p);
    d.lines = lines; d.nz = m.ord.nz; d.nx = m.ord.nx;
    for (int c = 0; c < 2; ++c) {
        d.ncols[c] = m.ord.ncols[c]; d.base[c] = m.ord.base[c];
        const int per_cta = (d.ncols[c] + h->sm_count - 1) / std::max(1, h->sm_count);
        d.rb[c] = std::min(kSpmvRows, std::max(4, (per_cta + 3) / 4 * 4));
    }
    d.valid = true;
    h->nlevL = h->nlevU = lines ? d.ncolours * d.nz : d.ncolours;
    h->have_pattern = true;
    return OPMGPU_OK;
}

int set_pattern(opmgpu_handle h, int N, int nnzb, const int* rowptr, const int* colidx)
{
    if (N < 1 || nnzb < N || rowptr[0] != 0 || rowptr[N] != nnzb) return h->bad("bad BCRS pattern sizes");
    for (int i = 0; i < N; ++i) {
        if (rowptr[i + 1] < rowptr[i]) return h->bad("rowptr not monotone");
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (colidx[k] < 0 || colidx[k] >= N) return h->bad("column index out of range");
            if (k > rowptr[i] && colidx[k] <= colidx[k - 1]) return h->bad("columns not strictly ascending in a row");
        }
    }
    h->have_pattern = h->have_values = h->have_factors = false;
    h->operator_only = false;
    if (h->ilu_order_req != OPMGPU_ILU_NATURAL) return set_pattern_mc(h, N, nnzb, rowptr, colidx);
    h->mc.release();
    analyse_pattern(N, rowptr, colidx, h->sweep_ctas, h->an, h->force_simple, &h->caps);
    h->cluster_size = h->an.cluster_size;
    if (h->an.missing_diag_row >= 0) {
        h->err = "diagonal entry missing in block row " + std::to_string(h->an.missing_diag_row);
        return OPMGPU_SINGULAR_BLOCK;
    }
    h->N = N; h->nnzb = nnzb;
    CK(h->d_rowptr.ensure((size_t)N + 1 + 8));      // + 8: the SpMV's bulk copies round sizes up to 16 bytes
    CK(h->d_colidx.ensure((size_t)nnzb + 8));
    CK(cudaMemcpyAsync(h->d_rowptr.p, rowptr, sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_colidx.p, colidx, sizeof(int) * (size_t)nnzb, cudaMemcpyHostToDevice, h->stream));
    int rc;
    if ((rc = upload(h, h->d_diag, h->an.diag))) return rc;
    if ((rc = upload(h, h->d_lvl_rows, h->an.lvl_rows))) return rc;
    h->use_pipe = h->an.pipeL.valid && h->an.pipeU.valid;
    h->N_for_upload = N;
    h->use_col = false;
#ifdef OPMGPU_EXPERIMENTS
    if (h->use_pipe && h->allow_col && h->an.grid_nx > 0) {
        // exact Cartesian stencil: column-owned sweeps (colprog.hpp) instead of the general pipelined ones
        ColProgram cp;
        build_col_program(N, rowptr, colidx, h->an.diag, h->an.grid_nx, h->an.grid_ny, h->an.grid_nz, h->sm_count,
                          (size_t)h->max_smem_optin, cp);
        if (cp.valid) {
            if ((rc = upload_col(h, cp, h->col))) return rc;
            h->use_col = h->col.valid;
        }
    }
#endif
    if (!h->use_col) h->col.release();
    if (h->use_pipe && !h->use_col) {
        if ((rc = upload_pipe(h, h->an.pipeL, h->pipeL))) return rc;
        if ((rc = upload_pipe(h, h->an.pipeU, h->pipeU))) return rc;
        if ((h->pipeL.nstages < 3 || h->pipeU.nstages < 3) && h->cluster_size > 1) {
            // not enough shared memory left for the record ring next to the intra-cluster entries
            ClusterCaps none;
            analyse_pattern(N, rowptr, colidx, h->sweep_ctas, h->an, h->force_simple, &none);
            h->cluster_size = 1;
            if ((rc = upload_pipe(h, h->an.pipeL, h->pipeL))) return rc;
            if ((rc = upload_pipe(h, h->an.pipeU, h->pipeU))) return rc;
        }
        if (h->pipeL.nstages < 3 || h->pipeU.nstages < 3) h->use_pipe = false;
    }
    if (h->use_col) { h->pipeL.release(); h->pipeU.release(); h->cluster_size = 1; }
    h->pipeF.valid = false;
    if (h->use_pipe && !h->factor_tile && !h->factor_by_levels) {
        if ((rc = upload_factor_pipe(h, h->an.pipeF, h->pipeF))) return rc;
    }
    if (!h->pipeF.valid) h->pipeF.release();
    if (!h->use_pipe) {
        if (h->an.upper.prow.empty()) analyse_pattern(N, rowptr, colidx, h->sweep_ctas, h->an, true);
        h->cluster_size = 1;
        if ((rc = upload_program(h, h->an.lower, h->progL, false))) return rc;
        if ((rc = upload_program(h, h->an.upper, h->progU, true))) return rc;
        h->pipeL.release(); h->pipeU.release();
    } else {
        if (!h->pipeF.valid) {          // the tile kernel factorises: its program is built on demand
            build_tile_factor_program(rowptr, colidx, h->an);
            if ((rc = upload_program(h, h->an.lower, h->progL, false, false))) return rc;
        } else {
            h->progL.release();
        }
        h->progU.release();
    }
    h->lvlU_ptr.clear();
    if (h->np_req != 3) {
        // level sets of the upper triangle (rows from the last to the first)
        std::vector<int> lev(N, 0);
        int nlev = 0;
        for (int i = N - 1; i >= 0; --i) {
            int l = 0;
            for (int k = rowptr[i + 1] - 1; k >= rowptr[i] && colidx[k] > i; --k) l = std::max(l, lev[colidx[k]] + 1);
            lev[i] = l; nlev = std::max(nlev, l + 1);
        }
        h->lvlU_ptr.assign((size_t)nlev + 1, 0);
        for (int i = 0; i < N; ++i) ++h->lvlU_ptr[(size_t)lev[i] + 1];
        for (int l = 0; l < nlev; ++l) h->lvlU_ptr[l + 1] += h->lvlU_ptr[l];
        std::vector<int> rows(N), fill(h->lvlU_ptr.begin(), h->lvlU_ptr.end() - 1);
        for (int i = 0; i < N; ++i) rows[fill[lev[i]]++] = i;
        if ((rc = upload(h, h->d_lvlU_rows, rows))) return rc;
        CK(cudaStreamSynchronize(h->stream));
    }
    CK(h->d_lu.ensure((size_t)nnzb * std::max(9, h->np_req * h->np_req)));
    if ((rc = ensure_vectors(h))) return rc;
    CK(h->d_flags.ensure(N));
    CK(cudaMemsetAsync(h->d_flags.p, 0, sizeof(int) * (size_t)N, h->stream));
    h->epoch = 0;
    CK(cudaStreamSynchronize(h->stream));      // host vectors of the analysis may now be dropped
    // keep only what the host still needs
    h->nlevL = h->an.nlevL; h->nlevU = h->an.nlevU;
    h->an.lower = SweepProgram(); h->an.upper = SweepProgram();
    h->an.owner_ = std::vector<int>(); h->an.level_lower_ = std::vector<int>();
    h->an.pipeL = PipeProgram(); h->an.pipeU = PipeProgram(); h->an.pipeF = FactorPipeProgram();
    h->have_pattern = true;
    return OPMGPU_OK;
}

// gather kernels of the distributed path
template <class T>
__global__ void __launch_bounds__(256)
pack_rows_kernel(int n, const int* __restrict__ rows, const T* __restrict__ x, T* __restrict__ buf)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * 3) return;
    const int k = e / 3;
    buf[e] = x[(size_t)rows[k] * 3 + (e - k * 3)];
}
template <class T>
__global__ void __launch_bounds__(256)
gather_blocks_kernel(size_t nblk, const int* __restrict__ src, const T* __restrict__ vals, double* __restrict__ lu)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nblk * 9) return;
    const size_t b = e / 9;
    lu[e] = enc(vals[(size_t)src[b] * 9 + (e - b * 9)]);
}

// x[N_local .. N_local + n_ghost) <- the owners' rows (ncclSend/ncclRecv over NVLink)
template <class T>
int halo_exchange(opmgpu_handle h, T* x, cudaStream_t stream = nullptr)
{
    if (h->world == 1) return 0;
    if (!stream) stream = h->stream;
    T* sendbuf = vec<T>(h->d_sendbuf);
    if (h->n_send) {
        pack_rows_kernel<T><<<(h->n_send * 3 + 255) / 256, 256, 0, stream>>>(h->n_send, h->d_send_rows.p, x, sendbuf);
        h->launches++;
    }
    NK(g_nccl.GroupStart());
    for (int p = 0; p < h->world; ++p) {
        if (h->send_cnt[p]) NK(g_nccl.Send(sendbuf + (size_t)h->send_off[p] * 3, (size_t)h->send_cnt[p] * 3, nccl_type<T>(), p, h->comm, stream));
        if (h->recv_cnt[p]) NK(g_nccl.Recv(x + ((size_t)h->N + h->recv_off[p]) * 3, (size_t)h->recv_cnt[p] * 3, nccl_type<T>(), p, h->comm, stream));
    }
    NK(g_nccl.GroupEnd());
    return 0;
}

// S[slot .. slot+count) <- sum over the ranks, through peer memory (see PeerScalDesc); optionally the
// whole scalar block goes to the host mailbox afterwards (what publish_scalars_kernel does)
__global__ void __launch_bounds__(128)
peer_allreduce_kernel(PeerScalDesc d, double* S, int slot, int count, unsigned long long epoch, HostBox hb, int publish, int* err)
{
    const int t = threadIdx.x, W = d.world, me = d.rank;
    const int par = (int)(epoch & 1ull);
    const int p = t / kScalSlots, i = t % kScalSlots;
    if (p < W && i < count) d.inbox[p][((size_t)par * kMaxPeers + me) * kScalSlots + i] = S[slot + i];
    __threadfence_system();
    __syncthreads();
    if (t < W && t != me) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(d.flags[t] + me), "l"(epoch) : "memory");
        unsigned long long v = 0;
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(d.flags[me] + t) : "memory");
            if (v >= epoch) break;
            if (++spins > (1u << 26)) { atomicExch(err, 11); break; }
        }
    }
    __syncthreads();
    if (t < count) {
        const double* in = d.inbox[me] + (size_t)par * kMaxPeers * kScalSlots;
        double sum = 0.0;
        for (int q = 0; q < W; ++q) {
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(in + (size_t)q * kScalSlots + t) : "memory");
            sum += v;
        }
        S[slot + t] = sum;
    }
    if (publish) {
        __syncthreads();
        if (t < S_COUNT) hb.hS[t] = S[t];
        __syncthreads();
        if (t == 0) {
            *hb.herr = S[S_ERRW] != 0.0 ? 9 : 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(hb.hseq) = hb.seq;
        }
    }
}
static_assert(kScalSlots == S_COUNT, "inbox rows hold the scalar block");
static_assert(kMaxPeers * kScalSlots <= 128, "one thread per (peer, slot)");

int allreduce_slots(opmgpu_handle h, int slot, int count, const HostBox* publish = nullptr)
{
    if (h->world == 1) return 0;
    if (h->peer.ready) {
        HostBox hb;
        if (publish) hb = *publish; else { hb.hS = nullptr; hb.herr = nullptr; hb.hseq = nullptr; hb.derr = nullptr; hb.seq = 0; }
        peer_allreduce_kernel<<<1, 128, 0, h->stream>>>(h->peer.sdesc, h->d_S.p, slot, count, ++h->peer.sepoch, hb, publish ? 1 : 0, h->d_err.p);
        h->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    NK(g_nccl.AllReduce(h->d_S.p + slot, h->d_S.p + slot, (size_t)count, ncclDouble, ncclSum, h->comm, h->stream));
    if (publish) {
        publish_scalars_kernel<<<1, 32, 0, h->stream>>>(h->d_S.p, *publish);
        h->launches++;
    }
    return 0;
}

template <class T>
int launch_spmv(opmgpu_handle h, int mode, const T* x, T* y, const T* w1)
{
    const T* vals = static_cast<const T*>(h->d_vals);
    const int* rowptr = h->world > 1 ? h->d_rowptr_full.p : h->d_rowptr.p;
    const int* colidx = h->world > 1 ? h->d_colidx_full.p : h->d_colidx.p;
    if (h->spmv_tma && (reinterpret_cast<uintptr_t>(h->d_vals) & 15) == 0) {
        const int nnzb = h->world > 1 ? h->nnzb_full : h->nnzb;
        const int ntiles = (h->N + kSpmvRows - 1) / kSpmvRows;
        const unsigned grid = (unsigned)std::min(ntiles, h->sm_count);
        static const int rowt_env = exp_env("OPMGPU_SPMV_ROWT") ? atoi(exp_env("OPMGPU_SPMV_ROWT")) : -1;
        const bool rowt = rowt_env >= 0 ? rowt_env != 0 : sizeof(T) == 4;      // thread per row: the float instance's default
        if (sizeof(T) == 4 && rowt_env < 0) {
            // float instance: 128-row tiles, one thread per row
            const unsigned grid2 = (unsigned)std::min((h->N + kSpmvRowsF32 - 1) / kSpmvRowsF32, h->sm_count);
            if (mode == 0) spmv3_tma_kernel<0, float, true, kSpmvRowsF32><<<grid2, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, (const float*)vals, (const float*)x, (float*)y, (const float*)w1, h->d_S.p, h->ws());
            else if (mode == 1) spmv3_tma_kernel<1, float, true, kSpmvRowsF32><<<grid2, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, (const float*)vals, (const float*)x, (float*)y, (const float*)w1, h->d_S.p, h->ws());
            else spmv3_tma_kernel<2, float, true, kSpmvRowsF32><<<grid2, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, (const float*)vals, (const float*)x, (float*)y, (const float*)w1, h->d_S.p, h->ws());
        } else if (rowt) {
            if (mode == 0) spmv3_tma_kernel<0, T, true><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
            else if (mode == 1) spmv3_tma_kernel<1, T, true><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
            else spmv3_tma_kernel<2, T, true><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
        }
        else if (mode == 0) spmv3_tma_kernel<0, T><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
        else if (mode == 1) spmv3_tma_kernel<1, T><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
        else spmv3_tma_kernel<2, T><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
        h->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    const long long threads = 3LL * h->N;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (mode != 0 && grid > kMaxRedBlocks) {
        // fused reduction epilogue has a bounded partial array: fall back to SpMV + dot kernels
        spmv3_kernel<0, T><<<grid, 256, 0, h->stream>>>(h->N, rowptr, colidx, vals, x, y, (const T*)nullptr, h->d_S.p, h->ws());
        h->launches++;
        return -100;
    }
    if (mode == 0) spmv3_kernel<0, T><<<grid, 256, 0, h->stream>>>(h->N, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
    else if (mode == 1) spmv3_kernel<1, T><<<grid, 256, 0, h->stream>>>(h->N, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
    else spmv3_kernel<2, T><<<grid, 256, 0, h->stream>>>(h->N, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

// separate (non-fused) dots for systems too large for the fused epilogue
template <class T>
__global__ void dot_to_slot_kernel(size_t n, const T* __restrict__ a, const T* __restrict__ b,
                                   double* S, int slot, ReduceWs ws)
{
    T v[1] = {T(0)};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v[0] = fma(a[i], b[i], v[0]);
    grid_reduce<1, T>(v, ws, [=](T (&t)[1]) { S[slot] = t[0]; });
}

// my boundary rows -> the neighbours' ghost rows (peer memory), then the epoch flag of every neighbour
template <class T>
__global__ void __launch_bounds__(256)
halo_push_kernel(PeerPushDesc d, int n_send, const int* __restrict__ rows, const T* __restrict__ x,
                 unsigned long long epoch, unsigned* ticket)
{
    __shared__ bool is_last;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)n_send * 3; e += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(e / 3), c = (int)(e - (size_t)k * 3);
        int p = 0;
        while (p + 1 < d.world && k >= d.send_off[p + 1]) ++p;
        T* dst = static_cast<T*>(d.y[p]) + (size_t)(d.dst_row0[p] + (k - d.send_off[p])) * 3 + c;
        *dst = x[(size_t)rows[k] * 3 + c];
    }
    __threadfence_system();                   // my stores before the ticket
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (is_last && threadIdx.x < d.world) {
        const int p = threadIdx.x;
        if (d.send_off[p + 1] > d.send_off[p]) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(d.flags[p] + d.rank), "l"(epoch) : "memory");
        }
    }
}
// wait until every neighbour has delivered its rows of this exchange (bounded; raises the watchdog word)
__global__ void halo_wait_kernel(const unsigned long long* flags_in, unsigned recv_mask, unsigned long long epoch, int* err)
{
    const int p = threadIdx.x;
    if (p < kMaxPeers && ((recv_mask >> p) & 1u)) {
        unsigned long long v = 0;
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags_in + p) : "memory");
            if (v >= epoch) break;
            if (++spins > (1u << 26)) { atomicExch(err, 10); break; }
        }
    }
}

template <class T>
int halo_exchange_peer(opmgpu_handle h, T* x)
{
    const unsigned long long epoch = ++h->peer.epoch;
    if (h->n_send) {
        const unsigned grid = (unsigned)std::min<size_t>(((size_t)h->n_send * 3 + 255) / 256, 64);
        halo_push_kernel<T><<<grid, 256, 0, h->stream>>>(h->peer.desc, h->n_send, h->d_send_rows.p, x, epoch, h->peer.ticket.p);
        h->launches++;
    }
    if (h->peer.recv_mask) {
        halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->peer.flags_in.p, h->peer.recv_mask, epoch, h->d_err.p);
        h->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

// Row-partitioned SpMV with the halo exchange in flight beside it ("fused
// SpMV-with-halo step"): pack + ncclSend/ncclRecv run on halo_stream while the tiles that need no
// ghost value are multiplied; the boundary tiles follow when the ghosts have arrived and add their
// share of the fused dot products.  A few SMs are left to the exchange kernels.
template <class T>
int spmv_overlapped(opmgpu_handle h, int mode, T* x, T* y, const T* w1)
{
    const T* vals = static_cast<const T*>(h->d_vals);
    CK(cudaEventRecord(h->ev_x_ready, h->stream));
    CK(cudaStreamWaitEvent(h->halo_stream, h->ev_x_ready, 0));
    if (int rc = halo_exchange<T>(h, x, h->halo_stream)) return rc;
    CK(cudaEventRecord(h->ev_halo_done, h->halo_stream));
    const int* rowptr = h->d_rowptr_full.p; const int* colidx = h->d_colidx_full.p;
    const int nnzb = h->nnzb_full;
    static const int reserve = exp_env("OPMGPU_HALO_RESERVE") ? atoi(exp_env("OPMGPU_HALO_RESERVE")) : 16;     // SMs left to the pack kernel and NCCL's send/recv kernel (32 p2p channels)
    const int ntiles = (h->N + kSpmvRows - 1) / kSpmvRows;
    const unsigned grid = (unsigned)std::max(1, std::min(ntiles, h->sm_count - reserve));
    const unsigned long long* skip = h->d_row_skip.p;
    if (mode == 0) spmv3_tma_kernel<0, T><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws(), skip);
    else if (mode == 1) spmv3_tma_kernel<1, T><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws(), skip);
    else spmv3_tma_kernel<2, T><<<grid, kSpmvThreads, kSpmvSmemBytes, h->stream>>>(h->N, nnzb, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws(), skip);
    CK(cudaStreamWaitEvent(h->stream, h->ev_halo_done, 0));
    const unsigned gridb = (unsigned)std::max<size_t>(1, std::min<size_t>(((size_t)h->n_bnd_rows * 3 + 255) / 256, (size_t)kVecBlocks));
    if (mode == 0) spmv3_rows_kernel<0, T><<<gridb, 256, 0, h->stream>>>(h->n_bnd_rows, h->d_bnd_rows.p, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
    else if (mode == 1) spmv3_rows_kernel<1, T><<<gridb, 256, 0, h->stream>>>(h->n_bnd_rows, h->d_bnd_rows.p, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
    else spmv3_rows_kernel<2, T><<<gridb, 256, 0, h->stream>>>(h->n_bnd_rows, h->d_bnd_rows.p, rowptr, colidx, vals, x, y, w1, h->d_S.p, h->ws());
    h->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

// np = 2 (generic_np.cuh): y = A x, then the dot products as separate passes
template <class T>
int np_spmv_with_dots(opmgpu_handle h, int mode, const T* x, T* y, const T* w1)
{
    const size_t n = (size_t)h->N * h->np;
    NP_DISPATCH(h->np, (np_spmv_kernel<NPV, T><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->N, h->d_rowptr.p, h->d_colidx.p,
                                                                            static_cast<const T*>(h->d_vals), x, y)));
    h->launches++;
    if (mode == 1) {
        dot_to_slot_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, w1, (const T*)y, h->d_S.p, S_H, h->ws());
        h->launches++;
    } else if (mode == 2) {
        dot_to_slot_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, (const T*)y, w1, h->d_S.p, S_TR, h->ws());
        dot_to_slot_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, (const T*)y, (const T*)y, h->d_S.p, S_TT, h->ws());
        h->launches += 2;
    }
    CK(cudaGetLastError());
    return 0;
}

template <class T>
int spmv_with_dots(opmgpu_handle h, int mode, T* x, T* y, const T* w1)
{
    int rc;
    if (h->np != 3) return np_spmv_with_dots<T>(h, mode, x, y, w1);
    if (h->world > 1 && h->overlap_halo && h->halo_stream && h->spmv_tma && h->n_bnd_rows > 0 &&
        (reinterpret_cast<uintptr_t>(h->d_vals) & 15) == 0) {
        rc = spmv_overlapped<T>(h, mode, x, y, w1);
        if (rc) return rc;
        if (mode == 1) rc = allreduce_slots(h, S_H, 1);
        if (mode == 2) rc = allreduce_slots(h, S_TR, 2);
        return rc;
    }
    // Peer-memory exchange for the solver's own SpMVs: x is the handle's d_y (the buffer the
    // neighbours have mapped) and an all-reduce follows every such SpMV, so no rank can overwrite
    // ghost rows a neighbour is still reading.  Everything else takes ncclSend/ncclRecv.
    if (h->world > 1 && h->peer.ready && mode != 0 && x == vec<T>(h->d_y)) rc = halo_exchange_peer<T>(h, x);
    else rc = halo_exchange<T>(h, x);
    if (rc) return rc;
    rc = launch_spmv<T>(h, mode, x, y, w1);
    if (rc == -100) {
        const size_t n = (size_t)h->N * 3;
        if (mode == 1) {
            dot_to_slot_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, w1, (const T*)y, h->d_S.p, S_H, h->ws());
            h->launches++;
        } else {
            dot_to_slot_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, (const T*)y, w1, h->d_S.p, S_TR, h->ws());
            dot_to_slot_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, (const T*)y, (const T*)y, h->d_S.p, S_TT, h->ws());
            h->launches += 2;
        }
        CK(cudaGetLastError());
        rc = 0;
    }
    if (rc == 0 && mode == 1) rc = allreduce_slots(h, S_H, 1);
    if (rc == 0 && mode == 2) rc = allreduce_slots(h, S_TR, 2);
    return rc;
}

int sweep_watchdog(opmgpu_handle h);

// T: scalar type of the instance (matrix values are stored as T, the arithmetic is T's; the factor
// array, the records and the pivots are 8-byte containers)
// np = 2: Dune::bilu0_decomposition level by level on a T copy of the values (generic_np.cuh)
template <class T>
int np_factor(opmgpu_handle h, int* bad_row)
{
    if (!h->have_values) return h->bad("no matrix values set");
    if (h->lvlU_ptr.empty()) return h->bad("the pattern was not prepared for this block size: opmgpu_set_block_size before opmgpu_set_pattern_bcrs");
    const size_t nv = (size_t)h->nnzb * h->np * h->np;
    T* lu = reinterpret_cast<T*>(h->d_lu.p);
    CK(cudaMemcpyAsync(lu, h->d_vals, nv * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    const int big = 0x7fffffff;
    CK(cudaMemcpyAsync(h->d_bad.p, &big, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    const std::vector<int>& lp = h->an.lvl_ptr;
    for (size_t l = 0; l + 1 < lp.size(); ++l) {
        const int n = lp[l + 1] - lp[l];
        if (n <= 0) continue;
        NP_DISPATCH(h->np, (np_factor_level_kernel<NPV, T><<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_lvl_rows.p, lp[l], lp[l + 1], h->d_rowptr.p,
                                                                              h->d_colidx.p, h->d_diag.p, lu, h->d_bad.p)));
        h->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h->h_flags2[1], h->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_flags2[1] != big) {
        if (bad_row) *bad_row = h->h_flags2[1];
        h->err = "singular diagonal block in ILU0 at block row " + std::to_string(h->h_flags2[1]);
        h->have_factors = false;
        return OPMGPU_SINGULAR_BLOCK;
    }
    if (bad_row) *bad_row = -1;
    h->have_factors = true;
    return OPMGPU_OK;
}

// Multicolour variant: A -> [ L | Dinv | U ] in permuted order, then bilu0_decomposition level by level
template <class T>
int mc_factor(opmgpu_handle h, int* bad_row)
{
    if (!h->have_values) return h->bad("no matrix values set");
    McDevMem& d = h->mc;
    T* uni = reinterpret_cast<T*>(d.uni.p);
    const size_t nscal = (size_t)h->nnzb * 9;
    mc_gather_values_kernel<T><<<(unsigned)((nscal + 255) / 256), 256, 0, h->stream>>>(nscal, d.psrc.p, d.ppos.p, static_cast<const T*>(h->d_vals), uni);
    h->launches++;
    const int big = 0x7fffffff;
    CK(cudaMemcpyAsync(h->d_bad.p, &big, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (d.lines) {
        for (int c = 0; c < d.ncolours; ++c) {
            if (d.ncols[c] <= 0) continue;
            mc_factor_lines_kernel<T><<<(d.ncols[c] + 31) / 32, 32, 0, h->stream>>>(d.base[c], d.ncols[c], d.nz, d.Lrowptr.p, d.Lcol.p, d.pair_ptr.p,
                                                                                    d.pair_jk.p, d.pair_ik.p, d.offD, uni, h->d_bad.p);
            h->launches++;
        }
    } else
    for (size_t l = 0; l + 1 < d.lvl_ptr.size(); ++l) {
        const int n = d.lvl_ptr[l + 1] - d.lvl_ptr[l];
        if (n <= 0) continue;
        mc_factor_level_kernel<T><<<(n + 127) / 128, 128, 0, h->stream>>>(d.lvl_rows.p, d.lvl_ptr[l], d.lvl_ptr[l + 1], d.Lrowptr.p, d.Lcol.p,
                                                                          d.pair_ptr.p, d.pair_jk.p, d.pair_ik.p, d.offD, uni, h->d_bad.p);
        h->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h->h_flags2[1], h->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_flags2[1] != big) {
        const int row = d.p2n_host[h->h_flags2[1]];          // reported in the caller's numbering
        if (bad_row) *bad_row = row;
        h->err = "singular diagonal block in ILU0 at block row " + std::to_string(row) + " (multicolour ordering)";
        h->have_factors = false;
        return OPMGPU_SINGULAR_BLOCK;
    }
    if (bad_row) *bad_row = -1;
    h->have_factors = true;
    return OPMGPU_OK;
}

template <class T>
int factor_t(opmgpu_handle h, int* bad_row)
{
    if (h->np != 3) return np_factor<T>(h, bad_row);
    if (h->mc.valid) return mc_factor<T>(h, bad_row);
    if (!h->have_values) return h->bad("no matrix values set");
    if (h->operator_only) return h->bad("the pattern was set operator-only (no ILU0 analysis): opmgpu_set_pattern_bcrs first");
    if (h->f32 && h->use_col) return h->bad("the column-owned sweeps (OPMGPU_COL=1) exist for the double instance only");
    const size_t nv = (size_t)h->nnzb * 9;
    const bool pipe_factor = h->pipeF.valid && !h->factor_by_levels;
    const T* vals = static_cast<const T*>(h->d_vals);
    h->lu_lazy = false;
    if (h->world > 1) {
        gather_blocks_kernel<T><<<(unsigned)((nv + 255) / 256), 256, 0, h->stream>>>((size_t)h->nnzb, h->d_lu_src.p, vals, h->d_lu.p);
        h->launches++;
    } else if (!pipe_factor) {
        if (sizeof(T) == 8) CK(cudaMemcpyAsync(h->d_lu.p, h->d_vals, nv * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        else if (int rc = encode<T>(h, nv, vals, h->d_lu.p)) return rc;
    }
    // the blocks of A the ILU0 is built on (the rank's diagonal block when partitioned: 8-byte containers)
    const bool A_in_lu = h->world > 1;
    const int big = 0x7fffffff;
    CK(cudaMemcpyAsync(h->d_bad.p, &big, sizeof(int), cudaMemcpyHostToDevice, h->stream));
#ifdef OPMGPU_EXPERIMENTS
    if (h->factor_by_levels) {
        const std::vector<int>& lp = h->an.lvl_ptr;
        for (size_t l = 0; l + 1 < lp.size(); ++l) {
            const int n = lp[l + 1] - lp[l];
            if (n <= 0) continue;
            ilu0_factor_level_kernel<T><<<(n + 127) / 128, 128, 0, h->stream>>>(
                h->d_lvl_rows.p, lp[l], lp[l + 1], h->d_rowptr.p, h->d_colidx.p, h->d_diag.p, h->d_lu.p, h->d_bad.p);
            h->launches++;
        }
    } else
#endif
    if (pipe_factor) {
        FactorPipeDevMem& d = h->pipeF;
        const size_t e = d.nval * 3;
        if (A_in_lu) pack_factor_records_kernel<double, true><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(d.nval, d.val_src.p, d.val_dst8.p, h->d_lu.p, (double*)d.buf.p);
        else pack_factor_records_kernel<T><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(d.nval, d.val_src.p, d.val_dst8.p, vals, (double*)d.buf.p);
        FactorPipeDev pg;
        pg.buf = d.buf.p; pg.cta_step_ptr = d.cta_step_ptr.p; pg.step_off16 = d.step_off16.p; pg.step_bytes = d.step_bytes.p;
        pg.cta_ext_base = d.cta_ext_base.p; pg.cta_row_base = d.cta_row_base.p; pg.ext = d.ext.p; pg.fout = d.fout.p;
        pg.stage_bytes = d.stage_bytes; pg.nstages = d.nstages;
        int* bad = h->d_bad.p; int* err = h->d_err.p;
        void* args[] = {&pg, &bad, &err};
        CK(cudaLaunchCooperativeKernel((void*)ilu0_factor_pipe_kernel<T>, dim3(d.P), dim3(kFThreads), args, d.smem, h->stream));
        h->launches += 2;
        h->lu_lazy = true;
    } else {
        FactorDev pg;
        pg.cta_step_ptr = h->progL.cta_step_ptr.p; pg.step_row_ptr = h->progL.step_row_ptr.p;
        pg.frow = (const int4*)h->progL.frow.p; pg.fent = (const int4*)h->progL.fent.p;
        pg.pair_jk = h->progL.pair_jk.p; pg.pair_ik = h->progL.pair_ik.p;
        pg.needs_flag = h->progL.needs_flag.p; pg.fpush_ptr = h->progL.fpush_ptr.p; pg.fpush_slot = h->progL.fpush_slot.p;
        pg.fslots = h->progL.fslots.p;
        double* lu = h->d_lu.p; int* flags = h->d_flags.p; int epoch = ++h->epoch; int* bad = h->d_bad.p; int* err = h->d_err.p;
        void* args[] = {&pg, &lu, &flags, &epoch, &bad, &err};
        CK(cudaLaunchCooperativeKernel((void*)ilu0_factor_tile_kernel<T>, dim3(h->progL.P), dim3(kFactorThreads), args, 0, h->stream));
        h->launches++;
    }
    CK(cudaGetLastError());
    // stream the factors into the sweep programs' layout
#ifdef OPMGPU_EXPERIMENTS
    if (h->use_col) {
        ColDevMem& c = h->col;
        const unsigned gridL = (unsigned)std::min<size_t>((c.nvalL * 3 + 255) / 256, (size_t)h->sm_count * 16);
        const unsigned gridU = (unsigned)std::min<size_t>((c.nvalU * 3 + 255) / 256, (size_t)h->sm_count * 16);
        if (pipe_factor) {
            const FactorPipeDevMem& f = h->pipeF;
            const double* ilu_A = A_in_lu ? h->d_lu.p : static_cast<const double*>(h->d_vals);      // (double instance only)
            if (c.nvalL) repack_col_kernel<true><<<gridL, 256, 0, h->stream>>>(c.nvalL, c.valL_src.p, c.valL_dst.p, h->d_colidx.p, f.fpos.p, ilu_A, f.fout.p, c.recL.p);
            repack_col_kernel<false><<<gridU, 256, 0, h->stream>>>(c.nvalU, c.valU_src.p, c.valU_dst.p, h->d_colidx.p, f.fpos.p, ilu_A, f.fout.p, c.recU.p);
        } else {
            if (c.nvalL) repack_col_from_lu_kernel<true><<<gridL, 256, 0, h->stream>>>(c.nvalL, c.valL_src.p, c.valL_dst.p, h->d_lu.p, c.recL.p);
            repack_col_from_lu_kernel<false><<<gridU, 256, 0, h->stream>>>(c.nvalU, c.valU_src.p, c.valU_dst.p, h->d_lu.p, c.recU.p);
        }
        h->launches += c.nvalL ? 2 : 1;
    } else
#endif
    if (h->use_pipe && pipe_factor) {
        // L_ij = A_ij * inv(D_j) is formed here, from A and the program-ordered pivots.  (Copying the
        // U blocks on a second stream beside the factorisation kernel was measured: it slows that
        // latency-bound kernel down by more than the copy costs, 0.84 -> 0.99 ms per factorisation.)
        const FactorPipeDevMem& f = h->pipeF;
        if (h->pipeL.nval) {
            const size_t e = h->pipeL.nval * 3;
            if (A_in_lu) repack_pipe2_kernel<true, 0, double, T, true><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->pipeL.nval, h->pipeL.val_src.p, h->pipeL.val_dst8.p,
                h->pipeL.val_stride.p, h->d_colidx.p, h->d_diag.p, f.fpos.p, h->d_lu.p, f.fout.p, (double*)h->pipeL.buf.p);
            else repack_pipe2_kernel<true, 0, T, T><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->pipeL.nval, h->pipeL.val_src.p, h->pipeL.val_dst8.p,
                h->pipeL.val_stride.p, h->d_colidx.p, h->d_diag.p, f.fpos.p, vals, f.fout.p, (double*)h->pipeL.buf.p);
            h->launches++;
        }
        if (h->pipeU.nval) {
            const size_t e = h->pipeU.nval * 3;
            if (A_in_lu) repack_pipe2_kernel<false, 0, double, T, true><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->pipeU.nval, h->pipeU.val_src.p, h->pipeU.val_dst8.p,
                h->pipeU.val_stride.p, h->d_colidx.p, h->d_diag.p, f.fpos.p, h->d_lu.p, f.fout.p, (double*)h->pipeU.buf.p);
            else repack_pipe2_kernel<false, 0, T, T><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->pipeU.nval, h->pipeU.val_src.p, h->pipeU.val_dst8.p,
                h->pipeU.val_stride.p, h->d_colidx.p, h->d_diag.p, f.fpos.p, vals, f.fout.p, (double*)h->pipeU.buf.p);
            h->launches++;
        }
    } else if (h->use_pipe) {
        for (PipeDevMem* d : {&h->pipeL, &h->pipeU}) {
            if (d->nval) {
                const size_t e = d->nval * 9;
                repack_pipe_kernel<<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(d->nval, d->val_src.p, d->val_dst8.p, d->val_stride.p, h->d_lu.p, (double*)d->buf.p);
                h->launches++;
            }
        }
    } else if (h->progL.nblk || h->progU.nblk || true) {
      if (h->progL.nblk) {
        const size_t e = h->progL.nblk * 9;
        repack_blocks_kernel<<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->progL.nblk, h->progL.psrc.p, h->d_lu.p, h->progL.pval.p);
        h->launches++;
    }
    if (h->progU.nblk) {
        const size_t e = h->progU.nblk * 9;
        repack_blocks_kernel<<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->progU.nblk, h->progU.psrc.p, h->d_lu.p, h->progU.pval.p);
        h->launches++;
    }
    {
        const size_t e = (size_t)h->N * 9;
        repack_dinv_kernel<<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->N, h->progU.prow.p, h->d_diag.p, h->d_lu.p, h->progU.pdinv.p);
        h->launches++;
    }
    }
    CK(cudaGetLastError());
    if (h->world > 1) {
        // a singular pivot or a watchdog trip on one rank must end the solve on every rank (the others
        // would otherwise wait for it in the next collective): agree on the verdict
        NK(g_nccl.AllReduce(h->d_bad.p, h->d_bad.p, 1, ncclInt32, ncclMin, h->comm, h->stream));
        NK(g_nccl.AllReduce(h->d_err.p, h->d_err.p, 1, ncclInt32, ncclMax, h->comm, h->stream));
    }
    CK(cudaMemcpyAsync(&h->h_flags2[1], h->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&h->h_flags2[0], h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_flags2[0]) return sweep_watchdog(h);
    if (h->h_flags2[1] != big) {
        if (bad_row) *bad_row = h->h_flags2[1];
        h->err = "singular diagonal block in ILU0 at block row " + std::to_string(h->h_flags2[1]);
        h->have_factors = false;
        return OPMGPU_SINGULAR_BLOCK;
    }
    if (bad_row) *bad_row = -1;
    h->have_factors = true;
    return OPMGPU_OK;
}
int factor(opmgpu_handle h, int* bad_row) { return h->f32 ? factor_t<float>(h, bad_row) : factor_t<double>(h, bad_row); }

// cooperative launch of a pipelined sweep, in thread-block clusters when the program asks for them
template <class T>
int launch_sweep(opmgpu_handle h, bool upper, const PipeDevMem& d, void** args)
{
    if (d.cluster_size > 1) {
        const void* fn = upper ? (const void*)ilu0_sweep_pipe_kernel<true, true, true, false, T> : (const void*)ilu0_sweep_pipe_kernel<false, true, true, false, T>;
#ifdef OPMGPU_EXPERIMENTS
        if (sizeof(T) == 8 && (h->gtrace_steps > 0 || (h->trace_cta >= 0 && h->d_trace.p)))                     // debug tools only
            fn = upper ? (const void*)ilu0_sweep_pipe_kernel<true, true, true, true> : (const void*)ilu0_sweep_pipe_kernel<false, true, true, true>;
#endif
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(d.P); cfg.blockDim = dim3(kPipeThreads); cfg.dynamicSmemBytes = d.smem; cfg.stream = h->stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = d.cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        // Co-residency: the grid was sized from cudaOccupancyMaxActiveClusters and the solver's
        // stream runs one kernel at a time.  The cooperative attribute would make the driver
        // verify that, but profilers (Nsight Compute 2025.2) cannot replay launches that carry both
        // the cluster and the cooperative attribute, so it is opt-in: OPMGPU_CLUSTER_COOP=1.
        // default: with the cooperative attribute, so the driver refuses the launch instead of letting
        // CTAs wait for peers that are not resident (shared stream, MPS neighbours).  OPMGPU_CLUSTER_COOP=0
        // drops it (profilers cannot replay such launches).  A refused cooperative cluster launch
        // is retried once without the attribute and remembered.
        static const bool coop_env = !(getenv("OPMGPU_CLUSTER_COOP") && atoi(getenv("OPMGPU_CLUSTER_COOP")) == 0);
        static bool coop_refused = false;
        cfg.attrs = at; cfg.numAttrs = coop_env && !coop_refused ? 2 : 1;
        cudaError_t le = cudaLaunchKernelExC(&cfg, fn, args);
        if (le != cudaSuccess && cfg.numAttrs == 2) {
            cudaGetLastError();
            coop_refused = true;
            cfg.numAttrs = 1;
            le = cudaLaunchKernelExC(&cfg, fn, args);
        }
        CK(le);
        return 0;
    }
    const void* fn = upper ? (d.lean ? (const void*)ilu0_sweep_pipe_kernel<true, true, false, false, T> : (const void*)ilu0_sweep_pipe_kernel<true, false, false, false, T>)
                           : (d.lean ? (const void*)ilu0_sweep_pipe_kernel<false, true, false, false, T> : (const void*)ilu0_sweep_pipe_kernel<false, false, false, false, T>);
#ifdef OPMGPU_EXPERIMENTS
    if (sizeof(T) == 8 && ((h->trace_cta >= 0 && h->d_trace.p) || h->gtrace_steps > 0))       // debug tools only
        fn = upper ? (d.lean ? (const void*)ilu0_sweep_pipe_kernel<true, true, false, true> : (const void*)ilu0_sweep_pipe_kernel<true, false, false, true>)
                   : (d.lean ? (const void*)ilu0_sweep_pipe_kernel<false, true, false, true> : (const void*)ilu0_sweep_pipe_kernel<false, false, false, true>);
#endif
    CK(cudaLaunchCooperativeKernel(fn, dim3(d.P), dim3(kPipeThreads), args, d.smem, h->stream));
    return 0;
}

// v = w U^-1 L^-1 d, all device pointers; asynchronous
// d_in_program_order: the producer of d already wrote it into pipeL.rhs_perm (fused permutation)
// np = 2: ParallelOverlappingILU0::apply level by level (generic_np.cuh)
template <class T>
int np_apply(opmgpu_handle h, double w, const T* d, T* v)
{
    const int scale = std::fabs(w - 1.0) > 1e-15 ? 1 : 0;
    const T* lu = reinterpret_cast<const T*>(h->d_lu.p);
    T* work = vec<T>(h->d_yL);
    const std::vector<int>& lp = h->an.lvl_ptr;
    for (size_t l = 0; l + 1 < lp.size(); ++l) {
        const int n = lp[l + 1] - lp[l];
        if (n <= 0) continue;
        NP_DISPATCH(h->np, (np_sweep_level_kernel<NPV, T, true><<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_lvl_rows.p, lp[l], lp[l + 1], h->d_rowptr.p,
            h->d_colidx.p, h->d_diag.p, lu, d, work, v, (T)w, scale)));
        h->launches++;
    }
    const std::vector<int>& up = h->lvlU_ptr;
    for (size_t l = 0; l + 1 < up.size(); ++l) {
        const int n = up[l + 1] - up[l];
        if (n <= 0) continue;
        NP_DISPATCH(h->np, (np_sweep_level_kernel<NPV, T, false><<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_lvlU_rows.p, up[l], up[l + 1], h->d_rowptr.p,
            h->d_colidx.p, h->d_diag.p, lu, d, work, v, (T)w, scale)));
        h->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

// Multicolour variant: v = w P^T U^-1 L^-1 P d, one pass per colour and direction; the first colour's
// lower pass is a gather, the last colour's lower and upper passes are one kernel
template <int KIND, class T>
int mc_launch(opmgpu_handle h, McSweepArgs& a)
{
    static const int rowt_env = exp_env("OPMGPU_MC_ROWT") ? atoi(exp_env("OPMGPU_MC_ROWT")) : -1;
    const void* fn;
    int rows;
    if (sizeof(T) == 4) { fn = (const void*)mc_sweep_tma_kernel<KIND, float, true, kSpmvRowsF32>; rows = kSpmvRowsF32; }
    else if (rowt_env == 1) { fn = (const void*)mc_sweep_tma_kernel<KIND, double, true, kSpmvRows>; rows = kSpmvRows; }
    else { fn = (const void*)mc_sweep_tma_kernel<KIND, double, false, kSpmvRows>; rows = kSpmvRows; }
    const int ntiles = (a.row1 + rows - 1) / rows - a.row0 / rows;
    if (ntiles <= 0) return 0;
    void* args[] = {&a};
    CK(cudaLaunchKernel(fn, dim3((unsigned)std::min(ntiles, h->sm_count)), dim3(kSpmvThreads), args, kMcSmemBytes, h->stream));
    h->launches++;
    return 0;
}

// k-line ordering: lower sweep colour by colour, upper sweep in the opposite order; a launch walks the planes
template <class T>
int mc_apply_lines(opmgpu_handle h, double w, const T* d, T* v)
{
    McDevMem& m = h->mc;
    const T* uni = reinterpret_cast<const T*>(m.uni.p);
    McLineArgs a;
    a.nz = m.nz; a.p2n = m.p2n.p; a.d = d; a.W = vec<T>(h->d_yL); a.dinv = uni + (size_t)m.offD * 9; a.out = v;
    a.w = w; a.scale = std::fabs(w - 1.0) > 1e-15 ? 1 : 0;
    for (int pass = 0; pass < 2 * m.ncolours; ++pass) {
        const bool upper = pass >= m.ncolours;
        const int c = upper ? 2 * m.ncolours - 1 - pass : pass;
        if (m.ncols[c] <= 0) continue;
        a.base = m.base[c]; a.ncols = m.ncols[c]; a.rb = m.rb[c];
        a.obase = m.base[1 - c]; a.oncols = m.ncolours == 2 ? m.ncols[1 - c] : 0; a.nx = m.nx; a.plane = m.ncols[0] + m.ncols[1];
        if (upper) { a.nnz = (int)m.nnzU; a.rowptr = m.Urowptr.p; a.colidx = m.Ucol.p; a.vals = uni + (size_t)m.offU * 9; }
        else { a.nnz = (int)m.nnzL; a.rowptr = m.Lrowptr.p; a.colidx = m.Lcol.p; a.vals = uni; }
        const void* fn = upper ? (const void*)mc_line_sweep_kernel<true, T> : (const void*)mc_line_sweep_kernel<false, T>;
        a.trace = nullptr;
        const bool trace = exp_env("OPMGPU_LINE_TRACE") != nullptr;          // (experiments build) per-step clock stamps of one warp
        if (trace) {
            CK(h->d_trace.ensure((size_t)m.nz * 8));
            CK(cudaMemsetAsync(h->d_trace.p, 0, sizeof(long long) * (size_t)m.nz * 8, h->stream));
            a.trace = h->d_trace.p;
        }
        void* args[] = {&a};
        CK(cudaLaunchKernel(fn, dim3((unsigned)((a.ncols + a.rb - 1) / a.rb)), dim3(kLineThreads), args, kLineSmemBytes, h->stream));
        h->launches++;
        if (trace) {
            std::vector<long long> tr((size_t)m.nz * 8);
            CK(cudaMemcpyAsync(tr.data(), h->d_trace.p, sizeof(long long) * tr.size(), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            fprintf(stderr, "[opmgpu] line sweep %s colour %d: step | start(rel) wait first-use chain exchange result arrive roll (cycles)\n", upper ? "upper" : "lower", c);
            for (int s2 = 20; s2 < std::min(m.nz, 28); ++s2) {
                const long long* q = &tr[(size_t)s2 * 8];
                fprintf(stderr, "  %3d | %7lld %5lld %5lld %5lld %5lld %5lld %5lld %5lld\n", s2, q[0] - tr[20 * 8], q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - q[4], q[6] - q[5], q[7] - q[6]);
            }
        }
    }
    return 0;
}

template <class T>
int mc_apply(opmgpu_handle h, double w, const T* d, T* v)
{
    McDevMem& m = h->mc;
    if (m.lines) return mc_apply_lines<T>(h, w, d, v);
    const int C = m.ncolours;
    const T* uni = reinterpret_cast<const T*>(m.uni.p);
    McSweepArgs a;
    a.N = h->N; a.p2n = m.p2n.p; a.d = d; a.W = vec<T>(h->d_yL); a.dinv = uni + (size_t)m.offD * 9; a.out = v;
    a.w = w; a.scale = std::fabs(w - 1.0) > 1e-15 ? 1 : 0;
    auto lower = [&](int c) { a.nnz = (int)m.nnzL; a.rowptr = m.Lrowptr.p; a.colidx = m.Lcol.p; a.vals = uni; a.row0 = m.colour_ptr[c]; a.row1 = m.colour_ptr[c + 1]; };
    auto upper = [&](int c) { a.nnz = (int)m.nnzU; a.rowptr = m.Urowptr.p; a.colidx = m.Ucol.p; a.vals = uni + (size_t)m.offU * 9; a.row0 = m.colour_ptr[c]; a.row1 = m.colour_ptr[c + 1]; };
    int rc;
    if (C > 1) {
        const size_t e = (size_t)(m.colour_ptr[1] - m.colour_ptr[0]) * 3;
        mc_copy_rows_kernel<T><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(m.colour_ptr[0], m.colour_ptr[1], m.p2n.p, d, vec<T>(h->d_yL));
        h->launches++;
        for (int c = 1; c + 1 < C; ++c) { lower(c); if ((rc = mc_launch<0, T>(h, a))) return rc; }
    }
    lower(C - 1);
    if ((rc = mc_launch<1, T>(h, a))) return rc;
    for (int c = C - 2; c >= 0; --c) { upper(c); if ((rc = mc_launch<2, T>(h, a))) return rc; }
    CK(cudaGetLastError());
    return 0;
}

template <class T>
int apply_precond(opmgpu_handle h, double w, const T* d, T* v, bool d_in_program_order = false)
{
    if (h->np != 3) return np_apply<T>(h, w, d, v);
    if (h->mc.valid) return mc_apply<T>(h, w, d, v);
    const int scale = std::fabs(w - 1.0) > 1e-15 ? 1 : 0;      // relaxation_ flag of the reference
    if (h->use_col && sizeof(T) == 4) return h->bad("the column-owned sweeps (OPMGPU_COL=1) exist for the double instance only");
#ifdef OPMGPU_EXPERIMENTS
    if (h->use_col) {
        const double* dd = reinterpret_cast<const double*>(d);      // (double instance only)
        double* vv = reinterpret_cast<double*>(v);
        ColDevMem& c = h->col;
        if (!d_in_program_order) {
            const size_t e = c.nperm * 3;
            permute_rows_kernel<double><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(c.nperm, c.perm_row.p, dd, c.rhsL.p);
        }
        for (int upper = 0; upper < 2; ++upper) {
            ColDev pg;
            pg.g = c.g; pg.rec = upper ? c.recU.p : c.recL.p; pg.cta_tile_ptr = c.tile_ptr.p;
            pg.cta_tiles = upper ? c.tilesU.p : c.tilesL.p; pg.ext = upper ? c.extU.p : c.extL.p;
            pg.nstages = upper ? c.nstagesU : c.nstagesL;
            static const int pf = exp_env("OPMGPU_COL_PF") ? atoi(exp_env("OPMGPU_COL_PF")) : 12;
            pg.pf_ahead = pf;
            // service warps: with W <= 3 every compute warp keeps a scheduler (SM sub-partition) to
            // itself and the producer (mostly asleep) shares one with the helper
            static const int roles = exp_env("OPMGPU_COL_ROLES") ? atoi(exp_env("OPMGPU_COL_ROLES")) : 1;
            int nwarps = c.g.W + 2;
            pg.producer_warp = c.g.W; pg.helper_warp = c.g.W + 1;
            if (roles == 1 && c.g.W <= 3) { nwarps = 8; pg.producer_warp = 3; pg.helper_warp = 7; }
            pg.prof = nullptr; pg.trace = nullptr; pg.trace_cta = -1;
            static const int trace_cta = exp_env("OPMGPU_COL_TRACE") ? atoi(exp_env("OPMGPU_COL_TRACE")) : -1;
            static const bool prof = exp_env("OPMGPU_COL_PROF") != nullptr;
            if (prof) {
                CK(h->d_trace.ensure((size_t)c.P * 8 * 4));
                CK(cudaMemsetAsync(h->d_trace.p, 0, sizeof(long long) * (size_t)c.P * 8 * 4, h->stream));
                pg.prof = h->d_trace.p;
                if (trace_cta >= 0) {
                    CK(h->d_gtrace.ensure((size_t)c.g.T * 8));
                    CK(cudaMemsetAsync(h->d_gtrace.p, 0, sizeof(long long) * (size_t)c.g.T * 8, h->stream));
                    pg.trace = h->d_gtrace.p; pg.trace_cta = trace_cta;
                }
            }
            const double* rhs = upper ? c.rhsU.p : c.rhsL.p; double* hand = upper ? nullptr : c.rhsU.p;
            double* out = upper ? vv : nullptr; int* err = h->d_err.p;
            void* args[] = {&pg, &rhs, &hand, &out, &w, (void*)&scale, &err};
            const void* fn = upper ? (prof ? (const void*)ilu0_sweep_col_kernel<true, true> : (const void*)ilu0_sweep_col_kernel<true, false>)
                                   : (prof ? (const void*)ilu0_sweep_col_kernel<false, true> : (const void*)ilu0_sweep_col_kernel<false, false>);
            CK(cudaLaunchCooperativeKernel(fn, dim3(c.P), dim3(32 * nwarps), args, upper ? c.smemU : c.smemL, h->stream));
            if (prof) {
                std::vector<long long> pr((size_t)c.P * 8 * 4);
                CK(cudaMemcpyAsync(pr.data(), h->d_trace.p, sizeof(long long) * pr.size(), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                double rec = 0, ring = 0, all = 0, steps = 0, worst = 0;
                for (size_t i = 0; i < pr.size(); i += 4) { rec += pr[i]; ring += pr[i + 1]; all += pr[i + 2]; steps += pr[i + 3]; worst = std::max(worst, (double)pr[i + 2]); }
                if (trace_cta >= 0) {
                    std::vector<long long> tr((size_t)c.g.T * 8);
                    CK(cudaMemcpy(tr.data(), h->d_gtrace.p, sizeof(long long) * tr.size(), cudaMemcpyDeviceToHost));
                    fprintf(stderr, "[opmgpu] col %s sweep, CTA %d warp 0: step: start(rel) | waits | loads+shuffles+chain | release+stores\n", upper ? "upper" : "lower", trace_cta);
                    for (int t = 0; t < c.g.T; ++t) {
                        const long long* q = &tr[(size_t)t * 8];
                        fprintf(stderr, "  %3d: %8lld | %5lld | %6lld | %5lld\n", t, q[0] - tr[0], q[1] - q[0], q[2] - q[1], q[3] - q[2]);
                    }
                }
                fprintf(stderr, "[opmgpu] col %s sweep: per step %.0f cycles (waits %.0f, loads+chain %.0f, rest %.0f); longest warp %.0f cycles\n",
                        upper ? "upper" : "lower", all / steps, rec / steps, ring / steps, (all - rec - ring) / steps, worst);
            }
        }
        h->launches += d_in_program_order ? 2 : 3;
        return 0;
    }
#endif
    if (h->use_pipe) {
        if (!d_in_program_order) {
            const size_t e = h->pipeL.nperm * 3;
            permute_rows_kernel<T><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->pipeL.nperm, h->pipeL.perm_row.p, d, h->pipeL.rhs_perm.p);
        }
        {
            PipeDev pg = pipe_dev(h->pipeL);
            if (h->trace_cta >= 0 && h->d_trace.p) { pg.trace = h->d_trace.p; pg.trace_cta = h->trace_cta; }
            if (h->gtrace_steps > 0) { pg.gtrace = h->d_gtrace.p; pg.gtrace_steps = h->gtrace_steps; }
            const double* rhs = h->pipeL.rhs_perm.p; double* work = h->d_yL.p; double* hand = h->pipeU.rhs_perm.p;
            T* out = nullptr; int* err = h->d_err.p;
            void* args[] = {&pg, &rhs, &work, &hand, &out, &w, (void*)&scale, &err};
            if (int rc = launch_sweep<T>(h, false, h->pipeL, args)) return rc;
        }
        {
            PipeDev pg = pipe_dev(h->pipeU);
            if (h->trace_cta >= 0 && h->d_trace.p) { pg.trace = h->d_trace.p + 512 * 16; pg.trace_cta = h->trace_cta; }
            if (h->gtrace_steps > 0) { pg.gtrace = h->d_gtrace.p + (size_t)h->pipeL.P * (h->gtrace_steps * 8 + 2048); pg.gtrace_steps = h->gtrace_steps; }
            const double* rhs = h->pipeU.rhs_perm.p; double* work = h->d_vU.p; double* hand = nullptr;
            T* out = v; int* err = h->d_err.p;
            void* args[] = {&pg, &rhs, &work, &hand, &out, &w, (void*)&scale, &err};
            if (int rc = launch_sweep<T>(h, true, h->pipeU, args)) return rc;
        }
        h->launches += d_in_program_order ? 2 : 3;
        return 0;
    }
    {
        SweepDev pg = sweep_dev(h->progL);
        const T* rhs = d; T* work = vec<T>(h->d_yL); T* out = nullptr;
        int* flags = h->d_flags.p; int epoch = ++h->epoch; int* err = h->d_err.p;
        void* args[] = {&pg, &rhs, &work, &out, &w, (void*)&scale, &flags, &epoch, &err};
        CK(cudaLaunchCooperativeKernel((void*)ilu0_sweep_kernel<true, T>, dim3(h->progL.P), dim3(256), args, 0, h->stream));
        h->launches++;
    }
    {
        SweepDev pg = sweep_dev(h->progU);
        const T* rhs = vec<T>(h->d_yL); T* work = vec<T>(h->d_vU); T* out = v;
        int* flags = h->d_flags.p; int epoch = ++h->epoch; int* err = h->d_err.p;
        void* args[] = {&pg, &rhs, &work, &out, &w, (void*)&scale, &flags, &epoch, &err};
        CK(cudaLaunchCooperativeKernel((void*)ilu0_sweep_kernel<false, T>, dim3(h->progU.P), dim3(256), args, 0, h->stream));
        h->launches++;
    }
    return 0;
}

int sweep_watchdog(opmgpu_handle h)
{
    h->err = "ILU0 sweep watchdog (code " + std::to_string(h->h_flags2[0]) + "): a dependency was never delivered";
    // re-arm: clear the error word and every push slot, so the handle stays usable
    cudaMemsetAsync(h->d_err.p, 0, sizeof(int), h->stream);
    for (PipeDevMem* d : {&h->pipeL, &h->pipeU})
        if (d->ext.p) cudaMemsetAsync(d->ext.p, 0xff, std::max<size_t>(d->next, 1) * 3 * sizeof(double), h->stream);
    if (h->col.valid) {
        cudaMemsetAsync(h->col.extL.p, 0xff, std::max<size_t>(h->col.next, 1) * 3 * sizeof(double), h->stream);
        cudaMemsetAsync(h->col.extU.p, 0xff, std::max<size_t>(h->col.next, 1) * 3 * sizeof(double), h->stream);
    }
    if (h->pipeF.ext.p) cudaMemsetAsync(h->pipeF.ext.p, 0xff, std::max<size_t>(h->pipeF.next, 1) * 9 * sizeof(double), h->stream);
    if (h->progL.fslots.p) cudaMemsetAsync(h->progL.fslots.p, 0xff, std::max<size_t>(h->progL.n_fslots, 1) * 9 * sizeof(double), h->stream);
    cudaStreamSynchronize(h->stream);
    return OPMGPU_CUDA_ERROR;
}

// mailbox of the next half-step-ending kernel.  Partitioned handles: the kernel itself gets no
// mailbox (its scalars are rank-local partial sums); publish_scalars_kernel writes it after the
// all-reduce (reduce_and_publish).
HostBox next_hostbox(opmgpu_handle h)
{
    HostBox hb;
    hb.hS = nullptr; hb.herr = nullptr; hb.hseq = nullptr; hb.derr = h->d_err.p; hb.seq = 0;
    if (h->world == 1 && h->use_hostbox) {
        hb.hS = h->h_S; hb.herr = &h->h_flags2[0]; hb.hseq = h->h_seq; hb.seq = ++h->seq;
    }
    return hb;
}

// partitioned handles: sum `count` scalars from S_NRM2 on over the ranks, then publish the whole
// scalar block into the host mailbox (returned in hb for wait_scalars)
int reduce_and_publish(opmgpu_handle h, int count, HostBox& hb)
{
    if (h->world == 1) return 0;
    if (!h->use_hostbox) return allreduce_slots(h, S_NRM2, count);
    hb.hS = h->h_S; hb.herr = &h->h_flags2[0]; hb.hseq = h->h_seq; hb.seq = ++h->seq;
    return allreduce_slots(h, S_NRM2, count, &hb);
}

int read_scalars(opmgpu_handle h);

// wait for the mailbox of `hb` (or fall back to copy + synchronise when it is disabled)
int wait_scalars(opmgpu_handle h, const HostBox& hb)
{
    if (!hb.hS) return read_scalars(h);
    volatile unsigned long long* seq = h->h_seq;
    for (unsigned spin = 1;; ++spin) {
        if (*seq == hb.seq) break;
        if ((spin & 0x3fffu) == 0) {                       // the kernel may have failed to run at all
            const cudaError_t q = cudaStreamQuery(h->stream);
            if (q == cudaSuccess) { if (*seq == hb.seq) break; return h->fail(cudaErrorUnknown, "host mailbox never written"); }
            if (q != cudaErrorNotReady) return h->fail(q, "BiCGStab half-step");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (h->h_flags2[0]) { CK(cudaStreamSynchronize(h->stream)); return sweep_watchdog(h); }
    return 0;
}

int read_scalars(opmgpu_handle h)
{
    CK(cudaMemcpyAsync(h->h_S, h->d_S.p, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&h->h_flags2[0], h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->world > 1) h->h_flags2[0] = h->h_S[S_ERRW] != 0.0 ? 9 : 0;      // the all-reduced verdict: the same on every rank
    if (h->h_flags2[0]) return sweep_watchdog(h);
    return 0;
}

// Dune::BiCGSTABSolver::apply on device vectors.  In: d_r = b (x0 = 0).  Out: d_x.
// T = float: every scalar of the recurrence is a float, as in Dune::BiCGSTABSolver<BlockVector<
// FieldVector<float,3>>> (field_type = real_type = float; EPSILON = 1e-80 rounds to 0).
template <class T>
int bicgstab(opmgpu_handle h, const opmgpu_params* prm, opmgpu_result* res)
{
    const T EPSILON = sizeof(T) == 8 ? (T)1e-80 : T(0);      // (float)1e-80
    const size_t n = (size_t)h->N * h->np;
    const T red = (T)prm->linear_solver_reduction;
    const double w = prm->ilu_relaxation;
    const int maxit = prm->linear_solver_maxiter, half_limit = prm->max_half_steps;
    int rc;
    h->history.clear();
    T* d_x = vec<T>(h->d_x); T* d_r = vec<T>(h->d_r); T* d_rt = vec<T>(h->d_rt); T* d_p = vec<T>(h->d_p);
    T* d_v = vec<T>(h->d_v); T* d_t = vec<T>(h->d_t); T* d_y = vec<T>(h->d_y);
    CK(cudaMemsetAsync(d_x, 0, n * sizeof(T), h->stream));
    CK(cudaMemcpyAsync(d_rt, d_r, n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    // the vector kernels write the next right-hand side of the lower sweep in program order
    const int* lpos = h->np == 3 && h->use_pipe && h->fuse_permute ? (h->use_col ? h->col.pos_of_row.p : h->pipeL.pos_of_row.p) : nullptr;
    double* lperm = h->use_pipe ? (h->use_col ? h->col.rhsL.p : h->pipeL.rhs_perm.p) : nullptr;
    HostBox hb = next_hostbox(h);
    bicg_init_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, d_r, h->d_S.p, h->ws(), hb);
    h->launches++;
    if ((rc = reduce_and_publish(h, 3, hb))) return rc;
    if ((rc = wait_scalars(h, hb))) return rc;
    const T norm0 = std::sqrt((T)h->h_S[S_NRM2]);
    T norm = norm0, rho = T(1), omega = T(1);
    double it = 0.0;
    // linear_solver_verbosity as dune's solvers read it: > 0 header and summary, > 1 a line per (half) iteration
    const int verbose = h->rank == 0 ? prm->linear_solver_verbosity : 0;
    T norm_old = norm0;
    if (verbose > 0) {
        std::printf("=== opmgpu BiCGSTABSolver (ILU0, %d block rows%s)\n", h->N, sizeof(T) == 4 ? ", single precision" : "");
        if (verbose > 1) std::printf(" Iter          Defect            Rate\n%5.1f %15.6e\n", 0.0, (double)norm0);
    }
    auto report = [&](double itv) {
        if (verbose > 1) std::printf("%5.1f %15.6e %15.6e\n", itv, (double)norm, norm_old > 0 ? (double)(norm / norm_old) : 0.0);
        norm_old = norm;
    };
    int half = 0, status = OPMGPU_OK, converged = 0;
    res->norm0 = norm0;
    if (norm < norm0 * red || norm < 1e-30) {
        res->converged = 1; res->iterations = 0; res->reduction = 0.0; res->half_steps = 0;
        return OPMGPU_OK;
    }
    for (it = 0.5; it < maxit; it += 0.5) {
        if (half_limit >= 0 && half >= half_limit) break;
        if (std::fabs(rho) <= EPSILON || std::fabs(omega) <= EPSILON) { status = OPMGPU_BREAKDOWN; break; }
        if (it < 1) {
            CK(cudaMemcpyAsync(d_p, d_r, n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
        } else {
            h->prof_begin(2);
            bicg_update_p_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, d_p, d_r, d_v, h->d_S.p, lpos, lperm);
            h->prof_end();
            h->launches++;
        }
        h->prof_begin(0);
        if ((rc = apply_precond<T>(h, w, d_p, d_y, lpos != nullptr && it >= 1))) return rc;
        h->prof_end();
        h->prof_begin(1);
        if ((rc = spmv_with_dots<T>(h, 1, d_y, d_v, d_rt))) return rc;
        h->prof_end();
        h->prof_begin(2);
        hb = next_hostbox(h);
        bicg_update1_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, d_x, d_r, d_y, d_v, h->d_S.p, h->ws(), hb, lpos, lperm);
        h->prof_end();
        h->launches++;
        if ((rc = reduce_and_publish(h, 2, hb))) return rc;
        if ((rc = wait_scalars(h, hb))) return rc;
        if (std::fabs((T)h->h_S[S_H]) < EPSILON) { status = OPMGPU_BREAKDOWN; break; }
        norm = std::sqrt((T)h->h_S[S_NRM2]);
        h->history.push_back(norm);
        report(it);
        ++half;
        if (norm < norm0 * red) { converged = 1; break; }
        it += 0.5;
        if (half_limit >= 0 && half >= half_limit) break;

        h->prof_begin(0);
        if ((rc = apply_precond<T>(h, w, d_r, d_y, lpos != nullptr))) return rc;
        h->prof_end();
        h->prof_begin(1);
        if ((rc = spmv_with_dots<T>(h, 2, d_y, d_t, d_r))) return rc;
        h->prof_end();
        h->prof_begin(2);
        hb = next_hostbox(h);
        bicg_update2_kernel<T><<<kVecBlocks, 256, 0, h->stream>>>(n, d_x, d_r, d_y, d_t, d_rt, h->d_S.p, h->ws(), hb);
        h->prof_end();
        h->launches++;
        if ((rc = reduce_and_publish(h, 3, hb))) return rc;
        if ((rc = wait_scalars(h, hb))) return rc;
        omega = (T)h->h_S[S_OMEGA];
        rho = (T)h->h_S[S_RHO_OLD];
        norm = std::sqrt((T)h->h_S[S_NRM2]);
        h->history.push_back(norm);
        report(it);
        ++half;
        if (norm < norm0 * red || norm < 1e-30) { converged = 1; break; }
    }
    if (it > maxit) it = maxit;
    res->iterations = (int)std::ceil(it);
    res->converged = converged;
    res->half_steps = half;
    res->reduction = (double)(norm / norm0);
    if (verbose > 0) std::printf("=== rate=%g, IT=%d, reduction=%g, %s\n", res->iterations > 0 ? std::pow((double)(norm / norm0), 1.0 / res->iterations) : 0.0,
                                 res->iterations, (double)(norm / norm0), converged ? "converged" : "NOT converged");
    if (status == OPMGPU_OK && !converged) status = OPMGPU_NOT_CONVERGED;
    if (status == OPMGPU_BREAKDOWN) h->err = "breakdown in BiCGSTAB (rho, omega or h below 1e-80)";
    return status;
}

// Dune::RestartedGMResSolver::apply on device vectors (newton_use_gmres, ISTLSolver.hpp:257-265):
// left preconditioned, modified Gram-Schmidt, Givens rotations on the host.  In: d_r = b (x0 = 0).
// Out: d_x.  One host synchronisation per Arnoldi step (the new Hessenberg column).
int gmres(opmgpu_handle h, const opmgpu_params* prm, opmgpu_result* res)
{
    const double EPSILON = 1e-80;
    const size_t n = (size_t)h->N * 3;
    const double red = prm->linear_solver_reduction, w = prm->ilu_relaxation;
    const int maxit = prm->linear_solver_maxiter, m = std::max(1, prm->linear_solver_restart);
    int rc;
    h->history.clear();
    CK(h->d_gmres_V.ensure((size_t)(m + 1) * n));
    CK(h->d_gmres_H.ensure((size_t)m + 2));
    double* V = h->d_gmres_V.p; double* Hd = h->d_gmres_H.p; double* wv = h->d_t.p; double* b = h->d_r.p; double* b2 = h->d_rt.p;
    std::vector<double> s(m + 1, 0.0), sn(m, 0.0), cs(m, 0.0), H((size_t)(m + 1) * m, 0.0), y(m + 1, 0.0), col(m + 2, 0.0);
    auto gen_rot = [](double dx, double dy, double& c, double& sv) {
        const double ndx = std::fabs(dx), ndy = std::fabs(dy);
        if (ndy < 1e-15) { c = 1.0; sv = 0.0; }
        else if (ndx < 1e-15) { c = 0.0; sv = 1.0; }
        else if (ndy > ndx) { const double t = ndx / ndy; c = 1.0 / std::sqrt(1.0 + t * t); sv = c; c *= t; sv *= dx / ndx; sv *= dy / ndy; }
        else { const double t = ndy / ndx; c = 1.0 / std::sqrt(1.0 + t * t); sv = c; sv *= dy / dx; }
    };
    auto app_rot = [](double& dx, double& dy, double c, double sv) { const double t = c * dx + sv * dy; dy = -sv * dx + c * dy; dx = t; };
    // H slots [lo, lo+cnt) hold rank-local partial sums: reduce them, then fetch [0, upto) to the host
    auto reduce_slots = [&](int lo, int cnt) -> int {
        if (h->world > 1) NK(g_nccl.AllReduce(Hd + lo, Hd + lo, (size_t)cnt, ncclDouble, ncclSum, h->comm, h->stream));
        return 0;
    };
    auto fetch = [&](int upto) -> int {
        CK(cudaMemcpyAsync(col.data(), Hd, sizeof(double) * upto, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&h->h_flags2[0], h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (h->h_flags2[0]) return sweep_watchdog(h);
        return 0;
    };
    auto precond_norm = [&](double* v0, double& nrm) -> int {          // v0 = W^-1 b, nrm = |v0|
        if ((rc = apply_precond<double>(h, w, b, v0))) return rc;
        gmres_dot_kernel<<<kVecBlocks, 256, 0, h->stream>>>(n, v0, v0, Hd, 0, h->ws());
        h->launches++;
        if ((rc = reduce_slots(0, 1))) return rc;
        if ((rc = fetch(1))) return rc;
        nrm = std::sqrt(col[0]);
        return 0;
    };
    CK(cudaMemsetAsync(h->d_x.p, 0, n * sizeof(double), h->stream));
    CK(cudaMemcpyAsync(b2, b, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    // b -= A x with x = 0 leaves b as it is
    double norm = 0.0, norm_0 = 0.0;
    if ((rc = precond_norm(V, norm_0))) return rc;
    norm = norm_0;
    res->norm0 = norm_0;
    int j = 1, status = OPMGPU_OK, converged = 0;
    if (norm_0 < EPSILON) { res->converged = 1; res->iterations = 0; res->reduction = 0.0; return OPMGPU_OK; }
    while (j <= maxit && !converged && status == OPMGPU_OK) {
        int i = 0;
        gmres_scale_kernel<<<kVecBlocks, 256, 0, h->stream>>>(n, V, V, 1.0 / norm);
        h->launches++;
        s[0] = norm;
        for (i = 1; i < m + 1; ++i) s[i] = 0.0;
        for (i = 0; i < m && j <= maxit && !converged; ++i, ++j) {
            double* vi = V + (size_t)i * n;
            double* vn = V + (size_t)(i + 1) * n;
            // v[i+1] = A v[i] (the operand needs room for the ghost rows when partitioned), w = W^-1 v[i+1]
            double* xin = vi;
            if (h->world > 1) { CK(cudaMemcpyAsync(h->d_tmp.p, vi, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream)); xin = h->d_tmp.p; }
            if ((rc = spmv_with_dots<double>(h, 0, xin, vn, nullptr))) return rc;
            if ((rc = apply_precond<double>(h, w, vn, wv))) return rc;
            // modified Gram-Schmidt: H[k][i] = v[k].w ; w -= H[k][i] v[k]
            gmres_dot_kernel<<<kVecBlocks, 256, 0, h->stream>>>(n, V, wv, Hd, 0, h->ws());
            h->launches++;
            if ((rc = reduce_slots(0, 1))) return rc;
            for (int k = 0; k <= i; ++k) {
                gmres_mgs_kernel<<<kVecBlocks, 256, 0, h->stream>>>(n, wv, V + (size_t)k * n, k < i ? V + (size_t)(k + 1) * n : nullptr, Hd, k, h->ws());
                h->launches++;
                if ((rc = reduce_slots(k + 1, 1))) return rc;
            }
            if ((rc = fetch(i + 2))) return rc;
            for (int k = 0; k <= i; ++k) H[(size_t)k * m + i] = col[k];
            const double hn = std::sqrt(col[i + 1]);
            H[(size_t)(i + 1) * m + i] = hn;
            if (std::fabs(hn) < EPSILON) { status = OPMGPU_BREAKDOWN; break; }
            gmres_scale_kernel<<<kVecBlocks, 256, 0, h->stream>>>(n, vn, wv, 1.0 / hn);
            h->launches++;
            for (int k = 0; k < i; ++k) app_rot(H[(size_t)k * m + i], H[(size_t)(k + 1) * m + i], cs[k], sn[k]);
            gen_rot(H[(size_t)i * m + i], H[(size_t)(i + 1) * m + i], cs[i], sn[i]);
            app_rot(H[(size_t)i * m + i], H[(size_t)(i + 1) * m + i], cs[i], sn[i]);
            app_rot(s[i], s[i + 1], cs[i], sn[i]);
            norm = std::fabs(s[i + 1]);
            h->history.push_back(norm);
            if (h->rank == 0 && prm->linear_solver_verbosity > 1) std::printf("%5d %15.6e\n", j, norm);
            if (norm < red * norm_0) converged = 1;
        }
        if (status != OPMGPU_OK) break;
        // update: back substitution, x += sum y[a] v[a]
        CK(cudaMemsetAsync(wv, 0, n * sizeof(double), h->stream));
        for (int a = 0; a < m + 1; ++a) y[a] = s[a];
        for (int a = i - 1; a >= 0; --a) {
            double rhs = s[a];
            for (int c = a + 1; c < i; ++c) rhs -= H[(size_t)a * m + c] * y[c];
            y[a] = rhs / H[(size_t)a * m + a];
            gmres_axpy_kernel<<<kVecBlocks, 256, 0, h->stream>>>(n, wv, y[a], V + (size_t)a * n);
            h->launches++;
        }
        gmres_add_kernel<<<kVecBlocks, 256, 0, h->stream>>>(n, h->d_x.p, wv);
        h->launches++;
        if (!converged && j <= maxit) {
            CK(cudaMemcpyAsync(b, b2, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            const double* xin = h->d_x.p;
            if (h->world > 1) {
                CK(cudaMemcpyAsync(h->d_tmp.p, h->d_x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
                if ((rc = halo_exchange<double>(h, h->d_tmp.p))) return rc;
                xin = h->d_tmp.p;
            }
            const int* rowptr = h->world > 1 ? h->d_rowptr_full.p : h->d_rowptr.p;
            const int* colidx = h->world > 1 ? h->d_colidx_full.p : h->d_colidx.p;
            residual3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->N, rowptr, colidx, static_cast<const double*>(h->d_vals), xin, b);
            h->launches++;
            if ((rc = precond_norm(V, norm))) return rc;
        }
    }
    CK(cudaGetLastError());
    res->iterations = j - 1;
    res->converged = converged;
    if (h->rank == 0 && prm->linear_solver_verbosity > 0)
        std::printf("=== opmgpu RestartedGMResSolver (ILU0, restart %d): IT=%d, reduction=%g, %s\n", m, j - 1, norm / norm_0, converged ? "converged" : "NOT converged");
    res->half_steps = (int)h->history.size();
    res->reduction = norm / norm_0;
    if (status == OPMGPU_OK && !converged) status = OPMGPU_NOT_CONVERGED;
    if (status == OPMGPU_BREAKDOWN) h->err = "breakdown in GMRes (new Krylov vector below 1e-80)";
    return status;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// factor + BiCGStab on the values/rhs already in place (d_vals, d_r); leaves the result in d_x
constexpr int kPatternChanged = -1000;     // internal: solve_resident -> CSC front end

int solve_resident(opmgpu_handle h, const opmgpu_params* prm, opmgpu_result* res)
{
    cudaEventRecord(h->ev[0], h->stream);
    int badrow = -1;
    h->prof_begin(3);
    int rc = factor(h, &badrow);
    h->prof_end();
    cudaEventRecord(h->ev[1], h->stream);
    // the CSC front end compares the caller's sparsity pattern with the cached one on host
    // threads while values are uploaded and factorised; a negative verdict that is already in
    // stops here (not waited for: a verdict that is still out is collected by the caller after the solve,
    // which is then repeated on the new pattern -- rare, patterns change at report steps)
    if (h->pattern_check.valid() && h->pattern_check.wait_for(std::chrono::seconds(0)) == std::future_status::ready &&
        !h->pattern_check.get()) return kPatternChanged;
    if (rc) { res->bad_row = badrow; return rc; }
    if (prm->newton_use_gmres && h->f32) {
        h->err = "restarted GMRES (newton_use_gmres) is available in the double-precision instance only";
        return OPMGPU_BAD_ARGUMENT;
    }
    rc = prm->newton_use_gmres ? gmres(h, prm, res) : (h->f32 ? bicgstab<float>(h, prm, res) : bicgstab<double>(h, prm, res));
    cudaEventRecord(h->ev[2], h->stream);
    cudaEventSynchronize(h->ev[2]);
    res->ms_factor = ev_ms(h->ev[0], h->ev[1]);
    res->ms_solve = ev_ms(h->ev[1], h->ev[2]);
    h->prof_collect();
    if (rc == OPMGPU_NOT_CONVERGED && prm->linear_solver_ignoreconvergencefailure) rc = OPMGPU_OK;
    return rc;
}

// memcmp of two large buffers on several host threads (the nine CSC index arrays are ~290 MB at
// 1M cells; a single-threaded compare would cost more than the whole GPU solve)
bool equal_parallel(const void* a, const void* b, size_t bytes)
{
    const size_t kMin = 4u << 20;
    unsigned nt = std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
    if (bytes < 2 * kMin) nt = 1;
    if (nt == 1) return std::memcmp(a, b, bytes) == 0;
    std::vector<std::thread> th;
    std::vector<int> diff(nt, 0);
    const size_t chunk = (bytes + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t lo = std::min(bytes, t * chunk), hi = std::min(bytes, lo + chunk);
        th.emplace_back([=, &diff]() { diff[t] = std::memcmp((const char*)a + lo, (const char*)b + lo, hi - lo) != 0; });
    }
    for (auto& x : th) x.join();
    for (int d : diff) if (d) return false;
    return true;
}

// O(1) part of the comparison: same sizes, so the cached upload plan cannot overrun the caller's arrays
bool same_csc_sizes(opmgpu_handle h, int N, const opmgpu_csc* b, bool full)
{
    if (!h->have_pattern || h->csc_colptr.size() != 9 || h->N != N || h->csc_full_pattern != full) return false;
    for (int q = 0; q < 9; ++q)
        if ((int)h->csc_colptr[q].size() != N + 1 || h->csc_rowidx[q].size() != (size_t)b[q].colptr[N] || b[q].colptr[0] != 0) return false;
    return true;
}

// Peer-memory halo plan (once per pattern, collective): every rank publishes its ghost-row target
// (the d_y vector) and its incoming-epoch array -- as CUDA IPC handles for ranks in other processes
// (torchrun: one process per GPU), as plain pointers plus peer access for ranks that are threads of
// this process (opmgpu_create_multi).  Any failure on any rank keeps ncclSend/ncclRecv on all of them.
void close_peer_halo(opmgpu_handle h)
{
    for (void* p : h->peer.opened) cudaIpcCloseMemHandle(p);
    h->peer.opened.clear();
    h->peer.ready = false;
}

int setup_peer_halo(opmgpu_handle h)
{
    struct Info {
        cudaIpcMemHandle_t y, flags, inbox, sflags;
        long long pid;
        void* raw_y; void* raw_flags; void* raw_inbox; void* raw_sflags;
        int device, N_local;
        int recv_off[kMaxPeers], recv_cnt[kMaxPeers];
    };
    close_peer_halo(h);
    const int W = h->world;
    if (W > kMaxPeers || !h->use_peer_halo) return 0;
    CK(h->peer.flags_in.ensure(kMaxPeers));
    CK(h->peer.ticket.ensure(1));
    const bool first = h->peer.inbox.p == nullptr;      // first plan of this handle (epochs only grow afterwards)
    CK(h->peer.inbox.ensure((size_t)2 * kMaxPeers * kScalSlots));
    CK(h->peer.sflags.ensure(kMaxPeers));
    if (first) {
        CK(cudaMemsetAsync(h->peer.flags_in.p, 0, sizeof(unsigned long long) * kMaxPeers, h->stream));
        CK(cudaMemsetAsync(h->peer.ticket.p, 0, sizeof(unsigned), h->stream));
        CK(cudaMemsetAsync(h->peer.inbox.p, 0, sizeof(double) * 2 * kMaxPeers * kScalSlots, h->stream));
        CK(cudaMemsetAsync(h->peer.sflags.p, 0, sizeof(unsigned long long) * kMaxPeers, h->stream));
    }
    Info mine;
    std::memset(&mine, 0, sizeof mine);
    int ok = 1;
    if (cudaIpcGetMemHandle(&mine.y, h->d_y.p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    if (cudaIpcGetMemHandle(&mine.flags, h->peer.flags_in.p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    if (cudaIpcGetMemHandle(&mine.inbox, h->peer.inbox.p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    if (cudaIpcGetMemHandle(&mine.sflags, h->peer.sflags.p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    mine.pid = (long long)getpid();
    mine.raw_y = h->d_y.p; mine.raw_flags = h->peer.flags_in.p;
    mine.raw_inbox = h->peer.inbox.p; mine.raw_sflags = h->peer.sflags.p;
    mine.device = h->device; mine.N_local = h->N;
    for (int p = 0; p < W; ++p) { mine.recv_off[p] = h->recv_off[p]; mine.recv_cnt[p] = h->recv_cnt[p]; }
    DevArr<unsigned char> d_mine, d_all;
    CK(d_mine.ensure(sizeof(Info))); CK(d_all.ensure(sizeof(Info) * (size_t)W));
    CK(cudaMemcpyAsync(d_mine.p, &mine, sizeof(Info), cudaMemcpyHostToDevice, h->stream));
    NK(g_nccl.AllGather(d_mine.p, d_all.p, sizeof(Info), ncclUint8, h->comm, h->stream));
    std::vector<Info> all((size_t)W);
    CK(cudaMemcpyAsync(all.data(), d_all.p, sizeof(Info) * (size_t)W, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    PeerPushDesc& d = h->peer.desc;
    std::memset(&d, 0, sizeof d);
    d.world = W; d.rank = h->rank;
    h->peer.recv_mask = 0;
    for (int p = 0; p < W; ++p) {
        d.send_off[p] = h->send_off[p];
        if (h->recv_cnt[p] > 0) h->peer.recv_mask |= 1u << p;
    }
    d.send_off[W] = h->n_send;
    PeerScalDesc& sd = h->peer.sdesc;
    std::memset(&sd, 0, sizeof sd);
    sd.world = W; sd.rank = h->rank;
    sd.inbox[h->rank] = h->peer.inbox.p; sd.flags[h->rank] = h->peer.sflags.p;
    auto map = [&](const cudaIpcMemHandle_t& hd, void*& out) {
        if (cudaIpcOpenMemHandle(&out, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return false; }
        h->peer.opened.push_back(out);
        return true;
    };
    for (int p = 0; p < W && ok; ++p) {
        if (p == h->rank) continue;
        const bool halo_peer = h->send_cnt[p] > 0;
        if (halo_peer) {
            if (all[p].recv_cnt[h->rank] != h->send_cnt[p]) { ok = 0; break; }
            d.dst_row0[p] = (long long)all[p].N_local + all[p].recv_off[h->rank];
        }
        if (all[p].pid == mine.pid) {
            // a thread of this process: plain pointers, peer access on
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[p].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ok = 0; cudaGetLastError(); break; }
            cudaGetLastError();
            if (halo_peer) { d.y[p] = all[p].raw_y; d.flags[p] = static_cast<unsigned long long*>(all[p].raw_flags); }
            sd.inbox[p] = static_cast<double*>(all[p].raw_inbox); sd.flags[p] = static_cast<unsigned long long*>(all[p].raw_sflags);
        } else {
            void* q = nullptr;
            if (halo_peer) {
                if (!map(all[p].y, q)) { ok = 0; break; }
                d.y[p] = q;
                if (!map(all[p].flags, q)) { ok = 0; break; }
                d.flags[p] = static_cast<unsigned long long*>(q);
            }
            if (!map(all[p].inbox, q)) { ok = 0; break; }
            sd.inbox[p] = static_cast<double*>(q);
            if (!map(all[p].sflags, q)) { ok = 0; break; }
            sd.flags[p] = static_cast<unsigned long long*>(q);
        }
    }
    // everybody or nobody
    DevArr<int> d_ok;
    CK(d_ok.ensure(1));
    CK(cudaMemcpyAsync(d_ok.p, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NK(g_nccl.AllReduce(d_ok.p, d_ok.p, 1, ncclInt32, ncclMin, h->comm, h->stream));
    CK(cudaMemcpyAsync(&ok, d_ok.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    d_mine.release(); d_all.release(); d_ok.release();
    if (!ok) { close_peer_halo(h); return 0; }
    h->peer.ready = true;
    if (getenv("OPMGPU_DEBUG")) fprintf(stderr, "[opmgpu] rank %d: halo exchange through peer memory (%s)\n", h->rank,
                                        h->peer.opened.empty() ? "same process" : "CUDA IPC");
    return 0;
}

// float instance: round the caller's doubles once into the handle's own float array (what the
// assignment to the reference's float matrix does, ...Interleaved.cpp:189)
int take_values_f32(opmgpu_handle h, const double* vals_dev)
{
    const size_t nv = (size_t)(h->world > 1 ? h->nnzb_full : h->nnzb) * 9;
    CK(h->d_vals32.ensure(nv + 16));
    if (int rc = convert<double, float>(h, nv, vals_dev, h->d_vals32.p)) return rc;
    h->d_vals = h->d_vals32.p;
    return 0;
}

bool same_csc_pattern(opmgpu_handle h, int N, const opmgpu_csc* b, bool full)
{
    if (!h->have_pattern || h->csc_colptr.size() != 9 || h->N != N || h->csc_full_pattern != full) return false;
    for (int q = 0; q < 9; ++q) {
        if ((int)h->csc_colptr[q].size() != N + 1) return false;
        if (!equal_parallel(h->csc_colptr[q].data(), b[q].colptr, sizeof(int) * ((size_t)N + 1))) return false;
        const size_t nnz = (size_t)b[q].colptr[N];
        if (h->csc_rowidx[q].size() != nnz) return false;
        if (nnz && !equal_parallel(h->csc_rowidx[q].data(), b[q].rowidx, sizeof(int) * nnz)) return false;
    }
    return true;
}

}  // namespace

// ================================================================================================
extern "C" {

void opmgpu_default_params(opmgpu_params* p)
{
    // FlowLinearSolverParameters::reset() of opm-simulators 2019.04 (not under /root/reference;
    // the same keys with CPR's defaults are visible at NewtonIterationBlackoilCPR.cpp:61-66).
    p->linear_solver_reduction = 1e-2;
    p->linear_solver_maxiter = 150;
    p->ilu_relaxation = 0.9;
    p->linear_solver_verbosity = 0;
    p->linear_solver_ignoreconvergencefailure = 0;
    p->require_full_sparsity_pattern = 0;
    p->max_half_steps = -1;
    p->newton_use_gmres = 0;
    p->linear_solver_restart = 40;
}

const char* opmgpu_last_error(opmgpu_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int opmgpu_create(int device, opmgpu_handle* out)
{
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                         " (this solver has no CPU fallback)";
        return OPMGPU_CUDA_ERROR;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return OPMGPU_BAD_ARGUMENT; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return OPMGPU_CUDA_ERROR;
    }
    if (prop.major != 10) {
        g_create_error = "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                         "; this library is built for sm_100a (B200) only";
        return OPMGPU_CUDA_ERROR;
    }
    opmgpu_handle h = new opmgpu_solver();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    int per_sm = 1;
    if (const char* s = exp_env("OPMGPU_SWEEP_CTAS_PER_SM")) per_sm = std::max(1, atoi(s));
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ilu0_sweep_kernel<true, double>, 256, 0);
    per_sm = std::min(per_sm, std::max(occ, 1));
    h->sweep_ctas = h->sm_count * per_sm;
    if (const char* s = exp_env("OPMGPU_SIMPLE_SWEEP")) h->force_simple = atoi(s) != 0;
    if (const char* s = exp_env("OPMGPU_FACTOR_BY_LEVELS")) h->factor_by_levels = atoi(s) != 0;
    if (const char* s = exp_env("OPMGPU_FACTOR_TILE")) h->factor_tile = atoi(s) != 0;
    if (const char* s = exp_env("OPMGPU_HOSTBOX")) h->use_hostbox = atoi(s) != 0;
    if (const char* s = exp_env("OPMGPU_FUSE_PERMUTE")) h->fuse_permute = atoi(s) != 0;
    if (const char* s = exp_env("OPMGPU_SPMV_SIMPLE")) h->spmv_tma = atoi(s) == 0;
    cudaFuncSetAttribute(spmv3_tma_kernel<0, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<1, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<2, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<0, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<1, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<2, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<0, float, true, kSpmvRowsF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<1, float, true, kSpmvRowsF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<2, float, true, kSpmvRowsF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<0, float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<1, float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<2, float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<0, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<1, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(spmv3_tma_kernel<2, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmvSmemBytes);
    cudaFuncSetAttribute(mc_line_sweep_kernel<false, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLineSmemBytes);
    cudaFuncSetAttribute(mc_line_sweep_kernel<true, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLineSmemBytes);
    cudaFuncSetAttribute(mc_line_sweep_kernel<false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLineSmemBytes);
    cudaFuncSetAttribute(mc_line_sweep_kernel<true, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLineSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<0, double, false, kSpmvRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<1, double, false, kSpmvRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<2, double, false, kSpmvRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<0, double, true, kSpmvRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<1, double, true, kSpmvRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<2, double, true, kSpmvRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<0, float, true, kSpmvRowsF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<1, float, true, kSpmvRowsF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    cudaFuncSetAttribute(mc_sweep_tma_kernel<2, float, true, kSpmvRowsF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMcSmemBytes);
    if (const char* s = exp_env("OPMGPU_HALO_OVERLAP")) h->overlap_halo = atoi(s) != 0;
    if (const char* s = getenv("OPMGPU_PEER_HALO")) h->use_peer_halo = atoi(s) != 0;
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaFuncSetAttribute(ilu0_factor_pipe_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_factor_pipe_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    // the float instance's sweeps (same records, float arithmetic)
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, true, true, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, true, true, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, false, false, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, false, false, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, true, false, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, true, false, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
#ifdef OPMGPU_EXPERIMENTS
    cudaFuncSetAttribute(ilu0_sweep_col_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_col_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_col_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_col_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
#endif
    if (const char* s = exp_env("OPMGPU_COL")) h->allow_col = atoi(s) != 0;
#ifdef OPMGPU_EXPERIMENTS
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
#endif
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
#ifdef OPMGPU_EXPERIMENTS
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
#endif
    {
        // how many CTAs of the cluster variants can be co-resident (one CTA per SM, all of its
        // shared memory): GPCs do not divide evenly into clusters of 4 or 8
        int want = 8;                                        // OPMGPU_CLUSTER=0: no clusters; 2/4/8: largest size tried
        if (const char* s = exp_env("OPMGPU_CLUSTER")) want = atoi(s);
        if (h->world > 1) want = 0;
        for (int lg = 1; lg <= 3; ++lg) {
            const int cs = 1 << lg;
            if (cs > want) break;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * h->sm_count); cfg.blockDim = dim3(kPipeThreads); cfg.dynamicSmemBytes = (size_t)h->max_smem_optin;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, (const void*)ilu0_sweep_pipe_kernel<true, true, true>, &cfg) == cudaSuccess)
                h->caps.max_ctas[lg] = std::min(ncl * cs, h->sm_count);
            else
                cudaGetLastError();
        }
        if (getenv("OPMGPU_DEBUG"))
            fprintf(stderr, "[opmgpu] co-resident CTAs at cluster size 2/4/8: %d/%d/%d\n", h->caps.max_ctas[1], h->caps.max_ctas[2], h->caps.max_ctas[3]);
    }
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    cudaFuncSetAttribute(ilu0_sweep_pipe_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
    if (!h->force_simple) h->sweep_ctas = h->sm_count;      // the pipelined sweep owns a whole SM per CTA
    if (const char* s = exp_env("OPMGPU_PIPE_CTAS")) h->sweep_ctas = std::max(1, std::min(h->sm_count, atoi(s)));
    if (getenv("OPMGPU_DEBUG")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, ilu0_sweep_pipe_kernel<false, true>);
        int occ0 = -1, occ1 = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, ilu0_sweep_pipe_kernel<false, true>, kPipeThreads, 100 * 1024);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, ilu0_sweep_pipe_kernel<false, true>, kPipeThreads, (size_t)h->max_smem_optin);
        fprintf(stderr, "[opmgpu] sweep kernel: regs %d, maxThreadsPerBlock %d, static smem %zu, local %zu, optin %d, occ(100KB) %d occ(max) %d, threads %d\n",
                fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, fa.localSizeBytes, h->max_smem_optin, occ0, occ1, kPipeThreads);
    }
    if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
        delete h;
        return OPMGPU_CUDA_ERROR;
    }
    h->stream = h->own_stream;
    bool ok = cudaMallocHost((void**)&h->h_S, sizeof(double) * S_COUNT) == cudaSuccess &&
              cudaMallocHost((void**)&h->h_flags2, sizeof(int) * 4) == cudaSuccess &&
              cudaMallocHost((void**)&h->h_seq, sizeof(unsigned long long) * 8) == cudaSuccess &&
              h->d_S.ensure(S_COUNT) == cudaSuccess && h->d_partials.ensure((size_t)4 * kMaxRedBlocks) == cudaSuccess &&
              h->d_ticket.ensure(1) == cudaSuccess && h->d_err.ensure(1) == cudaSuccess && h->d_bad.ensure(1) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreate(&h->ev[i]) == cudaSuccess;
    if (ok) {
        h->h_seq[0] = 0; h->h_flags2[0] = 0;
        cudaMemset(h->d_S.p, 0, sizeof(double) * S_COUNT);
        cudaMemset(h->d_ticket.p, 0, sizeof(unsigned));
        cudaMemset(h->d_err.p, 0, sizeof(int));
    } else {
        g_create_error = "allocation of solver workspace failed";
        delete h;
        return OPMGPU_CUDA_ERROR;
    }
    *out = h;
    return OPMGPU_OK;
}

int opmgpu_destroy(opmgpu_handle h)
{
    if (!h) return OPMGPU_OK;
    if (h->multi) { multi_destroy(h->multi); delete h; return OPMGPU_OK; }
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    h->d_rowptr.release(); h->d_colidx.release(); h->d_diag.release(); h->d_lvl_rows.release();
    h->progL.release(); h->progU.release(); h->pipeL.release(); h->pipeU.release(); h->pipeF.release(); h->col.release();
    h->mc.release();
    h->d_vals_own.release(); h->d_lu.release(); h->d_stage.release(); h->d_vals32.release();
    h->d_x.release(); h->d_r.release(); h->d_rt.release(); h->d_p.release(); h->d_v.release();
    h->d_t.release(); h->d_y.release(); h->d_yL.release(); h->d_vU.release(); h->d_tmp.release(); h->d_tmp2.release();
    h->d_S.release(); h->d_partials.release(); h->d_ticket.release(); h->d_flags.release();
    h->d_rowptr_full.release(); h->d_colidx_full.release(); h->d_lu_src.release(); h->d_send_rows.release(); h->d_sendbuf.release();
    h->d_bnd_rows.release(); h->d_row_skip.release(); h->d_lvlU_rows.release(); h->d_map_np.release();
    close_peer_halo(h); h->peer.flags_in.release(); h->peer.ticket.release(); h->peer.inbox.release(); h->peer.sflags.release();
    if (h->halo_stream) { cudaStreamDestroy(h->halo_stream); cudaEventDestroy(h->ev_x_ready); cudaEventDestroy(h->ev_halo_done); h->halo_stream = nullptr; }
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    h->d_err.release(); h->d_bad.release(); h->d_map9.release(); h->d_cscval.release(); h->d_rhs_stage.release();
    if (h->h_S) cudaFreeHost(h->h_S);
    if (h->h_flags2) cudaFreeHost(h->h_flags2);
    if (h->h_seq) cudaFreeHost(h->h_seq);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return OPMGPU_OK;
}

int opmgpu_set_stream(opmgpu_handle h, void* cuda_stream)
{
    if (!h) return OPMGPU_BAD_ARGUMENT;
    cudaStreamSynchronize(h->stream);
    h->stream = (cudaStream_t)cuda_stream;      // NULL is the legacy default stream
    return OPMGPU_OK;
}

int opmgpu_set_pattern_bcrs(opmgpu_handle h, int N, int nnzb, const int* rowptr, const int* colidx)
{
    if (!h || !rowptr || !colidx) return OPMGPU_BAD_ARGUMENT;
    if (h->multi) return multi_set_pattern(h->multi, N, nnzb, rowptr, colidx, h->err);
    if (h->world > 1) return h->bad("distributed handle: use opmgpu_set_pattern_bcrs_distributed");
    CK(cudaSetDevice(h->device));
    h->csc_colptr.clear(); h->csc_rowidx.clear();
    return set_pattern(h, N, nnzb, rowptr, colidx);
}

// Operator only: the pattern is uploaded for y = A x, no ILU0 analysis is run (micro-benchmarks of
// the SpMV at sizes where the factor records would not fit beside the matrix; BASELINE.json config 5).
int opmgpu_set_pattern_bcrs_operator_only(opmgpu_handle h, int N, int nnzb, const int* rowptr, const int* colidx)
{
    if (!h || !rowptr || !colidx) return OPMGPU_BAD_ARGUMENT;
    if (h->multi || h->world > 1) return h->bad("operator-only patterns exist for plain single-GPU handles");
    if (N < 1 || nnzb < N || rowptr[0] != 0 || rowptr[N] != nnzb) return h->bad("bad BCRS pattern sizes");
    CK(cudaSetDevice(h->device));
    h->have_pattern = h->have_values = h->have_factors = false;
    h->csc_colptr.clear(); h->csc_rowidx.clear();
    h->use_pipe = false; h->use_col = false;
    h->pipeL.release(); h->pipeU.release(); h->pipeF.release(); h->progL.release(); h->progU.release(); h->d_lu.release();
    h->mc.release();
    h->N = N; h->nnzb = nnzb; h->n_ghost = 0;
    CK(h->d_rowptr.ensure((size_t)N + 1 + 8));
    CK(h->d_colidx.ensure((size_t)nnzb + 8));
    CK(cudaMemcpyAsync(h->d_rowptr.p, rowptr, sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_colidx.p, colidx, sizeof(int) * (size_t)nnzb, cudaMemcpyHostToDevice, h->stream));
    CK(h->d_tmp.ensure((size_t)N * 3)); CK(h->d_tmp2.ensure((size_t)N * 3));
    CK(cudaStreamSynchronize(h->stream));
    h->operator_only = true;
    h->have_pattern = true;
    return OPMGPU_OK;
}

int opmgpu_set_values_bcrs3(opmgpu_handle h, const double* vals)
{
    if (!h || !vals) return OPMGPU_BAD_ARGUMENT;
    if (!h->have_pattern) return h->bad("set the pattern first");
    CK(cudaSetDevice(h->device));
    const size_t nv = (size_t)(h->world > 1 ? h->nnzb_full : h->nnzb) * 9;
    CK(h->d_vals_own.ensure(nv));
    CK(cudaMemcpyAsync(h->d_vals_own.p, vals, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (h->f32) { if (int rc = take_values_f32(h, h->d_vals_own.p)) return rc; }
    else h->d_vals = h->d_vals_own.p;
    CK(cudaStreamSynchronize(h->stream));
    h->have_values = true; h->have_factors = false;
    return OPMGPU_OK;
}

int opmgpu_set_values_bcrs3_dev(opmgpu_handle h, const double* vals_dev)
{
    if (!h || !vals_dev) return OPMGPU_BAD_ARGUMENT;
    if (!h->have_pattern) return h->bad("set the pattern first");
    if (h->f32) {          // the float instance keeps its own (rounded) copy: nothing is borrowed
        CK(cudaSetDevice(h->device));
        if (int rc = take_values_f32(h, vals_dev)) return rc;
    } else h->d_vals = vals_dev;
    h->have_values = true; h->have_factors = false;
    return OPMGPU_OK;
}

int opmgpu_set_precision(opmgpu_handle h, int single_precision)
{
    if (!h) return OPMGPU_BAD_ARGUMENT;
    if (h->multi) return multi_set_precision(h->multi, single_precision, h->err);
    const bool f32 = single_precision != 0;
    if (f32 != h->f32) {          // values and factors belong to the other instance
        h->f32 = f32;
        h->have_values = false; h->have_factors = false; h->d_vals = nullptr;
    }
    return OPMGPU_OK;
}
int opmgpu_get_precision(opmgpu_handle h) { return h && h->f32 ? 1 : 0; }

int opmgpu_spmv_dev(opmgpu_handle h, const double* x_dev, double* y_dev)
{
    if (!h || !h->have_values) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->N * 3;
    if (h->f32) {                // doubles at the ABI, floats inside
        float* xf = vec<float>(h->d_tmp); float* yf = vec<float>(h->d_tmp2);
        int rc = convert<double, float>(h, n, x_dev, xf);
        if (!rc) rc = spmv_with_dots<float>(h, 0, xf, yf, nullptr);
        if (!rc) rc = convert<float, double>(h, n, yf, y_dev);
        return rc;
    }
    if (h->world > 1) {          // x_dev has no ghost rows: stage it
        CK(cudaMemcpyAsync(h->d_tmp.p, x_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return spmv_with_dots<double>(h, 0, h->d_tmp.p, y_dev, nullptr);
    }
    return launch_spmv<double>(h, 0, x_dev, y_dev, nullptr);
}

int opmgpu_spmv(opmgpu_handle h, const double* x, double* y)
{
    if (!h || !h->have_values) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->N * 3;
    if (h->f32) {
        CK(h->d_stage.ensure(n));
        CK(cudaMemcpyAsync(h->d_stage.p, x, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        int rc = opmgpu_spmv_dev(h, h->d_stage.p, h->d_stage.p);
        if (rc) return rc;
        CK(cudaMemcpyAsync(y, h->d_stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return OPMGPU_OK;
    }
    CK(cudaMemcpyAsync(h->d_tmp.p, x, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int rc = spmv_with_dots<double>(h, 0, h->d_tmp.p, h->d_tmp2.p, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(y, h->d_tmp2.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return OPMGPU_OK;
}

int opmgpu_ilu0_factor(opmgpu_handle h, int* bad_row)
{
    if (!h) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    return factor(h, bad_row);
}

int opmgpu_ilu0_get_factors(opmgpu_handle h, double* lu)
{
    if (!h || !h->have_factors) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    if (h->mc.valid && h->np == 3) {
        // multicolour variant: the factors of P A P^T, returned in the slots of the caller's BCRS pattern
        // (slot of (i,j) holds the factor block of (p(i), p(j)): L where p(j) < p(i), U where p(j) > p(i))
        const size_t nscal = (size_t)h->nnzb * 9;
        CK(h->d_stage.ensure(nscal));
        if (h->f32) mc_scatter_factors_kernel<float><<<(unsigned)((nscal + 255) / 256), 256, 0, h->stream>>>(nscal, h->mc.psrc.p, h->mc.ppos.p, reinterpret_cast<const float*>(h->mc.uni.p), h->d_stage.p);
        else mc_scatter_factors_kernel<double><<<(unsigned)((nscal + 255) / 256), 256, 0, h->stream>>>(nscal, h->mc.psrc.p, h->mc.ppos.p, h->mc.uni.p, h->d_stage.p);
        h->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(lu, h->d_stage.p, nscal * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return OPMGPU_OK;
    }
    if (h->lu_lazy && !h->have_values)
        return h->bad("the factorised matrix values were only borrowed for the solve call (opmgpu_solve_bcrs3_dev): "
                      "set them again (opmgpu_set_values_bcrs3[_dev]) and factorise before asking for the factors");
    if (h->lu_lazy) {
        // the pipelined factorisation keeps only the pivots: build the BCRS factor array now
        // (in place on a copy of A; the matrix values must still be the ones that were factorised)
        if (h->world == 1) {
            if (h->f32) { if (int rc = encode<float>(h, (size_t)h->nnzb * 9, static_cast<const float*>(h->d_vals), h->d_lu.p)) return rc; }
            else CK(cudaMemcpyAsync(h->d_lu.p, h->d_vals, (size_t)h->nnzb * 9 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        }
        const size_t e = (size_t)h->N * 3;
        if (h->f32) materialise_lu_kernel<float><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->N, h->d_rowptr.p, h->d_colidx.p, h->pipeF.fpos.p, h->pipeF.fout.p, h->d_lu.p);
        else materialise_lu_kernel<double><<<(unsigned)((e + 255) / 256), 256, 0, h->stream>>>(h->N, h->d_rowptr.p, h->d_colidx.p, h->pipeF.fpos.p, h->pipeF.fout.p, h->d_lu.p);
        h->launches++;
        CK(cudaGetLastError());
        h->lu_lazy = false;
    }
    const double* src = h->d_lu.p;
    if (h->f32) {             // the float instance's containers -> plain doubles for the caller
        const size_t nv = (size_t)h->nnzb * 9;
        CK(h->d_stage.ensure(nv));
        const unsigned grid = (unsigned)std::min<size_t>((nv + 255) / 256, (size_t)h->sm_count * 16);
        decode_kernel<float><<<grid, 256, 0, h->stream>>>(nv, h->d_lu.p, h->d_stage.p);
        h->launches++;
        CK(cudaGetLastError());
        src = h->d_stage.p;
    }
    CK(cudaMemcpyAsync(lu, src, (size_t)h->nnzb * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return OPMGPU_OK;
}

int opmgpu_ilu0_apply_dev(opmgpu_handle h, double w, const double* d_dev, double* v_dev)
{
    if (!h || !h->have_factors) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    if (h->f32) {                // doubles at the ABI, floats inside
        const size_t n = (size_t)h->N * 3;
        float* df = vec<float>(h->d_tmp); float* vf = vec<float>(h->d_tmp2);
        int rc = convert<double, float>(h, n, d_dev, df);
        if (!rc) rc = apply_precond<float>(h, w, df, vf);
        if (!rc) rc = convert<float, double>(h, n, vf, v_dev);
        return rc;
    }
    return apply_precond<double>(h, w, d_dev, v_dev);
}

int opmgpu_ilu0_apply(opmgpu_handle h, double w, const double* d, double* v)
{
    if (!h || !h->have_factors) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->N * 3;
    int rc;
    if (h->f32) {
        CK(h->d_stage.ensure(n));
        CK(cudaMemcpyAsync(h->d_stage.p, d, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if ((rc = opmgpu_ilu0_apply_dev(h, w, h->d_stage.p, h->d_stage.p))) return rc;
        CK(cudaMemcpyAsync(h->d_tmp2.p, h->d_stage.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    } else {
        CK(cudaMemcpyAsync(h->d_tmp.p, d, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        rc = apply_precond<double>(h, w, h->d_tmp.p, h->d_tmp2.p);
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(v, h->d_tmp2.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&h->h_flags2[0], h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_flags2[0]) return sweep_watchdog(h);
    return OPMGPU_OK;
}

int opmgpu_dot(opmgpu_handle h, const double* x, const double* y, int n, double* out)
{
    if (!h || n < 0) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    DevArr<double> a, b;
    CK(a.ensure(n)); CK(b.ensure(n));
    CK(cudaMemcpyAsync(a.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(b.p, y, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    DevArr<float> af, bf;
    if (h->f32) {             // rounded to float, then the float scalar product
        CK(af.ensure(n)); CK(bf.ensure(n));
        int rc = convert<double, float>(h, (size_t)n, a.p, af.p);
        if (!rc) rc = convert<double, float>(h, (size_t)n, b.p, bf.p);
        if (rc) return rc;
        dot_kernel<float><<<kVecBlocks, 256, 0, h->stream>>>((size_t)n, af.p, bf.p, h->d_S.p, h->ws());
    } else dot_kernel<double><<<kVecBlocks, 256, 0, h->stream>>>((size_t)n, a.p, b.p, h->d_S.p, h->ws());
    h->launches++;
    CK(cudaMemcpyAsync(h->h_S, h->d_S.p, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *out = h->h_S[S_DOT];
    a.release(); b.release(); af.release(); bf.release();
    return OPMGPU_OK;
}

int opmgpu_solve_bcrs3_dev(opmgpu_handle h, const double* vals_dev, const double* rhs_dev, double* x_dev,
                           const opmgpu_params* params, opmgpu_result* result)
{
    if (!h || !rhs_dev || !x_dev || !params || !result) return OPMGPU_BAD_ARGUMENT;
    if (!h->have_pattern) return h->bad("set the pattern first");
    if (h->operator_only) return h->bad("the pattern was set operator-only (no ILU0 analysis): opmgpu_set_pattern_bcrs first");
    if (!vals_dev && !h->have_values) return h->bad("vals_dev == NULL: no matrix values are resident (opmgpu_set_values_bcrs3[_dev])");
    std::memset(result, 0, sizeof *result);
    result->bad_row = -1;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->N * 3;
    if (h->f32 || !vals_dev) {
        // values resident in the handle (vals_dev == NULL), or rounded into the float instance's own
        // array (kept: they stay valid after the call)
        int rc = 0;
        if (vals_dev) rc = take_values_f32(h, vals_dev);
        if (!rc && h->f32) rc = convert<double, float>(h, n, rhs_dev, vec<float>(h->d_r));
        if (rc) return rc;
        if (!h->f32) CK(cudaMemcpyAsync(h->d_r.p, rhs_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        h->have_values = true; h->have_factors = false;
        rc = solve_resident(h, params, result);
        if (!h->f32) {
            if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED) {
                CK(cudaMemcpyAsync(x_dev, h->d_x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
                CK(cudaStreamSynchronize(h->stream));
            }
            return rc;
        }
        if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED) {
            if (int rc2 = convert<float, double>(h, n, vec<float>(h->d_x), x_dev)) return rc2;
            CK(cudaStreamSynchronize(h->stream));
        }
        return rc;
    }
    h->d_vals = vals_dev;
    h->have_values = true; h->have_factors = false;
    CK(cudaMemcpyAsync(h->d_r.p, rhs_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    int rc = solve_resident(h, params, result);
    if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED) {
        CK(cudaMemcpyAsync(x_dev, h->d_x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    // the caller's buffer was only borrowed for this call: nothing may read it later (the lazy path
    // of opmgpu_ilu0_get_factors would); the factors in the sweep records stay usable
    h->have_values = false;
    h->d_vals = nullptr;
    return rc;
}

int opmgpu_solve_bcrs3(opmgpu_handle h, const double* vals, const double* rhs, double* x,
                       const opmgpu_params* params, opmgpu_result* result)
{
    if (!h || !vals || !rhs || !x || !params || !result) return OPMGPU_BAD_ARGUMENT;
    if (h->multi) return multi_solve_bcrs3(h->multi, vals, rhs, x, params, result, h->err);
    if (!h->have_pattern) return h->bad("set the pattern first");
    if (h->operator_only) return h->bad("the pattern was set operator-only (no ILU0 analysis): opmgpu_set_pattern_bcrs first");
    std::memset(result, 0, sizeof *result);
    result->bad_row = -1;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->N * 3, nv = (size_t)(h->world > 1 ? h->nnzb_full : h->nnzb) * 9;
    CK(h->d_vals_own.ensure(nv));
    if (h->f32) CK(h->d_stage.ensure(n));
    cudaEventRecord(h->ev[3], h->stream);
    CK(cudaMemcpyAsync(h->d_vals_own.p, vals, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->f32 ? h->d_stage.p : h->d_r.p, rhs, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    cudaEventRecord(h->ev[4], h->stream);
    if (h->f32) {
        int rc = take_values_f32(h, h->d_vals_own.p);
        if (!rc) rc = convert<double, float>(h, n, h->d_stage.p, vec<float>(h->d_r));
        if (rc) return rc;
    } else h->d_vals = h->d_vals_own.p;
    h->have_values = true; h->have_factors = false;
    int rc = solve_resident(h, params, result);
    result->ms_h2d = ev_ms(h->ev[3], h->ev[4]);
    if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED) {
        cudaEventRecord(h->ev[5], h->stream);
        if (h->f32) {
            if (int rc2 = convert<float, double>(h, n, vec<float>(h->d_x), h->d_stage.p)) return rc2;
            CK(cudaMemcpyAsync(x, h->d_stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        } else
        CK(cudaMemcpyAsync(x, h->d_x.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        cudaEventRecord(h->ev[6], h->stream);
        CK(cudaStreamSynchronize(h->stream));
        result->ms_d2h = ev_ms(h->ev[5], h->ev[6]);
    }
    return rc;
}

int opmgpu_solve_from_csc_blocks(opmgpu_handle h, int N, const opmgpu_csc blocks[9],
                                 const double matbalscale[3], const double* rhs_eqmajor,
                                 double* dx_varmajor, const opmgpu_params* params, opmgpu_result* result)
{
    if (!h || !blocks || !matbalscale || !rhs_eqmajor || !dx_varmajor || !params || !result || N < 1)
        return OPMGPU_BAD_ARGUMENT;
    if (h->multi) return multi_solve_from_csc_blocks(h->multi, N, blocks, matbalscale, rhs_eqmajor, dx_varmajor, params, result, h->err);
    std::memset(result, 0, sizeof *result);
    result->bad_row = -1;
    CK(cudaSetDevice(h->device));
    const bool full = params->require_full_sparsity_pattern != 0;
    if (h->pattern_check.valid()) h->pattern_check.get();
    // the compare threads read the caller's arrays: never return while they run
    struct WaitForCheck { opmgpu_handle h; ~WaitForCheck() { if (h->pattern_check.valid()) h->pattern_check.wait(); } } wait_for_check{h};
    // Same sizes as last time: assume the pattern is unchanged, start uploading and factorising,
    // and verify the ~290 MB of index arrays on host threads meanwhile (the reference rebuilds
    // the pattern on every call, NewtonIterationBlackoilInterleaved.cpp:110-194).
    bool speculative = same_csc_sizes(h, N, blocks, full);
    if (speculative)
        h->pattern_check = std::async(std::launch::async, [h, N, blocks, full]() { return same_csc_pattern(h, N, blocks, full); });
    for (int attempt = 0; attempt < 2; ++attempt) {
    if (!speculative) {
        cudaEventRecord(h->ev[3], h->stream);
        // pattern = union of the pressure-derivative patterns (all nine when required)
        std::vector<CscView> sel;
        for (int p1 = 0; p1 < 3; ++p1) sel.push_back({blocks[p1 * 3].colptr, blocks[p1 * 3].rowidx});
        if (full)
            for (int p1 = 0; p1 < 3; ++p1)
                for (int p2 = 1; p2 < 3; ++p2) sel.push_back({blocks[p1 * 3 + p2].colptr, blocks[p1 * 3 + p2].rowidx});
        // the index arrays are about to be trusted by host and device code: validate them once per pattern
        for (int q = 0; q < 9; ++q) {
            const int* cp = blocks[q].colptr;
            const int* ri = blocks[q].rowidx;
            if (!cp || cp[0] != 0) return h->bad("CSC block: colptr[0] != 0");
            for (int c = 0; c < N; ++c) {
                if (cp[c + 1] < cp[c]) return h->bad("CSC block: colptr not monotone");
                for (int k = cp[c]; k < cp[c + 1]; ++k) {
                    if (ri[k] < 0 || ri[k] >= N) return h->bad("CSC block: row index out of range");
                    if (k > cp[c] && ri[k] <= ri[k - 1]) return h->bad("CSC block: row indices not strictly ascending in a column");
                }
            }
        }
        std::vector<int> rowptr, colidx;
        union_pattern_from_csc(N, sel.data(), (int)sel.size(), rowptr, colidx);
        int rc = set_pattern(h, N, rowptr[N], rowptr.data(), colidx.data());
        if (rc) return rc;
        // cache the CSC index arrays (host, for the next call's comparison) and build the gather map
        h->csc_colptr.assign(9, {}); h->csc_rowidx.assign(9, {}); h->csc_base.assign(10, 0);
        for (int q = 0; q < 9; ++q) {
            h->csc_colptr[q].assign(blocks[q].colptr, blocks[q].colptr + N + 1);
            h->csc_rowidx[q].assign(blocks[q].rowidx, blocks[q].rowidx + blocks[q].colptr[N]);
            h->csc_base[q + 1] = h->csc_base[q] + blocks[q].colptr[N];
        }
        h->csc_total = h->csc_base[9];
        h->csc_full_pattern = full;
        const size_t nmap = (size_t)h->nnzb * 9;
        CK(h->d_map9.ensure(nmap));
        CK(cudaMemsetAsync(h->d_map9.p, 0xff, nmap * sizeof(long long), h->stream));
        CK(cudaMemsetAsync(h->d_bad.p, 0, sizeof(int), h->stream));
        DevArr<int> d_cp, d_ri;
        CK(d_cp.ensure((size_t)N + 1));
        size_t maxnnz = 1;
        for (int q = 0; q < 9; ++q) maxnnz = std::max(maxnnz, h->csc_rowidx[q].size());
        CK(d_ri.ensure(maxnnz));
        for (int q = 0; q < 9; ++q) {
            CK(cudaMemcpyAsync(d_cp.p, h->csc_colptr[q].data(), sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, h->stream));
            if (!h->csc_rowidx[q].empty())
                CK(cudaMemcpyAsync(d_ri.p, h->csc_rowidx[q].data(), sizeof(int) * h->csc_rowidx[q].size(), cudaMemcpyHostToDevice, h->stream));
            build_gather_map_kernel<<<(N + 255) / 256, 256, 0, h->stream>>>(N, q, d_cp.p, d_ri.p, h->csc_base[q],
                                                                            h->d_rowptr.p, h->d_colidx.p, h->d_map9.p, h->d_bad.p);
            h->launches++;
        }
        CK(cudaMemcpyAsync(&h->h_flags2[1], h->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        cudaEventRecord(h->ev[4], h->stream);
        CK(cudaStreamSynchronize(h->stream));
        d_cp.release(); d_ri.release();
        result->ms_analysis = ev_ms(h->ev[3], h->ev[4]);
        if (h->h_flags2[1]) {
            h->have_pattern = false;
            h->err = "Jacobian entry outside the interleaved sparsity pattern (set require_full_sparsity_pattern)";
            return OPMGPU_BAD_PATTERN;
        }
        CK(h->d_cscval.ensure((size_t)std::max<long long>(h->csc_total, 1)));
        CK(h->d_rhs_stage.ensure((size_t)N * 3));
    }
    if (h->f32) CK(h->d_vals32.ensure((size_t)h->nnzb * 9 + 16));
    else CK(h->d_vals_own.ensure((size_t)h->nnzb * 9));
    const size_t n = (size_t)N * 3, nv = (size_t)h->nnzb * 9;
    cudaEventRecord(h->ev[3], h->stream);
    for (int q = 0; q < 9; ++q) {
        const size_t nnz = (size_t)(h->csc_base[q + 1] - h->csc_base[q]);
        if (nnz) CK(cudaMemcpyAsync(h->d_cscval.p + h->csc_base[q], blocks[q].val, nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaMemcpyAsync(h->d_rhs_stage.p, rhs_eqmajor, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    cudaEventRecord(h->ev[4], h->stream);
    if (h->f32) {
        interleave_gather_kernel<float><<<(unsigned)((nv + 255) / 256), 256, 0, h->stream>>>(nv, h->d_map9.p, h->d_cscval.p,
            matbalscale[0], matbalscale[1], matbalscale[2], h->d_vals32.p);
        interleave_rhs_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(N, h->d_rhs_stage.p,
            matbalscale[0], matbalscale[1], matbalscale[2], vec<float>(h->d_r));
    } else {
        interleave_gather_kernel<double><<<(unsigned)((nv + 255) / 256), 256, 0, h->stream>>>(nv, h->d_map9.p, h->d_cscval.p,
            matbalscale[0], matbalscale[1], matbalscale[2], h->d_vals_own.p);
        interleave_rhs_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(N, h->d_rhs_stage.p,
            matbalscale[0], matbalscale[1], matbalscale[2], h->d_r.p);
    }
    h->launches += 2;
    CK(cudaGetLastError());
    cudaEventRecord(h->ev[5], h->stream);
    h->d_vals = h->f32 ? (const void*)h->d_vals32.p : (const void*)h->d_vals_own.p;
    h->have_values = true; h->have_factors = false;
    int rc = solve_resident(h, params, result);
    if (rc == kPatternChanged) {                  // the speculation failed: analyse the new pattern, start over
        CK(cudaStreamSynchronize(h->stream));
        speculative = false;
        continue;
    }
    if (h->pattern_check.valid() && !h->pattern_check.get() && speculative) {   // factor() failed before the verdict
        CK(cudaStreamSynchronize(h->stream));
        speculative = false;
        continue;
    }
    result->ms_h2d = ev_ms(h->ev[3], h->ev[4]);
    result->ms_interleave = ev_ms(h->ev[4], h->ev[5]);
    if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED) {
        cudaEventRecord(h->ev[5], h->stream);
        if (h->f32) deinterleave_x_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(N, vec<float>(h->d_x), h->d_tmp.p);
        else deinterleave_x_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(N, h->d_x.p, h->d_tmp.p);
        h->launches++;
        CK(cudaMemcpyAsync(dx_varmajor, h->d_tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        cudaEventRecord(h->ev[6], h->stream);
        CK(cudaStreamSynchronize(h->stream));
        result->ms_d2h = ev_ms(h->ev[5], h->ev[6]);
    }
    return rc;
    }
    return h->bad("sparsity pattern changed during the call");
}

// ---- block sizes other than 3 (the reference's Impl<np,Scalar>, np = 2..6) ---------------------------
}  // extern "C"
namespace {
struct NpGuard {          // the block size of the call in flight; back to 3 on every way out
    opmgpu_handle h;
    NpGuard(opmgpu_handle h_, int np) : h(h_) { h->np = np; }
    ~NpGuard() { h->np = 3; h->have_values = false; h->have_factors = false; h->d_vals = nullptr; }
};
int np_check(opmgpu_handle h, int np)
{
    if (np < 2 || np > 6) { h->err = "block size np = " + std::to_string(np) + " is not built (the reference instantiates np = 2..6)"; return OPMGPU_BAD_ARGUMENT; }
    if (h->multi || h->world > 1) return h->bad("np != 3 exists for plain single-GPU handles");
    if (!h->have_pattern || h->operator_only) return h->bad("set the pattern first");
    if (np != h->np_req) return h->bad("the pattern was prepared for another block size: opmgpu_set_block_size before opmgpu_set_pattern_bcrs");
    if (h->lvlU_ptr.empty()) return h->bad("the pattern was not prepared for this block size: opmgpu_set_block_size before opmgpu_set_pattern_bcrs");
    return 0;
}
// host doubles -> the instance's values (T) on the device, through d_stage
template <class T>
int np_upload(opmgpu_handle h, const double* src, size_t n, T* dst)
{
    if (sizeof(T) == 8) { CK(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream)); return 0; }
    CK(h->d_stage.ensure(n));
    CK(cudaMemcpyAsync(h->d_stage.p, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return convert<double, T>(h, n, h->d_stage.p, dst);
}
template <class T>
int np_download(opmgpu_handle h, const T* src, size_t n, double* dst)
{
    if (sizeof(T) == 8) { CK(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); }
    else {
        CK(h->d_stage.ensure(n));
        if (int rc = convert<T, double>(h, n, src, h->d_stage.p)) return rc;
        CK(cudaMemcpyAsync(dst, h->d_stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
// the values as the instance stores them: T array in d_vals_own (double) / d_vals32 (float)
template <class T>
T* np_vals_buffer(opmgpu_handle h, size_t nv)
{
    if (sizeof(T) == 8) { if (h->d_vals_own.ensure(nv) != cudaSuccess) return nullptr; return reinterpret_cast<T*>(h->d_vals_own.p); }
    if (h->d_vals32.ensure(nv + 16) != cudaSuccess) return nullptr;
    return reinterpret_cast<T*>(h->d_vals32.p);
}
template <class T>
int np_solve_bcrs(opmgpu_handle h, int np, const double* vals, const double* rhs, double* x,
                  const opmgpu_params* params, opmgpu_result* result)
{
    const size_t n = (size_t)h->N * np, nv = (size_t)h->nnzb * np * np;
    T* dv = np_vals_buffer<T>(h, nv);
    if (!dv) return h->fail(cudaErrorMemoryAllocation, "np values");
    int rc = np_upload<T>(h, vals, nv, dv);
    if (!rc) rc = np_upload<T>(h, rhs, n, vec<T>(h->d_r));
    if (rc) return rc;
    h->d_vals = dv;
    h->have_values = true; h->have_factors = false;
    rc = solve_resident(h, params, result);
    if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED) { if (int rc2 = np_download<T>(h, vec<T>(h->d_x), n, x)) return rc2; }
    return rc;
}
}  // namespace
extern "C" {

int opmgpu_set_block_size(opmgpu_handle h, int np)
{
    if (!h) return OPMGPU_BAD_ARGUMENT;
    if (np < 2 || np > 6) { h->err = "block size np = " + std::to_string(np) + " is not built (the reference instantiates np = 2..6)"; return OPMGPU_BAD_ARGUMENT; }
    if (h->multi || h->world > 1) return np == 3 ? OPMGPU_OK : h->bad("np != 3 exists for plain single-GPU handles");
    if (np != h->np_req) { h->np_req = np; h->have_pattern = false; h->have_values = false; h->have_factors = false; }
    return OPMGPU_OK;
}

int opmgpu_solve_bcrs_np(opmgpu_handle h, int np, const double* vals, const double* rhs, double* x,
                         const opmgpu_params* params, opmgpu_result* result)
{
    if (!h || !vals || !rhs || !x || !params || !result) return OPMGPU_BAD_ARGUMENT;
    if (np == 3) return opmgpu_solve_bcrs3(h, vals, rhs, x, params, result);
    if (int rc = np_check(h, np)) return rc;
    if (params->newton_use_gmres) return h->bad("restarted GMRES exists for np = 3 only");
    std::memset(result, 0, sizeof *result);
    result->bad_row = -1;
    CK(cudaSetDevice(h->device));
    NpGuard guard(h, np);
    return h->f32 ? np_solve_bcrs<float>(h, np, vals, rhs, x, params, result) : np_solve_bcrs<double>(h, np, vals, rhs, x, params, result);
}

// kernel-level entry points of the np path (parity tests): y = A x; factors and v = w U^-1 L^-1 d
int opmgpu_spmv_np(opmgpu_handle h, int np, const double* vals, const double* x, double* y)
{
    if (!h || !vals || !x || !y) return OPMGPU_BAD_ARGUMENT;
    if (int rc = np_check(h, np)) return rc;
    CK(cudaSetDevice(h->device));
    NpGuard guard(h, np);
    const size_t n = (size_t)h->N * np, nv = (size_t)h->nnzb * np * np;
    auto run = [&](auto tag) -> int {
        using T = decltype(tag);
        T* dv = np_vals_buffer<T>(h, nv);
        if (!dv) return h->fail(cudaErrorMemoryAllocation, "np values");
        int rc = np_upload<T>(h, vals, nv, dv);
        if (!rc) rc = np_upload<T>(h, x, n, vec<T>(h->d_tmp));
        if (rc) return rc;
        h->d_vals = dv;
        if ((rc = np_spmv_with_dots<T>(h, 0, vec<T>(h->d_tmp), vec<T>(h->d_tmp2), (const T*)nullptr))) return rc;
        return np_download<T>(h, vec<T>(h->d_tmp2), n, y);
    };
    return h->f32 ? run(float()) : run(double());
}

int opmgpu_ilu0_np(opmgpu_handle h, int np, const double* vals, double* lu_out, double w, const double* d, double* v, int* bad_row)
{
    if (!h || !vals) return OPMGPU_BAD_ARGUMENT;
    if (int rc = np_check(h, np)) return rc;
    CK(cudaSetDevice(h->device));
    NpGuard guard(h, np);
    const size_t n = (size_t)h->N * np, nv = (size_t)h->nnzb * np * np;
    auto run = [&](auto tag) -> int {
        using T = decltype(tag);
        T* dv = np_vals_buffer<T>(h, nv);
        if (!dv) return h->fail(cudaErrorMemoryAllocation, "np values");
        int rc = np_upload<T>(h, vals, nv, dv);
        if (rc) return rc;
        h->d_vals = dv; h->have_values = true;
        if ((rc = np_factor<T>(h, bad_row))) return rc;
        if (lu_out && (rc = np_download<T>(h, reinterpret_cast<const T*>(h->d_lu.p), nv, lu_out))) return rc;
        if (d && v) {
            if ((rc = np_upload<T>(h, d, n, vec<T>(h->d_tmp)))) return rc;
            if ((rc = np_apply<T>(h, w, vec<T>(h->d_tmp), vec<T>(h->d_tmp2)))) return rc;
            if ((rc = np_download<T>(h, vec<T>(h->d_tmp2), n, v))) return rc;
        }
        return OPMGPU_OK;
    };
    return h->f32 ? run(float()) : run(double());
}

int opmgpu_solve_from_csc_blocks_np(opmgpu_handle h, int N, int np, const opmgpu_csc* blocks,
                                    const double* matbalscale, const double* rhs_eqmajor,
                                    double* dx_varmajor, const opmgpu_params* params, opmgpu_result* result)
{
    if (!h || !blocks || !matbalscale || !rhs_eqmajor || !dx_varmajor || !params || !result || N < 1) return OPMGPU_BAD_ARGUMENT;
    if (np == 3) return opmgpu_solve_from_csc_blocks(h, N, blocks, matbalscale, rhs_eqmajor, dx_varmajor, params, result);
    if (np < 2 || np > 6) { h->err = "block size np = " + std::to_string(np) + " is not built (the reference instantiates np = 2..6)"; return OPMGPU_BAD_ARGUMENT; }
    if (h->multi || h->world > 1) return h->bad("np != 3 exists for plain single-GPU handles");
    if (params->newton_use_gmres) return h->bad("restarted GMRES exists for np = 3 only");
    std::memset(result, 0, sizeof *result);
    result->bad_row = -1;
    CK(cudaSetDevice(h->device));
    const int bb = np * np;
    const bool full = params->require_full_sparsity_pattern != 0;
    // the pattern: compared with the cached one on the host (np = 2 systems are small next to C3)
    bool same = h->have_pattern && h->np_req == np && h->N == N && (int)h->npcsc_colptr.size() == bb && !h->lvlU_ptr.empty() &&
                h->csc_full_pattern == full;
    for (int q = 0; q < bb && same; ++q) {
        same = (int)h->npcsc_colptr[q].size() == N + 1 && std::memcmp(h->npcsc_colptr[q].data(), blocks[q].colptr, sizeof(int) * ((size_t)N + 1)) == 0 &&
               h->npcsc_rowidx[q].size() == (size_t)blocks[q].colptr[N] &&
               std::memcmp(h->npcsc_rowidx[q].data(), blocks[q].rowidx, sizeof(int) * h->npcsc_rowidx[q].size()) == 0;
    }
    if (!same) {
        cudaEventRecord(h->ev[3], h->stream);
        for (int q = 0; q < bb; ++q) {
            const int* cp = blocks[q].colptr;
            const int* ri = blocks[q].rowidx;
            if (!cp || cp[0] != 0) return h->bad("CSC block: colptr[0] != 0");
            for (int c = 0; c < N; ++c) {
                if (cp[c + 1] < cp[c]) return h->bad("CSC block: colptr not monotone");
                for (int k = cp[c]; k < cp[c + 1]; ++k) {
                    if (ri[k] < 0 || ri[k] >= N) return h->bad("CSC block: row index out of range");
                    if (k > cp[c] && ri[k] <= ri[k - 1]) return h->bad("CSC block: row indices not strictly ascending in a column");
                }
            }
        }
        std::vector<CscView> sel;
        for (int p1 = 0; p1 < np; ++p1) sel.push_back({blocks[p1 * np].colptr, blocks[p1 * np].rowidx});
        if (full)
            for (int p1 = 0; p1 < np; ++p1)
                for (int p2 = 1; p2 < np; ++p2) sel.push_back({blocks[p1 * np + p2].colptr, blocks[p1 * np + p2].rowidx});
        std::vector<int> rowptr, colidx;
        union_pattern_from_csc(N, sel.data(), (int)sel.size(), rowptr, colidx);
        h->np_req = np;
        h->csc_colptr.clear(); h->csc_rowidx.clear();
        int rc = set_pattern(h, N, rowptr[N], rowptr.data(), colidx.data());
        if (rc) return rc;
        h->npcsc_colptr.assign(bb, {}); h->npcsc_rowidx.assign(bb, {}); h->npcsc_base.assign((size_t)bb + 1, 0);
        for (int q = 0; q < bb; ++q) {
            h->npcsc_colptr[q].assign(blocks[q].colptr, blocks[q].colptr + N + 1);
            h->npcsc_rowidx[q].assign(blocks[q].rowidx, blocks[q].rowidx + blocks[q].colptr[N]);
            h->npcsc_base[q + 1] = h->npcsc_base[q] + blocks[q].colptr[N];
        }
        h->csc_full_pattern = full;
        const size_t nmap = (size_t)h->nnzb * bb;
        CK(h->d_map_np.ensure(nmap));
        CK(cudaMemsetAsync(h->d_map_np.p, 0xff, nmap * sizeof(long long), h->stream));
        CK(cudaMemsetAsync(h->d_bad.p, 0, sizeof(int), h->stream));
        DevArr<int> d_cp, d_ri;
        CK(d_cp.ensure((size_t)N + 1));
        size_t maxnnz = 1;
        for (int q = 0; q < bb; ++q) maxnnz = std::max(maxnnz, h->npcsc_rowidx[q].size());
        CK(d_ri.ensure(maxnnz));
        for (int q = 0; q < bb; ++q) {
            CK(cudaMemcpyAsync(d_cp.p, h->npcsc_colptr[q].data(), sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, h->stream));
            if (!h->npcsc_rowidx[q].empty())
                CK(cudaMemcpyAsync(d_ri.p, h->npcsc_rowidx[q].data(), sizeof(int) * h->npcsc_rowidx[q].size(), cudaMemcpyHostToDevice, h->stream));
            np_build_gather_map_kernel<<<(N + 255) / 256, 256, 0, h->stream>>>(N, q, bb, d_cp.p, d_ri.p, h->npcsc_base[q],
                                                                               h->d_rowptr.p, h->d_colidx.p, h->d_map_np.p, h->d_bad.p);
            h->launches++;
            CK(cudaStreamSynchronize(h->stream));          // d_cp / d_ri are reused by the next block
        }
        CK(cudaMemcpyAsync(&h->h_flags2[1], h->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        cudaEventRecord(h->ev[4], h->stream);
        CK(cudaStreamSynchronize(h->stream));
        d_cp.release(); d_ri.release();
        result->ms_analysis = ev_ms(h->ev[3], h->ev[4]);
        if (h->h_flags2[1]) {
            h->have_pattern = false;
            h->err = "Jacobian entry outside the interleaved sparsity pattern (set require_full_sparsity_pattern)";
            return OPMGPU_BAD_PATTERN;
        }
    }
    NpGuard guard(h, np);
    const long long total = h->npcsc_base[bb];
    const size_t n = (size_t)N * np, nv = (size_t)h->nnzb * bb;
    CK(h->d_cscval.ensure((size_t)std::max<long long>(total, 1)));
    CK(h->d_rhs_stage.ensure(n));
    for (int q = 0; q < bb; ++q) {
        const size_t nnz = (size_t)(h->npcsc_base[q + 1] - h->npcsc_base[q]);
        if (nnz) CK(cudaMemcpyAsync(h->d_cscval.p + h->npcsc_base[q], blocks[q].val, nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaMemcpyAsync(h->d_rhs_stage.p, rhs_eqmajor, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    NpScale sc;
    for (int p = 0; p < 6; ++p) sc.s[p] = p < np ? matbalscale[p] : 1.0;
    auto run = [&](auto tag) -> int {
        using T = decltype(tag);
        T* dv = np_vals_buffer<T>(h, nv);
        if (!dv) return h->fail(cudaErrorMemoryAllocation, "np values");
        np_interleave_gather_kernel<T><<<(unsigned)((nv + 255) / 256), 256, 0, h->stream>>>(nv, np, h->d_map_np.p, h->d_cscval.p, sc, dv);
        np_interleave_rhs_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(N, np, h->d_rhs_stage.p, sc, vec<T>(h->d_r));
        h->launches += 2;
        CK(cudaGetLastError());
        h->d_vals = dv;
        h->have_values = true; h->have_factors = false;
        int rc = solve_resident(h, params, result);
        if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED) {
            np_deinterleave_x_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(N, np, vec<T>(h->d_x), h->d_tmp.p);
            h->launches++;
            CK(cudaMemcpyAsync(dx_varmajor, h->d_tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        return rc;
    };
    return h->f32 ? run(float()) : run(double());
}

// ---- multicolour ILU0 variant (flagged: NOT the reference's preconditioner) ----------------------
int opmgpu_set_ilu_ordering(opmgpu_handle h, int ordering)
{
    if (!h) return OPMGPU_BAD_ARGUMENT;
    if (ordering != OPMGPU_ILU_NATURAL && ordering != OPMGPU_ILU_MULTICOLOUR && ordering != OPMGPU_ILU_MULTICOLOUR_LINES)
        return h->bad("unknown ILU0 ordering");
    if (ordering != OPMGPU_ILU_NATURAL && (h->multi || h->world > 1 || h->comm))
        return h->bad("the multicolour ILU0 variant exists for plain single-GPU handles");
    if (ordering != h->ilu_order_req) {          // takes effect with the next pattern
        h->ilu_order_req = ordering;
        h->have_pattern = h->have_values = h->have_factors = false;
        h->csc_colptr.clear(); h->csc_rowidx.clear();
    }
    return OPMGPU_OK;
}
int opmgpu_get_ilu_ordering(opmgpu_handle h) { return h ? h->ilu_order_req : OPMGPU_ILU_NATURAL; }

int opmgpu_multicolour_order(int N, const int* rowptr, const int* colidx, int* ncolours, int* colour, int* n2p)
{
    if (N < 1 || !rowptr || !colidx) return OPMGPU_BAD_ARGUMENT;
    McOrder o;
    multicolour_order(N, rowptr, colidx, o);
    if (ncolours) *ncolours = o.ncolours;
    if (colour) std::copy(o.colour.begin(), o.colour.end(), colour);
    if (n2p) std::copy(o.n2p.begin(), o.n2p.end(), n2p);
    return OPMGPU_OK;
}

int opmgpu_line_order(int nx, int ny, int nz, int* n2p)
{
    if (nx < 1 || ny < 1 || nz < 1 || !n2p) return OPMGPU_BAD_ARGUMENT;
    McOrder o;
    line_order(nx, ny, nz, o);
    std::copy(o.n2p.begin(), o.n2p.end(), n2p);
    return OPMGPU_OK;
}

int opmgpu_get_ilu_permutation(opmgpu_handle h, int* ncolours, int* n2p)
{
    if (!h || !h->have_pattern) return OPMGPU_BAD_ARGUMENT;
    if (!h->mc.valid) return h->bad("the current pattern uses the natural ordering");
    if (ncolours) *ncolours = h->mc.ncolours;
    if (n2p) std::copy(h->mc.n2p_host.begin(), h->mc.n2p_host.end(), n2p);
    return OPMGPU_OK;
}

int opmgpu_num_levels(opmgpu_handle h, int* lower_levels, int* upper_levels)
{
    if (!h || !h->have_pattern) return OPMGPU_BAD_ARGUMENT;
    if (lower_levels) *lower_levels = h->nlevL;
    if (upper_levels) *upper_levels = h->nlevU;
    return OPMGPU_OK;
}

long long opmgpu_launch_count(opmgpu_handle h) { return h ? h->launches : 0; }

int opmgpu_residual_history(opmgpu_handle h, double* out, int cap, int* n)
{
    if (!h || !n) return OPMGPU_BAD_ARGUMENT;
    *n = (int)h->history.size();
    for (int i = 0; i < cap && i < *n; ++i) out[i] = h->history[i];
    return OPMGPU_OK;
}

int opmgpu_set_profiling(opmgpu_handle h, int on)
{
    if (!h) return OPMGPU_BAD_ARGUMENT;
    h->profile = on != 0;
    for (int k = 0; k < 4; ++k) { h->prof_ms[k] = 0; h->prof_cnt[k] = 0; }
    return OPMGPU_OK;
}

int opmgpu_get_profile(opmgpu_handle h, double ms[4], long long count[4])
{
    if (!h) return OPMGPU_BAD_ARGUMENT;
    for (int k = 0; k < 4; ++k) { ms[k] = h->prof_ms[k]; count[k] = h->prof_cnt[k]; }
    return OPMGPU_OK;
}

// Test hook (declared in the header; tests of the recovery path): sets the device watchdog word
// as a sweep kernel does when a dependency is never delivered.  The next call that collects it
// must fail with OPMGPU_CUDA_ERROR ("sweep watchdog"), re-arm the push slots and leave the handle usable.
int opmgpu_debug_set_watchdog_word(opmgpu_handle h, int code)
{
    if (!h || h->multi) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->d_err.p, &code, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return OPMGPU_OK;
}

#ifdef OPMGPU_EXPERIMENTS
// Debug (not in the public header): clock64 stamps of one CTA of the next pipelined apply.
// out[2][512][8]: lower then upper sweep; per step {enter, landed, computed, after barrier,
// nrows, bulk issued, rhs gather issued, -}.
int opmgpu_debug_trace_apply(opmgpu_handle h, int cta, double w, const double* d_dev, double* v_dev, long long* out)
{
    if (!h || !h->have_factors || !h->use_pipe || h->use_col) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    CK(h->d_trace.ensure(2 * 512 * 16));
    CK(cudaMemsetAsync(h->d_trace.p, 0, sizeof(long long) * 2 * 512 * 16, h->stream));
    h->trace_cta = cta;
    int rc = apply_precond(h, w, d_dev, v_dev);
    h->trace_cta = -1;
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, h->d_trace.p, sizeof(long long) * 2 * 512 * 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return OPMGPU_OK;
}

// Debug (not in the public header): %globaltimer stamps (ns) of every CTA of the next pipelined
// apply.  Per sweep: out[P][steps][8], per step {record landed, bar.sync passed, pushed inputs
// ready, chain done, ext_end, rows}; the last step slot holds {kernel entry, helper's first
// delivery, its polls}; then the helper's deliveries [P][1024][2] = {time, polls<<32 | ext_ready}.
int opmgpu_debug_gtrace_apply(opmgpu_handle h, int steps, double w, const double* d_dev, double* v_dev, long long* out, int* P_out)
{
    if (!h || !h->have_factors || !h->use_pipe || h->use_col || steps < 2) return OPMGPU_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)(h->pipeL.P + h->pipeU.P) * (steps * 8 + 2048);
    if (P_out) { P_out[0] = h->pipeL.P; P_out[1] = h->pipeU.P; }
    if (!out) return OPMGPU_OK;
    CK(h->d_gtrace.ensure(n));
    CK(cudaMemsetAsync(h->d_gtrace.p, 0, sizeof(long long) * n, h->stream));
    h->gtrace_steps = steps;
    int rc = apply_precond(h, w, d_dev, v_dev);
    h->gtrace_steps = 0;
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, h->d_gtrace.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return OPMGPU_OK;
}

#endif

// ---- multi-GPU entry points -------------------------------------------------------------------
int opmgpu_nccl_unique_id(void* id128)
{
    if (!id128) return OPMGPU_BAD_ARGUMENT;
    if (!g_nccl.load(g_create_error)) return OPMGPU_NCCL_ERROR;
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return OPMGPU_NCCL_ERROR; }
    std::memcpy(id128, &id, sizeof id);
    return OPMGPU_OK;
}

int opmgpu_create_distributed(int device, int rank, int world, const void* nccl_unique_id, opmgpu_handle* out)
{
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world || !nccl_unique_id) { g_create_error = "bad rank / world / id"; return OPMGPU_BAD_ARGUMENT; }
    if (!g_nccl.load(g_create_error)) return OPMGPU_NCCL_ERROR;
    opmgpu_handle h = nullptr;
    int rc = opmgpu_create(device, &h);
    if (rc) return rc;
    h->rank = rank; h->world = world;
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof id);
    const ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, id, rank);
    if (r != ncclSuccess) {
        g_create_error = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
        opmgpu_destroy(h);
        return OPMGPU_NCCL_ERROR;
    }
    // NCCL sets its channels up lazily, on the first collective / point-to-point call that uses them
    // (seconds on an 8-GPU box): do that once here, not inside the first pattern analysis or solve.
    {
        DevArr<double> warm;
        bool ok = warm.ensure((size_t)2 * world) == cudaSuccess && cudaMemsetAsync(warm.p, 0, sizeof(double) * 2 * world, h->stream) == cudaSuccess;
        ok = ok && g_nccl.AllReduce(warm.p, warm.p, 1, ncclDouble, ncclSum, h->comm, h->stream) == ncclSuccess;
        if (ok && world > 1) {
            ok = g_nccl.GroupStart() == ncclSuccess;
            for (int p = 0; p < world && ok; ++p) {
                if (p == rank) continue;
                ok = g_nccl.Send(warm.p + p, 1, ncclDouble, p, h->comm, h->stream) == ncclSuccess &&
                     g_nccl.Recv(warm.p + world + p, 1, ncclDouble, p, h->comm, h->stream) == ncclSuccess;
            }
            ok = g_nccl.GroupEnd() == ncclSuccess && ok;
        }
        ok = ok && cudaStreamSynchronize(h->stream) == cudaSuccess;
        warm.release();
        if (!ok) { g_create_error = "NCCL warm-up (all-reduce + send/recv) failed"; opmgpu_destroy(h); return OPMGPU_NCCL_ERROR; }
    }
    *out = h;
    return OPMGPU_OK;
}

int opmgpu_create_multi(int ngpus, const int* device_ids, opmgpu_handle* out)
{
    if (!out) return OPMGPU_BAD_ARGUMENT;
    *out = nullptr;
    MultiSolver* m = multi_create(ngpus, device_ids, g_create_error);
    if (!m) return g_create_error.find("NCCL") != std::string::npos ? OPMGPU_NCCL_ERROR : OPMGPU_CUDA_ERROR;
    opmgpu_handle h = new opmgpu_solver();
    h->multi = m;
    *out = h;
    return OPMGPU_OK;
}

int opmgpu_multi_partition(opmgpu_handle h, int* axis, long long* row_offsets)
{
    if (!h || !h->multi) return OPMGPU_BAD_ARGUMENT;
    return multi_partition_info(h->multi, axis, row_offsets);
}

int opmgpu_set_pattern_bcrs_distributed(opmgpu_handle h, int N_local, int nnzb_local, const int* rowptr,
                                        const long long* colidx_global, const long long* row_offsets)
{
    if (!h || !rowptr || !colidx_global || !row_offsets) return OPMGPU_BAD_ARGUMENT;
    if (!h->comm) return h->bad("handle was not created with opmgpu_create_distributed");
    if (N_local != (int)(row_offsets[h->rank + 1] - row_offsets[h->rank]) || rowptr[N_local] != nnzb_local)
        return h->bad("local row count does not match row_offsets");
    CK(cudaSetDevice(h->device));
    if (h->peer.ready) {
        // neighbours have this rank's vectors mapped: nobody may free or remap anything before every
        // rank has dropped its mappings
        CK(cudaStreamSynchronize(h->stream));
        close_peer_halo(h);
        NK(g_nccl.AllReduce(h->d_bad.p, h->d_bad.p, 1, ncclInt32, ncclMax, h->comm, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    LocalPartition lp;
    partition_local_rows(N_local, rowptr, colidx_global, row_offsets, h->world, h->rank, lp);
    h->n_ghost = lp.n_ghost;
    h->nnzb_full = nnzb_local;
    h->csc_colptr.clear(); h->csc_rowidx.clear();
    int rc = set_pattern(h, N_local, (int)lp.colidx_diag.size(), lp.rowptr_diag.data(), lp.colidx_diag.data());
    if (rc) return rc;
    h->have_pattern = false;
    CK(h->d_rowptr_full.ensure((size_t)N_local + 1 + 8));
    CK(cudaMemcpyAsync(h->d_rowptr_full.p, rowptr, sizeof(int) * ((size_t)N_local + 1), cudaMemcpyHostToDevice, h->stream));
    lp.colidx_full.resize(lp.colidx_full.size() + 8, 0);      // padding for the SpMV's 16-byte bulk copies
    if ((rc = upload(h, h->d_colidx_full, lp.colidx_full))) return rc;
    if ((rc = upload(h, h->d_lu_src, lp.lu_src))) return rc;
    {
        // rows that reference a ghost column wait for the halo; the others do not
        static_assert(kSpmvRows == 64, "one 64-bit word of boundary-row flags per SpMV tile");
        std::vector<int> bnd;
        std::vector<unsigned long long> skip((size_t)(N_local + kSpmvRows - 1) / kSpmvRows + 1, 0ull);
        for (int r = 0; r < N_local; ++r) {
            bool ghost = false;
            for (int k = rowptr[r]; k < rowptr[r + 1] && !ghost; ++k) ghost = lp.colidx_full[k] >= N_local;
            if (ghost) { bnd.push_back(r); skip[r / kSpmvRows] |= 1ull << (r % kSpmvRows); }
        }
        h->n_bnd_rows = (int)bnd.size();
        if (bnd.empty()) bnd.push_back(0);
        if ((rc = upload(h, h->d_bnd_rows, bnd))) return rc;
        if ((rc = upload(h, h->d_row_skip, skip))) return rc;
        if (!h->halo_stream) {
            int prio_lo = 0, prio_hi = 0;
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
            CK(cudaStreamCreateWithPriority(&h->halo_stream, cudaStreamNonBlocking, prio_hi));      // its small kernels go first when SMs free up
            CK(cudaEventCreateWithFlags(&h->ev_x_ready, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->ev_halo_done, cudaEventDisableTiming));
        }
        if (getenv("OPMGPU_DEBUG"))
            fprintf(stderr, "[opmgpu] rank %d: %d of %d rows reference ghost columns (computed after the halo exchange)\n", h->rank, h->n_bnd_rows, N_local);
    }
    // halo plan: tell every owner which of its rows this rank needs
    const int W = h->world;
    h->recv_cnt = lp.recv_cnt; h->recv_off = lp.recv_off;
    DevArr<int> d_cnt_mine, d_cnt_all;
    CK(d_cnt_mine.ensure(W)); CK(d_cnt_all.ensure((size_t)W * W));
    CK(cudaMemcpyAsync(d_cnt_mine.p, lp.recv_cnt.data(), sizeof(int) * W, cudaMemcpyHostToDevice, h->stream));
    NK(g_nccl.AllGather(d_cnt_mine.p, d_cnt_all.p, (size_t)W, ncclInt32, h->comm, h->stream));
    std::vector<int> cnt_all((size_t)W * W);
    CK(cudaMemcpyAsync(cnt_all.data(), d_cnt_all.p, sizeof(int) * W * W, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->send_cnt.assign(W, 0); h->send_off.assign(W, 0);
    for (int p = 0; p < W; ++p) h->send_cnt[p] = cnt_all[(size_t)p * W + h->rank];      // what p wants from me
    for (int p = 1; p < W; ++p) h->send_off[p] = h->send_off[p - 1] + h->send_cnt[p - 1];
    h->n_send = h->send_off[W - 1] + h->send_cnt[W - 1];
    DevArr<long long> d_want, d_asked;
    CK(d_want.ensure(std::max(lp.n_ghost, 1)));
    CK(d_asked.ensure(std::max(h->n_send, 1)));
    if (lp.n_ghost) CK(cudaMemcpyAsync(d_want.p, lp.ghost_global.data(), sizeof(long long) * lp.n_ghost, cudaMemcpyHostToDevice, h->stream));
    NK(g_nccl.GroupStart());
    for (int p = 0; p < W; ++p) {
        if (lp.recv_cnt[p]) NK(g_nccl.Send(d_want.p + lp.recv_off[p], (size_t)lp.recv_cnt[p], ncclInt64, p, h->comm, h->stream));
        if (h->send_cnt[p]) NK(g_nccl.Recv(d_asked.p + h->send_off[p], (size_t)h->send_cnt[p], ncclInt64, p, h->comm, h->stream));
    }
    NK(g_nccl.GroupEnd());
    std::vector<long long> asked(std::max(h->n_send, 1));
    if (h->n_send) CK(cudaMemcpyAsync(asked.data(), d_asked.p, sizeof(long long) * h->n_send, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::vector<int> send_rows(std::max(h->n_send, 1), 0);
    for (int k = 0; k < h->n_send; ++k) {
        const long long loc = asked[k] - row_offsets[h->rank];
        if (loc < 0 || loc >= N_local) return h->bad("a peer asked for a row this rank does not own");
        send_rows[k] = (int)loc;
    }
    if ((rc = upload(h, h->d_send_rows, send_rows))) return rc;
    CK(h->d_sendbuf.ensure((size_t)std::max(h->n_send, 1) * 3));
    if ((rc = ensure_vectors(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    d_cnt_mine.release(); d_cnt_all.release(); d_want.release(); d_asked.release();
    if ((rc = setup_peer_halo(h))) return rc;
    h->have_pattern = true;
    return OPMGPU_OK;
}

}  // extern "C"
