// Host-side analysis of the BCRS sparsity pattern, done once per pattern and cached in the
// solver handle (the reference rebuilds everything every Newton iteration,
// opm/autodiff/NewtonIterationBlackoilInterleaved.cpp:256; the pattern only changes when
// wells change).
//
// Produces, for the forward (L) and backward (U) sweeps of the natural-order block ILU0:
//   * the dependency level of every row (level(i) = 1 + max level of the rows it needs);
//   * a partition of the rows over P persistent CTAs -- (i,j) column tiles when the pattern
//     is a Cartesian stencil (so that most dependencies stay inside one CTA and the CTAs
//     form a pipelined wavefront), level-round-robin otherwise;
//   * per CTA a "program": its rows in ascending level order, grouped into steps of equal
//     level, with the off-diagonal blocks of each row laid out contiguously in the order
//     the reference visits them (ascending columns in L, descending in U --
//     ParallelOverlappingILU0::apply), so a CTA streams its part of the factors linearly.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>

// tuning switches of experiments exist only in builds with -DOPMGPU_EXPERIMENTS (libopmgpu_exp.so)
#ifdef OPMGPU_EXPERIMENTS
static inline const char* exp_env(const char* name) { return std::getenv(name); }
#else
static inline const char* exp_env(const char*) { return nullptr; }
#endif

namespace opmgpu {

constexpr int kExtBit = 0x40000000;          // pcol flag: dependency owned by another CTA

struct SweepProgram {
    int P = 0;                               // CTAs
    int nlevels = 0;
    int max_step_rows = 0;
    std::vector<int> cta_step_ptr;           // [P+1]   -> steps
    std::vector<int> step_row_ptr;           // [nsteps+1] -> program rows
    std::vector<int> prow;                   // [N]   original row of program row
    std::vector<int> pblk_ptr;               // [N+1] -> program blocks
    std::vector<int> pcol;                   // [nblk] dependency row | kExtBit
    std::vector<int> psrc;                   // [nblk] BCRS slot the block comes from
    std::vector<unsigned char> publish;      // [N]   program row has a consumer in another CTA
    // factorisation only (lower program): for every block (i,j) the pairs (slot of A_jk, slot of
    // A_ik), k > j present in both rows, that A_ik -= L_ij * A_jk touches
    std::vector<int> pair_ptr, pair_jk, pair_ik;
    // the same, packed for the factorisation kernel: frow[q] = {row, diag slot, first entry,
    // entry count | kFactorSimple}, fent[b] = {slot ij, diag slot of j, dep row | kExtBit,
    // pair count, first pair's jk, first pair's ik, pair_ptr, -}
    std::vector<int> frow, fent;
    // inverted pivot blocks travel between CTAs through self-validating push slots (9 doubles,
    // all-ones = empty) when both rows are "simple": fent[b][7] = slot or -1 (then the per-row
    // flag is used), fpush_ptr/fpush_slot = slots a row's pivot is pushed to, needs_flag[q] = some
    // consumer in another CTA still relies on the flag
    std::vector<int> fpush_ptr, fpush_slot;
    std::vector<unsigned char> needs_flag;
    int n_fslots = 0;
};
constexpr int kFactorSimple = 1 << 30;       // <= 3 entries, each touching only the row's diagonal

// ---- pipelined sweep program (the fast path) ---------------------------------------------------
// Per CTA a byte stream of step records, one record per step, consumed linearly through a
// shared-memory ring filled by bulk async copies (TMA).  A record is laid out for the thread
// that consumes it: thread slot j = 3*r + c owns component c of the step's r-th row and finds
// everything it needs at fixed offsets from one base address.  Record layout:
//   int hdr[8]: nrows, qbase (program row of the first row, CTA-local), ext_end, ext entries of
//               this step, ntail, off_lists/8, off_tailvals/8, has_lists
//   double cf[3n][9]       cf[j][k*3+e]: entry [c][e] of the k-th off-diagonal block of row r
//                          (k < kFastBlocks, reference visiting order; absent blocks are 0)
//   double dinv[3n][3]     dinv[j][e]: entry [c][e] of the inverted diagonal block (upper only)
//   int rowints[n][8]      rowinfo (global row | kRowWriteGlobal | kRowSlow), dep[3], upos,
//                          push[2], own window slot (doubles).   dep[k]: dependency of block k = index (in doubles)
//                          into the CTA's shared dependency array, or kDepGlobalBit | global
//                          row; upos (lower only): position of the row in the upper sweep's
//                          program order; push: global push slot ids (-1: none)
//   -- only when the step has slow rows (more than 3 blocks / more than 2 pushes):
//   int tail_end[n], xpush_end[n]; int tail_dep[ntail]; int xpush_slot[npushx];
//   double tail_vals[ntail][9]
// The shared dependency array holds 3-double entries: [0, kWindowRows) the CTA's own recent
// results (slot = program row % kWindowRows), [kWindowRows, kWindowRows+kExtRing) results
// pushed by other CTAs (slot = kWindowRows + ordinal % kExtRing), then one all-zero entry that
// absent blocks point to.
constexpr int kFastBlocks = 3;
constexpr int kDepGlobalBit = (int)0x80000000;
constexpr int kDepValueMask = 0x3fffffff;
constexpr int kRowWriteGlobal = 1 << 30;
constexpr int kRowSlow = (int)0x80000000;
constexpr int kRowMask = (1 << 30) - 1;
constexpr int kWindowRows = 512;          // W: results kept in shared memory per CTA
constexpr int kExtRing = 512;             // R: pushed results staged in shared memory per CTA
constexpr int kDepZeroSlot = kWindowRows + kExtRing;
constexpr int kMaxStepRows = 240;
constexpr int kMaxStepBytes = 64 * 1024;
constexpr int kMaxStepExt = kExtRing - 96;
constexpr int kLeanStepRows = 96;         // rows one pass of the compute warps covers (sweep_pipe.cuh: kPipeRowsPerPass)
// Thread-block clusters (lean programs only): CTAs cs*c .. cs*c+cs-1 form a cluster; a result
// needed by another CTA of the same cluster is written by its producer straight into that
// CTA's shared memory (distributed shared memory), into a per-CTA array of 3-double entries that
// follows the dependency array: dep code = (kCxBase + index) * 3.  Entries are written once per
// sweep (no reuse), initialised to all-ones by the consumer before the cluster starts.
// push ids: kPushDsmem | rank in cluster << 20 | entry index.
constexpr int kCxBase = kDepZeroSlot + 2;
constexpr int kPushDsmem = 0x40000000;
constexpr int kMaxCxEntries = 3072;       // 72 KB of shared memory at most
struct ClusterCaps { int max_ctas[4] = {0, 0, 0, 0}; };   // co-resident CTAs at cluster size 1, 2, 4, 8

struct PipeProgram {
    bool valid = false;
    bool lean = false;                        // no slow rows, no own-global reads, steps <= one pass
    int P = 0, nlevels = 0;
    int cluster_size = 1, max_cx = 0;            // CTAs per cluster; most intra-cluster entries of a CTA
    int max_step_bytes = 0, max_step_rows = 0;   // rows padded to even
    long long total_ext = 0;
    long long nperm = 0;                      // length (rows) of vectors in this program's order
    // The records are mostly factor VALUES, which arrive on the device with every factorisation;
    // the host only produces their integer parts (header, rowints, lists: ~11 % of the bytes) as a
    // compact stream, which a device kernel spreads into the zero-filled record buffer
    // (expand_records_kernel).  buf = the full records, only materialised for the host interpreter.
    std::vector<unsigned char> buf;           // all records, CTA after CTA (materialise_records)
    std::vector<unsigned char> ibuf;          // per step: header (32 bytes) | integer region of the record
    std::vector<unsigned> step_ioff16;        // [nsteps] offset of the step in ibuf / 16
    std::vector<unsigned> step_ilen;          // [nsteps] bytes of the step in ibuf (header included)
    std::vector<unsigned> step_roff;          // [nsteps] offset of the integer region inside the record
    size_t total_bytes = 0;                   // size of the record buffer
    std::vector<int> cta_step_ptr;            // [P+1]
    std::vector<unsigned> step_off16;         // [nsteps] record offset / 16
    std::vector<unsigned> step_bytes;         // [nsteps]
    std::vector<unsigned> step_rhs_row;       // [nsteps] first row of the step in program order (even)
    std::vector<unsigned> step_rhs_bytes;     // [nsteps] bytes of its rhs segment (rows padded to even)
    std::vector<int> perm_row;                // [nperm] natural row at each program position (-1 pad)
    std::vector<long long> cta_ext_base;      // [P+1] first global push slot of the CTA
    // where the factorisation's values go: (BCRS slot, element) -> double index in buf
    std::vector<int> val_src;                 // [nval] BCRS slot
    std::vector<unsigned> val_dst8;           // [nval] double index of element [0][0]...
    std::vector<int> val_stride;              // [nval] element [c][e] goes to dst8 + c*stride + e
};

// ---- pipelined factorisation program (stencil-like patterns) -----------------------------------
// Same tiles, steps (= dependency levels) and push-slot discipline as the lower sweep program,
// but the value that travels between rows is the inverted 3x3 pivot block (9 doubles), and a
// row's record carries the original A blocks it needs: A_ii, and per lower block k the pair
// A_ij (row i) and A_ji (row j; the only entry of row j that A_ij touches in row i -- every row
// must be "simple": at most kFastBlocks lower blocks, each updating only the diagonal).
// Record layout:
//   int hdr[8]:        nrows, qbase, ext_end, ext entries of this step, rows of the previous step, 0, 0, 0
//   int rowints[n][12] row, dep[3] (entry index in the CTA's shared dependency array), mask
//                      (bit k: lower block k present, bit 4+k: A_ji present), push[2] (global
//                      push slot ids, -1: none), own window entry, BCRS slot of L_ij [3], BCRS
//                      slot of the diagonal
//   double vals[n][63] A_ii | k = 0..2: A_ij, A_ji   (row-major blocks, 0 when absent)
// Shared dependency array: kFEntry-double entries, [0, kFWindow) own recent pivots (entry =
// program row % kFWindow), [kFWindow, kFWindow + kFRing) pivots pushed by other CTAs (ordinal %
// kFRing).  The kernel's only output is the inverted pivots in PROGRAM order (fpos[row] = the
// row's position, kFEntry doubles per row), written with bulk stores straight from the window;
// L_ij = A_ij * inv(D_j) is formed by the consumers (repack kernels) from A and the pivots.
constexpr int kFWindow = 320;               // >= 3 steps of kLeanStepRows rows (see the store protocol in factor_pipe.cuh)
constexpr int kFRing = 256;
constexpr int kFEntry = 10;                 // doubles per entry: 9 values + 1 pad, so entries are 16-byte aligned
constexpr int kFPoll = 64;                 // slots the helper warp examines per poll
constexpr int kFMaxStepExt = 80;           // per step; the ring must hold the entries of the three steps in flight
constexpr int kFRowInts = 12;
constexpr int kFRowVals = 63;

struct FactorPipeProgram {
    bool valid = false;
    int P = 0;
    int max_step_bytes = 0, max_step_rows = 0;
    long long total_ext = 0;
    std::vector<unsigned char> buf;           // all records, CTA after CTA (materialise_records)
    std::vector<unsigned char> ibuf;          // compact integer parts, as in PipeProgram
    std::vector<unsigned> step_ioff16, step_ilen, step_roff;
    size_t total_bytes = 0;
    std::vector<int> cta_step_ptr;            // [P+1]
    std::vector<unsigned> step_off16;         // [nsteps] record offset / 16
    std::vector<unsigned> step_bytes;         // [nsteps]
    std::vector<long long> cta_ext_base;      // [P+1] first global push slot of the CTA
    std::vector<int> cta_row_base;            // [P+1] program position of the CTA's first row
    std::vector<int> fpos;                    // [N] program position of every natural row
    std::vector<int> val_src;                 // [nval] BCRS slot of an A block ...
    std::vector<unsigned> val_dst8;           // ... and the double index in buf of its 9 values
};

struct PatternAnalysis {
    int N = 0, nnzb = 0;
    std::vector<int> diag;                   // [N] BCRS slot of the diagonal block
    // natural-order level sets of the lower triangle (factorisation)
    std::vector<int> lvl_ptr, lvl_rows;
    int grid_nx = 0, grid_ny = 0, grid_nz = 0;   // inferred Cartesian structure (0 = none)
    int P = 0;                                   // CTAs every program is laid out for (launch size)
    int cluster_size = 1;                        // sweeps are launched in clusters of this many CTAs
    int tiles_a = 0, tiles_b = 0;                // column tiling (Cartesian patterns)
    std::vector<int> owner_, level_lower_;       // partition and L levels (build_tile_factor_program)
    SweepProgram lower, upper;
    PipeProgram pipeL, pipeU;
    FactorPipeProgram pipeF;
    int missing_diag_row = -1;
    int nlevL = 0, nlevU = 0;
};

// Cartesian structure of a pattern in natural ordering (0 0 0 when there is none): nx, ny, nz.
void infer_cartesian_grid(int N, const int* rowptr, const int* colidx, int& nx, int& ny, int& nz);

// P = number of persistent CTAs the sweeps will be launched with.
void analyse_pattern(int N, const int* rowptr, const int* colidx, int P, PatternAnalysis& out,
                     bool force_simple = false, const ClusterCaps* caps = nullptr);

// SweepProgram `lower` (the flag-synchronised tile factorisation kernel's program) on demand.
void build_tile_factor_program(const int* rowptr, const int* colidx, PatternAnalysis& an);

// Full record buffer of a program from its compact integer stream (host interpreters / CPU tests).
void materialise_records(PipeProgram& pg);
void materialise_records(FactorPipeProgram& pg);

// Sequential interpreter of a pipelined program (debug / CPU tests of the host analysis).
bool interpret_pipe_program(const PipeProgram& pg, bool upper, const double* rhs_perm, double* work,
                            double* hand_off, double* out, double w, int scale);

// Sequential interpreter of the pipelined factorisation program (debug / CPU tests): vals and lu
// in BCRS layout; lu must hold a copy of vals on entry (U blocks are not touched).  Returns the
// first singular row or -1; -2 on a deadlock.
int interpret_factor_program(const FactorPipeProgram& pg, const double* vals, double* lu);

// Row-partitioned system: split a rank's local rows (global column ids) into the local
// operator pattern (own columns first, then ghost columns in ascending global order, grouped
// by owner) and the diagonal block the rank's ILU0 is built on (block-Jacobi).
struct LocalPartition {
    int N_local = 0, n_ghost = 0;
    std::vector<int> colidx_full;            // [nnzb_local] local ids; ghosts are N_local + g
    std::vector<long long> ghost_global;     // [n_ghost] ascending
    std::vector<int> recv_cnt, recv_off;     // [world] ghosts owned by each rank (contiguous)
    std::vector<int> rowptr_diag, colidx_diag;   // diagonal block (ascending columns)
    std::vector<int> lu_src;                 // [nnzb_diag] slot in the full local pattern
};
void partition_local_rows(int N_local, const int* rowptr, const long long* colidx_global,
                          const long long* row_offsets, int world, int rank, LocalPartition& out);

// Union pattern of the pressure-derivative CSC blocks -> row-major ascending
// (formInterleavedSystem, ...Interleaved.cpp:118-155).
struct CscView { const int* colptr; const int* rowidx; };
void union_pattern_from_csc(int N, const CscView* blocks, int nblocks,
                            std::vector<int>& rowptr, std::vector<int>& colidx);

}  // namespace opmgpu
