// Column-owned sweep program for exact Cartesian 7-/5-/3-point stencils in natural ordering
// (every synthetic configuration of BASELINE.json; the general pipelined program of analysis.hpp
// stays the path for every other pattern).
//
// The natural-order ILU0 sweeps (Opm::ParallelOverlappingILU0::apply, call site
// opm/autodiff/ISTLSolver.hpp:201-211) walk a dependency wavefront of nx+ny+nz-2 levels.  Here
// a LANE owns an (i,j) column of cells and walks it along k, one cell per step:
//   * the (i,j,k-1) result never leaves the lane's registers;
//   * the (i-1,j,k) and (i,j-1,k) results come from neighbouring lanes of the same warp by
//     shuffle -- a warp owns a pw x ph PATCH of columns (pw*ph <= 32), lane = li + pw*lj, and at
//     warp step t lane (li,lj) works on k = t - li - lj, so both neighbours finished exactly one
//     step earlier;
//   * results crossing a patch edge travel through self-validating shared-memory ring entries
//     (3 doubles, all-ones = empty) when the neighbouring patch belongs to the same CTA TILE
//     (ta x tb patches, one warp each), and through self-validating push slots in L2, staged into
//     the same rings by a helper warp, when it belongs to another tile.
// The upper sweep is the same program mirrored (i,j,k) -> (nx-1-i, ny-1-j, nz-1-k) on the SAME
// patches (lane numbering stays natural), so the row a lane finishes at lower step t is the row
// it starts from at upper step T-1-t: the lower sweep hands its result over as one contiguous
// 768-byte segment per warp step.
//
// Record stream (per sweep): rec[(tile*W + warp)*T + t][q/2][lane][q%2], q < NC, doubles (a lane
// fetches its coefficients as 16-byte pairs, conflict-free);
//   q = c*9 + kb*3 + e : element [c][e] of the kb-th off-diagonal block of the lane's row in the
//                        reference's visiting order (lower: k-1, j-1, i-1; upper: k+1, j+1, i+1),
//                        zero when the block (or the whole lane-step) does not exist;
//   q = 27 + c*3 + e   : element [c][e] of the inverted diagonal block (upper sweep only).
// Right-hand sides / hand-over in program order: rhs[((tile*W + warp)*T + t)*32 + lane][3].
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#if defined(__CUDACC__)
#define OPMGPU_HD __host__ __device__
#else
#define OPMGPU_HD
#endif

namespace opmgpu {

constexpr int kColNCL = 28, kColNCU = 36;       // doubles per lane-step of a lower / upper record (even: the kernel loads pairs)
constexpr int kColMaxWarps = 6;                 // compute warps (patches) per CTA tile
constexpr int kColRing = 16;                    // ring entries per incoming edge lane (k mod R)
constexpr int kColMaxStages = 8;
constexpr int kColMaxEdges = 64;                // incoming edge lanes of a tile the helper warp can serve

struct ColGeom {
    int nx, ny, nz;
    int pw, ph;            // patch: pw x ph columns, lane = li + pw*lj
    int ta, tb;            // tile: ta x tb patches
    int npa, npb;          // patches along i, j
    int nta, ntb;          // tiles along i, j
    int W;                 // ta*tb
    int T;                 // steps of a patch: nz + pw + ph - 2
    int NE;                // incoming edge lanes of a tile: tb*ph (from the i neighbour) + ta*pw (from the j neighbour)
};

// lane-step index of (tile, warp, t, lane): rows of rhs vectors in program order, and the
// (q = 0, lane) position of a record divided by NC
OPMGPU_HD inline size_t col_lane_step(const ColGeom& g, int tile, int warp, int t, int lane)
{
    return (((size_t)tile * g.W + warp) * g.T + t) * 32 + lane;
}
// double index of coefficient q of a lane-step in a record stream with NC doubles per lane-step
OPMGPU_HD inline size_t col_rec_index(size_t lane_step, int NC, int q)
{
    return (lane_step >> 5) * (size_t)NC * 32 + (size_t)(q >> 1) * 64 + (lane_step & 31) * 2 + (q & 1);
}
// distance of a lane from the upstream corner of its patch (steps before its first cell)
OPMGPU_HD inline int col_lane_delay(const ColGeom& g, bool upper, int li, int lj)
{
    return upper ? (g.pw - 1 - li) + (g.ph - 1 - lj) : li + lj;
}

struct ColProgram {
    bool valid = false;
    ColGeom g = {};
    int P = 0;                                  // CTAs
    std::vector<int> cta_tile_ptr;              // [P+1]
    std::vector<int> cta_tilesL, cta_tilesU;    // a CTA's tiles in processing (wavefront) order
    long long nperm = 0;                        // lane-steps = ntiles*W*T*32
    std::vector<int> perm_rowL, perm_rowU;      // [nperm] natural row of a lane-step, -1 = none
    // factor values -> records: BCRS slot, and (lane_step << 2 | kb), kb = 3: inverted diagonal
    std::vector<int> valL_src, valU_src;
    std::vector<unsigned long long> valL_dst, valU_dst;
    long long next = 0;                         // push slots per sweep: ntiles*NE*nz entries of 3 doubles
    double pad_factor = 1.0;                    // streamed lane-steps / rows
};

// True when the BCRS pattern is exactly the stencil of an nx x ny x nz grid in natural ordering.
bool is_exact_stencil(int N, const int* rowptr, const int* colidx, int nx, int ny, int nz);

// Lays the program out for at most P co-resident CTAs; smem_limit = dynamic shared memory a CTA
// may use.  Leaves out.valid false when the grid does not fit the scheme.
void build_col_program(int N, const int* rowptr, const int* colidx, const std::vector<int>& diag,
                       int nx, int ny, int nz, int P, size_t smem_limit, ColProgram& out);

// shared memory of the sweep kernel (also used by the host to size the stage ring)
OPMGPU_HD inline size_t col_smem_fixed(const ColGeom& g)
{
    return 2048 + (size_t)g.W * (size_t)(g.pw + g.ph) * kColRing * 32;
}
OPMGPU_HD inline size_t col_stage_bytes(bool upper) { return 768 + (size_t)(upper ? kColNCU : kColNCL) * 256; }
int col_stage_count(const ColGeom& g, bool upper, size_t smem_limit);

// Sequential interpreter (CPU tests of the layout): executes both sweeps from the record
// streams exactly as the kernel indexes them.  d and v in natural order.
void interpret_col_program(const ColProgram& pg, const double* recL, const double* recU,
                           const double* d, double* v, double w, int scale);

}  // namespace opmgpu
