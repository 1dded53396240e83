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
#include <vector>

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
};

struct PatternAnalysis {
    int N = 0, nnzb = 0;
    std::vector<int> diag;                   // [N] BCRS slot of the diagonal block
    // natural-order level sets of the lower triangle (factorisation)
    std::vector<int> lvl_ptr, lvl_rows;
    int grid_nx = 0, grid_ny = 0, grid_nz = 0;   // inferred Cartesian structure (0 = none)
    SweepProgram lower, upper;
    int missing_diag_row = -1;
};

// P = number of persistent CTAs the sweeps will be launched with.
void analyse_pattern(int N, const int* rowptr, const int* colidx, int P, PatternAnalysis& out);

// Union pattern of the pressure-derivative CSC blocks -> row-major ascending
// (formInterleavedSystem, ...Interleaved.cpp:118-155).
struct CscView { const int* colptr; const int* rowidx; };
void union_pattern_from_csc(int N, const CscView* blocks, int nblocks,
                            std::vector<int>& rowptr, std::vector<int>& colidx);

}  // namespace opmgpu
