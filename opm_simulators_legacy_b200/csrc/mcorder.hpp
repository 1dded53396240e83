// Multicolour ordering of the block ILU0 (the "level-set (or multicolour) scheduling" option of the
// north star; a FLAGGED VARIANT: it is the reference's ILU0 -- Dune::bilu0_decomposition and
// Opm::ParallelOverlappingILU0::apply -- of the symmetric permutation P A P^T, so the preconditioner,
// and with it the iteration counts, differ from the natural-order solve.  Never reported as parity
// with the reference's solve; parity is with the oracle run on the permuted system).
//
// Colouring: greedy in natural row order over the symmetrised block graph (row i takes the smallest
// colour none of its already coloured neighbours has); rows are then ordered by (colour, natural
// index).  A 7-point Cartesian stencil gets the red-black ordering ((i+j+k) odd / even).
#pragma once
#include <vector>

namespace opmgpu {

struct McOrder {
    // line mode (k-lines): colours over the (i,j) COLUMNS of a Cartesian grid, natural order along k inside
    // a column; permuted index = base[c] + k * ncols[c] + rank of the column among those of its colour
    bool lines = false;
    int nx = 0, nz = 0;
    int ncols[2] = {0, 0}, base[2] = {0, 0};
    int ncolours = 0;
    std::vector<int> colour;        // [N] colour of natural row i
    std::vector<int> p2n, n2p;      // permuted position <-> natural row
    std::vector<int> colour_ptr;    // [ncolours+1] permuted row ranges of the colours
};

struct McProgram {
    McOrder ord;
    // permuted full pattern, ascending permuted columns (what bilu0_decomposition walks)
    std::vector<int> prowptr, pcol, pdiag;
    std::vector<int> psrc;          // [nnzb] natural BCRS slot of permuted entry
    std::vector<int> ppos;          // [nnzb] block index into the unified factor array [L | Dinv | U]
    // the sweeps' operands: strictly lower part (ascending permuted columns) and strictly upper part
    // in DESCENDING column order (the order ParallelOverlappingILU0::apply visits them)
    std::vector<int> Lrowptr, Lcol, Urowptr, Ucol;
    long long offD = 0, offU = 0, total_blocks = 0;     // block offsets of Dinv and U in the unified array
    // factorisation: for lower entry l (= its block index in the L region) the pairs (block of A_jk, block
    // of A_ik), k > j present in both rows, that A_ik -= L_ij * A_jk touches, in ascending k
    std::vector<int> pair_ptr, pair_jk, pair_ik;
    // level sets of the permuted lower triangle (factorisation: one launch per level)
    std::vector<int> lvl_ptr, lvl_rows;
};

// colours + permutation only (pure host code, also exported through the C ABI)
void multicolour_order(int N, const int* rowptr, const int* colidx, McOrder& out);
// k-line ordering of an nx x ny x nz grid in natural numbering: red-black over the columns
void line_order(int nx, int ny, int nz, McOrder& out);
// everything the device path needs.  lines: the k-line ordering; returns false (out unusable) when the
// pattern is not a Cartesian stencil whose only same-colour couplings are the vertical ones
bool build_mc_program(int N, const int* rowptr, const int* colidx, McProgram& out, bool lines = false);

}  // namespace opmgpu
