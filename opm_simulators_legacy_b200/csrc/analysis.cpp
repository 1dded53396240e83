// See analysis.hpp.  Plain C++17, no CUDA: everything here runs once per sparsity pattern.
#include "analysis.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <numeric>

namespace opmgpu {

namespace {

// Recognise a Cartesian 7-/5-/3-point stencil in natural ordering from the dominant
// positive column offsets {1, nx, nx*ny}.  Extra couplings (Schur complements of
// multi-perforation wells, NewtonIterationUtilities.cpp:98-115) are tolerated: the
// partition is only a performance heuristic, correctness never depends on it.
void infer_grid(int N, const int* rowptr, const int* colidx, int& nx, int& ny, int& nz)
{
    nx = ny = nz = 0;
    if (N < 8) return;
    std::map<int, long long> hist;
    long long total = 0;
    const int stride = std::max(1, N / 200000);          // sample rows; the stencil is uniform
    for (int i = 0; i < N; i += stride)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int off = colidx[k] - i;
            if (off > 0) { ++hist[off]; ++total; }
        }
    if (total == 0) return;
    std::vector<std::pair<long long, int>> top;
    for (auto& kv : hist) top.push_back({kv.second, kv.first});
    std::sort(top.rbegin(), top.rend());
    std::vector<int> offs;
    long long covered = 0;
    for (size_t q = 0; q < top.size() && q < 3; ++q) {
        if (top[q].first * 20 < top[0].first) break;      // rare offset: not part of the stencil
        offs.push_back(top[q].second);
        covered += top[q].first;
    }
    if (covered * 10 < total * 9) return;
    std::sort(offs.begin(), offs.end());
    if (offs[0] != 1) return;
    int gx = N, gy = 1, gz = 1;
    if (offs.size() >= 2) {
        gx = offs[1];
        if (gx < 2 || N % gx) return;
        gy = N / gx;
        if (offs.size() == 3) {
            const int plane = offs[2];
            if (plane % gx || N % plane) return;
            gy = plane / gx;
            gz = N / plane;
        }
    }
    nx = gx; ny = gy; nz = gz;
}

void choose_tiling(int nx, int ny, int P, int& pa, int& pb)
{
    long long best = -1;
    double best_sq = 0;
    pa = pb = 1;
    for (int a = 1; a <= std::min(nx, P); ++a) {
        const int b = std::min(ny, P / a);
        if (b < 1) break;
        const long long cnt = (long long)a * b;
        const double sq = std::fabs(std::log((double(nx) / a) / (double(ny) / b)));
        if (cnt > best || (cnt == best && sq < best_sq)) { best = cnt; best_sq = sq; pa = a; pb = b; }
    }
}

void build_program(int N, const int* rowptr, const int* colidx, const std::vector<int>& level,
                   int nlevels, const std::vector<int>& owner, int P, bool lower,
                   SweepProgram& prog)
{
    prog.P = P;
    prog.nlevels = nlevels;
    // rows of each CTA in ascending (level, row) order: counting sort by (owner, level)
    std::vector<int> order(N);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (owner[a] != owner[b]) return owner[a] < owner[b];
        return level[a] < level[b];
    });
    prog.prow = order;
    prog.cta_step_ptr.assign(P + 1, 0);
    prog.step_row_ptr.clear();
    prog.pblk_ptr.assign(N + 1, 0);
    prog.max_step_rows = 0;
    int cur_cta = -1, cur_level = -1;
    for (int q = 0; q < N; ++q) {
        const int r = order[q];
        if (owner[r] != cur_cta || level[r] != cur_level) {
            // close steps of skipped CTAs
            while (cur_cta < owner[r]) { ++cur_cta; prog.cta_step_ptr[cur_cta] = (int)prog.step_row_ptr.size(); }
            cur_level = level[r];
            prog.step_row_ptr.push_back(q);
        }
    }
    while (cur_cta < P) { ++cur_cta; prog.cta_step_ptr[cur_cta] = (int)prog.step_row_ptr.size(); }
    prog.step_row_ptr.push_back(N);
    for (size_t s = 0; s + 1 < prog.step_row_ptr.size(); ++s)
        prog.max_step_rows = std::max(prog.max_step_rows, prog.step_row_ptr[s + 1] - prog.step_row_ptr[s]);

    std::vector<unsigned char> pub_row(N, 0);
    size_t nblk = 0;
    for (int q = 0; q < N; ++q) {
        const int r = order[q];
        int cnt = 0;
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k)
            cnt += lower ? (colidx[k] < r) : (colidx[k] > r);
        prog.pblk_ptr[q + 1] = prog.pblk_ptr[q] + cnt;
        nblk += cnt;
    }
    prog.pcol.resize(nblk);
    prog.psrc.resize(nblk);
    for (int q = 0; q < N; ++q) {
        const int r = order[q];
        int w = prog.pblk_ptr[q];
        auto emit = [&](int k) {
            const int j = colidx[k];
            int c = j;
            if (owner[j] != owner[r]) { c |= kExtBit; pub_row[j] = 1; }
            prog.pcol[w] = c;
            prog.psrc[w] = k;
            ++w;
        };
        if (lower) {
            for (int k = rowptr[r]; k < rowptr[r + 1] && colidx[k] < r; ++k) emit(k);
        } else {
            for (int k = rowptr[r + 1] - 1; k >= rowptr[r] && colidx[k] > r; --k) emit(k);
        }
    }
    prog.publish.resize(N);
    for (int q = 0; q < N; ++q) prog.publish[q] = pub_row[order[q]];
}

}  // namespace

void analyse_pattern(int N, const int* rowptr, const int* colidx, int P, PatternAnalysis& out)
{
    out = PatternAnalysis();
    out.N = N;
    out.nnzb = rowptr[N];
    out.diag.assign(N, -1);
    for (int i = 0; i < N; ++i) {
        const int* b = colidx + rowptr[i];
        const int* e = colidx + rowptr[i + 1];
        const int* it = std::lower_bound(b, e, i);
        if (it == e || *it != i) { out.missing_diag_row = i; return; }
        out.diag[i] = (int)(it - colidx);
    }
    // dependency levels
    std::vector<int> lvlL(N, 0), lvlU(N, 0);
    int nL = 0, nU = 0;
    for (int i = 0; i < N; ++i) {
        int l = 0;
        for (int k = rowptr[i]; k < out.diag[i]; ++k) l = std::max(l, lvlL[colidx[k]] + 1);
        lvlL[i] = l;
        nL = std::max(nL, l + 1);
    }
    for (int i = N - 1; i >= 0; --i) {
        int l = 0;
        for (int k = out.diag[i] + 1; k < rowptr[i + 1]; ++k) l = std::max(l, lvlU[colidx[k]] + 1);
        lvlU[i] = l;
        nU = std::max(nU, l + 1);
    }
    if (N == 0) { nL = nU = 0; }
    // level sets of L in natural order inside a level (factorisation kernel)
    out.lvl_ptr.assign(nL + 1, 0);
    for (int i = 0; i < N; ++i) out.lvl_ptr[lvlL[i] + 1]++;
    for (int l = 0; l < nL; ++l) out.lvl_ptr[l + 1] += out.lvl_ptr[l];
    out.lvl_rows.resize(N);
    {
        std::vector<int> fill(out.lvl_ptr.begin(), out.lvl_ptr.end() - (nL ? 1 : 0));
        for (int i = 0; i < N; ++i) out.lvl_rows[fill[lvlL[i]]++] = i;
    }
    // partition
    std::vector<int> owner(N, 0);
    int nx, ny, nz;
    infer_grid(N, rowptr, colidx, nx, ny, nz);
    out.grid_nx = nx; out.grid_ny = ny; out.grid_nz = nz;
    if (P < 1) P = 1;
    if (nx > 0 && (long long)nx * ny >= 4) {
        int pa, pb;
        choose_tiling(nx, ny, P, pa, pb);
        for (int r = 0; r < N; ++r) {
            const int i = r % nx, j = (r / nx) % ny;
            const int a = (int)((long long)i * pa / nx), b = (int)((long long)j * pb / ny);
            owner[r] = a + pa * b;
        }
    } else {
        // generic: contiguous share of every level per CTA
        for (int l = 0; l < nL; ++l) {
            const int b = out.lvl_ptr[l], sz = out.lvl_ptr[l + 1] - b;
            for (int q = 0; q < sz; ++q) owner[out.lvl_rows[b + q]] = (int)((long long)q * P / sz);
        }
    }
    build_program(N, rowptr, colidx, lvlL, nL, owner, P, true, out.lower);
    build_program(N, rowptr, colidx, lvlU, nU, owner, P, false, out.upper);
}

void union_pattern_from_csc(int N, const CscView* blocks, int nblocks,
                            std::vector<int>& rowptr, std::vector<int>& colidx)
{
    // Visit columns ascending and append to rows: every row receives ascending columns,
    // which is the row-major conversion at ...Interleaved.cpp:137.
    std::vector<int> cnt(N + 1, 0), tmp;
    rowptr.assign(N + 1, 0);
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            for (int r = 0; r < N; ++r) rowptr[r + 1] = rowptr[r] + cnt[r];
            colidx.resize(rowptr[N]);
            for (int r = 0; r < N; ++r) cnt[r] = rowptr[r];
        }
        for (int c = 0; c < N; ++c) {
            tmp.clear();
            for (int b = 0; b < nblocks; ++b)
                for (int k = blocks[b].colptr[c]; k < blocks[b].colptr[c + 1]; ++k)
                    tmp.push_back(blocks[b].rowidx[k]);
            if (nblocks > 1) {
                std::sort(tmp.begin(), tmp.end());
                tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            }
            for (int r : tmp) {
                if (pass == 0) cnt[r]++;
                else colidx[cnt[r]++] = c;
            }
        }
    }
}

}  // namespace opmgpu
