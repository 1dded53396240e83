// See colprog.hpp.  Plain C++17, no CUDA: runs once per sparsity pattern.
#include "colprog.hpp"
#include "analysis.hpp"      // exp_env

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>

namespace opmgpu {

bool is_exact_stencil(int N, const int* rowptr, const int* colidx, int nx, int ny, int nz)
{
    if (nx < 1 || ny < 1 || nz < 1 || (long long)nx * ny * nz != N) return false;
    const long long plane = (long long)nx * ny;
    for (int r = 0; r < N; ++r) {
        const int i = r % nx, j = (int)((r / nx) % ny), k = (int)(r / plane);
        int p = rowptr[r];
        const int e = rowptr[r + 1];
        auto take = [&](long long c) { return p < e && colidx[p] == c ? (++p, true) : false; };
        if (k > 0 && !take(r - plane)) return false;
        if (j > 0 && !take(r - nx)) return false;
        if (i > 0 && !take(r - 1)) return false;
        if (!take(r)) return false;
        if (i < nx - 1 && !take(r + 1)) return false;
        if (j < ny - 1 && !take(r + nx)) return false;
        if (k < nz - 1 && !take(r + plane)) return false;
        if (p != e) return false;
    }
    return true;
}

int col_stage_count(const ColGeom& g, bool upper, size_t smem_limit)
{
    const size_t fixed = col_smem_fixed(g);
    if (smem_limit <= fixed) return 0;
    const size_t per = (size_t)g.W * col_stage_bytes(upper);
    return (int)std::min<size_t>((smem_limit - fixed) / per, kColMaxStages);
}

namespace {

struct Shape { int pw, ph, ta, tb; };

// Cost model (microseconds) of one sweep: the record stream at HBM speed (padded lane-steps),
// against the dependency chain -- one step per level, a shared-memory hand-over per patch
// boundary, an L2 round trip per tile boundary -- and the imbalance of dealing tiles to P CTAs.
// Constants from tools/ubench/colstep.cu, xwarp*.cu on B200 (profiles/r02a_ubench_*.txt).
double shape_cost(int nx, int ny, int nz, int P, size_t smem_limit, const Shape& s, ColGeom& g)
{
    g.nx = nx; g.ny = ny; g.nz = nz; g.pw = s.pw; g.ph = s.ph; g.ta = s.ta; g.tb = s.tb;
    g.npa = (nx + s.pw - 1) / s.pw; g.npb = (ny + s.ph - 1) / s.ph;
    g.nta = (g.npa + s.ta - 1) / s.ta; g.ntb = (g.npb + s.tb - 1) / s.tb;
    g.W = s.ta * s.tb; g.T = nz + s.pw + s.ph - 2; g.NE = s.tb * s.ph + s.ta * s.pw;
    if (g.W > kColMaxWarps || g.NE > kColMaxEdges) return 1e300;
    if (col_stage_count(g, true, smem_limit) < 3) return 1e300;
    const double ntiles = (double)g.nta * g.ntb;
    const double rounds = std::ceil(ntiles / P);
    const double lane_steps = (double)g.npa * g.npb * 32.0 * g.T;
    const double bytes = lane_steps * (kColNCU * 8 + 24 + 24);
    const double active = std::min(ntiles, (double)P);
    // a CTA streams ~200 GB/s at most (tools/ubench/smbw.cu); HBM 6.5 TB/s
    const double stream_us = bytes / std::min(6.5e6, active * 0.2e6) * (rounds * std::min(ntiles, (double)P) / ntiles);
    const double levels = (double)g.npa * s.pw + (double)g.npb * s.ph + nz;
    const double chain_us = levels * 0.15 + (g.nta + g.ntb - 2) * 0.9 + ((g.npa - g.nta) + (g.npb - g.ntb)) * 0.12;
    return std::max(stream_us, chain_us) + 0.35 * std::min(stream_us, chain_us);
}

}  // namespace

void build_col_program(int N, const int* rowptr, const int* colidx, const std::vector<int>& diag,
                       int nx, int ny, int nz, int P, size_t smem_limit, ColProgram& out)
{
    out = ColProgram();
    if (P < 1 || !is_exact_stencil(N, rowptr, colidx, nx, ny, nz)) return;
    static const int patches[][2] = {{8, 4}, {4, 8}, {16, 2}, {2, 16}, {6, 5}, {5, 6}, {10, 3}, {3, 10}, {7, 4}, {4, 7},
                                     {32, 1}, {1, 32}, {5, 5}, {4, 4}, {3, 3}, {2, 2}, {1, 1}};
    ColGeom best_g = {};
    double best = 1e300;
    Shape forced = {0, 0, 0, 0};
    if (const char* e = exp_env("OPMGPU_COL_SHAPE")) {          // experiments: "PWxPH/TAxTB"
        if (std::sscanf(e, "%dx%d/%dx%d", &forced.pw, &forced.ph, &forced.ta, &forced.tb) != 4) forced = {0, 0, 0, 0};
        if (forced.pw < 1 || forced.ph < 1 || forced.pw * forced.ph > 32 || forced.ta < 1 || forced.tb < 1) forced = {0, 0, 0, 0};
    }
    if (forced.pw) {
        ColGeom g;
        best = shape_cost(nx, ny, nz, P, smem_limit, forced, g);
        best_g = g;
    } else {
        for (const auto& pt : patches)
            for (int ta = 1; ta <= kColMaxWarps; ++ta)
                for (int tb = 1; ta * tb <= kColMaxWarps; ++tb) {
                    ColGeom g;
                    const double c = shape_cost(nx, ny, nz, P, smem_limit, Shape{pt[0], pt[1], ta, tb}, g);
                    if (c < best) { best = c; best_g = g; }
                }
    }
    if (!(best < 1e299)) return;
    const ColGeom g = best_g;
    const long long ntiles = (long long)g.nta * g.ntb;
    const long long nperm = ntiles * g.W * g.T * 32;
    if (nperm >= (1LL << 31) || ntiles * g.NE * g.nz >= (1LL << 31)) return;       // int positions on the device
    out.g = g;
    out.P = (int)std::min<long long>(P, ntiles);
    out.nperm = nperm;
    out.next = ntiles * g.NE * g.nz;
    out.pad_factor = (double)nperm / (double)N;
    if (std::getenv("OPMGPU_DEBUG"))
        std::fprintf(stderr, "[opmgpu] column sweeps: patch %d x %d, tile %d x %d patches, %d x %d tiles on %d CTAs, %d steps per patch, padding %.3f, model %.1f us\n",
                     g.pw, g.ph, g.ta, g.tb, g.nta, g.ntb, out.P, g.T, out.pad_factor, best);
    // tiles in wavefront order, dealt round-robin
    std::vector<int> order((size_t)ntiles);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return a % g.nta + a / g.nta < b % g.nta + b / g.nta; });
    out.cta_tile_ptr.assign(out.P + 1, 0);
    for (long long r = 0; r < ntiles; ++r) out.cta_tile_ptr[r % out.P + 1]++;
    for (int c = 0; c < out.P; ++c) out.cta_tile_ptr[c + 1] += out.cta_tile_ptr[c];
    out.cta_tilesL.resize((size_t)ntiles); out.cta_tilesU.resize((size_t)ntiles);
    {
        std::vector<int> fill(out.cta_tile_ptr.begin(), out.cta_tile_ptr.end() - 1);
        for (long long r = 0; r < ntiles; ++r) out.cta_tilesL[fill[r % out.P]++] = order[r];
        fill.assign(out.cta_tile_ptr.begin(), out.cta_tile_ptr.end() - 1);
        for (long long r = 0; r < ntiles; ++r) out.cta_tilesU[fill[r % out.P]++] = order[ntiles - 1 - r];
    }
    // lane-steps and value maps
    out.perm_rowL.assign((size_t)nperm, -1); out.perm_rowU.assign((size_t)nperm, -1);
    const size_t nlow = ((size_t)rowptr[N] - N) / 2;
    out.valL_src.resize(nlow); out.valL_dst.resize(nlow);
    out.valU_src.resize(nlow + N); out.valU_dst.resize(nlow + N);
    // every patch writes a disjoint, precomputable range: count first (prefix sums), then fill in parallel
    const long long npatch = ntiles * g.W;
    std::vector<size_t> offL((size_t)npatch + 1, 0), offU((size_t)npatch + 1, 0);
    auto patch_cols = [&](long long pid, int& i0, int& j0, int& ni, int& nj) {
        const int tile = (int)(pid / g.W), warp = (int)(pid % g.W);
        const int pa = (tile % g.nta) * g.ta + warp % g.ta, pb = (tile / g.nta) * g.tb + warp / g.ta;
        i0 = pa * g.pw; j0 = pb * g.ph;
        ni = pa < g.npa ? std::min(g.pw, nx - i0) : 0;
        nj = pb < g.npb ? std::min(g.ph, ny - j0) : 0;
        if (ni < 0) ni = 0;
        if (nj < 0) nj = 0;
    };
    for (long long pid = 0; pid < npatch; ++pid) {
        int i0, j0, ni, nj;
        patch_cols(pid, i0, j0, ni, nj);
        size_t nl = 0, nu = 0;
        for (int b = 0; b < nj; ++b)
            for (int a = 0; a < ni; ++a) {
                const int i = i0 + a, j = j0 + b;
                // lower blocks of the column's cells: k-1 (nz-1 of them), j-1, i-1 (nz each when present)
                nl += (size_t)(nz - 1) + (j > 0 ? nz : 0) + (i > 0 ? nz : 0);
                nu += (size_t)(nz - 1) + (j < ny - 1 ? nz : 0) + (i < nx - 1 ? nz : 0) + nz;
            }
        offL[pid + 1] = offL[pid] + nl; offU[pid + 1] = offU[pid] + nu;
    }
    const long long plane = (long long)nx * ny;
    auto fill_patch = [&](long long pid) {
        int i0, j0, ni, nj;
        patch_cols(pid, i0, j0, ni, nj);
        if (ni == 0 || nj == 0) return;
        const int tile = (int)(pid / g.W), warp = (int)(pid % g.W);
        size_t pl = offL[pid], pu = offU[pid];
        // k outermost, lanes innermost: consecutive entries of the maps go to consecutive lanes
        for (int k = 0; k < nz; ++k)
            for (int b = 0; b < nj; ++b)
                for (int a = 0; a < ni; ++a) {
                    const int i = i0 + a, j = j0 + b, lane = a + g.pw * b;
                    const int r = (int)(i + (long long)nx * j + plane * k);
                    const int tL = col_lane_delay(g, false, a, b) + k;
                    const int tU = col_lane_delay(g, true, a, b) + (nz - 1 - k);
                    const size_t lsL = col_lane_step(g, tile, warp, tL, lane), lsU = col_lane_step(g, tile, warp, tU, lane);
                    out.perm_rowL[lsL] = r; out.perm_rowU[lsU] = r;
                    int p = rowptr[r];
                    if (k > 0) { out.valL_src[pl] = p++; out.valL_dst[pl++] = (unsigned long long)lsL << 2 | 0; }
                    if (j > 0) { out.valL_src[pl] = p++; out.valL_dst[pl++] = (unsigned long long)lsL << 2 | 1; }
                    if (i > 0) { out.valL_src[pl] = p++; out.valL_dst[pl++] = (unsigned long long)lsL << 2 | 2; }
                    out.valU_src[pu] = diag[r]; out.valU_dst[pu++] = (unsigned long long)lsU << 2 | 3;
                    p = rowptr[r + 1] - 1;
                    if (k < nz - 1) { out.valU_src[pu] = p--; out.valU_dst[pu++] = (unsigned long long)lsU << 2 | 0; }
                    if (j < ny - 1) { out.valU_src[pu] = p--; out.valU_dst[pu++] = (unsigned long long)lsU << 2 | 1; }
                    if (i < nx - 1) { out.valU_src[pu] = p--; out.valU_dst[pu++] = (unsigned long long)lsU << 2 | 2; }
                }
    };
    unsigned nt = std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
    if (N < 200000) nt = 1;
    if (nt == 1) {
        for (long long pid = 0; pid < npatch; ++pid) fill_patch(pid);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t]() { for (long long pid = t; pid < npatch; pid += nt) fill_patch(pid); });
        for (auto& x : th) x.join();
    }
    out.valid = offL[npatch] == nlow && offU[npatch] == nlow + (size_t)N;
}

void interpret_col_program(const ColProgram& pg, const double* recL, const double* recU,
                           const double* d, double* v, double w, int scale)
{
    const ColGeom& g = pg.g;
    const long long plane = (long long)g.nx * g.ny;
    const size_t N = (size_t)plane * g.nz;
    std::vector<double> y(N * 3, 0.0), hand((size_t)pg.nperm * 3, 0.0), rhs((size_t)pg.nperm * 3, 0.0);
    for (long long q = 0; q < pg.nperm; ++q)
        if (pg.perm_rowL[q] >= 0)
            for (int c = 0; c < 3; ++c) rhs[(size_t)q * 3 + c] = d[(size_t)pg.perm_rowL[q] * 3 + c];
    const double zero[3] = {0.0, 0.0, 0.0};
    for (int upper = 0; upper < 2; ++upper) {
        const double* rec = upper ? recU : recL;
        const int NC = upper ? kColNCU : kColNCL;
        const std::vector<double>& in = upper ? hand : rhs;
        // natural order (descending for the upper sweep) respects every dependency
        for (size_t n = 0; n < N; ++n) {
            const size_t r = upper ? N - 1 - n : n;
            const int i = (int)(r % g.nx), j = (int)((r / g.nx) % g.ny), k = (int)(r / plane);
            const int pa = i / g.pw, pb = j / g.ph, li = i % g.pw, lj = j % g.ph;
            const int tile = pa / g.ta + g.nta * (pb / g.tb), warp = pa % g.ta + g.ta * (pb % g.tb), lane = li + g.pw * lj;
            const int kl = upper ? g.nz - 1 - k : k;
            const int t = col_lane_delay(g, upper != 0, li, lj) + kl;
            const size_t ls = col_lane_step(g, tile, warp, t, lane);
            const int s = upper ? -1 : 1;
            const double* dep[3];
            dep[0] = kl > 0 ? &y[(r - s * plane) * 3] : zero;
            dep[1] = (upper ? j < g.ny - 1 : j > 0) ? &y[(r - s * (long long)g.nx) * 3] : zero;
            dep[2] = (upper ? i < g.nx - 1 : i > 0) ? &y[(r - s) * 3] : zero;
            double acc[3];
            for (int c = 0; c < 3; ++c) {
                double a = in[ls * 3 + c];
                for (int q = 0; q < 9; ++q) a = std::fma(-rec[col_rec_index(ls, NC, c * 9 + q)], dep[q / 3][q % 3], a);
                acc[c] = a;
            }
            if (upper) {
                double o[3];
                for (int c = 0; c < 3; ++c) {
                    double tt = 0.0;
                    for (int e = 0; e < 3; ++e) tt = std::fma(rec[col_rec_index(ls, NC, 27 + c * 3 + e)], acc[e], tt);
                    o[c] = tt;
                }
                for (int c = 0; c < 3; ++c) { y[r * 3 + c] = o[c]; v[r * 3 + c] = scale ? o[c] * w : o[c]; }
            } else {
                for (int c = 0; c < 3; ++c) y[r * 3 + c] = acc[c];
                const size_t lsU = col_lane_step(g, tile, warp, g.T - 1 - t, lane);
                for (int c = 0; c < 3; ++c) hand[lsU * 3 + c] = acc[c];
            }
        }
    }
}

}  // namespace opmgpu

// Debug entry (CPU tests of the column program's layout, no GPU): scatters BCRS factors into the
// record streams through the value maps and runs the sequential interpreter.  info[0..7] = pw, ph,
// ta, tb, number of tiles, CTAs, steps per patch, stages of the upper sweep.  Returns 0, -2 when
// the pattern has no column program.
extern "C" int opmgpu_debug_host_col_apply(int N, const int* rowptr, const int* colidx, int nx, int ny, int nz,
                                           const double* lu, int P, double w, const double* d, double* v, int* info)
{
    using namespace opmgpu;
    std::vector<int> diag(N, -1);
    for (int i = 0; i < N; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
            if (colidx[k] == i) diag[i] = k;
    ColProgram pg;
    build_col_program(N, rowptr, colidx, diag, nx, ny, nz, P, 227 * 1024, pg);
    if (!pg.valid) return -2;
    std::vector<double> recL((size_t)pg.nperm * kColNCL, 0.0), recU((size_t)pg.nperm * kColNCU, 0.0);
    for (int upper = 0; upper < 2; ++upper) {
        const std::vector<int>& src = upper ? pg.valU_src : pg.valL_src;
        const std::vector<unsigned long long>& dst = upper ? pg.valU_dst : pg.valL_dst;
        std::vector<double>& rec = upper ? recU : recL;
        const int NC = upper ? kColNCU : kColNCL;
        for (size_t b = 0; b < src.size(); ++b) {
            const size_t ls = (size_t)(dst[b] >> 2);
            const int kb = (int)(dst[b] & 3);
            for (int c = 0; c < 3; ++c)
                for (int e = 0; e < 3; ++e)
                    rec[col_rec_index(ls, NC, kb == 3 ? 27 + c * 3 + e : c * 9 + kb * 3 + e)] = lu[(size_t)src[b] * 9 + c * 3 + e];
        }
    }
    // the hand-over position formula needs perm_rowU consistent with it: checked here
    for (long long q = 0; q < pg.nperm; ++q) {
        const int r = pg.perm_rowL[q];
        if (r < 0) continue;
        const long long t = (q / 32) % pg.g.T, base = q / 32 - t;
        if (pg.perm_rowU[(size_t)((base + (pg.g.T - 1 - t)) * 32 + q % 32)] != r) return -5;
    }
    const int scale = std::fabs(w - 1.0) > 1e-15 ? 1 : 0;
    interpret_col_program(pg, recL.data(), recU.data(), d, v, w, scale);
    if (info) {
        info[0] = pg.g.pw; info[1] = pg.g.ph; info[2] = pg.g.ta; info[3] = pg.g.tb;
        info[4] = pg.g.nta * pg.g.ntb; info[5] = pg.P; info[6] = pg.g.T; info[7] = col_stage_count(pg.g, true, 227 * 1024);
    }
    return 0;
}
