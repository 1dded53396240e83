// See analysis.hpp.  Plain C++17, no CUDA: everything here runs once per sparsity pattern.
#include "analysis.hpp"

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <future>
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

// Cost model of one sweep (measured on B200, profiles/r01_summary.md): the critical path runs
// through nx+ny+nz-2 levels and crosses pa+pb-2 tile boundaries -- a store -> poll round trip
// through L2 between clusters, a distributed-shared-memory store inside one; every CTA also has
// to stream its own tile from HBM.  A tile level wider than one pass (kLeanStepRows rows) costs
// extra passes.  caps.max_ctas[log2 cs] = CTAs that can be co-resident at cluster size cs.
void choose_tiling(int nx, int ny, int nz, const ClusterCaps& caps, int& pa, int& pb, int& ca, int& cb)
{
    // (constants reproduce the measured ranking of tilings at 100^3, profiles/r01_summary.md)
    constexpr double kCrossUs = 1.7, kCrossClusterUs = 0.9, kLevelUs = 0.37, kRowUs = 0.004;
    static const int shapes[][2] = {{1, 1}, {1, 2}, {2, 1}, {2, 2}, {1, 4}, {4, 1}, {2, 4}, {4, 2}};
    double best = 1e300;
    pa = pb = ca = cb = 1;
    for (const auto& sh : shapes) {
        const int sa = sh[0], sb = sh[1], cs = sa * sb;
        const int lg = cs == 1 ? 0 : (cs == 2 ? 1 : (cs == 4 ? 2 : 3));
        const int P = caps.max_ctas[lg];
        if (P < cs) continue;
        for (int a = sa; a <= std::min(nx, P); a += sa)
            for (int b = sb; b <= std::min(ny, P / a); b += sb) {
                const int wa = (nx + a - 1) / a, wb = (ny + b - 1) / b;
                const int passes = (wa * wb + kLeanStepRows - 1) / kLeanStepRows;
                if (cs > 1 && passes > 1) continue;            // clusters only with lean programs
                const int cross_a = a - 1, cross_b = b - 1;
                const int glob = (a / sa - 1) + (b / sb - 1), intra = cross_a + cross_b - glob;
                double t = glob * kCrossUs + intra * kCrossClusterUs + (double)(nx + ny + nz - 2) * kLevelUs * passes +
                           (double)wa * wb * nz * kRowUs;
                if (passes > 1) t *= 1.25;          // the general kernel variant is slower than the lean one
                if (t < best) { best = t; pa = a; pb = b; ca = sa; cb = sb; }
            }
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
    if (lower) {
        // update lists of Dune::bilu0_decomposition: merge row j (right of its diagonal) with row i
        prog.pair_ptr.assign(nblk + 1, 0);
        prog.pair_jk.clear(); prog.pair_ik.clear();
        for (size_t b = 0; b < nblk; ++b) {
            const int ij = prog.psrc[b], j = colidx[ij];
            int i = 0;
            {   // row of slot ij
                int lo = 0, hi = N;
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (rowptr[mid] <= ij) lo = mid; else hi = mid; }
                i = lo;
            }
            int jk = rowptr[j], ik = ij + 1;
            while (jk < rowptr[j + 1] && colidx[jk] <= j) ++jk;
            const int jend = rowptr[j + 1], iend = rowptr[i + 1];
            while (ik < iend && jk < jend) {
                if (colidx[ik] == colidx[jk]) { prog.pair_jk.push_back(jk); prog.pair_ik.push_back(ik); ++ik; ++jk; }
                else if (colidx[ik] < colidx[jk]) ++ik;
                else ++jk;
            }
            prog.pair_ptr[b + 1] = (int)prog.pair_jk.size();
        }
        prog.frow.assign((size_t)N * 4, 0);
        prog.fent.assign(nblk * 8, 0);
        for (int q = 0; q < N; ++q) {
            const int r = order[q];
            int dslot = rowptr[r];
            while (colidx[dslot] != r) ++dslot;
            const int b0 = prog.pblk_ptr[q], b1 = prog.pblk_ptr[q + 1];
            bool simple = b1 - b0 <= 3;
            for (int b = b0; b < b1; ++b) {
                const int j = prog.pcol[b] & ~kExtBit;
                int jd = rowptr[j];
                while (colidx[jd] != j) ++jd;
                const int np = prog.pair_ptr[b + 1] - prog.pair_ptr[b];
                int* e = &prog.fent[(size_t)b * 8];
                e[0] = prog.psrc[b]; e[1] = jd; e[2] = prog.pcol[b]; e[3] = np;
                e[4] = np ? prog.pair_jk[prog.pair_ptr[b]] : -1;
                e[5] = np ? prog.pair_ik[prog.pair_ptr[b]] : -1;
                e[6] = prog.pair_ptr[b]; e[7] = 0;
                if (np > 1 || (np == 1 && e[5] != dslot)) simple = false;
            }
            int* fr = &prog.frow[(size_t)q * 4];
            fr[0] = r; fr[1] = dslot; fr[2] = b0; fr[3] = (b1 - b0) | (simple ? kFactorSimple : 0);
        }
        // push slots for the pivots: consumer entry (i <- j) gets one when i and j are both simple
        std::vector<int> qof(N);
        for (int q = 0; q < N; ++q) qof[order[q]] = q;
        std::vector<std::vector<int>> pushes(N);
        prog.needs_flag.assign(N, 0);
        int nslots = 0;
        for (int q = 0; q < N; ++q) {
            const bool isimple = (prog.frow[(size_t)q * 4 + 3] & kFactorSimple) != 0;
            for (int b = prog.pblk_ptr[q]; b < prog.pblk_ptr[q + 1]; ++b) {
                int* e = &prog.fent[(size_t)b * 8];
                e[7] = -1;
                if (!(prog.pcol[b] & kExtBit)) continue;
                const int j = prog.pcol[b] & ~kExtBit, qj = qof[j];
                const bool jsimple = (prog.frow[(size_t)qj * 4 + 3] & kFactorSimple) != 0;
                if (isimple && jsimple) { e[7] = nslots; pushes[qj].push_back(nslots); ++nslots; }
                else prog.needs_flag[qj] = 1;
            }
        }
        prog.n_fslots = nslots;
        prog.fpush_ptr.assign(N + 1, 0);
        for (int q = 0; q < N; ++q) prog.fpush_ptr[q + 1] = prog.fpush_ptr[q] + (int)pushes[q].size();
        prog.fpush_slot.clear();
        for (int q = 0; q < N; ++q) prog.fpush_slot.insert(prog.fpush_slot.end(), pushes[q].begin(), pushes[q].end());
    }
}


// ---- pipelined program ---------------------------------------------------------------------
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct RecLayout {
    size_t cf, dinv, rowints, lists, tail_end, xpush_end, tail_dep, xpush_slot, tail_vals, total;
};
RecLayout rec_layout(int n, int ntail, int npushx, bool upper, bool has_lists)
{
    RecLayout L;
    size_t o = 32;
    L.cf = o; o += (size_t)3 * n * 72;
    L.dinv = o; if (upper) o += (size_t)3 * n * 24;
    o = align_up(o, 16);
    L.rowints = o; o += (size_t)n * 32;
    L.lists = o;
    L.tail_end = o; if (has_lists) o += (size_t)n * 4;
    L.xpush_end = o; if (has_lists) o += (size_t)n * 4;
    L.tail_dep = o; o += (size_t)ntail * 4;
    L.xpush_slot = o; o += (size_t)npushx * 4;
    o = align_up(o, 16);
    L.tail_vals = o; o += (size_t)ntail * 72;
    L.total = align_up(o, 16);
    return L;
}

// Host threads for the per-CTA parts of the program builders (OPMGPU_ANALYSIS_THREADS, default: the
// hardware's, at most 16).  Work items are handed out through an atomic counter; every item writes
// to its own part of the output, so the result does not depend on the thread count.
inline int analysis_threads()
{
    static const int n = [] {
        if (const char* e = std::getenv("OPMGPU_ANALYSIS_THREADS")) return std::max(1, std::atoi(e));
        return (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    }();
    return n;
}
template <class F>
void parallel_for(int n, F&& f)
{
    const int T = std::min(analysis_threads(), n);
    if (T <= 1) { for (int i = 0; i < n; ++i) f(i); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    th.reserve(T);
    for (int t = 0; t < T; ++t)
        th.emplace_back([&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i); });
    for (auto& x : th) x.join();
}
// keys whose leading field is the owner: bucket by owner, sort the buckets side by side
inline void sort_keys_by_owner(std::vector<unsigned long long>& key, const std::vector<int>& owner, int max_owner)
{
    const int N = (int)key.size();
    std::vector<size_t> bptr((size_t)max_owner + 2, 0);
    for (int r = 0; r < N; ++r) ++bptr[(size_t)owner[r] + 1];
    for (int c = 0; c <= max_owner; ++c) bptr[c + 1] += bptr[c];
    std::vector<unsigned long long> skey(N);
    {
        std::vector<size_t> fill(bptr.begin(), bptr.end() - 1);
        for (int r = 0; r < N; ++r) skey[fill[owner[r]]++] = key[r];
    }
    parallel_for(max_owner + 1, [&](int c) { std::sort(skey.begin() + bptr[c], skey.begin() + bptr[c + 1]); });
    key.swap(skey);
}

// upos_of_row: for the lower program, position of every natural row in the upper program's
// order (the lower sweep hands its result to the upper sweep in that order); null for upper.
// owner[r]: CTA of row r; tile[r]: its tile.  A CTA may own several tiles (large grids), which it
// walks one after the other in ascending tile id; a dependency on another TILE travels through
// a push slot even when the same CTA owns both.
void build_pipe_program(int N, const int* rowptr, const int* colidx, const std::vector<int>& diag,
                        const std::vector<int>& level, int nlevels, const std::vector<int>& owner,
                        const std::vector<int>& tile,
                        int P, bool lower, const std::vector<int>* upos_of_row, PipeProgram& pg, int cs = 1,
                        std::vector<unsigned>* ri4_index = nullptr)
{
    // ri4_index (lower program built beside the upper one): where each row's "position in the upper
    // sweep's order" lives in ibuf (int index), to be filled in once the upper program is known
    if (ri4_index) ri4_index->assign(N, 0);
    pg = PipeProgram();
    pg.P = P;
    pg.cluster_size = cs;
    const bool timing__ = std::getenv("OPMGPU_DEBUG2") != nullptr;
    auto t_last__ = std::chrono::steady_clock::now();
    auto tick__ = [&](const char* what) {
        if (!timing__) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[opmgpu]   pipe %s: %-28s %8.1f ms\n", lower ? "L" : "U", what, std::chrono::duration<double, std::milli>(now - t_last__).count());
        t_last__ = now;
    };
    // a dependency on another CTA of the same cluster travels through distributed shared memory
    auto same_cluster = [&](int ca_, int cb_) { return cs > 1 && ca_ != cb_ && ca_ / cs == cb_ / cs; };
    pg.nlevels = nlevels;
    const bool upper = !lower;
    // rows by (CTA, tile in walking order, level, row): one 64-bit key per row, so the sort runs
    // over a contiguous array (tile ids ascend with the L wavefront, the upper sweep walks them backwards)
    std::vector<int> order(N);
    {
        int max_tile = 0, max_level = 0, max_owner = 0;
        for (int r = 0; r < N; ++r) { max_tile = std::max(max_tile, tile[r]); max_level = std::max(max_level, level[r]); max_owner = std::max(max_owner, owner[r]); }
        auto bits = [](long long v) { int b = 1; while ((1LL << b) <= v) ++b; return b; };
        const int br = bits(N), bl = bits(max_level), bt = bits(max_tile), bo = bits(max_owner);
        if (br + bl + bt + bo <= 64) {
            std::vector<unsigned long long> key(N);
            for (int r = 0; r < N; ++r) {
                const unsigned long long t = lower ? (unsigned long long)tile[r] : (unsigned long long)(max_tile - tile[r]);
                key[r] = ((((unsigned long long)owner[r] << bt | t) << bl | (unsigned long long)level[r]) << br) | (unsigned long long)r;
            }
            sort_keys_by_owner(key, owner, max_owner);
            const unsigned long long mask = (1ULL << br) - 1;
            for (int q = 0; q < N; ++q) order[q] = (int)(key[q] & mask);
        } else {
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                if (owner[a] != owner[b]) return owner[a] < owner[b];
                if (tile[a] != tile[b]) return lower ? tile[a] < tile[b] : tile[a] > tile[b];
                return level[a] < level[b];
            });
        }
    }
    tick__("sort");
    std::vector<int> nblk(N, 0), next(N, 0), npush(N, 0);
    auto for_deps = [&](int r, auto&& fn) {
        if (lower) { for (int k = rowptr[r]; k < diag[r]; ++k) fn(k); }
        else       { for (int k = rowptr[r + 1] - 1; k > diag[r]; --k) fn(k); }
    };
    const int nchunk = std::max(1, std::min(256, N / 4096));
    parallel_for(nchunk, [&](int ch) {          // (the push counts of other rows are the only shared writes)
        const int r0 = (int)((long long)N * ch / nchunk), r1 = (int)((long long)N * (ch + 1) / nchunk);
        for (int r = r0; r < r1; ++r)
            for_deps(r, [&](int k) {
                ++nblk[r];
                if (tile[colidx[k]] != tile[r]) {
                    if (!same_cluster(owner[colidx[k]], owner[r])) ++next[r];
                    __atomic_fetch_add(&npush[colidx[k]], 1, __ATOMIC_RELAXED);
                }
            });
    });
    auto tail_of = [&](int r) { return std::max(0, nblk[r] - kFastBlocks); };
    auto pushx_of = [&](int r) { return std::max(0, npush[r] - 2); };
    // steps: split CTA row lists at level changes and at the record limits
    struct Step { int q0, q1; };
    std::vector<Step> steps;
    std::vector<int> qlocal(N, 0);
    pg.cta_step_ptr.assign(P + 1, 0);
    {
        int q = 0;
        for (int c = 0; c < P; ++c) {
            pg.cta_step_ptr[c] = (int)steps.size();
            const int cbeg = q;
            int cur_level = -1, cur_tile = -1, rows = 0, tail = 0, ext = 0, pushx = 0;
            while (q < N && owner[order[q]] == c) {
                const int r = order[q];
                const bool fits = rows > 0 && level[r] == cur_level && tile[r] == cur_tile && rows + 1 <= kMaxStepRows &&
                                  ext + next[r] <= kMaxStepExt &&
                                  rec_layout(rows + 1, tail + tail_of(r), pushx + pushx_of(r), upper, true).total <= (size_t)kMaxStepBytes;
                if (!fits) {
                    if (rows > 0) steps.back().q1 = q;
                    steps.push_back({q, q});
                    cur_level = level[r]; cur_tile = tile[r]; rows = tail = ext = pushx = 0;
                    if (rec_layout(1, tail_of(r), pushx_of(r), upper, true).total > (size_t)kMaxStepBytes || next[r] > kMaxStepExt)
                        return;                           // a single row exceeds a record: not pipelinable
                }
                ++rows; tail += tail_of(r); ext += next[r]; pushx += pushx_of(r);
                qlocal[r] = q - cbeg;
                ++q;
            }
            if (rows > 0) steps.back().q1 = q;
        }
        pg.cta_step_ptr[P] = (int)steps.size();
    }
    tick__("counts + steps");
    // push slots: ordinals in consumption order per CTA
    pg.cta_ext_base.assign(P + 1, 0);
    std::vector<int> push_ptr(N + 1, 0);
    for (int r = 0; r < N; ++r) push_ptr[r + 1] = push_ptr[r] + npush[r];
    std::vector<long long> push_slot(push_ptr[N]);
    std::vector<int> push_fill(push_ptr.begin(), push_ptr.end() - 1);
    {
        long long base = 0;
        int q = 0;
        for (int c = 0; c < P; ++c) {
            pg.cta_ext_base[c] = base;
            long long e = 0;
            int cx = 0;
            while (q < N && owner[order[q]] == c) {
                const int r = order[q];
                for_deps(r, [&](int k) {
                    const int j = colidx[k];
                    if (tile[j] == tile[r]) return;
                    if (same_cluster(owner[j], c)) { push_slot[push_fill[j]++] = kPushDsmem | ((c % cs) << 20) | cx; ++cx; }
                    else { push_slot[push_fill[j]++] = base + e; ++e; }
                });
                ++q;
            }
            base += e;
            pg.max_cx = std::max(pg.max_cx, cx);
        }
        pg.cta_ext_base[P] = base;
        pg.total_ext = base;
        if (base >= (1LL << 30) || pg.max_cx > kMaxCxEntries) return;
    }
    tick__("push slots");
    // program-order positions (rhs segments are fetched by bulk copies: even row counts)
    pg.step_rhs_row.resize(steps.size());
    pg.step_rhs_bytes.resize(steps.size());
    std::vector<int> step_end_q(N), pos_of_row(N);
    {
        long long pos = 0;
        for (int c = 0; c < P; ++c) {
            int cbeg = -1;
            for (int sidx = pg.cta_step_ptr[c]; sidx < pg.cta_step_ptr[c + 1]; ++sidx) {
                if (cbeg < 0) cbeg = steps[sidx].q0;
                const int rows = steps[sidx].q1 - steps[sidx].q0;
                const int padded = (rows + 1) & ~1;
                pg.step_rhs_row[sidx] = (unsigned)pos;
                pg.step_rhs_bytes[sidx] = (unsigned)padded * 24u;
                for (int q = steps[sidx].q0; q < steps[sidx].q1; ++q) {
                    step_end_q[order[q]] = steps[sidx].q1 - cbeg;
                    pos_of_row[order[q]] = (int)(pos + (q - steps[sidx].q0));
                }
                pos += padded;
                pg.max_step_rows = std::max(pg.max_step_rows, padded);
            }
        }
        if (pos >= (1LL << 31)) return;
        pg.nperm = pos;
        pg.perm_row.assign((size_t)pos, -1);
        for (int r = 0; r < N; ++r) pg.perm_row[pos_of_row[r]] = r;
    }
    // own results older than the window must also live in HBM (natural order)
    std::vector<unsigned char> write_global(N, 0);
    parallel_for(nchunk, [&](int ch) {          // (concurrent writers of an entry all store 1)
        const int r0 = (int)((long long)N * ch / nchunk), r1 = (int)((long long)N * (ch + 1) / nchunk);
        for (int r = r0; r < r1; ++r)
            for_deps(r, [&](int k) {
                const int j = colidx[k];
                if (tile[j] == tile[r] && qlocal[j] + kWindowRows < step_end_q[r]) write_global[j] = 1;
            });
    });
    tick__("positions + write_global");
    // record sizes
    size_t total_bytes = 0, total_ibytes = 0;
    pg.step_off16.resize(steps.size());
    pg.step_bytes.resize(steps.size());
    pg.step_ioff16.resize(steps.size()); pg.step_ilen.resize(steps.size()); pg.step_roff.resize(steps.size());
    std::vector<RecLayout> lay(steps.size());
    for (size_t sidx = 0; sidx < steps.size(); ++sidx) {
        int rows = steps[sidx].q1 - steps[sidx].q0, tail = 0, pushx = 0;
        for (int q = steps[sidx].q0; q < steps[sidx].q1; ++q) { tail += tail_of(order[q]); pushx += pushx_of(order[q]); }
        lay[sidx] = rec_layout(rows, tail, pushx, upper, tail + pushx > 0);
        pg.step_off16[sidx] = (unsigned)(total_bytes / 16);
        pg.step_bytes[sidx] = (unsigned)lay[sidx].total;
        total_bytes += lay[sidx].total;
        pg.max_step_bytes = std::max(pg.max_step_bytes, (int)lay[sidx].total);
        // compact stream: header | [rowints, tail_vals) of the record
        pg.step_ioff16[sidx] = (unsigned)(total_ibytes / 16);
        pg.step_ilen[sidx] = (unsigned)(32 + (lay[sidx].tail_vals - lay[sidx].rowints));
        pg.step_roff[sidx] = (unsigned)lay[sidx].rowints;
        total_ibytes += pg.step_ilen[sidx];
    }
    if (total_bytes / 16 >= (1ull << 32) || total_bytes / 8 >= (1ull << 32) || total_ibytes / 16 >= (1ull << 32)) return;
    pg.total_bytes = total_bytes + 16;
    pg.ibuf.assign(total_ibytes + 16, 0);
    // where each CTA's (source block, destination) pairs go: known from the block counts
    std::vector<size_t> val_off((size_t)P + 1, 0);
    for (int c = 0; c < P; ++c) {
        size_t cnt = 0;
        if (pg.cta_step_ptr[c + 1] > pg.cta_step_ptr[c])
            for (int q = steps[pg.cta_step_ptr[c]].q0; q < steps[pg.cta_step_ptr[c + 1] - 1].q1; ++q) cnt += (size_t)nblk[order[q]] + (upper ? 1 : 0);
        val_off[c + 1] = val_off[c] + cnt;
    }
    pg.val_src.resize(val_off[P]); pg.val_dst8.resize(val_off[P]); pg.val_stride.resize(val_off[P]);
    std::atomic<bool> any_slow{false}, any_global{false};
    tick__("sizes + buffer");
    // emit (CTAs side by side: every CTA writes its own records and its own range of the value lists)
    parallel_for(P, [&](int c) {
        long long e = 0;
        int cx = 0;
        size_t vo = val_off[c];
        bool cta_slow = false, cta_global = false;
        for (int sidx = pg.cta_step_ptr[c]; sidx < pg.cta_step_ptr[c + 1]; ++sidx) {
            const size_t rec_off = (size_t)pg.step_off16[sidx] * 16;
            unsigned char* irec = pg.ibuf.data() + (size_t)pg.step_ioff16[sidx] * 16;     // header | integer region
            const long long e_before = e;
            const int cta_q0 = steps[pg.cta_step_ptr[c]].q0;
            const RecLayout& L = lay[sidx];
            const int n = steps[sidx].q1 - steps[sidx].q0;
            const bool has_lists = L.tail_dep > L.lists;      // the cumulative arrays are present
            int* hdr = (int*)irec;
            unsigned char* ireg = irec + 32 - L.rowints;          // record offsets >= L.rowints map into the region
            int* rowints = (int*)(ireg + L.rowints);
            int* tail_end = (int*)(ireg + L.tail_end);
            int* xpush_end = (int*)(ireg + L.xpush_end);
            int* tail_dep = (int*)(ireg + L.tail_dep);
            int* xpush_slot = (int*)(ireg + L.xpush_slot);
            int nt = 0, npx = 0, rr = 0;
            for (int q = steps[sidx].q0; q < steps[sidx].q1; ++q, ++rr) {
                const int r = order[q];
                int* ri = rowints + 8 * rr;
                int kb = 0;
                for (int k3 = 0; k3 < kFastBlocks; ++k3) ri[1 + k3] = kDepZeroSlot * 3;
                for_deps(r, [&](int k) {
                    const int j = colidx[k];
                    int code;
                    if (tile[j] != tile[r] && same_cluster(owner[j], c)) code = (kCxBase + cx++) * 3;
                    else if (tile[j] != tile[r]) code = (kWindowRows + (int)((e++) % kExtRing)) * 3;
                    else if (qlocal[j] + kWindowRows >= step_end_q[r]) code = (qlocal[j] % kWindowRows) * 3;
                    else code = kDepGlobalBit | j;
                    if (kb < kFastBlocks) {
                        ri[1 + kb] = code;
                        pg.val_src[vo] = k;
                        pg.val_dst8[vo] = (unsigned)((rec_off + L.cf) / 8 + (size_t)(3 * rr) * 9 + kb * 3);
                        pg.val_stride[vo] = 9;
                        ++vo;
                    } else {
                        tail_dep[nt] = code;
                        pg.val_src[vo] = k;
                        pg.val_dst8[vo] = (unsigned)((rec_off + L.tail_vals) / 8 + (size_t)nt * 9);
                        pg.val_stride[vo] = 3;
                        ++vo;
                        ++nt;
                    }
                    ++kb;
                });
                int np = 0;
                ri[5] = -1; ri[6] = -1;
                ri[7] = ((steps[sidx].q0 - cta_q0 + rr) % kWindowRows) * 3;      // own window slot (doubles)
                for (int t = push_ptr[r]; t < push_ptr[r + 1]; ++t, ++np) {
                    if (np < 2) ri[5 + np] = (int)push_slot[t];
                    else xpush_slot[npx++] = (int)push_slot[t];
                }
                if (has_lists) { tail_end[rr] = nt; xpush_end[rr] = npx; }
                const bool slow = kb > kFastBlocks || np > 2;
                cta_slow |= slow; cta_global |= write_global[r] != 0;
                ri[0] = r | (write_global[r] ? kRowWriteGlobal : 0) | (slow ? kRowSlow : 0);
                ri[4] = (lower && upos_of_row) ? (*upos_of_row)[r] : 0;
                if (ri4_index) (*ri4_index)[r] = (unsigned)((ri + 4) - (int*)pg.ibuf.data());
                if (upper) {
                    pg.val_src[vo] = diag[r];
                    pg.val_dst8[vo] = (unsigned)((rec_off + L.dinv) / 8 + (size_t)(3 * rr) * 3);
                    pg.val_stride[vo] = 3;
                    ++vo;
                }
            }
            hdr[0] = n; hdr[1] = steps[sidx].q0 - cta_q0; hdr[2] = (int)e; hdr[3] = (int)(e - e_before);
            hdr[4] = nt; hdr[5] = (int)(L.lists / 8); hdr[6] = (int)(L.tail_vals / 8);
            hdr[7] = has_lists ? 1 : 0;
        }
        if (cta_slow) any_slow = true;
        if (cta_global) any_global = true;
    });
    tick__("emit");
    pg.lean = !any_slow.load() && !any_global.load() && pg.max_step_rows <= kLeanStepRows;
    pg.valid = true;
}


// ---- pipelined factorisation program ---------------------------------------------------------
inline size_t factor_rec_bytes(int n) { return align_up(32 + (size_t)n * (kFRowInts * 4 + kFRowVals * 8), 16); }

void build_factor_pipe_program(int N, const int* rowptr, const int* colidx, const std::vector<int>& diag,
                               const std::vector<int>& level, const std::vector<int>& owner,
                               const std::vector<int>& tile, int P, FactorPipeProgram& pg)
{
    pg = FactorPipeProgram();
    pg.P = P;
    std::vector<int> order(N);
    {
        int max_tile = 0, max_level = 0, max_owner = 0;
        for (int r = 0; r < N; ++r) { max_tile = std::max(max_tile, tile[r]); max_level = std::max(max_level, level[r]); max_owner = std::max(max_owner, owner[r]); }
        auto bits = [](long long v) { int b = 1; while ((1LL << b) <= v) ++b; return b; };
        const int br = bits(N), bl = bits(max_level), bt = bits(max_tile), bo = bits(max_owner);
        if (br + bl + bt + bo <= 64) {
            std::vector<unsigned long long> key(N);
            for (int r = 0; r < N; ++r)
                key[r] = ((((unsigned long long)owner[r] << bt | (unsigned long long)tile[r]) << bl | (unsigned long long)level[r]) << br) | (unsigned long long)r;
            sort_keys_by_owner(key, owner, max_owner);
            const unsigned long long mask = (1ULL << br) - 1;
            for (int q = 0; q < N; ++q) order[q] = (int)(key[q] & mask);
        } else {
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                if (owner[a] != owner[b]) return owner[a] < owner[b];
                if (tile[a] != tile[b]) return tile[a] < tile[b];
                return level[a] < level[b];
            });
        }
    }
    // every row simple?  slot of A_ji for every lower block (i,j), -1 when absent
    std::vector<int> ji_slot(rowptr[N], -1), next(N, 0), npush(N, 0);
    {
        // (row chunks side by side; the push counts of other rows are the only shared writes)
        std::atomic<bool> not_simple{false};
        const int nchunk = std::max(1, std::min(256, N / 4096));
        parallel_for(nchunk, [&](int ch) {
            const int r0 = (int)((long long)N * ch / nchunk), r1 = (int)((long long)N * (ch + 1) / nchunk);
            for (int r = r0; r < r1 && !not_simple.load(std::memory_order_relaxed); ++r) {
                if (diag[r] - rowptr[r] > kFastBlocks) { not_simple = true; return; }
                for (int k = rowptr[r]; k < diag[r]; ++k) {
                    const int j = colidx[k];
                    for (int kk = diag[j] + 1; kk < rowptr[j + 1]; ++kk) {
                        const int c2 = colidx[kk];
                        if (c2 == r) { ji_slot[k] = kk; continue; }
                        if (std::binary_search(colidx + rowptr[r], colidx + rowptr[r + 1], c2)) { not_simple = true; return; }   // fill off the diagonal
                    }
                    if (tile[j] != tile[r]) { ++next[r]; __atomic_fetch_add(&npush[j], 1, __ATOMIC_RELAXED); }
                }
            }
        });
        if (not_simple) return;
    }
    for (int r = 0; r < N; ++r) if (npush[r] > 2) return;
    // steps = levels of the CTA's tile
    struct Step { int q0, q1; };
    std::vector<Step> steps;
    std::vector<int> qlocal(N, 0), step_end_q(N, 0);
    pg.cta_step_ptr.assign(P + 1, 0);
    pg.cta_row_base.assign(P + 1, N);
    pg.fpos.assign(N, 0);
    for (int q = 0; q < N; ++q) pg.fpos[order[q]] = q;
    {
        int q = 0;
        for (int c = 0; c < P; ++c) {
            pg.cta_step_ptr[c] = (int)steps.size();
            pg.cta_row_base[c] = q;
            const int cbeg = q;
            int cur_level = -1, cur_tile = -1, rows = 0, ext = 0;
            while (q < N && owner[order[q]] == c) {
                const int r = order[q];
                if (rows == 0 || level[r] != cur_level || tile[r] != cur_tile) {
                    if (rows > 0) steps.back().q1 = q;
                    steps.push_back({q, q});
                    cur_level = level[r]; cur_tile = tile[r]; rows = ext = 0;
                }
                ++rows; ext += next[r];
                if (rows > kLeanStepRows || ext > kFMaxStepExt) return;
                qlocal[r] = q - cbeg;
                ++q;
            }
            if (rows > 0) steps.back().q1 = q;
            for (int sidx = pg.cta_step_ptr[c]; sidx < (int)steps.size(); ++sidx)
                for (int qq = steps[sidx].q0; qq < steps[sidx].q1; ++qq) step_end_q[order[qq]] = steps[sidx].q1 - cbeg;
        }
        pg.cta_step_ptr[P] = (int)steps.size();
    }
    for (int r = 0; r < N; ++r)
        for (int k = rowptr[r]; k < diag[r]; ++k) {
            const int j = colidx[k];
            if (tile[j] == tile[r] && qlocal[j] + kFWindow < step_end_q[r]) return;      // pivot left the window
        }
    // push slots: ordinals in consumption order per CTA
    pg.cta_ext_base.assign(P + 1, 0);
    std::vector<int> push_ptr(N + 1, 0);
    for (int r = 0; r < N; ++r) push_ptr[r + 1] = push_ptr[r] + npush[r];
    std::vector<long long> push_slot(push_ptr[N]);
    std::vector<int> push_fill(push_ptr.begin(), push_ptr.end() - 1);
    {
        long long base = 0;
        int q = 0;
        for (int c = 0; c < P; ++c) {
            pg.cta_ext_base[c] = base;
            long long e = 0;
            while (q < N && owner[order[q]] == c) {
                const int r = order[q];
                for (int k = rowptr[r]; k < diag[r]; ++k) {
                    const int j = colidx[k];
                    if (tile[j] != tile[r]) { push_slot[push_fill[j]++] = base + e; ++e; }
                }
                ++q;
            }
            base += e;
        }
        pg.cta_ext_base[P] = base;
        pg.total_ext = base;
        if (base >= (1LL << 31)) return;
    }
    size_t total_bytes = 0;
    pg.step_off16.resize(steps.size());
    pg.step_bytes.resize(steps.size());
    for (size_t sidx = 0; sidx < steps.size(); ++sidx) {
        const int n = steps[sidx].q1 - steps[sidx].q0;
        const size_t b = factor_rec_bytes(n);
        pg.step_off16[sidx] = (unsigned)(total_bytes / 16);
        pg.step_bytes[sidx] = (unsigned)b;
        total_bytes += b;
        pg.max_step_bytes = std::max(pg.max_step_bytes, (int)b);
        pg.max_step_rows = std::max(pg.max_step_rows, n);
    }
    if (total_bytes / 8 >= (1ull << 32)) return;
    pg.total_bytes = total_bytes + 16;
    {
        size_t total_ibytes = 0;
        pg.step_ioff16.resize(steps.size()); pg.step_ilen.resize(steps.size()); pg.step_roff.assign(steps.size(), 32u);
        for (size_t sidx = 0; sidx < steps.size(); ++sidx) {
            const int n = steps[sidx].q1 - steps[sidx].q0;
            pg.step_ioff16[sidx] = (unsigned)(total_ibytes / 16);
            pg.step_ilen[sidx] = (unsigned)(32 + (size_t)n * kFRowInts * 4);
            total_ibytes += pg.step_ilen[sidx];
        }
        pg.ibuf.assign(total_ibytes + 16, 0);
    }
    // where each CTA's (source block, destination) pairs go
    std::vector<size_t> val_off((size_t)P + 1, 0);
    for (int c = 0; c < P; ++c) {
        size_t cnt = 0;
        if (pg.cta_step_ptr[c + 1] > pg.cta_step_ptr[c])
            for (int q = steps[pg.cta_step_ptr[c]].q0; q < steps[pg.cta_step_ptr[c + 1] - 1].q1; ++q) {
                const int r = order[q];
                cnt += 1 + (size_t)(diag[r] - rowptr[r]);
                for (int k = rowptr[r]; k < diag[r]; ++k) cnt += ji_slot[k] >= 0 ? 1 : 0;
            }
        val_off[c + 1] = val_off[c] + cnt;
    }
    pg.val_src.resize(val_off[P]); pg.val_dst8.resize(val_off[P]);
    parallel_for(P, [&](int c) {
        long long e = 0;
        size_t vo = val_off[c];
        if (pg.cta_step_ptr[c] == pg.cta_step_ptr[c + 1]) return;
        const int cta_q0 = steps[pg.cta_step_ptr[c]].q0;
        for (int sidx = pg.cta_step_ptr[c]; sidx < pg.cta_step_ptr[c + 1]; ++sidx) {
            const size_t rec_off = (size_t)pg.step_off16[sidx] * 16;
            unsigned char* rec = pg.ibuf.data() + (size_t)pg.step_ioff16[sidx] * 16;      // header | rowints
            const int n = steps[sidx].q1 - steps[sidx].q0;
            const long long e_before = e;
            int* hdr = (int*)rec;
            int* rowints = (int*)(rec + 32);
            const size_t vals_off = rec_off + 32 + (size_t)n * kFRowInts * 4;
            int rr = 0;
            for (int q = steps[sidx].q0; q < steps[sidx].q1; ++q, ++rr) {
                const int r = order[q];
                int* ri = rowints + kFRowInts * rr;
                const size_t v8 = vals_off / 8 + (size_t)rr * kFRowVals;
                ri[0] = r; ri[1] = ri[2] = ri[3] = 0; ri[4] = 0; ri[5] = ri[6] = -1;
                ri[7] = (steps[sidx].q0 - cta_q0 + rr) % kFWindow;
                ri[8] = ri[9] = ri[10] = -1; ri[11] = diag[r];
                pg.val_src[vo] = diag[r]; pg.val_dst8[vo] = (unsigned)v8; ++vo;
                int kb = 0;
                for (int k = rowptr[r]; k < diag[r]; ++k, ++kb) {
                    const int j = colidx[k];
                    ri[1 + kb] = tile[j] != tile[r] ? kFWindow + (int)((e++) % kFRing) : qlocal[j] % kFWindow;
                    ri[4] |= 1 << kb;
                    ri[8 + kb] = k;
                    pg.val_src[vo] = k; pg.val_dst8[vo] = (unsigned)(v8 + 9 + kb * 18); ++vo;
                    if (ji_slot[k] >= 0) {
                        ri[4] |= 1 << (4 + kb);
                        pg.val_src[vo] = ji_slot[k]; pg.val_dst8[vo] = (unsigned)(v8 + 18 + kb * 18); ++vo;
                    }
                }
                int np = 0;
                for (int t = push_ptr[r]; t < push_ptr[r + 1]; ++t, ++np) ri[5 + np] = (int)push_slot[t];
            }
            hdr[0] = n; hdr[1] = steps[sidx].q0 - cta_q0; hdr[2] = (int)e; hdr[3] = (int)(e - e_before);
            hdr[4] = sidx > pg.cta_step_ptr[c] ? steps[sidx - 1].q1 - steps[sidx - 1].q0 : 0;      // rows of the previous step
        }
    });
    pg.valid = true;
}

}  // namespace

namespace {
void materialise(std::vector<unsigned char>& buf, size_t total_bytes, const std::vector<unsigned char>& ibuf,
                 const std::vector<unsigned>& off16, const std::vector<unsigned>& ioff16,
                 const std::vector<unsigned>& ilen, const std::vector<unsigned>& roff)
{
    buf.assign(total_bytes, 0);
    for (size_t s = 0; s < off16.size(); ++s) {
        unsigned char* rec = buf.data() + (size_t)off16[s] * 16;
        const unsigned char* src = ibuf.data() + (size_t)ioff16[s] * 16;
        std::copy(src, src + 32, rec);
        std::copy(src + 32, src + ilen[s], rec + roff[s]);
    }
}
}  // namespace
void infer_cartesian_grid(int N, const int* rowptr, const int* colidx, int& nx, int& ny, int& nz) { infer_grid(N, rowptr, colidx, nx, ny, nz); }
void materialise_records(PipeProgram& pg) { materialise(pg.buf, pg.total_bytes, pg.ibuf, pg.step_off16, pg.step_ioff16, pg.step_ilen, pg.step_roff); }
void materialise_records(FactorPipeProgram& pg) { materialise(pg.buf, pg.total_bytes, pg.ibuf, pg.step_off16, pg.step_ioff16, pg.step_ilen, pg.step_roff); }

void analyse_pattern(int N, const int* rowptr, const int* colidx, int P, PatternAnalysis& out,
                     bool force_simple, const ClusterCaps* caps_in)
{
    out = PatternAnalysis();
    out.N = N;
    out.nnzb = rowptr[N];
    out.diag.assign(N, -1);
    const bool timing = std::getenv("OPMGPU_DEBUG") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto tick = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[opmgpu] analysis: %-34s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
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
    std::vector<int> owner(N, 0), tile(N, 0);
    int nx, ny, nz;
    infer_grid(N, rowptr, colidx, nx, ny, nz);
    out.grid_nx = nx; out.grid_ny = ny; out.grid_nz = nz;
    if (P < 1) P = 1;
    ClusterCaps caps;
    if (caps_in) caps = *caps_in;
    caps.max_ctas[0] = P;
    if (force_simple) caps.max_ctas[1] = caps.max_ctas[2] = caps.max_ctas[3] = 0;
    const bool cartesian = nx > 0 && (long long)nx * ny >= 4;
    for (int attempt = 0; attempt < 2; ++attempt) {
        int cs = 1, Pl = P;
        if (cartesian) {
            int pa, pb, ca, cb;
            choose_tiling(nx, ny, nz, caps, pa, pb, ca, cb);
            if (const char* e = exp_env("OPMGPU_TILING")) {       // experiments: "AxB" or "AxB/CxD" (cluster shape)
                int a = 0, b = 0, c = 1, d = 1;
                const int got = std::sscanf(e, "%dx%d/%dx%d", &a, &b, &c, &d);
                if (got < 4) { c = d = 1; }
                const int lg = c * d == 1 ? 0 : (c * d == 2 ? 1 : (c * d == 4 ? 2 : (c * d == 8 ? 3 : -1)));
                if (got >= 2 && a >= 1 && b >= 1 && a <= nx && b <= ny && lg >= 0 && a % c == 0 && b % d == 0 &&
                    a * b <= caps.max_ctas[lg]) { pa = a; pb = b; ca = c; cb = d; }
            }
            cs = ca * cb;
            // Large grids: no tiling with one tile per CTA keeps a tile level within one pass of the
            // compute warps.  Then cut smaller tiles and give every CTA several, dealt out in
            // wavefront order (ascending i0 + j0), so that a CTA walks its tiles one after the other
            // while the wavefront moves on; tile-to-tile dependencies all go through push slots.
            int rounds = 1;
            if (cs == 1 && ((nx + pa - 1) / pa) * ((ny + pb - 1) / pb) > kLeanStepRows && !exp_env("OPMGPU_TILING") &&
                !exp_env("OPMGPU_ONE_TILE_PER_CTA")) {
                double best = 1e300;
                for (int a = 1; a <= nx; ++a)
                    for (int b = 1; b <= ny; ++b) {
                        const int wa = (nx + a - 1) / a, wb = (ny + b - 1) / b;
                        if (wa * wb > kLeanStepRows || (long long)a * b > 64LL * P) continue;
                        const int rnd = (a * b + P - 1) / P;
                        // per-CTA work (rows) first, then the length of the tile chain
                        const double t = (double)rnd * wa * wb * nz * 0.007 + (double)rnd * (wa + wb + nz) * 0.37 * 0.25 + (a + b) * 0.05;
                        if (t < best) { best = t; pa = a; pb = b; rounds = rnd; }
                    }
            }
            const int ncx = pa / ca, ncy = pb / cb;
            Pl = cs > 1 ? ncx * ncy * cs : P;
            if (std::getenv("OPMGPU_DEBUG"))
                std::fprintf(stderr, "[opmgpu] sweep tiling %d x %d column tiles, clusters of %d x %d, %d CTAs, %d tile(s) per CTA\n", pa, pb, ca, cb, Pl, rounds);
            out.tiles_a = pa; out.tiles_b = pb;
            std::vector<int> tile_id((size_t)pa * pb);
            if (rounds > 1) {
                // tile ids in wavefront order; CTA = id % P, so a CTA's tiles ascend with the wavefront
                std::vector<int> tl((size_t)pa * pb);
                std::iota(tl.begin(), tl.end(), 0);
                auto first = [](int t, int parts, int n) { int i = (int)(((long long)t * n + parts - 1) / parts); return i; };
                std::stable_sort(tl.begin(), tl.end(), [&](int x, int y) {
                    const int lx = first(x % pa, pa, nx) + first(x / pa, pb, ny), ly = first(y % pa, pa, nx) + first(y / pa, pb, ny);
                    return lx < ly;
                });
                for (size_t k = 0; k < tl.size(); ++k) tile_id[tl[k]] = (int)k;
            }
            for (int r = 0; r < N; ++r) {
                const int i = r % nx, j = (r / nx) % ny;
                const int a = (int)((long long)i * pa / nx), b = (int)((long long)j * pb / ny);
                if (rounds > 1) { tile[r] = tile_id[a + pa * b]; owner[r] = tile[r] % P; continue; }
                // cluster-major CTA numbering: the cs tiles of a ca x cb block are consecutive CTAs
                owner[r] = cs > 1 ? ((a / ca) + ncx * (b / cb)) * cs + (a % ca) + ca * (b % cb) : a + pa * b;
                tile[r] = owner[r];
            }
        } else {
            // generic: contiguous share of every level per CTA
            for (int l = 0; l < nL; ++l) {
                const int b = out.lvl_ptr[l], sz = out.lvl_ptr[l + 1] - b;
                for (int q = 0; q < sz; ++q) owner[out.lvl_rows[b + q]] = (int)((long long)q * P / sz);
            }
            tile = owner;
        }
        out.P = Pl; out.cluster_size = cs;
        tick("levels + tiling");
        // the factorisation program only depends on the partition: build it beside the sweep programs
        std::future<void> fjob;
        if (!force_simple)
            fjob = std::async(std::launch::async, [&, Pl]() {
                build_factor_pipe_program(N, rowptr, colidx, out.diag, lvlL, owner, tile, Pl, out.pipeF);
            });
        // the two sweep programs are built side by side; the lower one needs every row's position in
        // the upper sweep's order, which is patched in afterwards
        std::vector<unsigned> ri4;
        std::future<void> ljob = std::async(std::launch::async, [&, Pl, cs]() {
            build_pipe_program(N, rowptr, colidx, out.diag, lvlL, nL, owner, tile, Pl, true, nullptr, out.pipeL, cs, &ri4);
        });
        build_pipe_program(N, rowptr, colidx, out.diag, lvlU, nU, owner, tile, Pl, false, nullptr, out.pipeU, cs);
        tick("upper sweep program");
        ljob.get();
        if (out.pipeU.valid && out.pipeL.valid) {
            int* li = (int*)out.pipeL.ibuf.data();
            for (size_t q = 0; q < out.pipeU.perm_row.size(); ++q)
                if (out.pipeU.perm_row[q] >= 0) li[ri4[out.pipeU.perm_row[q]]] = (int)q;
        } else {
            out.pipeL = PipeProgram();
        }
        tick("lower sweep program (concurrent)");
        if (fjob.valid()) fjob.get();
        tick("factorisation program (concurrent)");
        // clusters are only supported by the lean kernels: otherwise lay everything out again without
        if (cs > 1 && !(out.pipeL.valid && out.pipeU.valid && out.pipeL.lean && out.pipeU.lean)) {
            caps.max_ctas[1] = caps.max_ctas[2] = caps.max_ctas[3] = 0;
            continue;
        }
        P = Pl;
        break;
    }
    if (!(out.pipeL.valid && out.pipeU.valid)) out.pipeF = FactorPipeProgram();
    out.nlevL = nL; out.nlevU = nU;
    // the flag-synchronised tile kernel's program is only needed when the pipelined factorisation
    // is not available (or switched off): built on demand from the partition kept here
    out.owner_ = owner; out.level_lower_ = lvlL;
    if (force_simple || !out.pipeF.valid || exp_env("OPMGPU_FACTOR_TILE") || exp_env("OPMGPU_FACTOR_BY_LEVELS")) {
        build_tile_factor_program(rowptr, colidx, out);
        tick("tile-kernel factorisation program");
    }
    if (force_simple || !out.pipeL.valid || !out.pipeU.valid) {
        out.pipeL = PipeProgram(); out.pipeU = PipeProgram();
        build_program(N, rowptr, colidx, lvlU, nU, owner, P, false, out.upper);
    }
}

void build_tile_factor_program(const int* rowptr, const int* colidx, PatternAnalysis& an)
{
    if (!an.lower.prow.empty() || an.owner_.empty()) return;
    build_program(an.N, rowptr, colidx, an.level_lower_, an.nlevL, an.owner_, an.P, true, an.lower);
}

// Sequential interpreter of the pipelined programs (debug / CPU tests): executes the records
// exactly as the kernel does -- thread-slot arrays, window slots, own-global reads, pushed
// results, program-order right-hand sides -- advancing each CTA as far as its pushed inputs
// allow.  rhs_perm is in this program's order; hand_off (lower only) receives the result in
// the upper program's order; out (upper only) is in natural order.  Returns false on a deadlock.
bool interpret_pipe_program(const PipeProgram& pg, bool upper, const double* rhs_perm, double* work,
                            double* hand_off, double* out, double w, int scale)
{
    const int P = pg.P;
    std::vector<double> ext((size_t)std::max<long long>(pg.total_ext, 1) * 3);
    std::vector<unsigned char> ext_valid((size_t)std::max<long long>(pg.total_ext, 1), 0);
    std::vector<int> cur(P);
    const int cs = pg.cluster_size;
    const size_t ncx = (size_t)std::max(pg.max_cx, 1);
    std::vector<std::vector<double>> dep(P, std::vector<double>((size_t)(kCxBase + ncx) * 3, 0.0));
    std::vector<std::vector<unsigned char>> cx_valid(P, std::vector<unsigned char>(ncx, 0));
    std::vector<long long> qbase(P, 0), ext_seen(P, 0);
    for (int c = 0; c < P; ++c) cur[c] = pg.cta_step_ptr[c];
    bool progress = true, done = false;
    while (progress && !done) {
        progress = false; done = true;
        for (int c = 0; c < P; ++c) {
            while (cur[c] < pg.cta_step_ptr[c + 1]) {
                done = false;
                const int sidx = cur[c];
                const unsigned char* rec = pg.buf.data() + (size_t)pg.step_off16[sidx] * 16;
                const int* hdr = (const int*)rec;
                const int n = hdr[0];
                const double* cf = (const double*)(rec + 32);
                const double* dinv = cf + (size_t)3 * n * 9;
                const int* rowints = (const int*)(rec + ((32 + (size_t)3 * n * 72 + (upper ? (size_t)3 * n * 24 : 0) + 15) / 16) * 16);
                const int* lists = (const int*)(rec + (size_t)hdr[5] * 8);
                const bool has_lists = hdr[7] != 0;
                const int* tail_end = lists;
                const int* xpush_end = lists + (has_lists ? n : 0);
                const int* tail_dep = lists + (has_lists ? 2 * n : 0);
                const int* xpush_slot = tail_dep + hdr[4];
                const double* tail_vals = (const double*)(rec + (size_t)hdr[6] * 8);
                const long long ext_begin = ext_seen[c], ext_end = hdr[2];
                bool ready = true;
                for (long long e = ext_begin; e < ext_end && ready; ++e)
                    if (!ext_valid[pg.cta_ext_base[c] + e]) ready = false;
                // results delivered through distributed shared memory are validated entry by entry
                for (int rr = 0; rr < n && ready; ++rr)
                    for (int k = 0; k < kFastBlocks; ++k) {
                        const int code = rowints[8 * rr + 1 + k];
                        if (code >= kCxBase * 3 && !cx_valid[c][code / 3 - kCxBase]) ready = false;
                    }
                if (!ready) break;
                // the helper stages pushed results into the ring part of the dependency array
                for (long long e = ext_begin; e < ext_end; ++e)
                    for (int t = 0; t < 3; ++t)
                        dep[c][(size_t)(kWindowRows + e % kExtRing) * 3 + t] = ext[(size_t)(pg.cta_ext_base[c] + e) * 3 + t];
                std::vector<double> res((size_t)n * 3);
                const double* rhs = rhs_perm + (size_t)pg.step_rhs_row[sidx] * 3;
                auto yptr = [&](int code) -> const double* {
                    return code < 0 ? &work[(size_t)(code & kDepValueMask) * 3] : &dep[c][(size_t)code];
                };
                for (int rr = 0; rr < n; ++rr) {
                    double acc[3];
                    for (int cc = 0; cc < 3; ++cc) {
                        const int j = 3 * rr + cc;
                        double a = rhs[j];
                        for (int k = 0; k < kFastBlocks; ++k) {
                            const double* y = yptr(rowints[8 * rr + 1 + k]);
                            for (int e = 0; e < 3; ++e) a = std::fma(-cf[(size_t)j * 9 + k * 3 + e], y[e], a);
                        }
                        if (rowints[8 * rr] & kRowSlow) {
                            for (int t = rr ? tail_end[rr - 1] : 0; t < tail_end[rr]; ++t) {
                                const double* y = yptr(tail_dep[t]);
                                const double* av = tail_vals + (size_t)t * 9 + cc * 3;
                                for (int e = 0; e < 3; ++e) a = std::fma(-av[e], y[e], a);
                            }
                        }
                        acc[cc] = a;
                    }
                    const int row = rowints[8 * rr] & kRowMask;
                    if (upper) {
                        double v[3];
                        for (int cc = 0; cc < 3; ++cc) {
                            const int j = 3 * rr + cc;
                            double t = 0.0;
                            for (int e = 0; e < 3; ++e) t = std::fma(dinv[(size_t)j * 3 + e], acc[e], t);
                            v[cc] = t;
                        }
                        for (int cc = 0; cc < 3; ++cc) { acc[cc] = v[cc]; out[(size_t)row * 3 + cc] = scale ? v[cc] * w : v[cc]; }
                    } else {
                        for (int cc = 0; cc < 3; ++cc) hand_off[(size_t)rowints[8 * rr + 4] * 3 + cc] = acc[cc];
                    }
                    for (int cc = 0; cc < 3; ++cc) res[(size_t)rr * 3 + cc] = acc[cc];
                }
                for (int rr = 0; rr < n; ++rr) {
                    const int rowinfo_rr = rowints[8 * rr];
                    const int row = rowinfo_rr & kRowMask;
                    const size_t slot = (size_t)((qbase[c] + rr) % kWindowRows) * 3;
                    for (int t = 0; t < 3; ++t) dep[c][slot + t] = res[(size_t)rr * 3 + t];
                    if (rowinfo_rr & kRowWriteGlobal)
                        for (int t = 0; t < 3; ++t) work[(size_t)row * 3 + t] = res[(size_t)rr * 3 + t];
                    auto push_to = [&](int slot_id) {
                        if (slot_id & kPushDsmem) {      // into the shared memory of a CTA of the same cluster
                            const int target = (c / cs) * cs + ((slot_id >> 20) & 0xf), idx = slot_id & 0xfffff;
                            for (int u = 0; u < 3; ++u) dep[target][(size_t)(kCxBase + idx) * 3 + u] = res[(size_t)rr * 3 + u];
                            cx_valid[target][idx] = 1;
                            return;
                        }
                        for (int u = 0; u < 3; ++u) ext[(size_t)slot_id * 3 + u] = res[(size_t)rr * 3 + u];
                        ext_valid[slot_id] = 1;
                    };
                    if (rowints[8 * rr + 5] >= 0) push_to(rowints[8 * rr + 5]);
                    if (rowints[8 * rr + 6] >= 0) push_to(rowints[8 * rr + 6]);
                    if (rowinfo_rr & kRowSlow)
                        for (int t = rr ? xpush_end[rr - 1] : 0; t < xpush_end[rr]; ++t) push_to(xpush_slot[t]);
                }
                qbase[c] += n;
                ext_seen[c] = ext_end;
                ++cur[c];
                progress = true;
            }
        }
    }
    return done;
}


namespace {
inline double host_mat3_invert(double* M)
{
    double A[9];
    for (int q = 0; q < 9; ++q) A[q] = M[q];
    const double t4 = A[0] * A[4], t6 = A[0] * A[5], t8 = A[1] * A[3];
    const double t10 = A[2] * A[3], t12 = A[1] * A[6], t14 = A[2] * A[6];
    const double det = (t4 * A[8] - t6 * A[7] - t8 * A[8] + t10 * A[7] + t12 * A[5] - t14 * A[4]);
    const double t17 = 1.0 / det;
    M[0] = (A[4] * A[8] - A[5] * A[7]) * t17;
    M[1] = -(A[1] * A[8] - A[2] * A[7]) * t17;
    M[2] = (A[1] * A[5] - A[2] * A[4]) * t17;
    M[3] = -(A[3] * A[8] - A[5] * A[6]) * t17;
    M[4] = (A[0] * A[8] - t14) * t17;
    M[5] = -(t6 - t10) * t17;
    M[6] = (A[3] * A[7] - A[4] * A[6]) * t17;
    M[7] = -(A[0] * A[7] - t12) * t17;
    M[8] = (t4 - t8) * t17;
    return det;
}
}  // namespace

// Executes the records exactly as ilu0_factor_pipe_kernel does (window entries, pushed pivots,
// arithmetic order of Dune::bilu0_decomposition), advancing each CTA as far as its inputs allow.
int interpret_factor_program(const FactorPipeProgram& pg, const double* vals, double* lu)
{
    std::vector<double> fout((size_t)pg.fpos.size() * kFEntry, 0.0);      // what the kernel writes
    const int P = pg.P;
    std::vector<unsigned char> buf(pg.buf);
    double* bufd = (double*)buf.data();
    for (size_t b = 0; b < pg.val_src.size(); ++b)
        for (int t = 0; t < 9; ++t) bufd[(size_t)pg.val_dst8[b] + t] = vals[(size_t)pg.val_src[b] * 9 + t];
    std::vector<double> ext((size_t)std::max<long long>(pg.total_ext, 1) * 9);
    std::vector<unsigned char> ext_valid((size_t)std::max<long long>(pg.total_ext, 1), 0);
    std::vector<std::vector<double>> dep(P, std::vector<double>((size_t)(kFWindow + kFRing) * 9, 0.0));
    std::vector<int> cur(P);
    std::vector<long long> ext_seen(P, 0);
    for (int c = 0; c < P; ++c) cur[c] = pg.cta_step_ptr[c];
    int bad = 0x7fffffff;
    bool progress = true, done = false;
    while (progress && !done) {
        progress = false; done = true;
        for (int c = 0; c < P; ++c) {
            while (cur[c] < pg.cta_step_ptr[c + 1]) {
                done = false;
                const int sidx = cur[c];
                const unsigned char* rec = buf.data() + (size_t)pg.step_off16[sidx] * 16;
                const int* hdr = (const int*)rec;
                const int n = hdr[0];
                const int* rowints = (const int*)(rec + 32);
                const double* rv = (const double*)(rec + 32 + (size_t)n * kFRowInts * 4);
                const long long ext_begin = ext_seen[c], ext_end = hdr[2];
                bool ready = true;
                for (long long e = ext_begin; e < ext_end && ready; ++e)
                    if (!ext_valid[pg.cta_ext_base[c] + e]) ready = false;
                if (!ready) break;
                for (long long e = ext_begin; e < ext_end; ++e)
                    for (int t = 0; t < 9; ++t)
                        dep[c][(size_t)(kFWindow + e % kFRing) * 9 + t] = ext[(size_t)(pg.cta_ext_base[c] + e) * 9 + t];
                std::vector<double> res((size_t)n * 9);
                for (int rr = 0; rr < n; ++rr) {
                    const int* ri = rowints + kFRowInts * rr;
                    const double* v = rv + (size_t)rr * kFRowVals;
                    double D[9];
                    for (int t = 0; t < 9; ++t) D[t] = v[t];
                    for (int kb = 0; kb < kFastBlocks; ++kb) {
                        if (!(ri[4] & (1 << kb))) continue;
                        const double* Aij = v + 9 + kb * 18;
                        const double* Aji = v + 18 + kb * 18;
                        const double* Dj = &dep[c][(size_t)ri[1 + kb] * 9];
                        double L[9];
                        for (int cc = 0; cc < 3; ++cc)
                            for (int j = 0; j < 3; ++j) {
                                double sacc = 0.0;
                                for (int k = 0; k < 3; ++k) sacc = std::fma(Aij[cc * 3 + k], Dj[k * 3 + j], sacc);
                                L[cc * 3 + j] = sacc;
                            }
                        for (int t = 0; t < 9; ++t) lu[(size_t)ri[8 + kb] * 9 + t] = L[t];
                        if (ri[4] & (1 << (4 + kb)))
                            for (int cc = 0; cc < 3; ++cc)
                                for (int j = 0; j < 3; ++j) {
                                    double sacc = 0.0;
                                    for (int k = 0; k < 3; ++k) sacc = std::fma(L[cc * 3 + k], Aji[k * 3 + j], sacc);
                                    D[cc * 3 + j] -= sacc;
                                }
                    }
                    const double det = host_mat3_invert(D);
                    if (!(det != 0.0) || std::isinf(det) || std::isnan(det)) bad = std::min(bad, ri[0]);
                    for (int t = 0; t < 9; ++t) { lu[(size_t)ri[11] * 9 + t] = D[t]; res[(size_t)rr * 9 + t] = D[t]; }
                }
                for (int rr = 0; rr < n; ++rr) {
                    const int* ri = rowints + kFRowInts * rr;
                    for (int t = 0; t < 9; ++t) dep[c][(size_t)ri[7] * 9 + t] = res[(size_t)rr * 9 + t];
                    // the kernel stores the step's window entries to program position cta_row_base + qbase + rr
                    for (int t = 0; t < 9; ++t) fout[(size_t)(pg.cta_row_base[c] + hdr[1] + rr) * kFEntry + t] = res[(size_t)rr * 9 + t];
                    if (pg.fpos[ri[0]] != pg.cta_row_base[c] + hdr[1] + rr) return -2;      // consumers address pivots through fpos
                    for (int u = 0; u < 2; ++u)
                        if (ri[5 + u] >= 0) {
                            for (int t = 0; t < 9; ++t) ext[(size_t)ri[5 + u] * 9 + t] = res[(size_t)rr * 9 + t];
                            ext_valid[ri[5 + u]] = 1;
                        }
                }
                ext_seen[c] = ext_end;
                ++cur[c];
                progress = true;
            }
        }
    }
    if (!done) return -2;
    return bad == 0x7fffffff ? -1 : bad;
}

void partition_local_rows(int N_local, const int* rowptr, const long long* colidx_global,
                          const long long* row_offsets, int world, int rank, LocalPartition& out)
{
    out = LocalPartition();
    out.N_local = N_local;
    const long long lo = row_offsets[rank], hi = row_offsets[rank + 1];
    const int nnz = rowptr[N_local];
    std::vector<long long> ghosts;
    for (int k = 0; k < nnz; ++k)
        if (colidx_global[k] < lo || colidx_global[k] >= hi) ghosts.push_back(colidx_global[k]);
    std::sort(ghosts.begin(), ghosts.end());
    ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
    out.ghost_global = ghosts;
    out.n_ghost = (int)ghosts.size();
    out.recv_cnt.assign(world, 0);
    out.recv_off.assign(world, 0);
    {
        int p = 0;
        for (size_t g = 0; g < ghosts.size(); ++g) {
            while (ghosts[g] >= row_offsets[p + 1]) ++p;
            out.recv_cnt[p]++;
        }
        for (int q = 1; q < world; ++q) out.recv_off[q] = out.recv_off[q - 1] + out.recv_cnt[q - 1];
    }
    out.colidx_full.resize(nnz);
    out.rowptr_diag.assign(N_local + 1, 0);
    for (int i = 0; i < N_local; ++i) {
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const long long c = colidx_global[k];
            if (c >= lo && c < hi) {
                out.colidx_full[k] = (int)(c - lo);
                out.colidx_diag.push_back((int)(c - lo));
                out.lu_src.push_back(k);
            } else {
                out.colidx_full[k] = N_local + (int)(std::lower_bound(ghosts.begin(), ghosts.end(), c) - ghosts.begin());
            }
        }
        out.rowptr_diag[i + 1] = (int)out.colidx_diag.size();
    }
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

// Debug entry (not part of the public ABI; used by the CPU test-suite to validate the host
// analysis without a GPU): runs both pipelined sweep programs through the sequential
// interpreter on factors given in BCRS layout.  Returns 0, or a negative code.
// Debug (host only, no GPU): wall time of analyse_pattern in ms, with the cluster shape of a B200
// (tools/analysis_time.py; OPMGPU_DEBUG=1 OPMGPU_DEBUG2=1 print the phases).
extern "C" double opmgpu_debug_analyse_only(int N, const int* rowptr, const int* colidx, int P)
{
    using namespace opmgpu;
    PatternAnalysis an;
    ClusterCaps caps;
    caps.max_ctas[0] = P; caps.max_ctas[1] = P; caps.max_ctas[2] = P / 4 * 4 - 16; caps.max_ctas[3] = P / 8 * 8 - 24;      // 148 / 132 / 120 at P = 148
    const auto t0 = std::chrono::steady_clock::now();
    analyse_pattern(N, rowptr, colidx, P, an, false, &caps);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

extern "C" int opmgpu_debug_host_program_apply(int N, const int* rowptr, const int* colidx,
                                               const double* lu, int P, double w,
                                               const double* d, double* v, int* info /*[8]*/)
{
    using namespace opmgpu;
    PatternAnalysis an;
    ClusterCaps caps;                     // OPMGPU_TEST_CLUSTER_CTAS="p2,p4,p8": co-resident CTAs per cluster size
    if (const char* e = std::getenv("OPMGPU_TEST_CLUSTER_CTAS"))
        std::sscanf(e, "%d,%d,%d", &caps.max_ctas[1], &caps.max_ctas[2], &caps.max_ctas[3]);
    analyse_pattern(N, rowptr, colidx, P, an, false, &caps);
    if (an.missing_diag_row >= 0) return -1;
    if (!an.pipeL.valid || !an.pipeU.valid) return -2;
    if (info) info[7] = 0;
    for (PipeProgram* pg : {&an.pipeL, &an.pipeU}) {
        materialise_records(*pg);
        double* base = (double*)pg->buf.data();
        for (size_t b = 0; b < pg->val_src.size(); ++b)
            for (int c = 0; c < 3; ++c)
                for (int e = 0; e < 3; ++e) {
                    const size_t dst = (size_t)pg->val_dst8[b] + (size_t)c * pg->val_stride[b] + e;
                    base[dst] = lu[(size_t)pg->val_src[b] * 9 + c * 3 + e];
                }
    }
    std::vector<double> dperm((size_t)an.pipeL.nperm * 3, 0.0), yLperm((size_t)an.pipeU.nperm * 3, 0.0);
    std::vector<double> workL((size_t)N * 3, 0.0), workU((size_t)N * 3, 0.0);
    for (size_t q = 0; q < an.pipeL.perm_row.size(); ++q)
        if (an.pipeL.perm_row[q] >= 0)
            for (int t = 0; t < 3; ++t) dperm[q * 3 + t] = d[(size_t)an.pipeL.perm_row[q] * 3 + t];
    const int scale = std::fabs(w - 1.0) > 1e-15 ? 1 : 0;
    if (!interpret_pipe_program(an.pipeL, false, dperm.data(), workL.data(), yLperm.data(), nullptr, w, scale)) return -3;
    if (!interpret_pipe_program(an.pipeU, true, yLperm.data(), workU.data(), nullptr, v, w, scale)) return -4;
    if (info) {
        info[0] = an.grid_nx; info[1] = an.grid_ny; info[2] = an.grid_nz;
        info[3] = an.pipeL.nlevels; info[4] = an.pipeL.max_step_rows; info[5] = an.pipeL.max_step_bytes;
        info[6] = (int)std::min<long long>(an.pipeL.total_ext, 0x7fffffff);
        info[7] = (int)(an.pipeL.cta_step_ptr[an.P]);
    }
    return 0;
}

// Debug entry (CPU tests of the row-partition logic, no GPU / NCCL needed): local pattern,
// ghost list and diagonal block of one rank.  Arrays are caller-allocated with nnzb_local
// (colidx_full, colidx_diag, lu_src) or N_local+1 / world entries.  Returns n_ghost, writes
// nnzb_diag to *nnzb_diag_out; ghost_global must hold nnzb_local entries.
extern "C" int opmgpu_debug_partition(int N_local, const int* rowptr, const long long* colidx_global,
                                      const long long* row_offsets, int world, int rank,
                                      int* colidx_full, long long* ghost_global, int* recv_cnt,
                                      int* rowptr_diag, int* colidx_diag, int* lu_src, int* nnzb_diag_out)
{
    using namespace opmgpu;
    LocalPartition lp;
    partition_local_rows(N_local, rowptr, colidx_global, row_offsets, world, rank, lp);
    std::copy(lp.colidx_full.begin(), lp.colidx_full.end(), colidx_full);
    std::copy(lp.ghost_global.begin(), lp.ghost_global.end(), ghost_global);
    std::copy(lp.recv_cnt.begin(), lp.recv_cnt.end(), recv_cnt);
    std::copy(lp.rowptr_diag.begin(), lp.rowptr_diag.end(), rowptr_diag);
    std::copy(lp.colidx_diag.begin(), lp.colidx_diag.end(), colidx_diag);
    std::copy(lp.lu_src.begin(), lp.lu_src.end(), lu_src);
    *nnzb_diag_out = (int)lp.colidx_diag.size();
    return lp.n_ghost;
}

// Debug entry (CPU tests): builds the pipelined factorisation program for P CTAs and runs it
// through the sequential interpreter.  lu receives the factors in BCRS layout.  Returns 0 and
// the first singular row (or -1) in *bad_row; -2 when the pattern has no pipelined program.
extern "C" int opmgpu_debug_host_factor_program(int N, const int* rowptr, const int* colidx,
                                                const double* vals, int P, double* lu, int* bad_row, int* info /*[4]*/)
{
    using namespace opmgpu;
    PatternAnalysis an;
    analyse_pattern(N, rowptr, colidx, P, an, false);
    if (an.missing_diag_row >= 0) return -1;
    if (!an.pipeF.valid) return -2;
    std::copy(vals, vals + (size_t)rowptr[N] * 9, lu);
    materialise_records(an.pipeF);
    const int rc = interpret_factor_program(an.pipeF, vals, lu);
    if (rc == -2) return -3;
    if (bad_row) *bad_row = rc;
    if (info) {
        info[0] = an.pipeF.max_step_rows; info[1] = an.pipeF.max_step_bytes;
        info[2] = (int)std::min<long long>(an.pipeF.total_ext, 0x7fffffff); info[3] = an.pipeF.cta_step_ptr[P];
    }
    return 0;
}
