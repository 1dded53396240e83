// See multi.hpp.  Host C++ on top of the library's own distributed C ABI (opmgpu_create_distributed,
// opmgpu_set_pattern_bcrs_distributed, opmgpu_solve_bcrs3): a partitioner and one thread per GPU.
//
// Partition (the part that used to live in Python, opm_simulators_legacy_b200/distributed.py):
// Cartesian patterns are cut into slabs along the grid axis with the weakest coupling (block-Jacobi
// ILU0 drops exactly the couplings a slab boundary cuts; measured by the mean |a_00| of the
// off-diagonal blocks per axis), rows renumbered so that a GPU's rows are contiguous and keep their
// natural relative order; any other pattern is cut into contiguous row blocks of the natural
// order.  Either way the halo plan comes from the pattern (partition_local_rows), not from a grid.
#include "multi.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "analysis.hpp"

namespace opmgpu {

namespace {
template <class T>
struct Pinned {
    T* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n)
    {
        if (n <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (cudaHostAlloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T), cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
        cap = n;
        return true;
    }
    ~Pinned() { if (p) cudaFreeHost(p); }
};
}  // namespace

struct MultiSolver {
    int G = 0;
    std::vector<int> dev;
    std::vector<opmgpu_handle> child;
    // global pattern
    int N = 0, nnzb = 0, nx = 0, ny = 0, nz = 0;
    std::vector<int> rowptr, colidx;
    bool partitioned = false;
    int axis = -1;
    std::vector<long long> perm, inv, offsets;          // perm[new row] = natural row
    struct Rank {
        std::vector<int> rowptr;                        // local rows
        std::vector<long long> colg;                    // global (new numbering) column ids
        std::vector<int> ent;                           // BCRS slot (global pattern) of every local entry
        Pinned<double> vals, rhs, x;
        opmgpu_result res;
        int rc = 0;
        std::string err;
    };
    std::vector<Rank> rk;
    // CSC front end (formInterleavedSystem on the host: index map once per pattern, values per call)
    std::vector<std::vector<int>> csc_colptr, csc_rowidx;
    std::vector<long long> map9;                        // [nnzb*9] index into block q's values, -1 = structural zero
    bool csc_full = false;

    template <class F>
    void on_ranks(F&& f)
    {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back([&, g]() { f(g); });
        for (auto& t : th) t.join();
    }
};

MultiSolver* multi_create(int ngpus, const int* device_ids, std::string& err)
{
    if (ngpus < 1 || !device_ids) { err = "opmgpu_create_multi: need at least one device"; return nullptr; }
    unsigned char id[128];
    if (opmgpu_nccl_unique_id(id) != OPMGPU_OK) { err = opmgpu_last_error(nullptr); return nullptr; }
    MultiSolver* m = new MultiSolver();
    m->G = ngpus;
    m->dev.assign(device_ids, device_ids + ngpus);
    m->child.assign(ngpus, nullptr);
    m->rk.resize(ngpus);
    std::vector<int> rc(ngpus, 0);
    std::vector<std::string> errs(ngpus);
    m->on_ranks([&](int g) {        // ncclCommInitRank returns when every rank has joined: all at once
        rc[g] = opmgpu_create_distributed(m->dev[g], g, ngpus, id, &m->child[g]);
        if (rc[g]) errs[g] = opmgpu_last_error(nullptr);
    });
    for (int g = 0; g < ngpus; ++g)
        if (rc[g]) { err = "GPU " + std::to_string(m->dev[g]) + ": " + errs[g]; multi_destroy(m); return nullptr; }
    return m;
}

void multi_destroy(MultiSolver* m)
{
    if (!m) return;
    m->on_ranks([&](int g) { if (m->child[g]) opmgpu_destroy(m->child[g]); });
    delete m;
}

int multi_set_precision(MultiSolver* m, int single_precision, std::string& err)
{
    for (int g = 0; g < m->G; ++g)
        if (int rc = opmgpu_set_precision(m->child[g], single_precision)) { err = opmgpu_last_error(m->child[g]); return rc; }
    return OPMGPU_OK;
}

int multi_set_pattern(MultiSolver* m, int N, int nnzb, const int* rowptr, const int* colidx, std::string& err)
{
    if (N < m->G) { err = "fewer block rows than GPUs"; return OPMGPU_BAD_ARGUMENT; }
    m->N = N; m->nnzb = nnzb;
    m->rowptr.assign(rowptr, rowptr + N + 1);
    m->colidx.assign(colidx, colidx + nnzb);
    infer_cartesian_grid(N, rowptr, colidx, m->nx, m->ny, m->nz);
    m->partitioned = false;         // the slab axis depends on the values: decided at the first solve
    m->csc_colptr.clear(); m->csc_rowidx.clear(); m->map9.clear();
    return OPMGPU_OK;
}

namespace {

// a00(slot): |entry [0][0]| of a BCRS block of the global pattern
int ensure_partition(MultiSolver* m, const std::function<double(int)>& a00, std::string& err)
{
    if (m->partitioned) return OPMGPU_OK;
    const int N = m->N, G = m->G;
    int axis = -1;
    if (m->nx > 0 && G > 1) {
        const long long stride[3] = {1, m->nx, (long long)m->nx * m->ny};
        const int len[3] = {m->nx, m->ny, m->nz};
        double sum[3] = {0, 0, 0};
        long long cnt[3] = {0, 0, 0};
        const int step = std::max(1, N / 200000);            // sample rows: the coupling pattern is uniform
        for (int r = 0; r < N; r += step)
            for (int k = m->rowptr[r]; k < m->rowptr[r + 1]; ++k) {
                const long long d = (long long)m->colidx[k] - r;
                for (int a = 0; a < 3; ++a)
                    if (d == stride[a] && len[a] > 1 && !(a > 0 && stride[a] == stride[a - 1])) { sum[a] += a00(k); ++cnt[a]; }
            }
        double best = 1e300;
        for (int a = 0; a < 3; ++a) {
            if (len[a] < G || cnt[a] == 0) continue;
            const double w = sum[a] / (double)cnt[a];
            if (w < best * (1 - 1e-12) || (std::fabs(w - best) <= 1e-12 * best && a > axis)) { best = w; axis = a; }
        }
    }
    m->axis = axis;
    m->perm.resize(N); m->inv.resize(N);
    m->offsets.assign(G + 1, 0);
    if (axis < 0) {
        for (int r = 0; r < N; ++r) m->perm[r] = r;
        for (int g = 0; g <= G; ++g) m->offsets[g] = (long long)N * g / G;
    } else {
        const int n = axis == 0 ? m->nx : (axis == 1 ? m->ny : m->nz);
        std::vector<int> owner_of_coord(n);
        for (int c = 0; c < n; ++c) {
            int g = (int)(((long long)c * G) / n);
            while ((long long)n * g / G > c) --g;                 // owner g has n*g/G <= c < n*(g+1)/G
            while ((long long)n * (g + 1) / G <= c) ++g;
            owner_of_coord[c] = g;
        }
        auto coord = [&](long long cell) { return axis == 0 ? (int)(cell % m->nx) : (axis == 1 ? (int)((cell / m->nx) % m->ny) : (int)(cell / ((long long)m->nx * m->ny))); };
        for (int r = 0; r < N; ++r) m->offsets[owner_of_coord[coord(r)] + 1]++;
        for (int g = 0; g < G; ++g) m->offsets[g + 1] += m->offsets[g];
        std::vector<long long> fill(m->offsets.begin(), m->offsets.end() - 1);
        for (int r = 0; r < N; ++r) m->perm[fill[owner_of_coord[coord(r)]]++] = r;      // stable: natural order inside a GPU
    }
    for (int q = 0; q < N; ++q) m->inv[m->perm[q]] = q;
    for (int g = 0; g < G; ++g) if (m->offsets[g + 1] == m->offsets[g]) { err = "a GPU would own no rows"; return OPMGPU_BAD_ARGUMENT; }
    // local patterns and the collective pattern set-up (halo plan, block-Jacobi diagonal blocks, analysis)
    m->on_ranks([&](int g) {
        MultiSolver::Rank& R = m->rk[g];
        const long long lo = m->offsets[g], hi = m->offsets[g + 1];
        R.rowptr.assign((size_t)(hi - lo) + 1, 0);
        size_t nnz = 0;
        for (long long q = lo; q < hi; ++q) nnz += (size_t)(m->rowptr[m->perm[q] + 1] - m->rowptr[m->perm[q]]);
        R.colg.resize(nnz); R.ent.resize(nnz);
        size_t e = 0;
        for (long long q = lo; q < hi; ++q) {
            const int r = (int)m->perm[q];
            for (int k = m->rowptr[r]; k < m->rowptr[r + 1]; ++k, ++e) { R.colg[e] = m->inv[m->colidx[k]]; R.ent[e] = k; }
            R.rowptr[q - lo + 1] = (int)e;
        }
        R.rc = opmgpu_set_pattern_bcrs_distributed(m->child[g], (int)(hi - lo), (int)nnz, R.rowptr.data(), R.colg.data(), m->offsets.data());
        if (R.rc) R.err = opmgpu_last_error(m->child[g]);
        if (!R.vals.ensure(nnz * 9) || !R.rhs.ensure((size_t)(hi - lo) * 3) || !R.x.ensure((size_t)(hi - lo) * 3)) { R.rc = OPMGPU_CUDA_ERROR; R.err = "page-locked host buffers"; }
    });
    for (int g = 0; g < G; ++g) if (m->rk[g].rc) { err = "GPU " + std::to_string(m->dev[g]) + ": " + m->rk[g].err; return m->rk[g].rc; }
    m->partitioned = true;
    return OPMGPU_OK;
}

// collective solve on the per-rank buffers; the result of rank 0 (statuses are agreed on by the ranks)
int solve_ranks(MultiSolver* m, const opmgpu_params* prm, opmgpu_result* res, std::string& err)
{
    m->on_ranks([&](int g) {
        MultiSolver::Rank& R = m->rk[g];
        R.rc = opmgpu_solve_bcrs3(m->child[g], R.vals.p, R.rhs.p, R.x.p, prm, &R.res);
        if (R.rc) R.err = opmgpu_last_error(m->child[g]);
    });
    *res = m->rk[0].res;
    int rc = OPMGPU_OK;
    for (int g = 0; g < m->G; ++g) {
        const MultiSolver::Rank& R = m->rk[g];
        res->ms_h2d = std::max(res->ms_h2d, R.res.ms_h2d); res->ms_factor = std::max(res->ms_factor, R.res.ms_factor);
        res->ms_solve = std::max(res->ms_solve, R.res.ms_solve); res->ms_d2h = std::max(res->ms_d2h, R.res.ms_d2h);
        if (R.rc == OPMGPU_SINGULAR_BLOCK && R.res.bad_row >= 0) res->bad_row = (int)m->perm[m->offsets[g] + R.res.bad_row];
        if (R.rc && (rc == OPMGPU_OK || R.rc < 0)) { rc = R.rc; err = "GPU " + std::to_string(m->dev[g]) + ": " + R.err; }
    }
    return rc;
}

}  // namespace

int multi_solve_bcrs3(MultiSolver* m, const double* vals, const double* rhs, double* x, const opmgpu_params* prm,
                      opmgpu_result* res, std::string& err)
{
    if (m->N == 0) { err = "set the pattern first"; return OPMGPU_BAD_ARGUMENT; }
    int rc = ensure_partition(m, [&](int k) { return std::fabs(vals[(size_t)k * 9]); }, err);
    if (rc) return rc;
    m->on_ranks([&](int g) {
        MultiSolver::Rank& R = m->rk[g];
        const long long lo = m->offsets[g], hi = m->offsets[g + 1];
        for (size_t e = 0; e < R.ent.size(); ++e) std::memcpy(R.vals.p + e * 9, vals + (size_t)R.ent[e] * 9, 72);
        for (long long q = lo; q < hi; ++q) std::memcpy(R.rhs.p + (size_t)(q - lo) * 3, rhs + (size_t)m->perm[q] * 3, 24);
    });
    rc = solve_ranks(m, prm, res, err);
    if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED)
        m->on_ranks([&](int g) {
            const long long lo = m->offsets[g], hi = m->offsets[g + 1];
            for (long long q = lo; q < hi; ++q) std::memcpy(x + (size_t)m->perm[q] * 3, m->rk[g].x.p + (size_t)(q - lo) * 3, 24);
        });
    return rc;
}

int multi_solve_from_csc_blocks(MultiSolver* m, int N, const opmgpu_csc blocks[9], const double scale[3],
                                const double* rhs_eqmajor, double* dx_varmajor, const opmgpu_params* prm,
                                opmgpu_result* res, std::string& err)
{
    const bool full = prm->require_full_sparsity_pattern != 0;
    // same index arrays as last time?  (the reference rebuilds the pattern on every call, ...Interleaved.cpp:110-194)
    bool same = m->N == N && m->csc_colptr.size() == 9 && m->csc_full == full;
    for (int q = 0; q < 9 && same; ++q) {
        const size_t nnz = (size_t)blocks[q].colptr[N];
        same = m->csc_rowidx[q].size() == nnz && std::memcmp(m->csc_colptr[q].data(), blocks[q].colptr, sizeof(int) * ((size_t)N + 1)) == 0 &&
               (nnz == 0 || std::memcmp(m->csc_rowidx[q].data(), blocks[q].rowidx, sizeof(int) * nnz) == 0);
    }
    if (!same) {
        for (int q = 0; q < 9; ++q) {
            const int* cp = blocks[q].colptr;
            if (!cp || cp[0] != 0) { err = "CSC block: colptr[0] != 0"; return OPMGPU_BAD_ARGUMENT; }
            for (int c = 0; c < N; ++c) {
                if (cp[c + 1] < cp[c]) { err = "CSC block: colptr not monotone"; return OPMGPU_BAD_ARGUMENT; }
                for (int k = cp[c]; k < cp[c + 1]; ++k)
                    if (blocks[q].rowidx[k] < 0 || blocks[q].rowidx[k] >= N) { err = "CSC block: row index out of range"; return OPMGPU_BAD_ARGUMENT; }
            }
        }
        std::vector<CscView> sel;
        for (int p1 = 0; p1 < 3; ++p1) sel.push_back({blocks[p1 * 3].colptr, blocks[p1 * 3].rowidx});
        if (full)
            for (int p1 = 0; p1 < 3; ++p1)
                for (int p2 = 1; p2 < 3; ++p2) sel.push_back({blocks[p1 * 3 + p2].colptr, blocks[p1 * 3 + p2].rowidx});
        std::vector<int> rowptr, colidx;
        union_pattern_from_csc(N, sel.data(), (int)sel.size(), rowptr, colidx);
        int rc = multi_set_pattern(m, N, rowptr[N], rowptr.data(), colidx.data(), err);
        if (rc) return rc;
        m->map9.assign((size_t)m->nnzb * 9, -1);
        bool outside = false;
        for (int q = 0; q < 9; ++q) {
            const opmgpu_csc& b = blocks[q];
            for (int c = 0; c < N; ++c)
                for (int k = b.colptr[c]; k < b.colptr[c + 1]; ++k) {
                    const int r = b.rowidx[k];
                    const int* lo = m->colidx.data() + m->rowptr[r];
                    const int* hi = m->colidx.data() + m->rowptr[r + 1];
                    const int* it = std::lower_bound(lo, hi, c);
                    if (it == hi || *it != c) { outside = true; continue; }
                    m->map9[(size_t)(it - m->colidx.data()) * 9 + q] = k;
                }
        }
        if (outside) { m->csc_colptr.clear(); err = "a Jacobian entry lies outside the sparsity pattern (set require_full_sparsity_pattern)"; return OPMGPU_BAD_PATTERN; }
        m->csc_colptr.assign(9, {}); m->csc_rowidx.assign(9, {});
        for (int q = 0; q < 9; ++q) {
            m->csc_colptr[q].assign(blocks[q].colptr, blocks[q].colptr + N + 1);
            m->csc_rowidx[q].assign(blocks[q].rowidx, blocks[q].rowidx + blocks[q].colptr[N]);
        }
        m->csc_full = full;
    }
    int rc = ensure_partition(m, [&](int k) { const long long s = m->map9[(size_t)k * 9]; return s < 0 ? 0.0 : std::fabs(blocks[0].val[s] * scale[0]); }, err);
    if (rc) return rc;
    m->on_ranks([&](int g) {        // formInterleavedSystem + scaling + rhs interleave of this GPU's rows
        MultiSolver::Rank& R = m->rk[g];
        const long long lo = m->offsets[g], hi = m->offsets[g + 1];
        for (size_t e = 0; e < R.ent.size(); ++e)
            for (int q = 0; q < 9; ++q) {
                const long long s = m->map9[(size_t)R.ent[e] * 9 + q];
                R.vals.p[e * 9 + q] = s < 0 ? 0.0 : blocks[q].val[s] * scale[q / 3];
            }
        for (long long q = lo; q < hi; ++q)
            for (int p = 0; p < 3; ++p) R.rhs.p[(size_t)(q - lo) * 3 + p] = rhs_eqmajor[(size_t)p * N + m->perm[q]] * scale[p];
    });
    rc = solve_ranks(m, prm, res, err);
    if (rc == OPMGPU_OK || rc == OPMGPU_NOT_CONVERGED)
        m->on_ranks([&](int g) {
            const long long lo = m->offsets[g], hi = m->offsets[g + 1];
            for (long long q = lo; q < hi; ++q)
                for (int p = 0; p < 3; ++p) dx_varmajor[(size_t)p * N + m->perm[q]] = m->rk[g].x.p[(size_t)(q - lo) * 3 + p];
        });
    return rc;
}

int multi_partition_info(MultiSolver* m, int* axis, long long* offsets)
{
    if (!m->partitioned) return OPMGPU_BAD_ARGUMENT;
    if (axis) *axis = m->axis;
    if (offsets) std::copy(m->offsets.begin(), m->offsets.end(), offsets);
    return OPMGPU_OK;
}

}  // namespace opmgpu
