// Host analysis of the multicolour ILU0 variant (mcorder.hpp).  Once per pattern.
#include "mcorder.hpp"
#include "analysis.hpp"      // infer_cartesian_grid

#include <algorithm>
#include <cmath>
#include <numeric>

namespace opmgpu {

void multicolour_order(int N, const int* rowptr, const int* colidx, McOrder& o)
{
    // entries (j, c) with c > j seen from row c: the neighbours of c that the transpose adds
    std::vector<int> tptr((size_t)N + 1, 0);
    for (int j = 0; j < N; ++j)
        for (int k = rowptr[j]; k < rowptr[j + 1]; ++k)
            if (colidx[k] > j) ++tptr[(size_t)colidx[k] + 1];
    for (int i = 0; i < N; ++i) tptr[i + 1] += tptr[i];
    std::vector<int> tcol((size_t)tptr[N]), fill(tptr.begin(), tptr.end() - 1);
    for (int j = 0; j < N; ++j)
        for (int k = rowptr[j]; k < rowptr[j + 1]; ++k)
            if (colidx[k] > j) tcol[fill[colidx[k]]++] = j;

    o.colour.assign(N, -1);
    std::vector<int> stamp;          // stamp[c] == i: colour c is taken by a neighbour of row i
    int nc = 0;
    for (int i = 0; i < N; ++i) {
        for (int k = rowptr[i]; k < rowptr[i + 1] && colidx[k] < i; ++k) stamp[o.colour[colidx[k]]] = i;
        for (int k = tptr[i]; k < tptr[i + 1]; ++k) stamp[o.colour[tcol[k]]] = i;
        int c = 0;
        while (c < nc && stamp[c] == i) ++c;
        if (c == nc) { ++nc; stamp.push_back(-1); }
        o.colour[i] = c;
    }
    o.ncolours = nc;
    o.colour_ptr.assign((size_t)nc + 1, 0);
    for (int i = 0; i < N; ++i) ++o.colour_ptr[(size_t)o.colour[i] + 1];
    for (int c = 0; c < nc; ++c) o.colour_ptr[c + 1] += o.colour_ptr[c];
    o.p2n.resize(N); o.n2p.resize(N);
    std::vector<int> pos(o.colour_ptr.begin(), o.colour_ptr.end() - 1);
    for (int i = 0; i < N; ++i) {
        const int q = pos[o.colour[i]]++;
        o.p2n[q] = i; o.n2p[i] = q;
    }
}

void line_order(int nx, int ny, int nz, McOrder& o)
{
    const int N = nx * ny * nz, plane = nx * ny;
    o = McOrder();
    o.lines = true; o.nx = nx; o.nz = nz; o.ncolours = plane > 1 ? 2 : 1;
    std::vector<int> rank(plane);          // rank of a column among the columns of its colour, natural (i + nx j) order
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) rank[i + nx * j] = o.ncols[(i + j) & 1]++;
    o.base[0] = 0; o.base[1] = nz * o.ncols[0];
    o.colour.resize(N); o.p2n.resize(N); o.n2p.resize(N);
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int cell = i + nx * (j + ny * k), c = (i + j) & 1;
                const int q = o.base[c] + k * o.ncols[c] + rank[i + nx * j];
                o.colour[cell] = c; o.n2p[cell] = q; o.p2n[q] = cell;
            }
    o.colour_ptr = {0, o.base[1]};
    if (o.ncolours == 2) o.colour_ptr.push_back(N);
}

bool build_mc_program(int N, const int* rowptr, const int* colidx, McProgram& m, bool lines)
{
    if (lines) {
        int nx = 0, ny = 0, nz = 0;
        infer_cartesian_grid(N, rowptr, colidx, nx, ny, nz);
        if (nx <= 0 || (long long)nx * ny * nz != N) return false;
        line_order(nx, ny, nz, m.ord);
    } else {
        multicolour_order(N, rowptr, colidx, m.ord);
    }
    const std::vector<int>& p2n = m.ord.p2n;
    const std::vector<int>& n2p = m.ord.n2p;
    const int nnzb = rowptr[N];
    m.prowptr.assign((size_t)N + 1, 0);
    m.pcol.resize(nnzb); m.psrc.resize(nnzb); m.ppos.resize(nnzb); m.pdiag.assign(N, -1);
    m.Lrowptr.assign((size_t)N + 1, 0); m.Urowptr.assign((size_t)N + 1, 0);
    std::vector<std::pair<int, int>> row;      // (permuted column, natural slot)
    for (int q = 0; q < N; ++q) {
        const int i = p2n[q];
        row.clear();
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) row.emplace_back(n2p[colidx[k]], k);
        std::sort(row.begin(), row.end());
        int b = m.prowptr[q], nl = 0, nu = 0;
        for (const auto& e : row) {
            m.pcol[b] = e.first; m.psrc[b] = e.second;
            if (e.first == q) m.pdiag[q] = b;
            else if (e.first < q) ++nl;
            else ++nu;
            ++b;
        }
        m.prowptr[q + 1] = b;
        m.Lrowptr[q + 1] = m.Lrowptr[q] + nl;
        m.Urowptr[q + 1] = m.Urowptr[q] + nu;
    }
    const long long nnzL = m.Lrowptr[N], nnzU = m.Urowptr[N];
    auto up4 = [](long long v) { return (v + 3) / 4 * 4; };        // 4 blocks = 288 / 144 bytes: 16-byte aligned in both precisions
    m.offD = up4(nnzL);
    m.offU = m.offD + up4(N);
    m.total_blocks = m.offU + up4(nnzU);
    m.Lcol.resize((size_t)nnzL); m.Ucol.resize((size_t)nnzU);
    for (int q = 0; q < N; ++q) {
        int l = m.Lrowptr[q];
        const int u_end = m.Urowptr[q + 1];
        int nu = 0;
        for (int b = m.prowptr[q]; b < m.prowptr[q + 1]; ++b) {
            const int c = m.pcol[b];
            if (c < q) { m.Lcol[l] = c; m.ppos[b] = l; ++l; }
            else if (c == q) m.ppos[b] = (int)(m.offD + q);
            else {              // ascending here, stored descending
                const int slot = u_end - 1 - nu;
                m.Ucol[slot] = c; m.ppos[b] = (int)(m.offU + slot); ++nu;
            }
        }
    }
    // update lists of the factorisation (what Dune::bilu0_decomposition finds by walking both rows)
    m.pair_ptr.assign((size_t)nnzL + 1, 0);
    m.pair_jk.clear(); m.pair_ik.clear();
    for (int q = 0; q < N; ++q) {
        const int iend = m.prowptr[q + 1];
        for (int ij = m.prowptr[q]; ij < m.pdiag[q]; ++ij) {
            const int j = m.pcol[ij];
            int jk = m.pdiag[j] + 1, ik = ij + 1;
            const int jend = m.prowptr[j + 1];
            while (ik < iend && jk < jend) {
                if (m.pcol[ik] == m.pcol[jk]) { m.pair_jk.push_back(m.ppos[jk]); m.pair_ik.push_back(m.ppos[ik]); ++ik; ++jk; }
                else if (m.pcol[ik] < m.pcol[jk]) ++ik;
                else ++jk;
            }
            m.pair_ptr[(size_t)m.ppos[ij] + 1] = (int)m.pair_jk.size();
        }
    }
    // level sets of the permuted lower triangle
    std::vector<int> lev(N, 0);
    int nlev = 0;
    for (int q = 0; q < N; ++q) {
        int l = 0;
        for (int k = m.Lrowptr[q]; k < m.Lrowptr[q + 1]; ++k) l = std::max(l, lev[m.Lcol[k]] + 1);
        lev[q] = l; nlev = std::max(nlev, l + 1);
    }
    m.lvl_ptr.assign((size_t)nlev + 1, 0);
    for (int q = 0; q < N; ++q) ++m.lvl_ptr[(size_t)lev[q] + 1];
    for (int l = 0; l < nlev; ++l) m.lvl_ptr[l + 1] += m.lvl_ptr[l];
    m.lvl_rows.resize(N);
    std::vector<int> f(m.lvl_ptr.begin(), m.lvl_ptr.end() - 1);
    for (int q = 0; q < N; ++q) m.lvl_rows[f[lev[q]]++] = q;
    if (lines) {
        // what the line kernels rely on, checked for EVERY row (the grid inference is only a heuristic): the
        // one same-colour lower block is the row's own column one plane down, the one same-colour upper
        // block its own column one plane up; being the highest lower / lowest upper column they are the
        // last blocks the sweeps visit
        for (int q = 0; q < N; ++q) {
            const int c = q >= m.ord.base[1] && m.ord.ncolours == 2 ? 1 : 0;
            const int lo = m.ord.base[c], hi = c == 0 && m.ord.ncolours == 2 ? m.ord.base[1] : N, nc = m.ord.ncols[c];
            for (int k = m.Lrowptr[q]; k < m.Lrowptr[q + 1]; ++k)
                if (m.Lcol[k] >= lo && m.Lcol[k] < hi && m.Lcol[k] != q - nc) return false;
            for (int k = m.Urowptr[q]; k < m.Urowptr[q + 1]; ++k)
                if (m.Ucol[k] >= lo && m.Ucol[k] < hi && m.Ucol[k] != q + nc) return false;
        }
    }
    return true;
}

}  // namespace opmgpu

using namespace opmgpu;

// ---- test hook (host only, no GPU): the multicolour program interpreted sequentially -------------------
// Gathers A into [ L | Dinv | U ] through psrc / ppos, factorises row by row with the update lists and
// applies v = w P^T U^-1 L^-1 P d walking the operands exactly as the kernels index them (L ascending, U in
// stored = descending order, the inverted pivot last).  Same arithmetic as mc_ilu.cuh (one fma per y -= a x,
// OPM's 3x3 adjugate inverse), so the result must be bit-identical to the oracle run on P A P^T.
// lu_out (may be NULL): factors in the slots of the caller's pattern.  Returns 0, -2 when the k-line
// ordering refuses the pattern, 1 + row for a singular pivot (caller's numbering); info = {colours, levels}.
namespace {
void host_mat3_mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s = std::fma(A[i * 3 + k], B[k * 3 + j], s);
            C[i * 3 + j] = s;
        }
}
double host_mat3_invert(double* M)
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

extern "C" int opmgpu_debug_host_mc_apply(int N, const int* rowptr, const int* colidx, const double* vals, int lines,
                                          double w, const double* d, double* v, double* lu_out, int* info /*[2]*/)
{
    McProgram m;
    if (!build_mc_program(N, rowptr, colidx, m, lines != 0)) return -2;
    const int nnzb = rowptr[N];
    std::vector<double> uni((size_t)m.total_blocks * 9, 0.0);
    for (int b = 0; b < nnzb; ++b)
        for (int t = 0; t < 9; ++t) uni[(size_t)m.ppos[b] * 9 + t] = vals[(size_t)m.psrc[b] * 9 + t];
    if (info) { info[0] = m.ord.ncolours; info[1] = (int)m.lvl_ptr.size() - 1; }
    // factorisation in level order (what the device launches), rows of a level in ascending order
    for (int q : m.lvl_rows) {
        for (int l = m.Lrowptr[q]; l < m.Lrowptr[q + 1]; ++l) {
            const int j = m.Lcol[l];
            double L[9];
            host_mat3_mul(&uni[(size_t)l * 9], &uni[(size_t)(m.offD + j) * 9], L);
            for (int t = 0; t < 9; ++t) uni[(size_t)l * 9 + t] = L[t];
            for (int p = m.pair_ptr[l]; p < m.pair_ptr[l + 1]; ++p) {
                double B[9];
                host_mat3_mul(L, &uni[(size_t)m.pair_jk[p] * 9], B);
                for (int t = 0; t < 9; ++t) uni[(size_t)m.pair_ik[p] * 9 + t] -= B[t];
            }
        }
        const double det = host_mat3_invert(&uni[(size_t)(m.offD + q) * 9]);
        if (!(det != 0.0) || std::isinf(det) || std::isnan(det)) return 1 + m.ord.p2n[q];
    }
    if (lu_out)
        for (int b = 0; b < nnzb; ++b)
            for (int t = 0; t < 9; ++t) lu_out[(size_t)m.psrc[b] * 9 + t] = uni[(size_t)m.ppos[b] * 9 + t];
    if (!d || !v) return 0;
    const int scale = std::fabs(w - 1.0) > 1e-15 ? 1 : 0;
    std::vector<double> W((size_t)N * 3);
    for (int q = 0; q < N; ++q) {
        double rb[3];
        for (int r = 0; r < 3; ++r) rb[r] = d[(size_t)m.ord.p2n[q] * 3 + r];
        for (int l = m.Lrowptr[q]; l < m.Lrowptr[q + 1]; ++l) {
            const double* a = &uni[(size_t)l * 9];
            const double* y = &W[(size_t)m.Lcol[l] * 3];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) rb[r] = std::fma(-a[r * 3 + c], y[c], rb[r]);
        }
        for (int r = 0; r < 3; ++r) W[(size_t)q * 3 + r] = rb[r];
    }
    for (int q = N - 1; q >= 0; --q) {
        double rb[3];
        for (int r = 0; r < 3; ++r) rb[r] = W[(size_t)q * 3 + r];
        for (int u = m.Urowptr[q]; u < m.Urowptr[q + 1]; ++u) {          // stored in descending column order
            const double* a = &uni[(size_t)(m.offU + u) * 9];
            const double* x = &W[(size_t)m.Ucol[u] * 3];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) rb[r] = std::fma(-a[r * 3 + c], x[c], rb[r]);
        }
        const double* di = &uni[(size_t)(m.offD + q) * 9];
        for (int r = 0; r < 3; ++r) {
            double y = 0.0;
            for (int c = 0; c < 3; ++c) y = std::fma(di[r * 3 + c], rb[c], y);
            W[(size_t)q * 3 + r] = y;
        }
        for (int r = 0; r < 3; ++r) v[(size_t)m.ord.p2n[q] * 3 + r] = scale ? W[(size_t)q * 3 + r] * w : W[(size_t)q * 3 + r];
    }
    return 0;
}
