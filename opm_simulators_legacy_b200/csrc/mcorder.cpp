// Host analysis of the multicolour ILU0 variant (mcorder.hpp).  Once per pattern.
#include "mcorder.hpp"
#include "analysis.hpp"      // infer_cartesian_grid

#include <algorithm>
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
