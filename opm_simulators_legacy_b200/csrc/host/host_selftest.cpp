// Self-test of the C++ host mirror: a 3-cell-wide synthetic residual with one well goes through
// NewtonIterationBlackoilGPU::computeNewtonIncrement and is checked against a dense solve of the
// full (cells + wells) system.  Needs a B200; run by tests/test_gpu_parity.py.
#include <cmath>
#include <cstdio>
#include <random>

#include "NewtonIterationBlackoilGPU.hpp"

using namespace Opm;

static SparseCSC fromDense(const std::vector<double>& d, int r, int c)
{
    SparseCSC m = SparseCSC::zero(r, c);
    for (int j = 0; j < c; ++j) {
        for (int i = 0; i < r; ++i)
            if (d[(size_t)i * c + j] != 0.0) { m.rowidx.push_back(i); m.val.push_back(d[(size_t)i * c + j]); }
        m.colptr[j + 1] = (int)m.rowidx.size();
    }
    return m;
}

// np material-balance equations (3: black oil; 2: a two-phase deck, the reference's Impl<2,Scalar>) + wells
static int run_case(const std::string& devices, const int np, bool redblack = false)
{
    const int N = 60, nw = 2, ne = np + 2;
    std::vector<int> sizes(ne, N);
    sizes[np] = nw * np; sizes[np + 1] = nw;
    std::vector<int> offs(ne + 1, 0);
    for (int i = 0; i < ne; ++i) offs[i + 1] = offs[i] + sizes[i];
    const int nt = offs[ne];
    std::mt19937 rng(7);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> A((size_t)nt * nt, 0.0), b(nt);
    for (int i = 0; i < nt; ++i) {
        for (int j = 0; j < nt; ++j) {
            const int ci = i < np * N ? i % N : -1, cj = j < np * N ? j % N : -1;
            const bool cellpair = ci >= 0 && cj >= 0 && std::abs(ci - cj) <= 1;      // 1-D stencil, all np*np blocks
            const bool wellcpl = (ci < 0 || cj < 0) && (ci < 0 ? (cj < 0 || cj % 17 == 0) : ci % 17 == 0);
            if (cellpair || wellcpl) A[(size_t)i * nt + j] = 0.2 * U(rng);
        }
        A[(size_t)i * nt + i] += 3.0;
        b[i] = U(rng);
    }
    LinearisedBlackoilResidual res;
    res.matbalscale = {1.1169, 1.0031, 0.0031};
    std::vector<ADB> all(ne);
    for (int e = 0; e < ne; ++e) {
        all[e].val.assign(b.begin() + offs[e], b.begin() + offs[e + 1]);
        for (int v = 0; v < ne; ++v) {
            std::vector<double> blk((size_t)sizes[e] * sizes[v]);
            for (int i = 0; i < sizes[e]; ++i)
                for (int j = 0; j < sizes[v]; ++j) blk[(size_t)i * sizes[v] + j] = A[(size_t)(offs[e] + i) * nt + offs[v] + j];
            all[e].jac.push_back(fromDense(blk, sizes[e], sizes[v]));
        }
    }
    res.material_balance_eq.assign(all.begin(), all.begin() + np);
    res.well_flux_eq = all[np];
    res.well_eq = all[np + 1];
    std::map<std::string, std::string> kv = {{"linear_solver_reduction", "1e-12"}, {"linear_solver_maxiter", "200"},
                                             {"require_full_sparsity_pattern", "true"}};
    if (!devices.empty()) kv["gpu_devices"] = devices;
    if (redblack) kv["ilu_redblack"] = "true";      // the multicolour variant behind the reference's key
    NewtonIterationBlackoilGPU solver{ParameterGroup(kv)};
    const auto dx = solver.computeNewtonIncrement(res);
    // dense reference: Gaussian elimination on the full system
    std::vector<double> M = A, x = b;
    for (int c = 0; c < nt; ++c) {
        int p = c;
        for (int r = c + 1; r < nt; ++r) if (std::fabs(M[(size_t)r * nt + c]) > std::fabs(M[(size_t)p * nt + c])) p = r;
        for (int k = 0; k < nt; ++k) std::swap(M[(size_t)p * nt + k], M[(size_t)c * nt + k]);
        std::swap(x[p], x[c]);
        for (int r = c + 1; r < nt; ++r) {
            const double f = M[(size_t)r * nt + c] / M[(size_t)c * nt + c];
            for (int k = c; k < nt; ++k) M[(size_t)r * nt + k] -= f * M[(size_t)c * nt + k];
            x[r] -= f * x[c];
        }
    }
    for (int r = nt - 1; r >= 0; --r) {
        for (int k = r + 1; k < nt; ++k) x[r] -= M[(size_t)r * nt + k] * x[k];
        x[r] /= M[(size_t)r * nt + r];
    }
    double err = 0.0, ref = 0.0;
    for (int i = 0; i < nt; ++i) { err = std::fmax(err, std::fabs(dx[i] - x[i])); ref = std::fmax(ref, std::fabs(x[i])); }
    std::printf("host_selftest[%s, np = %d%s]: size %zu iterations %d max_abs_err %.3e (ref %.3e)\n", devices.empty() ? "1 GPU" : devices.c_str(), np, redblack ? ", ilu_redblack" : "", dx.size(), solver.iterations(), err, ref);
    // error contract: not converged -> LinearSolverProblem, iterations still reported
    bool threw = false;
    std::map<std::string, std::string> kv2 = {{"linear_solver_reduction", "1e-14"}, {"linear_solver_maxiter", "1"},
                                              {"require_full_sparsity_pattern", "true"}};
    if (!devices.empty()) kv2["gpu_devices"] = devices;
    NewtonIterationBlackoilGPU s2{ParameterGroup(kv2)};
    try { s2.computeNewtonIncrement(res); } catch (const LinearSolverProblem&) { threw = true; }
    std::printf("host_selftest: LinearSolverProblem thrown %d, iterations() %d\n", (int)threw, s2.iterations());
    return (dx.size() == (size_t)nt && err <= 1e-8 * ref && threw && s2.iterations() == 1) ? 0 : 1;
}

// usage: host_selftest [gpu_devices, e.g. 0,1]   (several devices: the multi-GPU handle beneath the same class)
int main(int argc, char** argv)
{
    const std::string devices = argc > 1 ? argv[1] : "";
    int rc = run_case(devices, 3);
    if (devices.find(',') == std::string::npos) {      // single-GPU handles: the other block sizes and the multicolour variant
        rc |= run_case(devices, 2);
        rc |= run_case(devices, 4);
        rc |= run_case(devices, 6);
        rc |= run_case(devices, 3, true);
    }
    return rc;
}
