// See NewtonIterationBlackoilGPU.hpp.  Host logic only; every number of the linear solve is
// produced by libopmgpu.so on the GPU.
#include "NewtonIterationBlackoilGPU.hpp"

#include <sstream>

#include <algorithm>
#include <cmath>

namespace Opm {

double SparseCSC::coeff(int r, int c) const
{
    for (int k = colptr[c]; k < colptr[c + 1]; ++k)
        if (rowidx[k] == r) return val[k];
    return 0.0;
}

namespace {

// dense helpers for the (small) well blocks: D is (nw*np)^2 or nw^2
std::vector<double> toDense(const SparseCSC& m)
{
    std::vector<double> d((size_t)m.rows * m.cols, 0.0);
    for (int c = 0; c < m.cols; ++c)
        for (int k = m.colptr[c]; k < m.colptr[c + 1]; ++k) d[(size_t)m.rowidx[k] * m.cols + c] = m.val[k];
    return d;
}

// inverse by Gauss-Jordan with partial pivoting (the reference solves D u = I with SparseLU)
std::vector<double> invertDense(std::vector<double> a, int n)
{
    std::vector<double> inv((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)p * n + c])) p = r;
        if (a[(size_t)p * n + c] == 0.0) throw NumericalIssue("eliminateVariable: singular well block");
        if (p != c) for (int k = 0; k < n; ++k) { std::swap(a[(size_t)p * n + k], a[(size_t)c * n + k]); std::swap(inv[(size_t)p * n + k], inv[(size_t)c * n + k]); }
        const double d = 1.0 / a[(size_t)c * n + c];
        for (int k = 0; k < n; ++k) { a[(size_t)c * n + k] *= d; inv[(size_t)c * n + k] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int k = 0; k < n; ++k) { a[(size_t)r * n + k] -= f * a[(size_t)c * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k]; }
        }
    }
    return inv;
}

// J - B * (Di * C), all sparse in / sparse out (column-major, rows ascending, exact zeros kept out)
SparseCSC schurBlock(const SparseCSC& J, const SparseCSC& B, const std::vector<double>& Di, int nd, const SparseCSC& Cm)
{
    SparseCSC out = SparseCSC::zero(J.rows, J.cols);
    std::vector<double> u(nd), acc(J.rows, 0.0);
    std::vector<char> mark(J.rows, 0);
    std::vector<int> touched;
    for (int c = 0; c < J.cols; ++c) {
        touched.clear();
        for (int k = J.colptr[c]; k < J.colptr[c + 1]; ++k) {
            const int r = J.rowidx[k];
            acc[r] = J.val[k]; mark[r] = 1; touched.push_back(r);
        }
        if (Cm.colptr[c + 1] > Cm.colptr[c]) {
            std::fill(u.begin(), u.end(), 0.0);                         // u = Di * C(:,c)
            for (int k = Cm.colptr[c]; k < Cm.colptr[c + 1]; ++k)
                for (int r = 0; r < nd; ++r) u[r] += Di[(size_t)r * nd + Cm.rowidx[k]] * Cm.val[k];
            for (int d = 0; d < nd; ++d) {
                if (u[d] == 0.0) continue;
                for (int k = B.colptr[d]; k < B.colptr[d + 1]; ++k) {
                    const int r = B.rowidx[k];
                    if (!mark[r]) { mark[r] = 1; acc[r] = 0.0; touched.push_back(r); }
                    acc[r] -= B.val[k] * u[d];
                }
            }
        }
        std::sort(touched.begin(), touched.end());
        for (int r : touched) { out.rowidx.push_back(r); out.val.push_back(acc[r]); mark[r] = 0; }
        out.colptr[c + 1] = (int)out.rowidx.size();
    }
    return out;
}

}  // namespace

// Schur complement of (A B; C D) wrt D: A - B inv(D) C, rhs b_i - B inv(D) b_n
std::vector<ADB> eliminateVariable(const std::vector<ADB>& eqs, const int n)
{
    const int num_eq = (int)eqs.size();
    if (num_eq != (int)eqs[0].jac.size())
        throw std::logic_error("eliminateVariable() requires the same number of variables and equations.");
    if (n >= num_eq) throw std::logic_error("Trying to eliminate variable from too small set of equations.");
    const std::vector<SparseCSC>& Jn = eqs[n].jac;
    const int nd = Jn[n].rows;
    const std::vector<double> Di = invertDense(toDense(Jn[n]), nd);
    std::vector<double> Dibn(nd, 0.0);
    for (int r = 0; r < nd; ++r)
        for (int k = 0; k < nd; ++k) Dibn[r] += Di[(size_t)r * nd + k] * eqs[n].val[k];
    std::vector<ADB> out;
    for (int eq = 0; eq < num_eq; ++eq) {
        if (eq == n) continue;
        const SparseCSC& B = eqs[eq].jac[n];
        ADB r;
        r.val = eqs[eq].val;
        for (int d = 0; d < nd; ++d)
            for (int k = B.colptr[d]; k < B.colptr[d + 1]; ++k) r.val[B.rowidx[k]] -= B.val[k] * Dibn[d];
        for (int var = 0; var < num_eq; ++var) {
            if (var == n) continue;
            r.jac.push_back(schurBlock(eqs[eq].jac[var], B, Di, nd, Jn[var]));
        }
        out.push_back(std::move(r));
    }
    return out;
}

// y = inv(D) (b - C x), spliced back at the eliminated offset
ADB::V recoverVariable(const ADB& equation, const ADB::V& partial_solution, const int n)
{
    const SparseCSC& D = equation.jac[n];
    const int nd = D.rows;
    std::vector<double> b = equation.val;
    int off = 0, start = 0;
    for (int v = 0; v < (int)equation.jac.size(); ++v) {
        if (v == n) continue;
        const SparseCSC& Cm = equation.jac[v];
        for (int c = 0; c < Cm.cols; ++c)
            for (int k = Cm.colptr[c]; k < Cm.colptr[c + 1]; ++k) b[Cm.rowidx[k]] -= Cm.val[k] * partial_solution[off + c];
        off += Cm.cols;
    }
    for (int i = 0; i < n; ++i) start += equation.jac[i].cols;
    const std::vector<double> Di = invertDense(toDense(D), nd);
    ADB::V sol;
    sol.reserve(partial_solution.size() + nd);
    sol.insert(sol.end(), partial_solution.begin(), partial_solution.begin() + start);
    for (int r = 0; r < nd; ++r) {
        double y = 0.0;
        for (int k = 0; k < nd; ++k) y += Di[(size_t)r * nd + k] * b[k];
        sol.push_back(y);
    }
    sol.insert(sol.end(), partial_solution.begin() + start, partial_solution.end());
    return sol;
}

NewtonIterationBlackoilGPU::NewtonIterationBlackoilGPU(const ParameterGroup& param, const std::any& parallelInformation, int device)
    : parallelInformation_(parallelInformation)
{
    // the keys FlowLinearSolverParameters reads (ISTLSolver.hpp:142,204-208,255-262,364)
    opmgpu_default_params(&parameters_);
    parameters_.linear_solver_reduction = param.getDefault("linear_solver_reduction", parameters_.linear_solver_reduction);
    parameters_.linear_solver_maxiter = param.getDefault("linear_solver_maxiter", parameters_.linear_solver_maxiter);
    parameters_.ilu_relaxation = param.getDefault("ilu_relaxation", parameters_.ilu_relaxation);
    parameters_.linear_solver_verbosity = param.getDefault("linear_solver_verbosity", parameters_.linear_solver_verbosity);
    parameters_.linear_solver_ignoreconvergencefailure = param.getDefault("linear_solver_ignoreconvergencefailure", false) ? 1 : 0;
    parameters_.require_full_sparsity_pattern = param.getDefault("require_full_sparsity_pattern", false) ? 1 : 0;
    parameters_.newton_use_gmres = param.getDefault("newton_use_gmres", false) ? 1 : 0;
    parameters_.linear_solver_restart = param.getDefault("linear_solver_restart", parameters_.linear_solver_restart);
    if (param.getDefault("linear_solver_use_amg", false) || param.getDefault("ilu_fillin_level", 0) != 0 ||
        param.getDefault("ilu_milu", std::string("ILU")) != "ILU")
        throw std::invalid_argument("solver_approach=gpu supports ILU0-preconditioned BiCGStab / restarted GMRes only "
                                    "(no AMG/CPR, no fill-in, no MILU)");
    // ilu_redblack (ISTLSolver.hpp:207-209): the ILU0 of a colour-sorted reordering.  Here: the multicolour
    // variant of the library (greedy natural-order colouring; the reference's Welsh-Powell colouring and
    // sphere reordering live in opm-simulators' GraphColoring.hpp, outside the reference tree, and are not
    // restated -- iteration counts are comparable with neither reference ordering)
    const bool redblack = param.getDefault("ilu_redblack", false);
    // gpu_devices=0,1,...: several GPUs of this process behind the same interface (the caller stays
    // unaware of the partition, like a caller of the reference's MPI-parallel ISTLSolver)
    std::vector<int> devs;
    {
        std::istringstream is(param.getDefault("gpu_devices", std::string()));
        std::string tok;
        while (std::getline(is, tok, ',')) if (!tok.empty()) devs.push_back(std::stoi(tok));
    }
    const int rc = devs.size() > 1 ? opmgpu_create_multi((int)devs.size(), devs.data(), &handle_)
                                   : opmgpu_create(devs.empty() ? device : devs[0], &handle_);
    if (rc != OPMGPU_OK)
        throw std::runtime_error(std::string("NewtonIterationBlackoilGPU: ") + opmgpu_last_error(nullptr));
    if (redblack && opmgpu_set_ilu_ordering(handle_, OPMGPU_ILU_MULTICOLOUR) != OPMGPU_OK) {
        const std::string msg = opmgpu_last_error(handle_);
        opmgpu_destroy(handle_);
        throw std::invalid_argument("NewtonIterationBlackoilGPU: ilu_redblack: " + msg);
    }
}

NewtonIterationBlackoilGPU::~NewtonIterationBlackoilGPU() { opmgpu_destroy(handle_); }

NewtonIterationBlackoilGPU::SolutionVector
NewtonIterationBlackoilGPU::computeNewtonIncrement(const LinearisedBlackoilResidual& residual) const
{
    const int np = (int)residual.material_balance_eq.size();
    // the reference's switch instantiates Impl<np,Scalar> for np = 2..6 (...Interleaved.cpp:467-487)
    if (np < 2 || np > 6) throw std::logic_error("NewtonIterationBlackoilGPU: np out of the reference's range 2..6");
    std::vector<ADB> eqs(residual.material_balance_eq.begin(), residual.material_balance_eq.end());
    const bool hasWells = residual.well_flux_eq.size() > 0;
    std::vector<ADB> elim_eqs;
    if (hasWells) {
        eqs.push_back(residual.well_flux_eq);
        eqs.push_back(residual.well_eq);
        elim_eqs.push_back(eqs[np]);
        eqs = eliminateVariable(eqs, np);       // well flux unknowns
        elim_eqs.push_back(eqs[np]);
        eqs = eliminateVariable(eqs, np);       // bhp unknowns
    }
    const int N = eqs[0].size();
    opmgpu_csc blocks[36];          // np <= 6
    for (int p1 = 0; p1 < np; ++p1)
        for (int p2 = 0; p2 < np; ++p2) {
            const SparseCSC& s = eqs[p1].jac[p2];
            blocks[p1 * np + p2] = opmgpu_csc{s.colptr.data(), s.rowidx.data(), s.val.data()};
        }
    std::vector<double> b;
    b.reserve((size_t)np * N);
    for (int p = 0; p < np; ++p) b.insert(b.end(), eqs[p].val.begin(), eqs[p].val.end());
    SolutionVector dx((size_t)np * N, 0.0);
    double scale[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
    for (int p = 0; p < np && p < (int)residual.matbalscale.size(); ++p) scale[p] = residual.matbalscale[p];
    // the reference's dispatcher (...Interleaved.cpp:467-487): the float instance when the residual asks
    // for it (restarted GMRES exists for the double instance only; that combination stays in double)
    if (opmgpu_set_precision(handle_, residual.singlePrecision && !parameters_.newton_use_gmres) != OPMGPU_OK)
        throw std::runtime_error(opmgpu_last_error(handle_));
    const int rc = opmgpu_solve_from_csc_blocks_np(handle_, N, np, blocks, scale, b.data(), dx.data(), &parameters_, &last_);
    iterations_ = last_.iterations;             // before any throw
    switch (rc) {
    case OPMGPU_OK: break;
    case OPMGPU_NOT_CONVERGED: throw LinearSolverProblem("Convergence failure for linear solver.");
    case OPMGPU_SINGULAR_BLOCK:
    case OPMGPU_BREAKDOWN: throw NumericalIssue(opmgpu_last_error(handle_));
    case OPMGPU_BAD_PATTERN:
    case OPMGPU_BAD_ARGUMENT: throw std::logic_error(opmgpu_last_error(handle_));
    default: throw std::runtime_error(opmgpu_last_error(handle_));
    }
    if (hasWells) {
        dx = recoverVariable(elim_eqs[1], dx, np);
        dx = recoverVariable(elim_eqs[0], dx, np);
    }
    return dx;
}

}  // namespace Opm
