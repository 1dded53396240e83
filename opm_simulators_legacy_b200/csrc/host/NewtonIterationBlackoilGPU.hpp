// NewtonIterationBlackoilGPU -- host-side C++ mirror of the reference's linear-solver boundary
// over the C ABI of include/opm_gpu_solver.h.
//
// Interface mirrored: Opm::NewtonIterationBlackoilInterface
//   (opm/autodiff/NewtonIterationBlackoilInterface.hpp:31-52) with the constructor of
//   NewtonIterationBlackoilInterleaved (opm/autodiff/NewtonIterationBlackoilInterleaved.hpp:55-56).
// Body mirrored: NewtonIterationBlackoilInterleavedImpl<3,double>::computeNewtonIncrement
//   (opm/autodiff/NewtonIterationBlackoilInterleaved.cpp:202-292): well Schur elimination on the
//   host (NewtonIterationUtilities.cpp:45-128), everything from the matbal scaling to the
//   de-interleave in ONE call into the GPU library, recovery on the host (:134-184).
//
// This image has neither OPM, Eigen nor Boost, so the file carries minimal stand-ins for the
// reference's input types (same members, same meaning: column-major compressed Jacobian
// blocks like Eigen::SparseMatrix<double>, std::any for boost::any).  INTEGRATION.md shows the
// three lines that change when it is compiled inside opm-simulators-legacy against the real
// headers.  Error contract: non-convergence -> LinearSolverProblem (ISTLSolver.hpp:358-368),
// singular ILU0 pivot / BiCGStab breakdown -> NumericalIssue, CUDA/NCCL -> std::runtime_error;
// iterations() is valid also on the exception path (BlackoilModelBase_impl.hpp:291-297).
#pragma once
#include <any>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/opm_gpu_solver.h"

namespace Opm {

// ---- stand-ins for reference types ---------------------------------------------------------
/// Eigen::SparseMatrix<double> (column-major): outerIndexPtr / innerIndexPtr / valuePtr.
struct SparseCSC {
    int rows = 0, cols = 0;
    std::vector<int> colptr{0};
    std::vector<int> rowidx;
    std::vector<double> val;
    static SparseCSC zero(int r, int c) { SparseCSC m; m.rows = r; m.cols = c; m.colptr.assign(c + 1, 0); return m; }
    double coeff(int r, int c) const;
};

/// AutoDiffBlock<double>: value() and derivative()[block] (opm/autodiff/AutoDiffBlock.hpp:99,458-461)
struct ADB {
    typedef std::vector<double> V;
    V val;
    std::vector<SparseCSC> jac;
    const V& value() const { return val; }
    const std::vector<SparseCSC>& derivative() const { return jac; }
    int size() const { return (int)val.size(); }
};

/// opm/autodiff/LinearisedBlackoilResidual.hpp:47-72
struct LinearisedBlackoilResidual {
    std::vector<ADB> material_balance_eq;
    ADB well_flux_eq;
    ADB well_eq;
    std::vector<double> matbalscale;
    bool singlePrecision = false;
};

/// Opm::ParameterGroup::getDefault(key, default) over key=value strings.
class ParameterGroup {
public:
    ParameterGroup() {}
    explicit ParameterGroup(const std::map<std::string, std::string>& kv) : kv_(kv) {}
    void insertParameter(const std::string& k, const std::string& v) { kv_[k] = v; }
    template <class T> T getDefault(const std::string& key, const T& dflt) const
    {
        auto it = kv_.find(key);
        if (it == kv_.end()) return dflt;
        std::istringstream is(it->second);
        T v; is >> std::boolalpha >> v;
        if (is.fail()) { std::istringstream is2(it->second); is2 >> v; }
        return v;
    }
private:
    std::map<std::string, std::string> kv_;
};

struct LinearSolverProblem : std::runtime_error { using std::runtime_error::runtime_error; };
struct NumericalIssue : std::runtime_error { using std::runtime_error::runtime_error; };

/// opm/autodiff/NewtonIterationBlackoilInterface.hpp:31-52
class NewtonIterationBlackoilInterface {
public:
    typedef ADB::V SolutionVector;
    virtual ~NewtonIterationBlackoilInterface() {}
    virtual SolutionVector computeNewtonIncrement(const LinearisedBlackoilResidual& residual) const = 0;
    virtual int iterations() const = 0;
    virtual const std::any& parallelInformation() const = 0;
};

// host-side well elimination / recovery, as NewtonIterationUtilities.cpp:45-184
std::vector<ADB> eliminateVariable(const std::vector<ADB>& eqs, int n);
ADB::V recoverVariable(const ADB& equation, const ADB::V& partial_solution, int n);

// ---- the drop-in ----------------------------------------------------------------------------
/// solver_approach=gpu (opm/autodiff/FlowMain.hpp:806-830).
class NewtonIterationBlackoilGPU : public NewtonIterationBlackoilInterface {
public:
    explicit NewtonIterationBlackoilGPU(const ParameterGroup& param, const std::any& parallelInformation = std::any(),
                                        int device = 0);
    ~NewtonIterationBlackoilGPU() override;
    NewtonIterationBlackoilGPU(const NewtonIterationBlackoilGPU&) = delete;
    NewtonIterationBlackoilGPU& operator=(const NewtonIterationBlackoilGPU&) = delete;

    SolutionVector computeNewtonIncrement(const LinearisedBlackoilResidual& residual) const override;
    int iterations() const override { return iterations_; }
    const std::any& parallelInformation() const override { return parallelInformation_; }
    const opmgpu_result& lastResult() const { return last_; }

private:
    opmgpu_params parameters_;
    std::any parallelInformation_;
    mutable opmgpu_handle handle_ = nullptr;     // like the reference's mutable Impl cache (.hpp:75-79)
    mutable int iterations_ = 0;
    mutable opmgpu_result last_{};
};

}  // namespace Opm
