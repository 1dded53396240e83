// Multi-GPU beneath the C ABI ("driver model"): one host process, one
// worker thread per GPU, the OPM caller unaware of ranks.  The counterpart in the reference is the
// MPI branch inside ISTLSolver::solve (opm/autodiff/ISTLSolver.hpp:283-306), which the caller of
// NewtonIterationBlackoilInterface never sees either.
#pragma once
#include <string>

#include "../../include/opm_gpu_solver.h"

namespace opmgpu {

struct MultiSolver;
MultiSolver* multi_create(int ngpus, const int* device_ids, std::string& err);
void multi_destroy(MultiSolver* m);
int multi_set_pattern(MultiSolver* m, int N, int nnzb, const int* rowptr, const int* colidx, std::string& err);
int multi_solve_bcrs3(MultiSolver* m, const double* vals, const double* rhs, double* x, const opmgpu_params* prm,
                      opmgpu_result* res, std::string& err);
int multi_solve_from_csc_blocks(MultiSolver* m, int N, const opmgpu_csc blocks[9], const double scale[3],
                                const double* rhs_eqmajor, double* dx_varmajor, const opmgpu_params* prm,
                                opmgpu_result* res, std::string& err);
// single_precision != 0: every GPU runs the reference's float instance
int multi_set_precision(MultiSolver* m, int single_precision, std::string& err);
// partition facts for tests / reports: axis (0 = i, 1 = j, 2 = k, -1 = contiguous row blocks), rows per GPU
int multi_partition_info(MultiSolver* m, int* axis, long long* offsets /*[ngpus+1]*/);

}  // namespace opmgpu
