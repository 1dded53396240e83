// f3  Restarted GMRES kernels (Dune::RestartedGMResSolver::apply, selected by newton_use_gmres with
// restart = linear_solver_restart: opm/autodiff/ISTLSolver.hpp:257-265).  The Krylov loop itself
// is host-driven (solver.cu: gmres()); these are its vector kernels.  Modified Gram-Schmidt is a
// chain of dependent (dot, axpy) pairs: one kernel does "w -= H[k] v_k" and the NEXT dot product
// in the same pass over w, so an Arnoldi step with i+1 orthogonalisations costs i+2 passes instead
// of 2i+3.  Dot products use the deterministic grid reduction of kernels.cuh.
#pragma once
#include "kernels.cuh"

namespace opmgpu {

// H[slot] = a . b
__global__ void __launch_bounds__(256)
gmres_dot_kernel(size_t n, const double* __restrict__ a, const double* __restrict__ b, double* H, int slot, ReduceWs ws)
{
    double v[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v[0] = fma(a[i], b[i], v[0]);
    grid_reduce<1>(v, ws, [=](double (&t)[1]) { H[slot] = t[0]; });
}

// w.axpy(-H[k], v_k), then H[k+1] = next . w   (next == nullptr: w . w, the square of the new column's last entry)
__global__ void __launch_bounds__(256)
gmres_mgs_kernel(size_t n, double* __restrict__ w, const double* __restrict__ vk, const double* __restrict__ next,
                 double* H, int k, ReduceWs ws)
{
    const double h = H[k];
    double v[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double wi = fma(-h, vk[i], w[i]);
        w[i] = wi;
        v[0] = fma(next ? next[i] : wi, wi, v[0]);
    }
    grid_reduce<1>(v, ws, [=](double (&t)[1]) { H[k + 1] = t[0]; });
}

// dst = src * alpha   (v[i+1] = w; v[i+1] *= 1/H[i+1][i]   and   v[0] *= 1/norm)
__global__ void __launch_bounds__(256)
gmres_scale_kernel(size_t n, double* dst, const double* src, double alpha)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i] * alpha;
}

// y.axpy(a, x)
__global__ void __launch_bounds__(256)
gmres_axpy_kernel(size_t n, double* __restrict__ y, double a, const double* __restrict__ x)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = fma(a, x[i], y[i]);
}

// x += w
__global__ void __launch_bounds__(256)
gmres_add_kernel(size_t n, double* __restrict__ x, const double* __restrict__ w)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] += w[i];
}

// b -= A x   (MatrixAdapter::applyscaleadd(-1, x, b) -> BCRSMatrix::usmv: block by block, ascending columns)
__global__ void __launch_bounds__(256)
residual3_kernel(int N, const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
                 const double* __restrict__ x, double* __restrict__ b)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)N * 3) return;
    const int row = (int)(t / 3), c = (int)(t - (size_t)row * 3);
    double r = b[t];
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
        const double* a = vals + (size_t)k * 9 + c * 3;
        const double* xj = x + (size_t)colidx[k] * 3;
        r = fma(-a[0], xj[0], r); r = fma(-a[1], xj[1], r); r = fma(-a[2], xj[2], r);
    }
    b[t] = r;
}

}  // namespace opmgpu
