// FP64 / LDS issue throughput per SM on sm_100a as a function of the number of warps.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(long long* out, double* sink, int iters)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    const double x = 0.999999;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, 1e-3); a1 = fma(a1, x, 1e-3); a2 = fma(a2, x, 1e-3); a3 = fma(a3, x, 1e-3);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockDim.x / 32] = (t1 - t0) * 100 / (iters * 4);       // centi-cycles per warp DFMA
    sink[threadIdx.x] = a0 + a1 + a2 + a3;
}
__global__ void k_dep_dfma(long long* out, double* sink, int iters)
{
    double a0 = threadIdx.x;
    const double x = 0.999999;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a0 = fma(a0, x, 1e-3); }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[32 + blockDim.x / 32] = (t1 - t0) * 100 / iters;
    sink[threadIdx.x] = a0;
}
__global__ void k_lds(long long* out, double* sink, int iters)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const double* p = sm + ((threadIdx.x * 3 + i * 7) & 2047);
        acc += p[0] + p[1] + p[2] + p[256] + p[257] + p[258];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[64 + blockDim.x / 32] = (t1 - t0) * 100 / (iters * 6);
    sink[threadIdx.x] = acc;
}
int main()
{
    long long* out; double* sink;
    cudaMalloc(&out, 128 * 8); cudaMalloc(&sink, 4096 * 8);
    long long h[128];
    for (int w : {1, 2, 4, 8, 16}) {
        k_dfma<<<1, 32 * w>>>(out, sink, 4000);
        k_dep_dfma<<<1, 32 * w>>>(out, sink, 4000);
        k_lds<<<1, 32 * w, 64 * 1024>>>(out, sink, 2000);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        printf("warps %2d: indep DFMA %.2f cyc/warp-instr | dependent DFMA %.2f | LDS.64 %.2f cyc/warp-instr\n",
               w, h[w] / 100.0, h[32 + w] / 100.0, h[64 + w] / 100.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
