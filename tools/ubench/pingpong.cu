// Ping-pong hand-over cost between two groups of 8 warps on named barriers (sm_100a):
// group g runs steps s = g, g+2, ...: bar.sync(2-g) -> 9 LDS.64 -> 9 dependent DFMA -> STS -> bar.arrive(1+g)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(576) k_pp(long long* out, double* sink, int nsteps, int extra)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = 1.0 / (i + 1);
    __syncthreads();
    if (warp < 2) return;
    const int g = (warp - 2) / 8, t = tid - 64 - g * 256;
    double acc = t;
    long long t0 = clock64();
    for (int s = g; s < nsteps; s += 2) {
        // "static phase" stand-in: extra independent work off the critical path
        double e = 0;
        for (int k = 0; k < extra; ++k) e += sm[(t * 7 + k * 13 + s) & 4095];
        if (s > 0) asm volatile("bar.sync %0, 512;" ::"r"(2 - g) : "memory");
        const double* y = sm + 4096 + ((t * 3 + s * 5) & 2047);
        double a = acc;
#pragma unroll
        for (int q = 0; q < 9; ++q) a = fma(y[q], 0.999, a);
        sm[4096 + 2048 + t] = a + e * 1e-30;
        asm volatile("bar.arrive %0, 512;" ::"r"(1 + g) : "memory");
        acc = a;
    }
    if (((nsteps - 1) & 1) != g) asm volatile("bar.sync %0, 512;" ::"r"(2 - g) : "memory");
    long long t1 = clock64();
    if (t == 0 && g == 0) out[0] = (t1 - t0) / nsteps;
    sink[tid] = acc;
}
int main()
{
    long long* out; double* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 4096 * 8);
    cudaFuncSetAttribute(k_pp, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int extra : {0, 10, 30, 60}) {
        k_pp<<<1, 576, 100 * 1024>>>(out, sink, 4000, extra);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("ping-pong step (9 LDS + 9 DFMA chain + STS + hand-over), %2d off-path loads per own step: %lld cycles/step\n", extra, h);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
