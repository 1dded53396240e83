// Micro-benchmarks of the latencies that bound one step of the pipelined ILU0 sweep on
// sm_100a: dependent LDS, dependent DFMA, named barrier over 8 warps, STS->LDS turnaround.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_lat(long long* out, double* sink, int iters)
{
    extern __shared__ double sm[];
    int* smi = reinterpret_cast<int*>(sm);
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += blockDim.x) smi[i] = (i * 17 + 1) & 4095;
    __syncthreads();
    long long t0, t1;
    // (a) dependent LDS chain (pointer chase), one warp
    if (tid < 32) {
        int p = tid;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) p = smi[p];
        t1 = clock64();
        if (tid == 0) out[0] = (t1 - t0) / iters;
        sink[tid] = p;
    }
    __syncthreads();
    // (b) dependent DFMA chain, one warp
    if (tid < 32) {
        double a = 1.0 + tid * 1e-9, x = 0.999999;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) a = fma(a, x, 1e-3);
        t1 = clock64();
        if (tid == 0) out[1] = (t1 - t0) / iters;
        sink[32 + tid] = a;
    }
    __syncthreads();
    // (c) barrier among 256 threads, back to back
    if (tid < 256) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) asm volatile("bar.sync 1, 256;" ::: "memory");
        t1 = clock64();
        if (tid == 0) out[2] = (t1 - t0) / iters;
    }
    __syncthreads();
    // (d) STS -> barrier -> LDS by another warp -> DFMA -> STS (one "step" skeleton), 256 threads
    if (tid < 256) {
        double v = tid;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            sm[2048 + tid] = v;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            double y = sm[2048 + ((tid + 33) & 255)];
            v = fma(v, 0.5, y);
        }
        t1 = clock64();
        if (tid == 0) out[3] = (t1 - t0) / iters;
        sink[64 + tid] = v;
    }
    __syncthreads();
    // (e) 9 dependent DFMA with operands from LDS.64 (stride 27 doubles per lane)
    if (tid < 32) {
        double acc = 0.0;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const double* a = sm + ((tid * 27 + i) & 1023);
#pragma unroll
            for (int k = 0; k < 9; ++k) acc = fma(a[k], 1.0000001, acc);
        }
        t1 = clock64();
        if (tid == 0) out[4] = (t1 - t0) / iters;
        sink[400 + tid] = acc;
    }
    // (f) bar.red.or
    __syncthreads();
    if (tid < 256) {
        unsigned r = 0;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, 1, 256, p;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(r) : "r"(r) : "memory");
        }
        t1 = clock64();
        if (tid == 0) out[5] = (t1 - t0) / iters;
        sink[700 + tid] = r;
    }
}

__global__ void k_step(long long* out, double* sink, double* g, int iters, int mode)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    double v = tid;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        sm[2048 + tid] = v;
        if (mode == 1) g[(size_t)((tid * 977 + i * 131) & 0xfffff) * 3] = v;                 // plain STG, scattered
        if (mode == 2) __stcg(g + (size_t)((tid * 977 + i * 131) & 0xfffff) * 3, v);         // st.cg
        if (mode == 3) { g[(size_t)((tid * 977 + i * 131) & 0xfffff) * 3] = v; __threadfence_block(); }
        if (mode == 4) { double q = *(volatile double*)(g + (size_t)((tid * 977 + i * 131) & 0xfffff) * 3); v += q * 1e-300; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        double y = sm[2048 + ((tid + 33) & 255)];
        v = fma(v, 0.5, y);
    }
    long long t1 = clock64();
    if (tid == 0) out[mode] = (t1 - t0) / iters;
    sink[tid] = v;
}

int main()
{
    long long* out; double* sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 4096 * 8);
    cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        k_lat<<<1, 320, 200 * 1024>>>(out, sink, 2000);
        cudaDeviceSynchronize();
    }
    long long h[8];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("dependent LDS %lld cyc | dependent DFMA %lld cyc | bar.sync(256) %lld cyc | STS-bar-LDS-DFMA step %lld cyc | 9 LDS.64+DFMA chain %lld cyc | bar.red.or %lld cyc\n",
           h[0], h[1], h[2], h[3], h[4], h[5]);
    double* g; cudaMalloc(&g, (size_t)(1 << 20) * 3 * 8 + 64);
    cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode = 0; mode < 5; ++mode) { k_step<<<1, 256, 200 * 1024>>>(out, sink, g, 2000, mode); cudaDeviceSynchronize(); }
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("step skeleton: plain %lld | +STG %lld | +st.cg %lld | +STG+fence_block %lld | +volatile LDG %lld cyc\n", h[0], h[1], h[2], h[3], h[4]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
