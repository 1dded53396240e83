// Visibility latency of a store (no fence, producer keeps doing ALU work afterwards) to a
// polling thread on another SM, measured with %globaltimer.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
template <int MODE>
__global__ void k_vis(long long* buf, unsigned long long* tput, unsigned long long* tsee, int iters, int peer, int busy)
{
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        double x = 1.0;
        for (int i = 1; i <= iters; ++i) {
            for (int k = 0; k < 2000; ++k) x = x * 1.0000001 + 1e-9;      // gap between pushes
            tput[i] = gtime();
            if (MODE == 0) __stcg(buf, (long long)i);
            else if (MODE == 1) { __stcg(buf, (long long)i); __threadfence(); }
            else asm volatile("st.relaxed.gpu.global.s64 [%0], %1;" ::"l"(buf), "l"((long long)i) : "memory");
            for (int k = 0; k < busy; ++k) x = x * 1.0000001 + 1e-9;      // ALU only, no memory ops
        }
        buf[8] = (long long)x;
    } else if (blockIdx.x == peer) {
        for (int i = 1; i <= iters; ++i) {
            long long v;
            do { asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(buf) : "memory"); } while (v < i);
            tsee[i] = gtime();
        }
    }
}
int main()
{
    long long* buf; unsigned long long *tput, *tsee;
    const int iters = 200;
    cudaMalloc(&buf, 4096); cudaMalloc(&tput, 8 * 256); cudaMalloc(&tsee, 8 * 256);
    unsigned long long hp[256], hs[256];
    for (int mode = 0; mode < 3; ++mode)
        for (int busy : {0, 20000}) {
            cudaMemset(buf, 0, 4096);
            if (mode == 0) k_vis<0><<<148, 32>>>(buf, tput, tsee, iters, 75, busy);
            if (mode == 1) k_vis<1><<<148, 32>>>(buf, tput, tsee, iters, 75, busy);
            if (mode == 2) k_vis<2><<<148, 32>>>(buf, tput, tsee, iters, 75, busy);
            cudaDeviceSynchronize();
            cudaMemcpy(hp, tput, sizeof hp, cudaMemcpyDeviceToHost); cudaMemcpy(hs, tsee, sizeof hs, cudaMemcpyDeviceToHost);
            double sum = 0, mx = 0; for (int i = 10; i <= iters; ++i) { double d = (double)hs[i] - (double)hp[i]; sum += d; if (d > mx) mx = d; }
            printf("mode %d (0 st.cg, 1 st.cg+fence, 2 st.relaxed.gpu) busy-after-store %5d: mean visibility %.0f ns, max %.0f ns\n", mode, busy, sum / (iters - 9), mx);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
