// One-way store -> poll latency between two SMs through L2 (sm_100a): ping-pong of a sequence
// number between CTA 0 and CTA k, round trip / 2, for several store / load flavours.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ void put(long long* p, long long v)
{
    if (MODE == 0) *(volatile long long*)p = v;
    else if (MODE == 1) __stcg(p, v);
    else asm volatile("st.relaxed.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <int MODE>
__device__ __forceinline__ long long get(const long long* p)
{
    long long v;
    if (MODE == 0) v = *(const volatile long long*)p;
    else if (MODE == 1) v = __ldcg(p);
    else asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE>
__global__ void k_pp(long long* buf, long long* out, int iters, int peer)
{
    long long* a = buf;            // written by CTA 0
    long long* b = buf + 64;       // written by the peer (different line)
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            put<MODE>(a, i);
            while (get<MODE>(b) != i) {}
        }
        out[MODE] = (clock64() - t0) / (2 * iters);
    } else if (blockIdx.x == peer) {
        for (int i = 1; i <= iters; ++i) {
            while (get<MODE>(a) != i) {}
            put<MODE>(b, i);
        }
    }
}
int main()
{
    long long *buf, *out;
    cudaMalloc(&buf, 4096); cudaMalloc(&out, 64);
    for (int peer : {1, 2, 75, 147}) {
        long long h[4];
        cudaMemset(buf, 0, 4096); k_pp<0><<<148, 32>>>(buf, out, 2000, peer); cudaDeviceSynchronize();
        cudaMemset(buf, 0, 4096); k_pp<1><<<148, 32>>>(buf, out, 2000, peer); cudaDeviceSynchronize();
        cudaMemset(buf, 0, 4096); k_pp<2><<<148, 32>>>(buf, out, 2000, peer); cudaDeviceSynchronize();
        cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
        printf("CTA 0 <-> CTA %3d one-way latency: volatile %lld | st.cg/ld.cg %lld | relaxed.gpu %lld cycles\n", peer, h[0], h[1], h[2]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
