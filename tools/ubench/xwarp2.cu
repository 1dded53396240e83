// Micro-benchmarks behind the column-owned sweep design (sm_100a):
//  1. dependent DFMA / SHFL / LDS latency of a lone warp;
//  2. warp-to-warp hand-over through a self-validating shared-memory entry with
//     (a) one polling lane per warp and no other lanes, (b) all lanes polling the same entry
//     (warp-uniform loop), (c) per-lane entries with a warp vote (the sweep kernel's scheme).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ bool valid3(double a, double b, double c)
{
    return __double_as_longlong(a) != -1LL && __double_as_longlong(b) != -1LL && __double_as_longlong(c) != -1LL;
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds(uint32_t a) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts(uint32_t a, double v) { asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

__global__ void k_lat(long long* out, double* sink, double c0)
{
    __shared__ double sm[64];
    double x = c0 + threadIdx.x, y = 1.0000001;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
        x = fma(x, y, c0); x = fma(x, y, c0); x = fma(x, y, c0); x = fma(x, y, c0);
        x = fma(x, y, c0); x = fma(x, y, c0); x = fma(x, y, c0); x = fma(x, y, c0);
    }
    long long t1 = clock64();
    double z = x;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
        z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31); z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31);
        z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31); z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31);
    }
    long long t2 = clock64();
    sm[threadIdx.x] = z; sm[threadIdx.x + 32] = 0.0;
    __syncwarp();
    uint32_t a = s32(sm) + 8 * threadIdx.x;
    double u = z;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
        sts(a, u); u = lds(a ^ 8) + 0.0 * u;   // store -> load of the neighbour's slot -> dependent use
        sts(a, u); u = lds(a ^ 8) + 0.0 * u;
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / 8; out[1] = (t2 - t1) / 4; out[2] = (t3 - t2) / 2; }
    sink[threadIdx.x] = x + z + u;
}

// (a) one thread per warp
__global__ void k_ring_a(long long* out, int hops, int nw)
{
    __shared__ double ent[32][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) { ent[threadIdx.x][0] = ent[threadIdx.x][1] = ent[threadIdx.x][2] = __longlong_as_double(-1LL); }
    __syncthreads();
    if (lane != 0) return;
    const uint32_t mine = s32(ent[warp]), next = s32(ent[(warp + 1) % nw]);
    const double E = __longlong_as_double(-1LL);
    long long t0 = clock64();
    if (warp == 0) { sts(next, 1.0); sts(next + 8, 2.0); sts(next + 16, 3.0); }
    for (int h = 0; h < hops; ++h) {
        double a, b, c;
        do { a = lds(mine); b = lds(mine + 8); c = lds(mine + 16); } while (!valid3(a, b, c));
        sts(mine, E); sts(mine + 8, E); sts(mine + 16, E);
        sts(next, a + 1.0); sts(next + 8, b); sts(next + 16, c);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / ((long long)hops * nw);
}
// (b) all lanes poll the same entry; lane 0 writes
__global__ void k_ring_b(long long* out, int hops, int nw)
{
    __shared__ double ent[32][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) { ent[threadIdx.x][0] = ent[threadIdx.x][1] = ent[threadIdx.x][2] = __longlong_as_double(-1LL); }
    __syncthreads();
    const uint32_t mine = s32(ent[warp]), next = s32(ent[(warp + 1) % nw]);
    const double E = __longlong_as_double(-1LL);
    long long t0 = clock64();
    if (warp == 0 && lane == 0) { sts(next, 1.0); sts(next + 8, 2.0); sts(next + 16, 3.0); }
    for (int h = 0; h < hops; ++h) {
        double a, b, c;
        do { a = lds(mine); b = lds(mine + 8); c = lds(mine + 16); } while (!valid3(a, b, c));
        __syncwarp();
        if (lane == 0) {
            sts(mine, E); sts(mine + 8, E); sts(mine + 16, E);
            sts(next, a + 1.0); sts(next + 8, b); sts(next + 16, c);
        }
        __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / ((long long)hops * nw);
}
// (c) per-lane entries (4 "edge" lanes per warp), warp vote decides; every edge lane forwards its own token
__global__ void k_ring_c(long long* out, int hops, int nw)
{
    __shared__ double ent[32][4][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 4 * 4; i += blockDim.x) (&ent[0][0][0])[i] = __longlong_as_double(-1LL);
    __syncthreads();
    const bool edge = lane < 4;
    const uint32_t mine = s32(ent[warp][lane & 3]), next = s32(ent[(warp + 1) % nw][lane & 3]);
    const double E = __longlong_as_double(-1LL);
    long long t0 = clock64();
    if (warp == 0 && edge) { sts(next, 1.0); sts(next + 8, 2.0); sts(next + 16, 3.0); }
    for (int h = 0; h < hops; ++h) {
        double a = 0, b = 0, c = 0;
        bool need = edge;
        for (;;) {
            if (need) { a = lds(mine); b = lds(mine + 8); c = lds(mine + 16); need = !valid3(a, b, c); }
            if (!__any_sync(0xffffffffu, need)) break;
        }
        if (edge) {
            sts(mine, E); sts(mine + 8, E); sts(mine + 16, E);
            sts(next, a + 1.0); sts(next + 8, b); sts(next + 16, c);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / ((long long)hops * nw);
}

int main()
{
    long long* out; double* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
    long long h[4];
    k_lat<<<1, 32>>>(out, sink, 0.5); cudaDeviceSynchronize();
    k_lat<<<1, 32>>>(out, sink, 0.5); cudaDeviceSynchronize();
    cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    printf("dependent DFMA %lld cycles, SHFL.64 round %lld cycles, STS->LDS->use %lld cycles [%s]\n", h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
    for (int nw : {2, 4, 8}) {
        k_ring_a<<<1, nw * 32>>>(out, 2000, nw); cudaDeviceSynchronize();
        k_ring_a<<<1, nw * 32>>>(out, 2000, nw); cudaDeviceSynchronize();
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("(a) one thread per warp, %d warps: %lld cycles per hop\n", nw, h[0]);
        k_ring_b<<<1, nw * 32>>>(out, 2000, nw); cudaDeviceSynchronize();
        k_ring_b<<<1, nw * 32>>>(out, 2000, nw); cudaDeviceSynchronize();
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("(b) uniform poll, %d warps: %lld cycles per hop\n", nw, h[0]);
        k_ring_c<<<1, nw * 32>>>(out, 2000, nw); cudaDeviceSynchronize();
        k_ring_c<<<1, nw * 32>>>(out, 2000, nw); cudaDeviceSynchronize();
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("(c) per-lane entries + vote, %d warps: %lld cycles per hop [%s]\n", nw, h[0], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
