// Micro-benchmark: latency of one warp-to-warp hand-over through a SELF-VALIDATING shared-memory
// entry (3 doubles, all-ones = empty; consumer polls the data itself and re-arms it, producer
// checks the entry is empty before writing).  A token travels round a ring of NW warps of one
// CTA; cycles per hop are reported for the same CTA, for two CTAs of a cluster (the entry lives
// in the consumer's shared memory and is written with st.shared::cluster) and for two CTAs on
// different SMs through L2 slots.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ bool valid3(double a, double b, double c)
{
    return __double_as_longlong(a) != -1LL && __double_as_longlong(b) != -1LL && __double_as_longlong(c) != -1LL;
}

__global__ void k_ring(long long* out, int hops, int nw)
{
    __shared__ double ent[32][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) { ent[threadIdx.x][0] = ent[threadIdx.x][1] = ent[threadIdx.x][2] = __longlong_as_double(-1LL); }
    __syncthreads();
    volatile double* mine = ent[warp];
    volatile double* next = ent[(warp + 1) % nw];
    long long t0 = clock64();
    if (warp == 0 && lane == 0) { next[0] = 1.0; next[1] = 2.0; next[2] = 3.0; }
    for (int h = 0; h < hops; ++h) {
        double a = 0, b = 0, c = 0;
        if (lane == 0) {
            do { a = mine[0]; b = mine[1]; c = mine[2]; } while (!valid3(a, b, c));
            mine[0] = mine[1] = mine[2] = __longlong_as_double(-1LL);
        }
        a = __shfl_sync(0xffffffffu, a, 0);
        if (lane == 0) {
            while (valid3(next[0], next[1], next[2])) { }
            next[0] = a + 1.0; next[1] = b; next[2] = c;
        }
        __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / ((long long)hops * nw);
}

// two CTAs of a cluster bounce a token through each other's shared memory
__global__ void __cluster_dims__(2, 1, 1) k_dsmem(long long* out, int hops)
{
    __shared__ double ent[4];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    if (threadIdx.x == 0) ent[0] = ent[1] = ent[2] = __longlong_as_double(-1LL);
    cl.sync();
    volatile double* mine = ent;
    volatile double* peer = (volatile double*)cl.map_shared_rank(ent, rank ^ 1);
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        if (rank == 0) { peer[0] = 1.0; peer[1] = 2.0; peer[2] = 3.0; }
        for (int h = 0; h < hops; ++h) {
            double a, b, c;
            do { a = mine[0]; b = mine[1]; c = mine[2]; } while (!valid3(a, b, c));
            mine[0] = mine[1] = mine[2] = __longlong_as_double(-1LL);
            peer[0] = a + 1.0; peer[1] = b; peer[2] = c;
        }
    }
    long long t1 = clock64();
    cl.sync();
    if (threadIdx.x == 0 && rank == 0) out[0] = (t1 - t0) / (2LL * hops);
}

// two CTAs on different SMs bounce a token through L2 slots (st.cg / ld.relaxed.gpu)
__global__ void k_l2(long long* out, double* slots, int hops)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    long long* mine = reinterpret_cast<long long*>(slots + 16 * me);
    double* peer = slots + 16 * (me ^ 1);
    long long t0 = clock64();
    if (me == 0) { __stcg(peer, 1.0); __stcg(peer + 1, 2.0); __stcg(peer + 2, 3.0); }
    for (int h = 0; h < hops; ++h) {
        long long a, b, c;
        do {
            asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(a) : "l"(mine) : "memory");
            asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(b) : "l"(mine + 1) : "memory");
            asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(c) : "l"(mine + 2) : "memory");
        } while (a == -1 || b == -1 || c == -1);
        __stcg(mine, -1LL); __stcg(mine + 1, -1LL); __stcg(mine + 2, -1LL);
        __stcg(peer, __longlong_as_double(a) + 1.0); __stcg(peer + 1, __longlong_as_double(b)); __stcg(peer + 2, __longlong_as_double(c));
    }
    long long t1 = clock64();
    if (me == 0) out[0] = (t1 - t0) / (2LL * hops);
}

int main()
{
    long long* out; double* slots;
    cudaMalloc(&out, 64); cudaMalloc(&slots, 4096);
    long long h;
    for (int nw : {2, 4, 8}) {
        k_ring<<<1, nw * 32>>>(out, 2000, nw); cudaDeviceSynchronize();
        k_ring<<<1, nw * 32>>>(out, 2000, nw); cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("smem hand-over, %d warps: %lld cycles per hop [%s]\n", nw, h, cudaGetErrorString(cudaGetLastError()));
    }
    k_dsmem<<<2, 32>>>(out, 2000); cudaDeviceSynchronize();
    k_dsmem<<<2, 32>>>(out, 2000); cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("dsmem hand-over (cluster of 2): %lld cycles per hop [%s]\n", h, cudaGetErrorString(cudaGetLastError()));
    cudaMemset(slots, 0xff, 4096);
    k_l2<<<2, 32>>>(out, slots, 2000); cudaDeviceSynchronize();
    cudaMemset(slots, 0xff, 4096);
    k_l2<<<2, 32>>>(out, slots, 2000); cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("L2 hand-over (2 CTAs): %lld cycles per hop [%s]\n", h, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
