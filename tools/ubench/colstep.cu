// Micro-benchmark of one step of a COLUMN-OWNED ILU0 sweep on sm_100a: every lane owns an (i,j)
// column of an 8x4 patch and walks k; the (i,j,k-1) result stays in registers, the (i-1,j,k) and
// (i,j-1,k) results come from neighbouring lanes of the same warp.  Measures cycles per step for
//   mode 0: neighbours through __shfl_sync (12 SHFL.32 per step)
//   mode 1: neighbours through a per-warp shared-memory row (3 STS.64, __syncwarp, 6 LDS.64)
//   mode 2: as 1 with 32-byte entries (STS.128 + STS.64, LDS.128 + LDS.64 per neighbour)
// with the 27 (lower) or 36 (upper) coefficients prefetched from a shared-memory stage one step
// ahead, for 1..8 warps per CTA (independent patches), 1 CTA per SM on all SMs.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, bool UPPER>
__global__ void __launch_bounds__(256, 1) k_colstep(long long* out, double* sink, int iters)
{
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NC = UPPER ? 39 : 30;                 // doubles per row of a step record
    constexpr int STAGES = 4;
    double* ring = sm + (size_t)warp * (STAGES * NC * 32 + 32 * 4);
    double* xch = ring + STAGES * NC * 32;              // exchange row of the warp
    for (int i = lane; i < STAGES * NC * 32; i += 32) ring[i] = 1e-3 * ((i * 7 + warp) % 13) - 5e-3;
    for (int i = lane; i < 128; i += 32) xch[i] = 0.0;
    __syncthreads();
    const int li = lane & 7, lj = lane >> 3;
    const int src_i = li > 0 ? lane - 1 : lane, src_j = lj > 0 ? lane - 8 : lane;
    double y0 = 1.0 + lane, y1 = 2.0, y2 = 3.0;
    double cf[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) cf[q] = ring[q * 32 + lane];
    long long t0 = clock64();
    for (int s = 0; s < iters; ++s) {
        // prefetch the next step's record (independent of the chain)
        double nf[NC];
        const double* nxt = ring + ((s + 1) % STAGES) * NC * 32;
#pragma unroll
        for (int q = 0; q < NC; ++q) nf[q] = nxt[q * 32 + lane];
        double a0 = cf[27], a1 = cf[28], a2 = cf[29];
        // block 0: own previous result (registers)
        a0 = fma(-cf[0], y0, a0); a1 = fma(-cf[9], y0, a1); a2 = fma(-cf[18], y0, a2);
        a0 = fma(-cf[1], y1, a0); a1 = fma(-cf[10], y1, a1); a2 = fma(-cf[19], y1, a2);
        a0 = fma(-cf[2], y2, a0); a1 = fma(-cf[11], y2, a1); a2 = fma(-cf[20], y2, a2);
        double j0, j1, j2, i0, i1, i2;
        if (MODE == 0) {
            j0 = __shfl_sync(0xffffffffu, y0, src_j); j1 = __shfl_sync(0xffffffffu, y1, src_j); j2 = __shfl_sync(0xffffffffu, y2, src_j);
            i0 = __shfl_sync(0xffffffffu, y0, src_i); i1 = __shfl_sync(0xffffffffu, y1, src_i); i2 = __shfl_sync(0xffffffffu, y2, src_i);
        } else if (MODE == 1) {
            xch[lane * 3] = y0; xch[lane * 3 + 1] = y1; xch[lane * 3 + 2] = y2;
            __syncwarp();
            j0 = xch[src_j * 3]; j1 = xch[src_j * 3 + 1]; j2 = xch[src_j * 3 + 2];
            i0 = xch[src_i * 3]; i1 = xch[src_i * 3 + 1]; i2 = xch[src_i * 3 + 2];
            __syncwarp();
        } else {
            *reinterpret_cast<double2*>(xch + lane * 4) = make_double2(y0, y1); xch[lane * 4 + 2] = y2;
            __syncwarp();
            const double2 ja = *reinterpret_cast<const double2*>(xch + src_j * 4); j0 = ja.x; j1 = ja.y; j2 = xch[src_j * 4 + 2];
            const double2 ia = *reinterpret_cast<const double2*>(xch + src_i * 4); i0 = ia.x; i1 = ia.y; i2 = xch[src_i * 4 + 2];
            __syncwarp();
        }
        a0 = fma(-cf[3], j0, a0); a1 = fma(-cf[12], j0, a1); a2 = fma(-cf[21], j0, a2);
        a0 = fma(-cf[4], j1, a0); a1 = fma(-cf[13], j1, a1); a2 = fma(-cf[22], j1, a2);
        a0 = fma(-cf[5], j2, a0); a1 = fma(-cf[14], j2, a1); a2 = fma(-cf[23], j2, a2);
        a0 = fma(-cf[6], i0, a0); a1 = fma(-cf[15], i0, a1); a2 = fma(-cf[24], i0, a2);
        a0 = fma(-cf[7], i1, a0); a1 = fma(-cf[16], i1, a1); a2 = fma(-cf[25], i1, a2);
        a0 = fma(-cf[8], i2, a0); a1 = fma(-cf[17], i2, a1); a2 = fma(-cf[26], i2, a2);
        if (UPPER) {
            double v0 = fma(cf[30], a0, 0.0), v1 = fma(cf[33], a0, 0.0), v2 = fma(cf[36], a0, 0.0);
            v0 = fma(cf[31], a1, v0); v1 = fma(cf[34], a1, v1); v2 = fma(cf[37], a1, v2);
            v0 = fma(cf[32], a2, v0); v1 = fma(cf[35], a2, v1); v2 = fma(cf[38], a2, v2);
            a0 = v0; a1 = v1; a2 = v2;
        }
        y0 = a0; y1 = a1; y2 = a2;
#pragma unroll
        for (int q = 0; q < NC; ++q) cf[q] = nf[q];
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[warp] = (t1 - t0) / iters;
    sink[blockIdx.x * 256 + tid] = y0 + y1 + y2;
}

template <int MODE, bool UPPER>
void run(const char* name, long long* out, double* sink)
{
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(k_colstep<MODE, UPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("%-28s", name);
    for (int nw : {1, 2, 3, 4, 6, 8}) {
        constexpr int NC = UPPER ? 39 : 30;
        if ((size_t)nw * (4 * NC * 32 + 128) * 8 > smem) { printf("  nw=%d: -", nw); continue; }
        for (int rep = 0; rep < 2; ++rep) { k_colstep<MODE, UPPER><<<148, nw * 32, smem>>>(out, sink, 4000); cudaDeviceSynchronize(); }
        long long h[8];
        cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        printf("  nw=%d: %lld cyc", nw, h[0]);
    }
    printf("  [%s]\n", cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    long long* out; double* sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 148 * 256 * 8);
    run<0, false>("lower, shuffles", out, sink);
    run<1, false>("lower, smem 24 B entries", out, sink);
    run<2, false>("lower, smem 32 B entries", out, sink);
    run<0, true>("upper, shuffles", out, sink);
    run<1, true>("upper, smem 24 B entries", out, sink);
    run<2, true>("upper, smem 32 B entries", out, sink);
    return 0;
}
