// How fast can ONE SM stream from HBM through a TMA-fed shared-memory ring, as a function of how
// many SMs stream at the same time?  (The ILU0 sweeps keep only about a third of the SMs busy at
// any time, so the per-SM rate, not the chip rate, bounds a step.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smbw smbw.cu ; run: ./smbw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int STAGES>
__global__ void __launch_bounds__(128, 1) stream_kernel(const unsigned char* src, size_t bytes_per_cta, int stage_bytes, unsigned long long* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem);
    unsigned char* stages = smem + 128;
    const unsigned char* mine = src + (size_t)blockIdx.x * bytes_per_cta;
    const int nsteps = (int)(bytes_per_cta / stage_bytes);
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long acc = 0;
    if (threadIdx.x == 0) {
        int issued = 0;
        for (; issued < STAGES && issued < nsteps; ++issued) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[issued])), "r"(stage_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(stages + (size_t)issued * stage_bytes)), "l"(mine + (size_t)issued * stage_bytes), "r"(stage_bytes), "r"(s32(&full[issued])) : "memory");
        }
        for (int i = 0; i < nsteps; ++i) {
            const int st = i % STAGES;
            const unsigned par = (i / STAGES) & 1;
            unsigned ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&full[st])), "r"(par) : "memory");
            acc += *reinterpret_cast<unsigned long long*>(stages + (size_t)st * stage_bytes);
            if (issued < nsteps) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(stage_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(stages + (size_t)st * stage_bytes)), "l"(mine + (size_t)issued * stage_bytes), "r"(stage_bytes), "r"(s32(&full[st])) : "memory");
                ++issued;
            }
        }
        sink[blockIdx.x] = acc;
    }
}

int main()
{
    const int stage_bytes = 24 * 1024, S = 8;
    const size_t per_cta = (size_t)stage_bytes * 1024;       // 24 MB per CTA
    unsigned char* src; unsigned long long* sink;
    cudaMalloc(&src, per_cta * 148); cudaMemset(src, 1, per_cta * 148); cudaMalloc(&sink, 148 * 8);
    const size_t smem = 128 + (size_t)S * stage_bytes;
    cudaFuncSetAttribute(stream_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int ctas : {1, 8, 24, 48, 74, 100, 148}) {
        stream_kernel<8><<<ctas, 128, smem>>>(src, per_cta, stage_bytes, sink);
        cudaEventRecord(a);
        stream_kernel<8><<<ctas, 128, smem>>>(src, per_cta, stage_bytes, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%3d streaming SMs: %7.1f GB/s per SM, %7.1f GB/s total (%s)\n", ctas, per_cta / ms / 1e6, per_cta * ctas / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
