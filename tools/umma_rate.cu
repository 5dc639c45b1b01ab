// Probe (GPU box): issue cost and completion time of back-to-back tcgen05.mma (no-swizzle K-major FP16 operands).
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../ancde_b200/csrc/pipeline.cuh"
#include "../ancde_b200/csrc/umma.cuh"
using namespace ancde;

__global__ void rate(long long* out, int N, int nmma, int mode)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // halfs of 1.0
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint32_t sbo = 1024;
    long long t0 = 0, t1 = 0, t2 = 0;
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {
        if (tid == 0) {
            const uint64_t ad0 = umma_smem_desc(smem_u32(smem), 128, sbo);
            const uint64_t bd0 = umma_smem_desc(smem_u32(smem) + 16384, 128, sbo);
            t0 = clock64();
            if (mode == 0) {
#pragma unroll 4
                for (int i = 0; i < nmma; ++i) umma_f16(tbase, ad0 + (uint64_t)((i & 3) * 16), bd0 + (uint64_t)((i & 3) * 16), idesc, i != 0);
            } else {
                // descriptors rebuilt from scratch per MMA (what the first engine version did)
                for (int i = 0; i < nmma; ++i)
                    umma_f16(tbase, umma_smem_desc(smem_u32(smem) + (i & 3) * 256, 128, sbo), umma_smem_desc(smem_u32(smem) + 16384 + (i & 3) * 256, 128, sbo), idesc, i != 0);
            }
            t1 = clock64();
            umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after();
        if (tid == 0) t2 = clock64();
        __syncthreads();
    }
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main()
{
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int Ns[] = {16, 32, 64, 128, 256};
    for (int mode = 0; mode < 2; ++mode)
        for (int N : Ns)
            for (int nmma : {1, 4, 12, 48}) {
                rate<<<1, 128, 64 * 1024>>>(d, N, nmma, mode);
                long long h[2];
                cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("mode %d N %3d nmma %2d: issue %6lld cyc (%.1f/mma)  issue->complete %6lld cyc (%.1f/mma)\n", mode, N, nmma, h[0], (double)h[0] / nmma, h[1], (double)h[1] / nmma);
            }
    return 0;
}
