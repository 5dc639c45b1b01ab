// Probe (GPU box): issue/completion time of one "layer" of products: n1 MMAs with N1 columns then n2 with N2, same D.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../ancde_b200/csrc/pipeline.cuh"
#include "../ancde_b200/csrc/umma.cuh"
using namespace ancde;

__global__ void __launch_bounds__(512, 1) rate(long long* out, int N1, int n1, int N2, int n2, uint32_t aoff, uint32_t boff, int dcol, int spin_warps, int sbo)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // halfs of 1.0
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    const uint32_t id1 = umma_idesc_f16(128, N1), id2 = umma_idesc_f16(128, N2);
    const uint64_t ad0 = umma_smem_desc(smem_u32(smem) + aoff, 128, sbo);
    const uint64_t bd0 = umma_smem_desc(smem_u32(smem) + boff, 128, sbo);
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    uint32_t parity = 0;
    for (int rep = 0; rep < 4; ++rep) {
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                t0 = clock64();
                tc_fence_after();
#pragma unroll 4
                for (int i = 0; i < n1; ++i) umma_f16(tbase + dcol, ad0 + (uint64_t)(i * 16), bd0 + (uint64_t)(i * 16), id1, i != 0);
#pragma unroll 4
                for (int i = 0; i < n2; ++i) umma_f16(tbase + dcol, ad0 + (uint64_t)(i * 16), bd0 + (uint64_t)(i * 16), id2, true);
                t1 = clock64();
                umma_commit(&bar);
                t2 = clock64();
            }
            __syncwarp();
        }
        if (warp < spin_warps) {
            mbar_wait(&bar, parity);
            tc_fence_after();
            if (tid == 0) t3 = clock64();
        }
        parity ^= 1;
    }
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main()
{
    long long* d; cudaMalloc(&d, 32);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int N1, n1, N2, n2; unsigned aoff, boff; int dcol, spin, sbo; const char* what; };
    const Cfg cfgs[] = {
        {32, 12, 32, 0, 0, 16384, 0, 1, 1024, "12 x N32 (first engine version)"},
        {64, 4, 32, 4, 0, 16384, 0, 1, 1024, "4 x N64 + 4 x N32"},
        {64, 4, 32, 4, 0, 16384, 256, 1, 1024, "same, D column 256"},
        {64, 4, 32, 4, 0, 190000 / 1024 * 1024, 256, 1, 1024, "same, B operand at 185 KB"},
        {64, 4, 32, 4, 0, 190000 / 1024 * 1024, 256, 4, 1024, "same, 4 warps waiting"},
        {64, 4, 32, 4, 0, 190000 / 1024 * 1024, 256, 16, 1024, "same, 16 warps waiting"},
        {64, 4, 64, 0, 0, 16384, 0, 1, 1024, "4 x N64"},
        {32, 4, 32, 0, 0, 16384, 0, 1, 1024, "4 x N32"},
        {64, 2, 32, 2, 0, 16384, 0, 1, 512, "2 x N64 + 2 x N32, sbo 512 (bottom hidden)"},
        {128, 4, 64, 4, 0, 16384, 0, 1, 1024, "4 x N128 + 4 x N64 (CS=8)"},
    };
    for (const Cfg& c : cfgs) {
        rate<<<1, 512, 200 * 1024>>>(d, c.N1, c.n1, c.N2, c.n2, c.aoff, c.boff, c.dcol, c.spin, c.sbo);
        long long h[3];
        cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("%-46s issue %5lld  commit %4lld  commit->woken %5lld   total %5lld\n", c.what, h[0], h[1], h[2], h[0] + h[1] + h[2]);
    }
    return 0;
}
