// Probe (GPU box): time for W warps to each read `nld` x 16 TMEM columns (tcgen05.ld 32x32b.x16), slowest warp.
#include <cstdio>
#include <cstdlib>
#include "../ancde_b200/csrc/pipeline.cuh"
#include "../ancde_b200/csrc/umma.cuh"
using namespace ancde;

__global__ void __launch_bounds__(512, 1) k(long long* out, int W, int nld, int same_cols)
{
    __shared__ uint32_t tslot;
    __shared__ long long tmax[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    float acc = 0.f;
    long long dt = 0;
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        if (warp < W) {
            const uint32_t tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
            for (int i = 0; i < nld; ++i) {
                float v[16];
                tmem_ld16(tl + (same_cols ? 0 : (warp >> 2) * 64) + 16 * (i & 3), v);
#pragma unroll
                for (int e = 0; e < 16; ++e) acc += v[e];
            }
        }
        dt = clock64() - t0;
    }
    if (lane == 0) tmax[warp] = (warp < W) ? dt : 0;
    if (acc == 123.f) out[5] = 1;
    __syncthreads();
    if (tid == 0) { long long m = 0; for (int w = 0; w < 16; ++w) m = tmax[w] > m ? tmax[w] : m; out[0] = m; out[1] = tmax[0]; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main()
{
    long long* d; cudaMalloc(&d, 64);
    for (int same = 0; same < 2; ++same)
        for (int W : {1, 2, 4, 8, 16})
            for (int nld : {1, 2, 4, 8}) {
                k<<<1, 512>>>(d, W, nld, same);
                long long h[2];
                if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
                printf("same_cols %d warps %2d loads/warp %d: slowest warp %5lld cyc, warp 0 %5lld cyc  (%.1f B/cyc)\n", same, W, nld, h[0], h[1], (double)W * nld * 32 * 64 / h[0]);
            }
    return 0;
}
