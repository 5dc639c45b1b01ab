// Probe (GPU box): cost of an (almost) empty barrier-separated phase in a 512-thread CTA, with and without the fences
// the UMMA engine uses, with and without a cluster launch.
#include <cstdio>
#include "../ancde_b200/csrc/pipeline.cuh"
#include "../ancde_b200/csrc/umma.cuh"
using namespace ancde;
__global__ void __launch_bounds__(512, 1) k(long long* out, int mode, int iters)
{
    extern __shared__ uint4 buf[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int wq = warp & 3, wg = warp >> 2;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (wq < 2 && wg < 2) {
            if (mode & 8) buf[tid] = make_uint4(it, tid, 0, 0);
            if (mode & 1) fence_async_smem();
            if (mode & 2) { tc_fence_after(); tc_fence_before(); }
            if (mode & 4) __threadfence_block();
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
}
int main()
{
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int cluster = 0; cluster < 2; ++cluster)
        for (int mode : {0, 1, 2, 3, 4, 8, 9, 11}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(128); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 200 * 1024;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster ? 4 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k, d, mode, 1000);
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("cluster %d mode %2d (1 fence.proxy.async, 2 tcgen05 fences, 4 threadfence_block, 8 sts): %5lld cycles per phase\n", cluster, mode, h);
        }
    return 0;
}
