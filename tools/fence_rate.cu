// Probe (GPU box): cost of fence.proxy.async.shared::cta / tcgen05 fences / __syncthreads after a few shared stores.
#include <cstdio>
#include "../ancde_b200/csrc/pipeline.cuh"
#include "../ancde_b200/csrc/umma.cuh"
using namespace ancde;
__global__ void __launch_bounds__(512, 1) k(long long* out, int mode)
{
    __shared__ uint4 buf[2048];
    const int tid = threadIdx.x;
    long long acc = 0;
    for (int rep = 0; rep < 8; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        buf[tid] = make_uint4(tid, rep, 0, 0);
        buf[tid + 512] = make_uint4(tid, rep, 1, 0);
        if (mode == 1 || mode == 3) fence_async_smem();
        if (mode == 2 || mode == 3) tc_fence_before();
        if (mode == 4) __threadfence_block();
        __syncthreads();
        if (mode == 2 || mode == 3) tc_fence_after();
        const uint4 x = buf[(tid + 33) & 511];
        acc += clock64() - t0 + (x.x == 12345678u);
    }
    if (tid == 0) out[0] = acc / 8;
}
int main()
{
    long long* d; cudaMalloc(&d, 8);
    const char* names[] = {"sts + syncthreads + lds", "+ fence.proxy.async", "+ tcgen05 fences", "+ both", "+ threadfence_block"};
    for (int mode = 0; mode < 5; ++mode) {
        k<<<1, 512>>>(d, mode);
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-28s %5lld cycles\n", names[mode], h);
    }
    return 0;
}
