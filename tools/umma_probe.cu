// Probe (GPU box): checks the tcgen05 plumbing of csrc/umma.cuh -- descriptors, TMEM layout, commit/mbarrier, tcgen05.ld --
// against a CPU product.  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../ancde_b200/csrc/pipeline.cuh"
#include "../ancde_b200/csrc/umma.cuh"
using namespace ancde;

// A: [rowsA][K] halfs, B: [N][K] halfs (global, row-major); D out: [128][N] floats
__global__ void probe(const __half* A, const __half* B, float* D, int rowsA, int N, int K, int lbo, int ntiles, int passes, int bmn, int amn)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kc = K / 8;
    const uint32_t sboA = kc * lbo, sboB = kc * lbo;
    unsigned char* sA = smem;
    unsigned char* sB = smem + (size_t)(rowsA / 8) * sboA;
    for (int idx = tid; idx < rowsA * K; idx += blockDim.x) {
        int r = idx / K, k = idx % K;
        if (amn) *reinterpret_cast<__half*>(sA + (r / 8) * sboA + (k / 8) * lbo + (k % 8) * 16 + (r % 8) * 2) = A[idx];   // MN-major
        else *reinterpret_cast<__half*>(sA + (r / 8) * sboA + (k / 8) * lbo + (r % 8) * 16 + (k % 8) * 2) = A[idx];
    }
    for (int idx = tid; idx < N * K; idx += blockDim.x) {
        int r = idx / K, k = idx % K;
        if (bmn) *reinterpret_cast<__half*>(sB + (r / 8) * sboB + (k / 8) * lbo + (k % 8) * 16 + (r % 8) * 2) = B[idx];   // MN-major: 8 rows contiguous
        else *reinterpret_cast<__half*>(sB + (r / 8) * sboB + (k / 8) * lbo + (r % 8) * 16 + (k % 8) * 2) = B[idx];
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    const uint32_t idesc = umma_idesc_f16(128, N) | (bmn ? (1u << 16) : 0u) | (amn ? (1u << 15) : 0u);
    if (tid == 0) {
        for (int t = 0; t < ntiles; ++t)
            for (int ps = 0; ps < passes; ++ps)
                for (int ks = 0; ks < K / 16; ++ks) {
                    uint64_t ad = umma_smem_desc(smem_u32(sA) + t * 16 * sboA + ks * 2 * lbo, lbo, sboA);
                    uint64_t bd = umma_smem_desc(smem_u32(sB) + ks * 2 * lbo, lbo, sboB);
                    umma_f16(tbase + t * N, ad, bd, idesc, (ps | ks) != 0);
                }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int t = 0; t < ntiles; ++t)
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tmem_ld16(tbase + ((uint32_t)(32 * (warp & 3)) << 16) + t * N + c0, v);
            for (int e = 0; e < 16; ++e) D[((size_t)t * 128 + tid) * N + c0 + e] = v[e];
        }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int run(int rowsA, int N, int K, int lbo, int ntiles, int passes, int bmn = 0, int amn = 0, int sbo_extra = 0)
{
    std::vector<__half> hA((size_t)rowsA * K), hB((size_t)N * K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    for (size_t i = 0; i < hA.size(); ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(x); fA[i] = __half2float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(x); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, (size_t)ntiles * 128 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, (size_t)ntiles * 128 * N * 4);
    size_t smem = (size_t)(rowsA / 8 + N / 8) * (K / 8) * lbo + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, rowsA, N, K, lbo, ntiles, passes, bmn, amn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rows %d N %d K %d lbo %d: CUDA error %s\n", rowsA, N, K, lbo, cudaGetErrorString(e)); return 1; }
    std::vector<float> hD((size_t)ntiles * 128 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int t = 0; t < ntiles; ++t)
        for (int m = 0; m < 128; ++m) {
            int r = t * 128 + m;
            if (r >= rowsA) continue;
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)fA[(size_t)r * K + k] * fB[(size_t)n * K + k];
                s *= passes;
                maxerr = fmax(maxerr, fabs(s - hD[((size_t)t * 128 + m) * N + n]));
            }
        }
    printf("amn %d bmn %d rows %d N %d K %d lbo %d tiles %d passes %d: max abs err %.3g %s\n", amn, bmn, rowsA, N, K, lbo, ntiles, passes, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr < 1e-3 ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad += run(128, 32, 64, 128, 1, 1);
    bad += run(128, 16, 32, 128, 1, 1);
    bad += run(256, 32, 64, 128, 2, 3);
    bad += run(128, 64, 64, 144, 1, 1);
    bad += run(384, 16, 16, 128, 3, 2);
    bad += run(128, 32, 64, 128, 1, 1, 1);   // B operand MN-major
    bad += run(256, 64, 48, 128, 2, 2, 1);
    bad += run(128, 16, 32, 128, 1, 1, 1);
    bad += run(128, 64, 32, 128, 1, 1, 0, 1);   // A operand MN-major (the weight-gradient kernel)
    bad += run(128, 128, 32, 128, 1, 2, 0, 1);
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
