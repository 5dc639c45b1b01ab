// Microbenchmarks that size the v2 solver design on a B200: legacy mma.sync TF32 rate per SM, cluster barrier +
// DSMEM exchange latency, MUFU rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mma ubench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC, bool BF16>
__global__ void mma_rate(float* out, int iters, long long* cyc)
{
    float d[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 0x3f800000u + threadIdx.x + i;
    b[0] = 0x3f000000u + threadIdx.x; b[1] = 0x3e800000u;
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) { if (BF16) mma_bf16(d[j], a, b); else mma_tf32(d[j], a, b); }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += d[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// 3xTF32 with on-the-fly split of A (W stays fp32 in registers) and pre-split B
template <int NT_>
__global__ void mma3_rate(float* out, int iters, long long* cyc)
{
    float d[NT_][4];
    float w[4];
    uint32_t bh[NT_][2], bl[NT_][2];
    for (int i = 0; i < 4; ++i) w[i] = 1.0f + 0.001f * (threadIdx.x + i);
    for (int j = 0; j < NT_; ++j) { bh[j][0] = 0x3f000000u + threadIdx.x + j; bh[j][1] = 0x3e800000u; bl[j][0] = 0x3a000000u; bl[j][1] = 0x39000000u + j; }
    for (int j = 0; j < NT_; ++j) for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(ah[i]) : "f"(w[i]));
            al[i] = __float_as_uint(w[i] - __uint_as_float(ah[i]));
            w[i] += 1e-7f;
        }
#pragma unroll
        for (int j = 0; j < NT_; ++j) { mma_tf32(d[j], al, bh[j]); mma_tf32(d[j], ah, bl[j]); mma_tf32(d[j], ah, bh[j]); }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < NT_; ++j) for (int i = 0; i < 4; ++i) s += d[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void mufu_rate(float* out, int iters, long long* cyc)
{
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = 0.5f + 0.01f * (threadIdx.x + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// cluster exchange: every CTA writes `nfl` floats into each peer's shared memory, then cluster barrier; repeated
template <int CS>
__global__ void cluster_xchg(float* out, int iters, int nfl, long long* cyc)
{
    extern __shared__ float sm[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    for (int i = threadIdx.x; i < CS * nfl * 2; i += blockDim.x) sm[i] = 0.f;
    cl.sync();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float* base = sm + (it & 1) * CS * nfl;
        for (int i = threadIdx.x; i < nfl * CS; i += blockDim.x) {
            const int peer = i / nfl, e = i - peer * nfl;
            float* dst = cl.map_shared_rank(base + rank * nfl + e, peer);
            *dst = (float)(it + e);
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = threadIdx.x; i < CS * nfl * 2; i += blockDim.x) s += sm[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    cl.sync();
}

template <int CS>
static void run_cluster(float* out, long long* cyc, int nt, int nfl)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * (148 / CS));
    cfg.blockDim = dim3(nt);
    cfg.dynamicSmemBytes = CS * nfl * 2 * 4;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const int iters = 2000;
    cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_xchg<CS>, out, iters, nfl, cyc);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("cluster CS=%d nt=%d nfl=%d : %s  %.0f cyc per exchange+barrier\n", CS, nt, nfl, cudaGetErrorString(e), (double)h / iters);
}

int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4 * 4); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    auto report = [&](const char* name, int nwarps, int mma_per_iter, double macs_per_mma) {
        cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double per_sm = (double)nwarps * mma_per_iter * iters / (double)h;
        printf("%-28s warps/SM=%2d : %.3f mma/clk/SM = %.0f MAC/clk/SM  (err %s)\n", name, nwarps, per_sm, per_sm * macs_per_mma,
               cudaGetErrorString(cudaGetLastError()));
    };
    for (int nw : {4, 8, 16}) {
        mma_rate<8, false><<<148, nw * 32>>>(out, iters, cyc); report("tf32 m16n8k8 x8 indep", nw, 8, 1024);
        mma_rate<2, false><<<148, nw * 32>>>(out, iters, cyc); report("tf32 m16n8k8 x2 indep", nw, 2, 1024);
        mma_rate<8, true><<<148, nw * 32>>>(out, iters, cyc);  report("bf16 m16n8k16 x8 indep", nw, 8, 2048);
        mma3_rate<4><<<148, nw * 32>>>(out, iters, cyc);       report("3xtf32 split-A 4 n-tiles", nw, 12, 1024);
        mma3_rate<2><<<148, nw * 32>>>(out, iters, cyc);       report("3xtf32 split-A 2 n-tiles", nw, 6, 1024);
    }
    for (int nw : {4, 8, 16}) {
        mufu_rate<<<148, nw * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mufu ex2 warps/SM=%d : %.2f lanes/clk/SM\n", nw, (double)nw * 32 * 8 * iters / (double)h);
    }
    run_cluster<2>(out, cyc, 256, 512);
    run_cluster<4>(out, cyc, 256, 512);
    run_cluster<4>(out, cyc, 448, 400);
    run_cluster<8>(out, cyc, 256, 256);
    run_cluster<4>(out, cyc, 256, 32);
    return 0;
}
