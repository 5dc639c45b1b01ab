// Microbenchmark: register-tiled 8x4 outer-product FMA loop (the output-layer loop of ncde_fwd_kernel) with scalar FFMA
// against packed FFMA2 (fma.rn.f32x2, sm_100), operands from shared memory as in the kernel.  Prints cycles per k-step.
#include <cstdio>
#include <cuda_runtime.h>

template <bool PACKED>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* cyc)
{
    __shared__ __align__(16) float sw[64 * 4 * 16];
    __shared__ __align__(16) float sx[64 * 8];
    for (int i = threadIdx.x; i < 64 * 4 * 16; i += blockDim.x) sw[i] = 1.0f + 1e-6f * i;
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) sx[i] = 1e-3f * i;
    __syncthreads();
    float acc[8][4];
    float2 acc2[4][4];
    for (int b = 0; b < 8; ++b) for (int c = 0; c < 4; ++c) acc[b][c] = 0.f;
    for (int q = 0; q < 4; ++q) for (int c = 0; c < 4; ++c) acc2[q][c] = make_float2(0.f, 0.f);
    const int lane16 = threadIdx.x & 15;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int kk = 0; kk < 64; ++kk) {
            const float4 w = *reinterpret_cast<const float4*>(sw + (kk * 16 + lane16) * 4);
            const float4 x0 = *reinterpret_cast<const float4*>(sx + kk * 8);
            const float4 x1 = *reinterpret_cast<const float4*>(sx + kk * 8 + 4);
            if (PACKED) {
                const float2 v[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y), make_float2(x1.z, x1.w)};
                const float2 wx = make_float2(w.x, w.x), wy = make_float2(w.y, w.y), wz = make_float2(w.z, w.z), ww = make_float2(w.w, w.w);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc2[q][0] = __ffma2_rn(v[q], wx, acc2[q][0]);
                    acc2[q][1] = __ffma2_rn(v[q], wy, acc2[q][1]);
                    acc2[q][2] = __ffma2_rn(v[q], wz, acc2[q][2]);
                    acc2[q][3] = __ffma2_rn(v[q], ww, acc2[q][3]);
                }
            } else {
                const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    acc[b][0] = fmaf(v[b], w.x, acc[b][0]);
                    acc[b][1] = fmaf(v[b], w.y, acc[b][1]);
                    acc[b][2] = fmaf(v[b], w.z, acc[b][2]);
                    acc[b][3] = fmaf(v[b], w.w, acc[b][3]);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int b = 0; b < 8; ++b) for (int c = 0; c < 4; ++c) s += acc[b][c];
    for (int q = 0; q < 4; ++q) for (int c = 0; c < 4; ++c) s += acc2[q][c].x + acc2[q][c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    const int iters = 200;
    for (int threads = 128; threads <= 512; threads += 64) {
        long long c0, c1;
        k<false><<<148, threads>>>(out, iters, cyc); cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        k<true><<<148, threads>>>(out, iters, cyc); cudaMemcpy(&c1, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %3d (%.1f warps/SMSP): FFMA %.1f cycles per k-step, FFMA2 %.1f  (32 FMA per thread and k-step)\n", threads,
               threads / 128.0, (double)c0 / (iters * 64), (double)c1 / (iters * 64));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
