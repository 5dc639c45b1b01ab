// Error reporting, launch accounting and the FP32 roofline microbenchmark of libancde_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ancde {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch() { ++g_launches; }
void reset_launch_count() { g_launches = 0; }

// 8 independent FFMA chains per thread: the issue-bound FP32 peak of the SM (128 FMA/clk).
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters)
{
    float a[8];
    const float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-3f * (blockIdx.x + 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (float)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chains alive
}

}  // namespace ancde

extern "C" int ancde_abi_version(void) { return 1; }

extern "C" int ancde_last_error(char* buf, size_t len)
{
    if (buf && len) {
        strncpy(buf, ancde::g_err, len - 1);
        buf[len - 1] = 0;
    }
    return (int)strlen(ancde::g_err);
}

extern "C" int ancde_last_launch_count(void) { return ancde::g_launches; }

extern "C" int ancde_fp32_peak_kernel(float* scratch, int iters, double* flops, void* stream)
{
    ANCDE_CHECK_ARG(scratch && iters > 0, "fp32_peak: bad arguments");
    const int blocks = 148 * 8, threads = 256;
    ancde::fp32_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters);
    ANCDE_CUDA(cudaGetLastError());
    if (flops) *flops = 2.0 * 64.0 * (double)iters * blocks * threads;
    return ANCDE_OK;
}
