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

// ---- per-kernel timing: event pairs recorded around every launch while profiling is enabled
constexpr int PROF_MAX = 8192;
struct ProfEntry { const char* name; cudaEvent_t e0, e1; };
static ProfEntry g_prof[PROF_MAX];
static int g_prof_n = 0, g_prof_made = 0;
static bool g_prof_on = false;

void prof_begin(const char* name, cudaStream_t stream)
{
    if (!g_prof_on || g_prof_n >= PROF_MAX) return;
    ProfEntry& e = g_prof[g_prof_n];
    if (g_prof_n >= g_prof_made) {
        cudaEventCreate(&e.e0);
        cudaEventCreate(&e.e1);
        g_prof_made = g_prof_n + 1;
    }
    e.name = name;
    cudaEventRecord(e.e0, stream);
}
void prof_end(cudaStream_t stream)
{
    if (!g_prof_on || g_prof_n >= PROF_MAX) return;
    cudaEventRecord(g_prof[g_prof_n].e1, stream);
    ++g_prof_n;
}

static long long* g_dbg = nullptr;
long long* debug_buffer() { return g_dbg; }

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

extern "C" int ancde_abi_version(void) { return 7; }

// development aid: device buffer of 64 int64 counters for -DANCDE_PHASE_TIMING builds
extern "C" int ancde_debug_set_buffer(long long* dev_ptr) { ancde::g_dbg = dev_ptr; return ANCDE_OK; }

extern "C" int ancde_last_error(char* buf, size_t len)
{
    if (buf && len) {
        strncpy(buf, ancde::g_err, len - 1);
        buf[len - 1] = 0;
    }
    return (int)strlen(ancde::g_err);
}

extern "C" int ancde_last_launch_count(void) { return ancde::g_launches; }

extern "C" int ancde_profile_enable(int on)
{
    ancde::g_prof_on = on != 0;
    ancde::g_prof_n = 0;
    return ANCDE_OK;
}

extern "C" int ancde_profile_count(void) { return ancde::g_prof_n; }

extern "C" int ancde_profile_get(int i, char* name, size_t len, float* ms)
{
    ANCDE_CHECK_ARG(i >= 0 && i < ancde::g_prof_n && name && ms, "profile_get: bad index %d", i);
    strncpy(name, ancde::g_prof[i].name, len - 1);
    name[len - 1] = 0;
    ANCDE_CUDA(cudaEventElapsedTime(ms, ancde::g_prof[i].e0, ancde::g_prof[i].e1));
    return ANCDE_OK;
}

extern "C" int ancde_fp32_peak_kernel(float* scratch, int iters, double* flops, void* stream)
{
    ANCDE_CHECK_ARG(scratch && iters > 0, "fp32_peak: bad arguments");
    const int blocks = 148 * 8, threads = 256;
    ancde::prof_begin("fp32_peak", (cudaStream_t)stream);
    ancde::fp32_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters);
    ancde::prof_end((cudaStream_t)stream);
    ANCDE_CUDA(cudaGetLastError());
    if (flops) *flops = 2.0 * 64.0 * (double)iters * blocks * threads;
    return ANCDE_OK;
}
