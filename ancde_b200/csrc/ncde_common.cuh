// Device helpers shared by the forward and backward NCDE kernels.
#pragma once
#include "ncde_layout.cuh"

namespace ancde {

// ---- small vectors of samples: a CTA integrates TS samples, activations are stored
// [feature][TS] so one LDS.128 broadcasts 4 samples of one feature to every lane.
template <int VW> struct Vec;
template <> struct Vec<4> {
    static __device__ __forceinline__ void ld(const float* p, float* v) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<2> {
    static __device__ __forceinline__ void ld(const float* p, float* v) {
        float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void st(float* p, const float* v) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
};
template <> struct Vec<1> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = *p; }
    static __device__ __forceinline__ void st(float* p, const float* v) { *p = v[0]; }
};

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the MUFU pipe (ex2 + rcp): 5 issue slots, |abs err| < 4e-7
// (tanh.approx.f32 is ~5e-4 relative and fails the 1e-4 parity bar after hundreds of stages).
__device__ __forceinline__ float tanh_mufu(float x)
{
#ifdef ANCDE_TANH_LIBM
    return tanhf(x);
#else
    float e, r;
    float a = x * 2.885390081777927f;   // 2 * log2(e)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
    float d = e + 1.0f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return fmaf(-2.0f, r, 1.0f);
#endif
}

// Time of grid point k: arange(k) * step + t_start, last point clipped (torchdiffeq
// _grid_constructor_from_step_size), in float32 without contraction.
// An explicit grid (options['grid_constructor'], or rk4 without a step size: the grid is `t` itself) is read from memory.
__device__ __forceinline__ float grid_time(int k, int n_steps, float t_start, float t_end, float step, const float* grid)
{
    if (grid) return __ldg(grid + k);
    float t = __fadd_rn(__fmul_rn((float)k, step), t_start);
    if (k == n_steps && t > t_end) t = t_end;
    return t;
}

// Stage times of torchdiffeq's rk4_alt_step_func: t, t + dt/3, t + dt*2/3, t + dt.
__device__ __forceinline__ float stage_time(float t0, float dt, int s)
{
    if (s == 0) return t0;
    if (s == 1) return __fadd_rn(t0, __fdiv_rn(dt, 3.0f));
    if (s == 2) return __fadd_rn(t0, __fdiv_rn(__fmul_rn(dt, 2.0f), 3.0f));
    return __fadd_rn(t0, dt);
}

// Index of control channel j, sample b inside a dU block (see NcdeLayout::DUS).
__device__ __forceinline__ int du_index(int j, int b, int TS, int DUS) { return (j >> 2) * DUS + (j & 3) * TS + b; }

struct SolveArgs {
    NcdeLayout lo;
    float t_start, t_end, step;
    const float* grid;         // explicit grid points (n_steps + 1) or NULL
    const float* times;
    const float* t_out;
    const float* cb;
    const float* cc2;
    const float* cd3;
    const float* z0;
    const float* attention;
    const float* h_prime;
    const float* wpack;
    float* z_out;
    float* k_out;
    float* save_act;
    float* save_G;
    // backward
    const float* grad_z_out;
    const float* grad_k_out;   // single-stage mode: cotangent of the one stage derivative
    float* grad_z0;
    float* grad_attention;
    float* grad_hidden;     // packed like the hidden part of wpack (Wt layout), accumulated with atomics
    // launch-time tiling of the output-layer weight stream (filled by the launchers)
    int ntc;                // consumer threads per CTA (a TMA producer warp may sit behind them)
    int ring_kc;            // k rows per ring slot
    int ring_slots;         // ring depth
    int ring_pw;            // floats per row in a slot (columns of one pass)
    long long* dbg;         // phase timing buffer (only read when built with -DANCDE_PHASE_TIMING)
};

// Phase timing for kernel development: thread 0 of CTA 0 accumulates clock64() deltas per phase.
#ifdef ANCDE_PHASE_TIMING
#define PHASE_T0() long long pt__ = clock64()
#define PHASE_MARK(i)                                                      \
    do {                                                                   \
        if (threadIdx.x == 0 && blockIdx.x == 0 && p.dbg) {                \
            const long long now__ = clock64();                             \
            p.dbg[8 + (i)] += now__ - pt__;                                      \
            pt__ = now__;                                                  \
        }                                                                  \
    } while (0)
#else
#define PHASE_T0() do { } while (0)
#define PHASE_MARK(i) do { } while (0)
#endif

}  // namespace ancde
