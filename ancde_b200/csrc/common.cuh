// Shared helpers for the sm_100a kernels of libancde_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ancde_b200.h"

namespace ancde {

void set_error(const char* fmt, ...);
void count_launch();
void reset_launch_count();
// Optional per-kernel timing (CUDA events on the launching stream), see ancde_profile_* in the header.
void prof_begin(const char* name, cudaStream_t stream);
void prof_end(cudaStream_t stream);
long long* debug_buffer();

#define ANCDE_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            ancde::set_error(__VA_ARGS__);         \
            return ANCDE_EINVAL;                   \
        }                                          \
    } while (0)

#define ANCDE_CUDA(call)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ancde::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                             __LINE__);                                                        \
            return ANCDE_ECUDA;                                                                \
        }                                                                                      \
    } while (0)

// IEEE single-rounding arithmetic without FMA contraction, to follow the reference's
// op-by-op CPU arithmetic (each torch elementwise op rounds once).
template <typename T> struct Ops;
template <> struct Ops<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Ops<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace ancde
