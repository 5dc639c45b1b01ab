// Warp-level tensor-core (mma.sync) and thread-block-cluster helpers shared by the solver kernels:
// FP16 hi/lo operand splitting, ldmatrix / stmatrix wrappers, cluster rank / DSMEM addressing, asynchronous remote
// stores, and the register-resident phase timers used during kernel development.
#pragma once
#include <cuda_fp16.h>

#include "ncde_common.cuh"
#include "pipeline.cuh"

namespace ancde {

// Phase timing for kernel development (-DANCDE_PHASE_TIMING): clock64() deltas are accumulated in registers and
// written once at the end, so that the probe itself does not stall the warp on a global round trip per phase.
#ifdef ANCDE_PHASE_TIMING
#define PH2_DECL() long long ph2_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long ph2_t = 0
#define PH2_T0() ph2_t = clock64()
#define PH2_MARK(i) do { const long long now__ = clock64(); ph2_acc[i] += now__ - ph2_t; ph2_t = now__; } while (0)
#define PH2_FLUSH(base) do { if (threadIdx.x == 0 && blockIdx.x == 0 && p.dbg) { for (int i__ = 0; i__ < 8; ++i__) p.dbg[(base) + i__] += ph2_acc[i__]; } } while (0)
#else
#define PH2_DECL() do { } while (0)
#define PH2_T0() do { } while (0)
#define PH2_MARK(i) do { } while (0)
#define PH2_FLUSH(base) do { } while (0)
#endif

#ifdef __CUDACC__
// ---- warp-level tensor-core primitives (mma.sync: operands in registers, M = 16, N = 8): used where a CTA owns only
// 8 samples (hidden layers of the backward, hidden weight gradients) ------------------------------------------
__device__ __forceinline__ void mma16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
// d += (Ahi + Alo) * (Bhi + Blo) without the lo*lo term; small terms first.
__device__ __forceinline__ void mma16x3(float (&d)[4], const uint4& ah, const uint4& al, uint32_t bh0, uint32_t bh1, uint32_t bl0,
                                        uint32_t bl1)
{
    mma16(d, al.x, al.y, al.z, al.w, bh0, bh1);
    mma16(d, ah.x, ah.y, ah.z, ah.w, bl0, bl1);
    mma16(d, ah.x, ah.y, ah.z, ah.w, bh0, bh1);
}
// k8 tail: a = {hi a0, hi a1, lo a0, lo a1}
__device__ __forceinline__ void mma8x3(float (&d)[4], const uint4& a, uint32_t bh0, uint32_t bl0)
{
    mma8(d, a.z, a.w, bh0);
    mma8(d, a.x, a.y, bl0);
    mma8(d, a.x, a.y, bh0);
}

// FP32 pair -> packed FP16 hi and lo halves (x = hi + lo up to 2^-22 |x|)
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo)
{
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// B-operand (activations) buffers: two FP16 matrices (hi, lo) [n][KP] with k contiguous and KP = 16*KSB + 8 halfs
// (row stride = 16 bytes mod 128: ldmatrix / stmatrix rows fall into distinct bank groups).
// Address of this lane's row for the x4 access of (k16 step ks, n-tile nt): matrices {hi k0-7, hi k8-15, lo k0-7, lo k8-15}.
__device__ __forceinline__ uint32_t bmat_lane_addr(uint32_t hi_base, uint32_t lo_minus_hi, int KP, int lane)
{
    return hi_base + (uint32_t)(((lane & 7) * KP + ((lane >> 3) & 1) * 8) * 2) + ((lane >> 4) ? lo_minus_hi : 0u);
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// Stores an accumulator tile (rows = features k, columns = samples n) transposed, i.e. as rows n of the [n][k] matrices.
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
// Splits an accumulator fragment {(g,2t),(g,2t+1),(g+8,2t),(g+8,2t+1)} into hi/lo halves and stores it as the B operand
// of the next product.
__device__ __forceinline__ void store_act_tile(uint32_t lane_addr, const float (&v)[4])
{
    uint32_t h0, l0, h1, l1;
    split_pair(v[0], v[1], h0, l0);
    split_pair(v[2], v[3], h1, l1);
    stmatrix_x4_trans(lane_addr, h0, h1, l0, l1);
}

// ---- cluster plumbing -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank)
{
    uint32_t d;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(d) : "r"(smem_addr), "r"(rank));
    return d;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v)
{
    asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
// Asynchronous remote store: writes `v` into a peer CTA's shared memory and counts 4 bytes on the peer's mbarrier.
// Unlike st.shared::cluster + barrier.cluster (release/acquire over ALL prior writes of the thread, global stores
// included), the completion is tracked per datum, so the consumer only waits for the data it needs.
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n" ::"r"(remote_addr), "r"(__float_as_uint(v)),
                 "r"(remote_mbar)
                 : "memory");
}
// 16 bytes into a peer CTA's shared memory, counted on the peer's mbarrier.
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float a, float b, float c, float d, uint32_t remote_mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(remote_addr),
                 "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(remote_mbar)
                 : "memory");
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// Sum over the 8 lanes that share lane%4 of V values per lane, halving the live values at every step: afterwards
// the lane with g = lane/4 holds, in v[0 .. V/8-1], the totals of indices base(g) .. base(g) + V/8 - 1 with
// base(g) = (g>>2)*(V/2) + ((g>>1)&1)*(V/4) + (g&1)*(V/8).  For V = 4 the last step is a plain butterfly and both
// lanes of a pair hold the total of index (g>>2)*2 + ((g>>1)&1).
template <int V>
__device__ __forceinline__ void reduce_scatter_g(float (&v)[V], int lane)
{
    int mask = 16;
#pragma unroll
    for (int half = V / 2; half >= 1 && mask >= 4; half /= 2) {
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
        mask >>= 1;
    }
    if (V == 4) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 4);
}
#endif

}  // namespace ancde
