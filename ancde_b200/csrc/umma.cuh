// sm_100a 5th-generation tensor-core plumbing as inline PTX: tensor-memory (TMEM) allocation, shared-memory matrix
// descriptors, tcgen05.mma issue / commit, and TMEM -> register loads.
//
// Operand layout used throughout (no swizzle, K-major, "interleaved" canonical layout): a matrix X[rows][K] of FP16 is
// stored as core matrices of 8 rows x 16 bytes (8 halfs of k), each core matrix 128 contiguous bytes:
//     byte address of X[r][k] = base + (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2
// LBO = distance between core matrices adjacent in k, SBO = distance between 8-row groups; both multiples of 16.
// One tcgen05.mma.kind::f16 consumes K = 16 (two core matrices in k), M = 128 rows of A, N rows of B and accumulates
// D[m][n] (FP32) into TMEM lane m, column n.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ancde {

// 64-bit shared-memory matrix descriptor (tcgen05 "version 1", no swizzle).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46);
}

// 32-bit instruction descriptor: FP16 A (K-major) and B (K-major, or MN-major when b_mn), FP32 accumulator, M x N tile.
// An MN-major B matrix X[rows n][K] keeps 8 consecutive ROWS contiguous instead:
//     byte address of X[n][k] = base + (n / 8) * SBO + (k / 8) * LBO + (k % 8) * 16 + (n % 8) * 2
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, bool b_mn = false)
{
    return (1u << 4) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Allocates `cols` (power of two >= 32) TMEM columns; the base address is written to *smem_slot. One full warp.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_slot)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// Orders generic-proxy shared-memory writes (st.shared) before async-proxy reads (tcgen05.mma operand fetch).
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// True in exactly one lane of a converged warp.  Issuing tcgen05.mma under `if (warp-uniform) if (elect_one())` lets the
// compiler keep descriptors and TMEM addresses in uniform registers; under `if (threadIdx.x == 0)` every product is
// wrapped in an R2UR waterfall loop that costs ~100 cycles of issue time.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate));
    // no "memory" clobber: the operands are made visible by fence.proxy.async + a barrier before the issue, and
    // asm volatile statements keep their order; a clobber would only make the single issuing thread reload every
    // loop-invariant value between two products
}
// The same with the A operand in TENSOR MEMORY (lane = row m, two FP16 of consecutive k per 32-bit column, 8 columns per
// K = 16 step): no shared-memory fetch of A, which bounds the small-N products of this library (measured on B200,
// M = 128, N = 64: ~110 cycles per product from shared memory).  A from tensor memory is always K-major.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate));
}
// Registers -> 8 consecutive 32-bit columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid in taddr[31:16]).
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// Arrives once on the mbarrier when every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

// 16 consecutive FP32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid, encoded in taddr[31:16]).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    // the registers are operands of the wait so that no use of them can be scheduled ahead of it
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Two 16-column loads of this thread's lane in flight behind ONE wait (the round trip of a tcgen05.ld is ~300 cycles).
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr0, uint32_t taddr1, float (&v0)[16], float (&v1)[16])
{
    uint32_t r[16], q[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr0)
        : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
          "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
        : "r"(taddr1)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(q[0]), "+r"(q[1]), "+r"(q[2]),
                   "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]), "+r"(q[8]), "+r"(q[9]), "+r"(q[10]), "+r"(q[11]),
                   "+r"(q[12]), "+r"(q[13]), "+r"(q[14]), "+r"(q[15])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) { v0[i] = __uint_as_float(r[i]); v1[i] = __uint_as_float(q[i]); }
}

// tcgen05.ld split into issue and wait, so that the load of the NEXT tile is in flight while the current one is processed
// (tcgen05.wait::ld waits for every outstanding load of the thread).  The registers are operands of the wait so that no
// use of them is scheduled ahead of it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

}  // namespace ancde
