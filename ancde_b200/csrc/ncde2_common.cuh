// Tensor-core ("v2") fused NCDE solve: shared definitions.
//
// Design: the vector field's output layer  G = tanh(W_out h + b).view(Hd, C)  is 88 % of the
// path's FLOPs.  With the batch spread over all SMs an SM only owns ~8 samples, so a plain FFMA or UMMA formulation
// re-reads all of W_out (84 k floats at the Sepsis shape) from shared memory for every 8 samples and is bound by
// shared-memory bandwidth.  Here a thread-block CLUSTER of CS CTAs integrates NS = 8*NT samples together: every CTA
// keeps a 1/CS slice of W_out's rows resident in shared memory for the whole solve (pre-split into FP16 hi/lo
// fragments, 4 bytes per weight), multiplies it with the activations of all NS samples on the tensor cores
// (mma.sync m16n8k16, 3 products per tile = FP32-grade accuracy: hi*hi + lo*hi + hi*lo), applies tanh and the
// contraction with dU in the accumulator registers, and the CTAs exchange the 1/CS slices of the stage derivative
// through distributed shared memory.  The small hidden layers are evaluated redundantly by every CTA of the cluster.
//
// Output-layer rows are grouped into TILES of 16 = 2 field rows (i-pair) x 8 control channels (j-chunk):
//   tile (ip, jc), row r < 8  -> (i = 2*ip,     j = 8*jc + r)
//                  row r >= 8 -> (i = 2*ip + 1, j = 8*jc + r - 8)
// so that in the m16n8 accumulator fragment a thread holds the same j for both field rows and the contraction over
// j is a reduction over the 8 lanes that share lane%4.
#pragma once
#include <cuda_fp16.h>

#include "ncde_common.cuh"
#include "pipeline.cuh"

namespace ancde {

constexpr float V2_WSCALE = 64.0f;          // weights are stored as fp16(64*W): keeps the lo halves out of the subnormals
constexpr float V2_WSCALE_INV = 1.0f / 64.0f;
constexpr int V2_MAX_KSTEPS = 4;            // output-layer reduction size K <= 64
constexpr int V2_RK_ITEMS = 6;              // (state element, sample) items a thread keeps in registers
constexpr int V2_MAX_CS = 8;

struct V2Layout {
    // ---- problem ----
    int B, L, C, Hd, n_hidden, mode, att_C, att_L, n_hp, n_steps, n_out, n_stages;
    int width[ANCDE_MAX_LAYERS], in_dim[ANCDE_MAX_LAYERS];
    int K;                      // width of the last hidden layer
    // ---- output-layer tiling ----
    int NIP, NJC, C8, ntiles;   // i-pairs, j-chunks, 8*NJC, NIP*NJC
    // ---- packed weights (wpack, 16-byte units) ----
    APack hf[ANCDE_MAX_LAYERS];     // hidden layer l forward:   A = 64*W_l      [width_l][in_l]
    APack hb[ANCDE_MAX_LAYERS];     // hidden layer l backward:  A = 64*W_l^T    [in_l][width_l]
    APack of;                       // output layer forward:     A = 64*W_out tiles, M = ntiles*16, K
    APack ob;                       // output layer backward:    per tile A = 64*W_out^T [K][16 rows]; M = K, K = 16, `off` per tile = ob.off + tile*ob_tile
    int ob_tile;                    // uint4 per tile of the backward pack = mtiles(K)*64
    int off_hbias[ANCDE_MAX_LAYERS];    // float offsets (from wpack as float*): 64*bias_l padded to 16*mtiles
    int off_obias;                      // 64*bias_out per tile row [ntiles][16]
    int hidden_u4, hidden_b_u4;     // total uint4 of all hf / hb packs (contiguous from hf[0].off / hb[0].off)
    int64_t wpack_floats;
    int KSB;                        // k16 steps of the widest B-operand (activations) buffer
    int MTz;                        // m-tiles of the state (ceil(Hd / 16))
    // ---- tiling of the batch (fixed per shape so that forward and backward agree on the saved layouts) ----
    int NT, NS, nclusters;          // n-tiles (8 samples each) and samples per cluster, clusters
    unsigned magicC;                // ceil(2^32 / C): item / C == __umulhi(item, magicC) for item * C < 2^32
    // ---- saved activations, per (stage, cluster) a block of act_block floats; matrices [rows][NS] are stored in
    // accumulator-FRAGMENT order: m-tile mt, n-tile nt, lane -> float4 {(g, 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)}
    int fo_z, fo_h[ANCDE_MAX_LAYERS], fo_du, fo_duda, fo_gbar, fo_delta[ANCDE_MAX_LAYERS], act_block;
    // dU / dU/da blocks are [item = n*C + j]
    // tanh outputs: [stage][cluster][tile][nt][lane] float4 in the same fragment order (rows = tile rows)
};

struct V2Launch {
    int CS, threads;
    int ip_begin[V2_MAX_CS + 1];    // i-pair range of every cluster rank
    size_t smem;
};

struct V2Args {
    V2Layout lo;
    V2Launch la;
    float t_start, t_end, step;
    const float* times;
    const float* t_out;
    const float* cb;
    const float* cc2;
    const float* cd3;
    const float* z0;
    const float* attention;
    const float* h_prime;
    const uint4* wpack;
    float* z_out;
    float* k_out;
    float* save_act;
    float* save_G;
    const float* grad_z_out;
    float* grad_z0;
    float* grad_attention;
    long long* dbg;
};

int make_layout2(const ancde_ncde_desc* d, V2Layout* lo);
int pack_weights2(const ancde_ncde_desc* d, const V2Layout& lo, cudaStream_t stream);
int fill_args2(const ancde_ncde_desc* d, const V2Layout& lo, V2Args* a);
int launch_fwd2(V2Args& a, cudaStream_t stream);
int choose_launch2_fwd(V2Layout* lo, V2Launch* la);

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
// ---- tensor-core primitives (legacy mma.sync path: the operands live in registers, M = 16 rows of weights,
// N = 8 samples -- the shape of this problem; see DESIGN.md for why not tcgen05) ------------------------------
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
