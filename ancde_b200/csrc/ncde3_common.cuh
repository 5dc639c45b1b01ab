// UMMA ("v3") fused NCDE forward solve: tcgen05 tensor cores with the accumulators in tensor memory.
//
// The vector field's output layer  G = tanh(W_out h + b).view(Hd, C)  is ~88 % of the path's FLOPs, but the batch spread
// over all SMs leaves an SM only ~8 samples -- too few columns for a 128-row tensor-core tile to pay, and W_out (in the
// FP16 hi/lo split that FP32-grade accuracy needs, 4 bytes per weight) does not fit one SM's shared memory.  So a
// thread-block CLUSTER of CS CTAs integrates NS = 8*CS samples together:
//   * every CTA keeps a 1/CS slice of W_out's rows RESIDENT in shared memory for the whole solve, as two FP16 matrices
//     (hi, lo) in the tensor core's canonical K-major layout (umma.cuh), bias folded in as an extra k column;
//   * per RK4 stage one thread issues tcgen05.mma (M = 128 rows of W_out, N = NS samples, K = 16) three times per k
//     step (lo*hi + hi*lo + hi*hi) into TMEM; the warps read the accumulators back with tcgen05.ld, apply tanh,
//     multiply with dU and reduce over the control channel with a shuffle butterfly;
//   * the hidden ReLU layers run on the same tensor cores (M = units, N = samples), each CTA redundantly for all NS
//     samples, with the activations written back to shared memory as the next layer's B operand;
//   * the CTAs exchange their 1/CS slices of the stage derivative through distributed shared memory.
//
// Row order of a CTA's W_out slice (rows are TMEM lanes, 32 per warp): the field rows i in [i_begin, i_end) of the
// CTA, each split over the control channel j into chunks of 32.  A FULL chunk (32 channels) is one 32-row block
// (lane = j - 32*chunk), so that the contraction over j is a reduction over the lanes of ONE warp.  The last, partial
// chunk of `rem` channels is padded to seg = 2^ceil(log2 rem) lanes and 32/seg field rows share a block.
#pragma once
#include <cuda_fp16.h>

#include "ncde_common.cuh"
#include "pipeline.cuh"

namespace ancde {

constexpr int V3_MAX_CS = 8;
constexpr int V3_THREADS = 512;
constexpr int V3_MAX_TILES = 16;
constexpr float V3_WSCALE = 64.0f;          // weights are stored as fp16(64*W): keeps the lo halves out of the subnormals
constexpr float V3_WSCALE_INV = 1.0f / 64.0f;

struct V3Layout {
    // ---- problem ----
    int B, L, C, Hd, n_hidden, mode, att_C, att_L, n_hp, n_steps, n_out, n_stages;
    int width[ANCDE_MAX_LAYERS], in_dim[ANCDE_MAX_LAYERS];
    int K;                          // width of the last hidden layer
    // ---- batch / row tiling ----
    int CS, NS, ns_log2, nclusters;
    int i_begin[V3_MAX_CS + 1];     // field-row range of every cluster rank
    int ni_max;                     // most field rows a rank owns
    int nfull, rem, seg, seg_log2, ipb, nchunks;   // control-channel chunks (see above); ipb = 32 / seg
    int mtiles;                     // 128-row tiles of a W_out slice (same for every rank; short slices are zero padded)
    // ---- operands: FP16 matrices in the canonical K-major layout, LBO = 128 bytes ----
    int KPo, SBOo;                  // output layer: padded k (multiple of 16, > K), bytes per 8-row group
    int KPh[ANCDE_MAX_LAYERS], SBOh[ANCDE_MAX_LAYERS], rows_h[ANCDE_MAX_LAYERS];   // hidden layer l (rows stored: 8*ceil(width/8))
    int KPa, SBOa;                  // activation operand buffer [NS][KPa]
    // ---- packed image (bytes from the start of the image buffer) ----
    int off_hhi[ANCDE_MAX_LAYERS], off_hlo[ANCDE_MAX_LAYERS];
    int hidden_bytes;               // all hidden layers; the same for every rank
    int wo_mat_bytes;               // one matrix (hi or lo) of a W_out slice = 16*mtiles*SBOo
    int64_t img_bytes;              // hidden_bytes + CS * 2 * wo_mat_bytes
    // ---- shared memory ----
    int DUSTR, du_floats;           // dU[j][DUSTR] (DUSTR = NS + 4: conflict-free 128-bit reads by lanes with consecutive j)
    int tcol_h;                     // TMEM column of the hidden layers' accumulator (W_out tile t sits at column t*2*NS)
    // ---- round 2: the W_out slice resident in TENSOR MEMORY (A operand of tcgen05.mma from TMEM), when the columns allow ----
    int ts;                         // 1: W_out hi/lo tiles live in TMEM from column a_col on (tile t: KPo/2 columns; hi tiles, then lo)
    int a_col;
    // ---- single-CTA variant (ncde3_fwd1.cu): CS = 1, 8 samples per CTA, hidden layers (<= 32 units) as the FP32 warp chain of
    // the streamed engine (its weights: c_hidden_floats floats at wpack[0], rows c_stride, see ncde_layout.cuh) ----
    int chain1, c_hidden_floats, c_HP;
    int c_off_w[ANCDE_MAX_LAYERS], c_off_b[ANCDE_MAX_LAYERS], c_stride[ANCDE_MAX_LAYERS];
    int sg_str, C4, nip_log2;       // ts: tanh outputs staged in shared memory [NS][sg_str = ni_max*C4], C4 = C rounded up to 4; items of
                                    // the contraction = (sample, field row < 2^nip_log2); dU is then kept as [NS][DUSTR = C4]
    // ---- buffers saved for the backward, in the streamed engine's layout with tiles of 8 samples (ncde_layout.cuh) ----
    int sv_ntiles, sv_act_floats, sv_off_h[ANCDE_MAX_LAYERS], sv_off_du, sv_off_duda, sv_DUS, sv_NCP, sv_C4;
};

struct V3Smem {
    size_t img, act, du, k, kpart, times, tout, bars, sg, total;
};

__host__ __device__ static inline V3Smem v3_smem(const V3Layout& lo)
{
    V3Smem s;
    size_t o = 0;
    s.img = o;    o += (size_t)lo.hidden_bytes + (lo.ts ? 0 : 2 * (size_t)lo.wo_mat_bytes);
    s.act = o;    o += 2 * (size_t)(lo.NS / 8) * lo.SBOa;                       // hi then lo
    s.du = o;     o += 2 * (size_t)lo.du_floats * 4;                            // double buffered by stage parity
    s.k = o;      o += 2 * (size_t)lo.Hd * (lo.NS + 4) * 4;                     // stage derivative [Hd][NS + 4], double buffered
    s.kpart = o;  o += lo.ts ? 0 : (size_t)lo.ni_max * lo.nchunks * lo.NS * 4;
    s.times = o;  o += (size_t)pad4(lo.L) * 4;
    s.tout = o;   o += (size_t)pad4(lo.n_out) * 4;
    s.bars = o;   o += 256;                                                     // mbarriers + TMEM base slot
    s.sg = o;     o += lo.ts ? (size_t)lo.NS * lo.sg_str * 4 : 0;
    // a hidden layer's A descriptor spans 128 rows: what it reads past the stored rows must stay inside the allocation
    for (int l = 0; l < lo.n_hidden; ++l)
        if ((size_t)lo.off_hlo[l] + 16 * (size_t)lo.SBOh[l] > o) o = (size_t)lo.off_hlo[l] + 16 * (size_t)lo.SBOh[l];
    s.total = (o + 15) & ~(size_t)15;
    return s;
}

// Shared-memory carve-up of the single-CTA variant
struct V3S1Smem {
    size_t hw, lp, z, last, act, du, sg, times, tout, bars, total;
};
__host__ __device__ static inline V3S1Smem v3_s1_smem(const V3Layout& lo)
{
    V3S1Smem s;
    size_t o = 0;
    s.hw = o;     o += (size_t)pad4(lo.c_hidden_floats) * 4;
    s.lp = o;     o += 8 * ANCDE_MAX_LAYERS * 4;
    s.z = o;      o += 3 * (size_t)8 * lo.c_HP * 4;                             // stage input, ping, pong: [sample][c_HP]
    s.last = o;   o += (size_t)pad4(lo.K * 8) * 4;                             // last hidden layer [unit][8]
    s.act = o;    o += 2 * (size_t)lo.SBOa;                                    // B operand: hi (8 samples) then lo
    s.du = o;     o += 2 * (size_t)8 * lo.C4 * 4;                              // [8][C4], double buffered by stage parity
    s.sg = o;     o += (size_t)8 * lo.sg_str * 4;
    s.times = o;  o += (size_t)pad4(lo.L) * 4;
    s.tout = o;   o += (size_t)pad4(lo.n_out) * 4;
    s.bars = o;   o += 256;
    s.total = (o + 15) & ~(size_t)15;
    return s;
}

struct V3Args {
    V3Layout lo;
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
    const unsigned char* img;
    const float* wpack;        // streamed engine's packed weights (the FP32 hidden block of the single-CTA variant)
    float* z_out;
    float* k_out;
    float* save_act;
    float* save_G;
    long long* dbg;
};

// tanh(x) = 1 - 2 / (exp(2x) + 1), x = acc / weight-scale; |abs err| < 4e-7 (MUFU ex2 + rcp)
__device__ __forceinline__ float tanh_scaled3(float acc, float c)
{
    float ex, r;
    const float a = acc * c;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(a));
    const float dd = ex + 1.0f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dd));
    return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo)
{
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

constexpr int CTL_BATCH = 6;        // control items a thread loads before the first use (one memory round trip per batch)

// Returns ANCDE_OK and fills the layout when the shape is supported by this engine (quietly ANCDE_EUNSUPPORTED otherwise).
int make_layout3(const ancde_ncde_desc* d, const NcdeLayout* streamed, V3Layout* lo);
int pack_weights3(const ancde_ncde_desc* d, const V3Layout& lo, unsigned char* img, cudaStream_t stream);
int launch_fwd3(const ancde_ncde_desc* d, const V3Layout& lo, const unsigned char* img, cudaStream_t stream);
int launch_fwd3_single(const V3Args& a, cudaStream_t stream);

// Row r (0 .. 128*mtiles) of rank's slice -> field row i, control channel j; false for padding rows.
__host__ __device__ static inline bool v3_row(const V3Layout& lo, int rank, int r, int* i, int* j)
{
    const int i0 = lo.i_begin[rank], ni = lo.i_begin[rank + 1] - i0;
    const int blk = r >> 5, lane = r & 31;
    if (blk < ni * lo.nfull) {
        const int il = blk / lo.nfull, c = blk - il * lo.nfull;
        *i = i0 + il;
        *j = 32 * c + lane;
        return true;
    }
    if (lo.rem == 0) return false;
    const int q = blk - ni * lo.nfull;
    const int il = q * lo.ipb + (lane >> lo.seg_log2), jj = lane & (lo.seg - 1);
    *i = i0 + il;
    *j = 32 * lo.nfull + jj;
    return il < ni && jj < lo.rem;
}

}  // namespace ancde
