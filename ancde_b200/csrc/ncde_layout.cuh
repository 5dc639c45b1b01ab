// Host+device description of how one fused NCDE solve is laid out: packed weights, shared
// memory carve-up and the buffers saved for the backward pass.
#pragma once
#include "common.cuh"

namespace ancde {

// A-operand pack of a matrix A[M][K]: m-tiles of 16 rows; per m-tile `full` k16 steps (64 uint4 each: 32 lanes of hi
// fragments then 32 lanes of lo fragments) and an optional k8 tail (32 uint4: {hi a0, hi a1, lo a0, lo a1} per lane)
// or k16 tail.
struct APack {
    int M, K;           // logical sizes
    int mtiles;         // ceil(M / 16)
    int full;           // k16 steps (including a k16 tail)
    int tail8;          // 1 if a final k8 step follows
    int stride;         // uint4 per m-tile = full*64 + tail8*32
    int off;            // offset in uint4 units from the start of wpack
};

__host__ __device__ static inline APack make_apack(int M, int K, int off)
{
    APack a;
    a.M = M; a.K = K;
    a.mtiles = (M + 15) / 16;
    const int rem = K % 16;
    a.full = K / 16 + (rem > 8 ? 1 : 0);
    a.tail8 = (rem > 0 && rem <= 8) ? 1 : 0;
    a.stride = a.full * 64 + a.tail8 * 32;
    a.off = off;
    return a;
}

struct NcdeLayout {
    // ---- problem ----
    int B, L, C, Hd, n_hidden, mode, att_C, att_L, n_hp, n_steps, n_out;
    int width[ANCDE_MAX_LAYERS];
    // ---- derived ----
    int TS;            // samples per CTA
    int ntiles;        // CTAs
    int C4;            // C rounded up to 4: the output layer's columns are (i, j) -> i*C4 + j
    int NCH;           // C4 / 4 column chunks per row
    int NCG;           // Hd * NCH column groups of 4
    int NCP;           // Hd * C4 padded columns
    int K;             // width of the last hidden layer = reduction size of the output layer
    int Hmax;          // widest activation
    // packed weights (floats from the start of wpack).  Hidden layer l for the forward chain: rows W[o][stride] (k
    // contiguous, zero padded) + bias; for the backward chain a second block holds the transpose Wt[k][tstride] (o
    // contiguous).  Both strides are 4 * (odd number): 16-byte rows that a warp reads with conflict-free LDS.128
    // when lane = row.
    int in_dim[ANCDE_MAX_LAYERS], stride[ANCDE_MAX_LAYERS], tstride[ANCDE_MAX_LAYERS];
    int off_w[ANCDE_MAX_LAYERS], off_b[ANCDE_MAX_LAYERS], off_wt[ANCDE_MAX_LAYERS];
    int hidden_floats;  // forward block: all hidden layers (row-major weights + biases), starts at wpack[0]
    int off_hiddenb, hiddenb_floats;   // backward block: the transposed weights of all hidden layers
    int off_wo, off_bo; // output layer Wt_out[k][NCP], bias_out[NCP]
    int off_wot, KPS;   // second copy for the backward: W_out[col][KPS] (k contiguous, KPS = 8*ceil(K/8) + 4)
    // tensor-core copy for the backward: the output layer as FP16 hi/lo A fragments of W_out^T, per TILE of 16 rows
    // (2 field rows x 8 control channels: row r < 8 -> (2*ip, 8*jc + r), r >= 8 -> (2*ip + 1, 8*jc + r - 8)); a tile holds
    // MTK m-tiles (16 k each) of 64 uint4 (32 lanes hi, 32 lanes lo); values are scaled by 64
    int NIP, NJC, MTK, n_wtiles, off_wob;
    // hidden layers as FP16 hi/lo A fragments (scaled by 64): forward A = W_l [width][in], backward A = W_l^T [in][width];
    // APack::off is in uint4 units from off_whf / off_whb (float offsets into wpack)
    APack hfp[ANCDE_MAX_LAYERS], hbp[ANCDE_MAX_LAYERS];
    int off_whf, off_whb, hf_u4, hb_u4, off_hbias64, KSB;   // KSB: k16 steps of the widest activation
    unsigned magicNJC;  // ceil(2^32 / NJC)
    int off_gscale;     // 4 floats of scratch: max |stage cotangent| of the running backward (bits, atomicMax)
    int64_t wpack_floats;
    // saved activations per (stage, tile): z_s | h_1 .. h_n | dU | dU/da   (each [feature][TS])
    int act_off_h[ANCDE_MAX_LAYERS];
    int act_off_du, act_off_duda, act_floats;
    int act_fwd_floats;                   // floats of a block the forward writes (everything before gbar)
    int act_off_gbar;                     // written by the backward: cotangent of the stage derivative [Hd][TS]
    int act_off_delta[ANCDE_MAX_LAYERS];  // written by the backward: masked cotangent of hidden layer l [width][TS]
    int DUS, du_floats; // dU is stored per column chunk: [chunk][4][TS] with a chunk stride of DUS = 4*TS + 4 floats
                        // (bank-conflict free 128-bit reads by lanes that own consecutive chunks)
    int n_stages;       // 4 * n_steps
};

__host__ __device__ static inline int pad4(int x) { return (x + 3) & ~3; }
// smallest 4 * (odd number) >= x: a row stride in floats whose 16-byte chunks of 8 consecutive rows fall in 8 distinct
// bank groups
__host__ __device__ static inline int pad4odd(int x) { int q = (x + 3) >> 2; if ((q & 1) == 0) ++q; return 4 * q; }

// Fills the layout from a descriptor; returns 0 or an ANCDE_E* code (message set).
int make_layout(const ancde_ncde_desc* d, NcdeLayout* lo);

}  // namespace ancde
