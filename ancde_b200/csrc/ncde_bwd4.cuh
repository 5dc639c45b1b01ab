// Cluster / tcgen05 ("v4") recurrent backward of the fused NCDE solve: shared definitions.
//
// The streamed backward (ncde_bwd.cu) re-streams all of W_out^T (~500 KB as FP16 hi/lo fragments) through shared memory
// every RK4 stage for the 8 samples of a CTA and is bound by that shared-memory round trip.  Here a thread-block CLUSTER
// of CS CTAs walks NS = 8*CS samples backwards together, weight stationary:
//   * every CTA keeps a 1/CS slice of W_out^T RESIDENT in TENSOR MEMORY for the whole solve (the A operand of
//     tcgen05.mma may come from tensor memory: no shared-memory fetch of the weights at all, and 100 KB of shared memory
//     less than a resident copy there).  The slice is cut along the REDUCTION index p = i*C4 + j of
//     hbar[u] = sum_p W_out[p][u] * prebar[p]  (k steps of 16 columns).  It is ONE FP16 matrix of 128 rows: groups of 8
//     rows alternate between the hi halves and the lo halves of 64 * W_out[.][u] for 8 hidden units u (row
//     16*(u/8) + u%8 and + 8), so that the two halves of a unit sit in lanes l and l + 8 of the same warp when the
//     accumulator is read back; row m is TMEM lane m, two consecutive p per 32-bit column;
//   * per stage every CTA builds the pre-activation cotangent  prebar[p][n] = gbar[i][n] * dU[j][n] * (1 - G[p][n]^2)  of
//     its slice for ALL NS samples straight from the saved tanh outputs (HBM / L2 -> registers), splits it into FP16
//     hi/lo and writes it as the B operand [prebar_hi ; prebar_lo] (2*NS rows) into shared memory, in one pass or, where
//     the operand does not fit, in a few chunks; a dedicated issuer warp multiplies each chunk as soon as all worker
//     warps have written it: tcgen05.mma (M = 128, N = 2*NS, K = 16) per k step, the four hi/lo products accumulate in
//     tensor memory (the lo*lo quadrant is simply not read);
//   * the accumulator is read back with tcgen05.ld, the three useful quadrants are summed with one shuffle, and the
//     partial sums are REDUCE-SCATTERED through distributed shared memory (st.async + mbarrier transaction counts): CTA r
//     receives the CS partial hbar of ITS 8 samples;
//   * the hidden ReLU layers backwards, the RK4 adjoint recurrence (state in registers) and the saved cotangents for the
//     weight-gradient kernels are per CTA for its own 8 samples, exactly as in the streamed kernel; the stage cotangent
//     gbar of the next stage is broadcast to the cluster (st.async) for the next operand build.
// Buffers read and written in HBM are those of the streamed engine (tiles of 8 samples, ncde_layout.cuh), so either
// forward engine and all weight-gradient kernels work unchanged.
#pragma once
#include "ncde_common.cuh"

namespace ancde {

constexpr int V4_WORKERS = 512;             // 16 worker warps
constexpr int V4_THREADS = V4_WORKERS + 32;  // + the MMA issuer warp
constexpr int V4_MAX_CS = 8;

struct V4Layout {
    int CS, NS, nclusters;
    int QW;                         // warps per tile in the operand build = 16 / CS
    int PASSK;                      // k steps one pass of the 512 workers converts = 32 / CS (one item per thread)
    int PPC, npass_max;             // passes per operand chunk; passes of the longest slice
    int CK;                         // k steps (of 16 columns) per operand chunk = PPC * PASSK
    int nslots;                     // chunk buffers (1 or 2)
    int ks_begin[V4_MAX_CS + 1];    // k-step range of every cluster rank
    int nks_max, nchunks_max;
    int a_bytes;                    // per-rank A image in global memory: 128 rows x 16*nks_max FP16, row-major
    int64_t img_bytes;              // CS * a_bytes
    int a_col;                      // first TMEM column of the resident A slice (8 columns per k step); the accumulator sits at column 0
    int SBOb, ring_bytes;           // one chunk buffer: 2*NS rows x 16*CK columns
    int tmem_cols;                  // power of two >= max(32, 2*NS)
    int GSTR;                       // row stride (floats) of the gathered stage cotangent [Hd][GSTR]
    int GROW;                       // row stride (floats) of the staged tanh outputs [NS][GROW] = slice + 16 bytes
    int dub;                        // floats of one tile's dU (| dU/da) block
    int act_ld;                     // floats of a saved-activation block staged in shared memory (z_s | h_1 .. h_n)
    int KP, HP;                     // hidden-layer buffers (see ncde_bwd.cu)
    int chain;                      // hidden-layer mode of ncde_bwd4_kernel<HM>: 0 FP32 warp chain, 1 mma.sync, 2 FP32 one output per thread
    int chain_warps;                // warps that run the hidden layers (chain: one per sample; mma.sync: one per 16 inputs)
    int need_att;                   // timewise attention cotangent requested
    // shared-memory carve-up (byte offsets)
    int o_bars, o_ring, o_hid, o_recv, o_recvatt, o_gall, o_du, o_act, o_zb, o_bd, o_max, o_lp, o_tab, o_gst, o_tout;
    int smem_total;
};

struct Bwd4Args {
    SolveArgs s;
    V4Layout v;
    const unsigned char* img;
};

// Fills the layout when the shape is supported by this engine (ANCDE_EUNSUPPORTED otherwise, quietly).
int make_layout4(const ancde_ncde_desc* d, const NcdeLayout& st, bool need_att, int force_cs, V4Layout* v);
// make_layout4 with the cluster size chosen for co-residency (force_cs = 0) or given.
int choose_layout4(const ancde_ncde_desc* d, const NcdeLayout& st, bool need_att, int force_cs, V4Layout* v);
int pack_weights4(const ancde_ncde_desc* d, const NcdeLayout& st, const V4Layout& v, unsigned char* img, cudaStream_t stream);
int launch_bwd4(const SolveArgs& a, const V4Layout& v, const unsigned char* img, cudaStream_t stream);

}  // namespace ancde
