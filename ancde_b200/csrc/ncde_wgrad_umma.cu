// Output-layer weight gradient of the fused NCDE solve on the tcgen05 tensor cores (tiles of 8 samples).
//
//   dW_out[(i,j)][k] = sum over every (stage n, sample b) of  pre_bar[n,b,(i,j)] * h_last[n,b,k]
//   pre_bar[n,b,(i,j)] = gbar[n,b,i] * dU[n,b,j] * (1 - G[n,b,i,j]^2)          (tanh' of the saved output)
//
// A GEMM whose reduction dimension r = (stage, sample) is 4*steps*B long (290 816 for the Sepsis batch) and whose left
// operand is produced on the fly from the tanh outputs G the forward saved in HBM: this kernel reads G exactly once and
// is bound by that HBM stream.  A CTA owns 128 output-layer columns and a slice of the rows; it keeps the accumulator
// D[128 columns][k] in tensor memory for its whole life and per chunk of 32 rows
//   1. loads its 32 x 128 slab of G straight into registers (coalesced 128-bit streaming loads, issued three chunks ahead:
//      the HBM latency is covered by registers, not by shared memory) and receives the small per-tile tensors (last
//      hidden activation, dU, stage cotangent: one contiguous 5.7 KB span of the saved block) by ONE TMA bulk copy per
//      tile into a 4-stage ring (completion on an mbarrier per stage; per-copy overhead of many small copies was the
//      limit of the first version),
//   2. all 16 warps turn G into pre_bar and split it and h into FP16 hi/lo operands in shared memory (scaled by a power
//      of two taken from max |gbar|, which the recurrent backward kernel tracks, so that the cotangents sit in FP16's
//      range), in the tensor core's canonical layouts: A = pre_bar^T MN-major (8 columns contiguous), B = h K-major
//      (the 8 samples of a tile contiguous), two operand buffers so that the conversion of chunk c+1 overlaps the
//      products of chunk c,
//   3. one elected thread issues  D += pre_bar_hi^T [h_hi ; h_lo]  (one tcgen05.mma with the two h matrices stacked
//      along N fills columns [0,64) and [64,128) of D) and  D[:, 0:64] += pre_bar_lo^T h_hi:  FP32-grade accuracy,
//      2 x 2 instructions per chunk; tcgen05.commit frees the operand buffer.
// The bias gradient is the extra k column h = 1.  At the end D is read back (tcgen05.ld) and added to the gradient
// with atomics (a column tile is shared by the CTAs of the row slices).
#include "mma_common.cuh"
#include "umma.cuh"

namespace ancde {

constexpr int WU_ROWS = 32;         // rows (stage, sample) per chunk = 4 tiles of 8 samples = two k16 steps
constexpr int WU_COLS = 128;        // output-layer columns per accumulator tile = TMEM lanes
constexpr int WU_TILES = 2;         // accumulator tiles per CTA (they share the h operand, the small loads and the barrier of a chunk)
constexpr int WU_CW = WU_COLS / 4;  // column groups of 4
constexpr int WU_STAGES = 4;
constexpr int WU_GDEPTH = 2;        // chunks of G in flight in registers
constexpr int WU_THREADS = 512;
constexpr int WU_NK = 64;           // padded k (hidden units + bias column)
constexpr int WU_LBO = 128;         // bytes between core matrices adjacent in r (the reduction dimension)
constexpr int WU_SBO = 4 * WU_LBO + 16;   // bytes between groups of 8 columns / 8 k (+16: spreads the groups over the banks)
constexpr int WU_A_BYTES = (WU_COLS / 8) * WU_SBO;          // one pre_bar matrix (hi or lo)
constexpr int WU_B_BYTES = (2 * WU_NK / 8) * WU_SBO;        // h_hi stacked on h_lo
constexpr int WU_OP_BYTES = WU_TILES * 2 * WU_A_BYTES + WU_B_BYTES;    // one operand buffer: per tile { pre_bar hi, lo }, then h

__global__ void __launch_bounds__(WU_THREADS, 1) ncde_wgrad_umma_kernel(const __grid_constant__ SolveArgs p, float* __restrict__ gW,
                                                                        float* __restrict__ gB, int chunks_per_cta,
                                                                        const unsigned* __restrict__ gmax_bits)
{
    constexpr int TS = 8, BPC = WU_ROWS / TS, NT = WU_THREADS;
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int Hd = lo.Hd, C4 = lo.C4, NCH = lo.NCH, NCG = lo.NCG, K = lo.K, C = lo.C;
    int cg[WU_TILES], i_row[WU_TILES], chunk[WU_TILES];      // this lane's column group in tile t of the transform pass
    bool cg_ok[WU_TILES];
#pragma unroll
    for (int t = 0; t < WU_TILES; ++t) {
        cg[t] = (blockIdx.x * WU_TILES + t) * WU_CW + lane;
        cg_ok[t] = cg[t] < NCG;
        i_row[t] = cg_ok[t] ? cg[t] / NCH : 0;
        chunk[t] = cg_ok[t] ? cg[t] - i_row[t] * NCH : 0;
    }

    // power-of-two scale that brings the largest cotangent to ~2^6
    float scale = 1.0f;
    {
        const float gm = __uint_as_float(__ldg(gmax_bits));
        if (gm > 0.f) {
            int e;
            frexpf(gm, &e);                    // gm = f * 2^e, f in [0.5, 1)
            scale = ldexpf(1.0f, 6 - e);
        }
    }

    // shared memory: WU_STAGES load stages of BPC spans { h_last | dU | (dU/da) | gbar } of the saved blocks, then two
    // operand buffers { pre_bar hi, pre_bar lo, h hi | h lo }
    __shared__ __align__(8) uint64_t ld_bar[WU_STAGES];    // TMA completion of a load stage
    __shared__ __align__(8) uint64_t mma_bar[2];           // products that read operand buffer b have completed
    __shared__ uint32_t tslot;
    extern __shared__ __align__(128) float smem[];
    const int off_hl = lo.act_off_h[lo.n_hidden - 1];
    const int span = lo.act_off_gbar + pad4(Hd * TS) - off_hl;     // floats of a block from h_last to the end of gbar
    const int stage_floats = BPC * span;
    unsigned char* sOp = reinterpret_cast<unsigned char*>(smem + WU_STAGES * stage_floats);

    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + BPC - 1) / BPC;
    const int64_t c0 = (int64_t)blockIdx.y * chunks_per_cta;
    int64_t c1 = c0 + chunks_per_cta;
    if (c1 > nchunks) c1 = nchunks;
    const int n_it = (c1 > c0) ? (int)(c1 - c0) : 0;
    // Small tensors of a chunk: one bulk copy per (stage, tile) block, issued by the first BPC lanes of warp 0.
    auto issue = [&](int64_t ch, int s) {
        const int64_t blk0 = ch * BPC;
        int64_t nvalid = nblocks - blk0;                      // (stage, tile) blocks of this chunk that exist
        if (nvalid > BPC) nvalid = BPC;
        if (lane == 0) mbar_arrive_expect_tx(ld_bar + s, (uint32_t)nvalid * (uint32_t)(span * 4));
        __syncwarp();
        if (lane < nvalid)
            bulk_g2s(smem + s * stage_floats + lane * span, p.save_act + (blk0 + lane) * lo.act_floats + off_hl, (uint32_t)(span * 4), ld_bar + s);
    };
    // G of a chunk: this thread's (row, column group) pieces -- 2 rows x WU_TILES tiles -- streamed past L1
    const float4* gsave = reinterpret_cast<const float4*>(p.save_G);
    auto load_g = [&](int64_t ch, float4 (&g)[2 * WU_TILES]) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = warp + rr * 16;
            const int64_t blk = ch * BPC + (r >> 3);
            const bool row_ok = ch < c1 && blk < nblocks;
#pragma unroll
            for (int t = 0; t < WU_TILES; ++t) {
                g[rr * WU_TILES + t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_ok && cg_ok[t]) g[rr * WU_TILES + t] = __ldcs(gsave + (blk * TS + (r & 7)) * NCG + cg[t]);
            }
        }
    };

    if (tid == 0) {
        for (int s = 0; s < WU_STAGES; ++s) mbar_init(ld_bar + s, 1);
        mbar_init(mma_bar + 0, 1);
        mbar_init(mma_bar + 1, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tslot, 128 * WU_TILES);
    // operand buffers: zero, then the bias column (k == K) of h_hi = 1.0 for every row
    for (int i = tid; i < 2 * WU_OP_BYTES / 16; i += NT) reinterpret_cast<uint4*>(sOp)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (tid < 2 * WU_ROWS) {
        const int ob = tid / WU_ROWS, r = tid - ob * WU_ROWS;
        unsigned char* sB = sOp + ob * WU_OP_BYTES + WU_TILES * 2 * WU_A_BYTES;
        *reinterpret_cast<__half*>(sB + (K >> 3) * WU_SBO + (r >> 3) * WU_LBO + (K & 7) * 16 + (r & 7) * 2) = __float2half_rn(1.0f);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;

    if (warp == 0)
        for (int s = 0; s < WU_STAGES; ++s)
            if (s < n_it) issue(c0 + s, s);
    float4 g0[2 * WU_TILES], g1[2 * WU_TILES];          // G of chunks it, it+1
    load_g(c0, g0);
    load_g(c0 + 1, g1);

    const uint32_t idesc_a = (1u << 4) | (1u << 15) | ((uint32_t)(2 * WU_NK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // A MN-major, N = 128
    const uint32_t idesc_b = (1u << 4) | (1u << 15) | ((uint32_t)(WU_NK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);       // N = 64

    for (int it = 0; it < n_it; ++it) {
        const int64_t ch = c0 + it;
        const int s = it % WU_STAGES, ob = it & 1;
        const float* sS = smem + s * stage_floats;          // block bl: h_last at sS + bl*span, dU at + d_off, gbar at + g_off
        const int d_off = lo.act_off_du - off_hl, g_off = lo.act_off_gbar - off_hl;
        unsigned char* sA = sOp + ob * WU_OP_BYTES;          // tile t: hi at sA + t*2*WU_A_BYTES, lo right behind
        unsigned char* sBh = sA + WU_TILES * 2 * WU_A_BYTES;
        unsigned char* sBl = sBh + (WU_NK / 8) * WU_SBO;
        float4 g2[2 * WU_TILES];
        load_g(ch + WU_GDEPTH, g2);                          // two chunks ahead
        mbar_wait(ld_bar + s, (uint32_t)((it / WU_STAGES) & 1));
        if (it >= 2) mbar_wait(mma_bar + ob, (uint32_t)(((it - 2) >> 1) & 1));       // the products of chunk it-2 have read this buffer
        const int64_t blk_left = nblocks - ch * BPC;            // rows of blocks beyond the end contribute nothing
        // ---- G -> scaled pre_bar, FP16 hi/lo, A[m][r] MN-major: this lane's 4 columns, rows warp and warp + 16
#pragma unroll
        for (int rr = 0; rr < WU_ROWS / 16; ++rr) {
            const int r = warp + rr * 16;
            const int bl = r / TS, b = r - bl * TS;
#pragma unroll
            for (int t = 0; t < WU_TILES; ++t) {
                const float4 gq = g0[rr * WU_TILES + t];
                const bool live = bl < blk_left && cg_ok[t];
                const float gb = live ? sS[bl * span + g_off + i_row[t] * TS + b] * scale : 0.f;
                const float* du = sS + bl * span + d_off + chunk[t] * lo.DUS + b;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                if (live) {
                    v0 = gb * du[0 * TS] * fmaf(-gq.x, gq.x, 1.0f);
                    v1 = gb * du[1 * TS] * fmaf(-gq.y, gq.y, 1.0f);
                    v2 = gb * du[2 * TS] * fmaf(-gq.z, gq.z, 1.0f);
                    v3 = gb * du[3 * TS] * fmaf(-gq.w, gq.w, 1.0f);
                }
                uint32_t h0, l0, h1, l1;
                split_pair(v0, v1, h0, l0);
                split_pair(v2, v3, h1, l1);
                const uint32_t o = (uint32_t)(t * 2 * WU_A_BYTES + (lane >> 1) * WU_SBO + (r >> 3) * WU_LBO + (r & 7) * 16 + (lane & 1) * 8);
                *reinterpret_cast<uint2*>(sA + o) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(sA + o + WU_A_BYTES) = make_uint2(l0, l1);
            }
        }
        // ---- h -> FP16 hi/lo, B[k][r] K-major: one (k, tile) item = the 8 samples of a tile = one 128-bit store
        for (int idx = tid; idx < K * BPC; idx += NT) {
            const int bl = idx / K, kk = idx - bl * K;
            float v[8];
            if (bl < blk_left) {
                const float4 a = *reinterpret_cast<const float4*>(sS + bl * span + kk * TS);
                const float4 c = *reinterpret_cast<const float4*>(sS + bl * span + kk * TS + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.f;
            }
            uint32_t hh[4], hl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_pair(v[2 * e], v[2 * e + 1], hh[e], hl[e]);
            const uint32_t o = (uint32_t)((kk >> 3) * WU_SBO + bl * WU_LBO + (kk & 7) * 16);
            *reinterpret_cast<uint4*>(sBh + o) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(sBl + o) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
        }
        fence_async_smem();
        __syncthreads();
        // ---- D += pre_bar^T h over the 32 rows (two k16 steps)
        if (warp_u == 1) {
            if (elect_one()) {
                tc_fence_after();
                const uint64_t dB = umma_smem_desc(smem_u32(sBh), WU_LBO, WU_SBO);
#pragma unroll
                for (int t = 0; t < WU_TILES; ++t) {
                    const uint64_t dAh = umma_smem_desc(smem_u32(sA) + t * 2 * WU_A_BYTES, WU_LBO, WU_SBO);
                    const uint64_t dAl = umma_smem_desc(smem_u32(sA) + t * 2 * WU_A_BYTES + WU_A_BYTES, WU_LBO, WU_SBO);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) umma_f16(tbase + t * 128, dAh + (uint64_t)(ks * 16), dB + (uint64_t)(ks * 16), idesc_a, (it | ks) != 0);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) umma_f16(tbase + t * 128, dAl + (uint64_t)(ks * 16), dB + (uint64_t)(ks * 16), idesc_b, true);
                }
                umma_commit(mma_bar + ob);
            }
        }
        if (warp == 0 && it + WU_STAGES < n_it) issue(ch + WU_STAGES, s);      // every thread has finished reading stage s
#pragma unroll
        for (int e = 0; e < 2 * WU_TILES; ++e) { g0[e] = g1[e]; g1[e] = g2[e]; }
    }

    // ---- results: D[col][k] + D[col][64 + k]; TMEM lane = column of the tile, this warp's quarter, 16 k per warp group
    if (n_it > 0) {
        for (int ob = 0; ob < 2; ++ob)
            if (n_it > ob) mbar_wait(mma_bar + ob, (uint32_t)(((n_it - 1 - ob) >> 1) & 1));
        tc_fence_after();
        const float inv = 1.0f / scale;
        const int wq = warp & 3, kg = warp >> 2;
#pragma unroll
        for (int t = 0; t < WU_TILES; ++t) {
            float v[16], v2[16];
            const uint32_t ta = tbase + ((uint32_t)(32 * wq) << 16) + t * 128 + 16 * kg;
            tmem_ld16x2(ta, ta + WU_NK, v, v2);
            const int col = (blockIdx.x * WU_TILES + t) * WU_COLS + 32 * wq + lane;
            const int i = col / C4, j = col - i * C4;
            if (i < Hd && j < C) {
                const int row = i * C + j;           // row of the torch (Hd*C, K) weight
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int kr = 16 * kg + e;
                    const float val = (v[e] + v2[e]) * inv;
                    if (kr < K) atomicAdd(gW + (int64_t)row * K + kr, val);
                    else if (kr == K) atomicAdd(gB + row, val);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, 128 * WU_TILES);
}

// Returns ANCDE_EUNSUPPORTED (quietly) when the shape does not fit this kernel; the caller then uses the mma.sync kernel.
int launch_wgrad_umma(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    if (lo.TS != 8 || lo.K + 1 > WU_NK) return ANCDE_EUNSUPPORTED;
    constexpr int BPC = WU_ROWS / 8;
    const int span = lo.act_off_gbar + pad4(lo.Hd * 8) - lo.act_off_h[lo.n_hidden - 1];
    if (span % 4 != 0 || lo.act_off_h[lo.n_hidden - 1] % 4 != 0 || lo.act_floats % 4 != 0) return ANCDE_EUNSUPPORTED;   // 16-byte bulk copies
    const size_t stage_floats = (size_t)BPC * span;
    const size_t smem = WU_STAGES * stage_floats * sizeof(float) + 2 * (size_t)WU_OP_BYTES;
    if (smem > 226 * 1024) return ANCDE_EUNSUPPORTED;   // 227 KB per CTA minus the static barriers (Sepsis OI top network: 214 KB)
    const unsigned* gmax = reinterpret_cast<const unsigned*>(a.wpack + lo.off_gscale);   // written by ncde_bwd_kernel
    const int gx = (lo.NCP + WU_TILES * WU_COLS - 1) / (WU_TILES * WU_COLS);
    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + BPC - 1) / BPC;
    ANCDE_CUDA(cudaFuncSetAttribute(ncde_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int num_sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // exactly one wave, one CTA per SM: as many row slices as fit next to each other
    int64_t gy = num_sms / gx;
    if (gy > nchunks) gy = nchunks;
    if (gy < 1) gy = 1;
    const int per = (int)((nchunks + gy - 1) / gy);
    gy = (nchunks + per - 1) / per;
    static const char* wnames[3] = {"ncde_wgrad_plain", "ncde_wgrad_bottom", "ncde_wgrad_top"};
    prof_begin(wnames[lo.mode], stream);
    ncde_wgrad_umma_kernel<<<dim3(gx, (unsigned)gy), WU_THREADS, smem, stream>>>(a, gW, gB, per, gmax);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

}  // namespace ancde
