// Output-layer weight gradient of the fused NCDE solve on the tensor cores.
//
//   dW_out[(i,j)][k] = sum over every (stage n, sample b) of  pre_bar[n,b,(i,j)] * h_last[n,b,k]
//   pre_bar[n,b,(i,j)] = gbar[n,b,i] * dU[n,b,j] * (1 - G[n,b,i,j]^2)          (tanh' of the saved output)
//
// A GEMM whose reduction dimension is (4*steps*B) rows long (290 816 for the Sepsis batch) and whose left operand
// is produced on the fly from the tanh outputs G the forward saved in HBM (read exactly once: this kernel is the
// main consumer of HBM bandwidth in the backward).  Per chunk of 32 rows a CTA
//   1. receives G (32 x 128 columns), the stage cotangents, dU and the last hidden activation by cp.async,
//   2. turns G into pre_bar and splits it and h into FP16 hi/lo halves in shared memory (scaled by a power of two
//      taken from max |gbar|, which the recurrent backward kernel tracks, so that the cotangents sit in FP16's range; contributions more than 2^-22 below the
//      largest one are rounded away, which is below FP32 round-off of the sum),
//   3. accumulates  D[16 cols per warp][all k] += pre_bar^T h  with mma.sync m16n8k16, three products per tile
//      (hi*hi + lo*hi + hi*lo), FP32 accumulators in registers for the whole kernel (8 warps = 128 columns).
// The bias gradient is the extra k column h = 1.
#include "mma_common.cuh"

namespace ancde {

constexpr int WM_ROWS = 32;      // rows (stage, sample) per pipeline chunk = two k16 steps
constexpr int WM_COLS = 128;     // output-layer columns per CTA = 8 m-tiles
constexpr int WM_CW = WM_COLS / 4;
constexpr int WM_AS = WM_COLS + 8;   // halfs per row of the pre_bar matrices [row][col]
constexpr int WM_HS = WM_ROWS + 8;   // halfs per row of the h matrices [k][row]

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}

template <int TS>
__global__ void __launch_bounds__(256) ncde_wgrad_mma_kernel(const __grid_constant__ SolveArgs p, float* __restrict__ gW,
                                                             float* __restrict__ gB, int chunks_per_cta, int NKT,
                                                             const unsigned* __restrict__ gmax_bits)
{
    constexpr int BPC = WM_ROWS / TS;          // (stage, tile) blocks per chunk
    const NcdeLayout& lo = p.lo;
    constexpr int NT = 256, NWARP = 8;     // fixed launch shape: the row loops below unroll
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int Hd = lo.Hd, C4 = lo.C4, NCH = lo.NCH, NCG = lo.NCG, K = lo.K, C = lo.C;
    const int KP = NKT * 8;
    const int cg = blockIdx.x * WM_CW + lane;  // this lane's column group in the transform pass
    const bool cg_ok = cg < NCG;
    const int i_row = cg_ok ? cg / NCH : 0;
    const int chunk = cg_ok ? cg - i_row * NCH : 0;

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

    // shared memory: two cp.async stages of { G [32][128] f32, gbar [BPC][Hd*TS], dU [BPC][du_floats], h [BPC][K*TS] },
    // then the FP16 operands { pre_bar hi, lo [32][WM_AS], h hi, lo [KP][WM_HS] }
    __shared__ __align__(8) uint64_t ld_bar[2];        // TMA completion of the two load stages (TS >= 4)
    extern __shared__ __align__(16) float smem[];
    const int szP = WM_ROWS * WM_COLS, szG = BPC * pad4(Hd * TS), szD = BPC * lo.du_floats, szH = pad4(BPC * K * TS);
    const int stage_floats = szP + szG + szD + szH;
    __half* sAh = reinterpret_cast<__half*>(smem + 2 * stage_floats);
    __half* sAl = sAh + WM_ROWS * WM_AS;
    __half* sHh = sAl + WM_ROWS * WM_AS;
    __half* sHl = sHh + KP * WM_HS;

    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + BPC - 1) / BPC;
    const int64_t c0 = (int64_t)blockIdx.y * chunks_per_cta;
    int64_t c1 = c0 + chunks_per_cta;
    if (c1 > nchunks) c1 = nchunks;
    const int off_hl = lo.act_off_h[lo.n_hidden - 1];

    const uint32_t smem_base = smem_u32(smem);
    const int r0 = tid / WM_CW, cc0 = tid - r0 * WM_CW;      // the pieces of a thread are rows r0 + m*NWARP of column group cc0
    const int gcg0 = blockIdx.x * WM_CW + cc0;
    const bool col_ok = gcg0 < NCG;
    auto cp16 = [](uint32_t dst, const void* src, bool ok) {
        const int sz = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    };
    auto cp4 = [](uint32_t dst, const void* src, bool ok) {
        const int sz = ok ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    };
    // Loads of a chunk.  TS >= 4: every piece is a contiguous run of >= 16 bytes, so one warp hands them to the TMA
    // engine (bulk copies that complete on the stage's mbarrier) and no thread spends issue slots on address
    // arithmetic; smaller tiles fall back to per-thread cp.async.
    auto issue = [&](int64_t ch, int s) {
        const uint32_t sP = smem_base + (uint32_t)(s * stage_floats) * 4u;
        const uint32_t sG = sP + (uint32_t)szP * 4u;
        const uint32_t sD = sG + (uint32_t)szG * 4u;
        const uint32_t sH = sD + (uint32_t)szD * 4u;
        const int64_t blk0 = ch * BPC;
        if (TS >= 4) {
            if (warp != 0) return;
            int64_t nvalid = nblocks - blk0;                      // (stage, tile) blocks of this chunk that exist
            if (nvalid > BPC) nvalid = BPC;
            const int ncg = (NCG - blockIdx.x * WM_CW < WM_CW) ? NCG - blockIdx.x * WM_CW : WM_CW;
            const uint32_t bytes_g = (uint32_t)ncg * 16u;
            const uint32_t bytes_o = (uint32_t)((Hd * TS + lo.du_floats + K * TS) * 4);
            if (lane == 0) mbar_arrive_expect_tx(ld_bar + s, (uint32_t)nvalid * (TS * bytes_g + bytes_o));
            __syncwarp();
            {
                const int r = lane, bl = r / TS;
                if (bl < nvalid)
                    bulk_g2s(smem + s * stage_floats + (size_t)r * WM_COLS,
                             reinterpret_cast<const float4*>(p.save_G) + ((blk0 + bl) * TS + (r % TS)) * NCG + blockIdx.x * WM_CW, bytes_g,
                             ld_bar + s);
            }
            if (lane < 3 * BPC) {
                const int bl = lane / 3, what = lane - bl * 3;
                if (bl < nvalid) {
                    const float* act = p.save_act + (blk0 + bl) * lo.act_floats;
                    float* base = smem + s * stage_floats + szP;
                    if (what == 0) bulk_g2s(base + bl * pad4(Hd * TS), act + lo.act_off_gbar, (uint32_t)(Hd * TS * 4), ld_bar + s);
                    else if (what == 1) bulk_g2s(base + szG + bl * lo.du_floats, act + lo.act_off_du, (uint32_t)(lo.du_floats * 4), ld_bar + s);
                    else bulk_g2s(base + szG + szD + bl * K * TS, act + off_hl, (uint32_t)(K * TS * 4), ld_bar + s);
                }
            }
            return;
        }
        const float4* gsrc = reinterpret_cast<const float4*>(p.save_G) + (col_ok ? gcg0 : 0);
#pragma unroll
        for (int r = r0; r < WM_ROWS; r += NWARP) {
            const int64_t blk = blk0 + r / TS;
            const bool ok = col_ok && (blk < nblocks);
            cp16(sP + (uint32_t)(r * WM_CW + cc0) * 16u, gsrc + ((ok ? blk : 0) * TS + (r % TS)) * NCG, ok);
        }
        for (int bl = 0; bl < BPC; ++bl) {
            const int64_t blk = blk0 + bl;
            const bool ok = blk < nblocks;
            const float* act = p.save_act + (ok ? blk : 0) * lo.act_floats;
            for (int i = tid; i < Hd * TS; i += NT) cp4(sG + (uint32_t)(bl * pad4(Hd * TS) + i) * 4u, act + lo.act_off_gbar + i, ok);
            for (int i = tid; i < lo.du_floats; i += NT) cp4(sD + (uint32_t)(bl * lo.du_floats + i) * 4u, act + lo.act_off_du + i, ok);
            for (int i = tid; i < K * TS; i += NT) cp4(sH + (uint32_t)(bl * K * TS + i) * 4u, act + off_hl + i, ok);
        }
    };
    if (TS >= 4) {
        if (tid == 0) {
            mbar_init(ld_bar + 0, 1);
            mbar_init(ld_bar + 1, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }

    // rows k >= K of the h operand: k == K is the bias column (1.0), the rest is padding
    for (int idx = tid; idx < KP * WM_HS; idx += NT) {
        const int kk = idx / WM_HS;
        sHh[idx] = __float2half_rn(kk == K ? 1.0f : 0.0f);
        sHl[idx] = __float2half_rn(0.0f);
    }

    float acc[8][4];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[m][e] = 0.f;

    if (c0 < c1) issue(c0, 0);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (c0 + 1 < c1) issue(c0 + 1, 1);
    asm volatile("cp.async.commit_group;\n" ::: "memory");

    // ldmatrix lane addresses: A (pre_bar^T) via .trans from [row][col], B (h) from [k][row]
    const int mat = lane >> 3, lr = lane & 7;
    const uint32_t aAh = smem_u32(sAh) + (uint32_t)((((mat >> 1) * 8 + lr) * WM_AS + warp * 16 + (mat & 1) * 8) * 2);
    const uint32_t aAl = aAh + (uint32_t)(WM_ROWS * WM_AS * 2);
    // B x4: matrices {hi rows 0-7 of the k16 step, hi rows 8-15, lo rows 0-7, lo rows 8-15}; smem rows = k index (n-tile 0)
    const uint32_t aB = smem_u32(sHh) + (uint32_t)((lr * WM_HS + (mat & 1) * 8) * 2) +
                        ((mat >> 1) ? (uint32_t)(KP * WM_HS * 2) : 0u);

    for (int64_t ch = c0; ch < c1; ++ch) {
        const int s = (int)((ch - c0) & 1);
        const float* sP = smem + s * stage_floats;
        const float* sG = sP + szP;
        const float* sD = sG + szG;
        const float* sH = sD + szD;
        if (TS >= 4) mbar_wait(ld_bar + s, (uint32_t)(((ch - c0) >> 1) & 1));
        else asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();
        const int64_t blk_left = nblocks - ch * BPC;            // rows of blocks beyond the end contribute nothing
        // ---- G -> scaled pre_bar, FP16 hi/lo: this lane's column group, rows warp, warp + NKT, ...
#pragma unroll
        for (int rr = 0; rr < WM_ROWS / NWARP; ++rr) {
            const int r = warp + rr * NWARP;
            const int bl = r / TS, b = r - bl * TS;
            const float4 gq = *reinterpret_cast<const float4*>(sP + ((size_t)r * WM_CW + lane) * 4);
            const bool live = bl < blk_left;
            const float gb = live ? sG[bl * pad4(Hd * TS) + i_row * TS + b] * scale : 0.f;
            const float* du = sD + bl * lo.du_floats + chunk * lo.DUS + b;
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
            *reinterpret_cast<uint2*>(sAh + r * WM_AS + lane * 4) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(sAl + r * WM_AS + lane * 4) = make_uint2(l0, l1);
        }
        // ---- h -> FP16 hi/lo [k][row]
#pragma unroll 4
        for (int idx = tid; idx < K * WM_ROWS; idx += NT) {
            const int kk = idx / WM_ROWS, r = idx - kk * WM_ROWS;
            const int bl = r / TS, b = r - bl * TS;
            const float v = (bl < blk_left) ? sH[bl * K * TS + kk * TS + b] : 0.f;
            const __half hi = __float2half_rn(v);
            sHh[kk * WM_HS + r] = hi;
            sHl[kk * WM_HS + r] = __float2half_rn(v - __half2float(hi));
        }
        __syncthreads();
        // ---- D[128 cols][8 k of this warp] += pre_bar^T h over the 32 rows (two k16 steps)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint4 ah, al;
            ldmatrix_x4_trans(aAh + (uint32_t)(ks * 16 * WM_AS * 2), ah.x, ah.y, ah.z, ah.w);
            ldmatrix_x4_trans(aAl + (uint32_t)(ks * 16 * WM_AS * 2), al.x, al.y, al.z, al.w);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (nt < NKT) {
                    uint32_t bh0, bh1, bl0, bl1;
                    ldmatrix_x4(aB + (uint32_t)((nt * 8 * WM_HS + ks * 16) * 2), bh0, bh1, bl0, bl1);
                    mma16x3(acc[nt], ah, al, bh0, bh1, bl0, bl1);
                }
            }
        }
        __syncthreads();
        if (ch + 2 < c1) issue(ch + 2, s);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");

    // ---- results: acc[nt] = {(col g, k 2t), (col g, k 2t+1), (col g+8, k 2t), (col g+8, k 2t+1)}, col relative to this warp's m-tile, k to 8*nt
    if (c0 < c1) {
        const float inv = 1.0f / scale;
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (m >= NKT) continue;
                const int col = blockIdx.x * WM_COLS + warp * 16 + g + (e >> 1) * 8;
                const int kr = m * 8 + 2 * t + (e & 1);
                const int i = col / C4, j = col - i * C4;
                if (i < Hd && j < C) {
                    const int row = i * C + j;           // row of the torch (Hd*C, K) weight
                    if (kr < K) atomicAdd(gW + (int64_t)row * K + kr, acc[m][e] * inv);
                    else if (kr == K) atomicAdd(gB + row, acc[m][e] * inv);
                }
            }
    }
}

template <int TS>
static int launch_wgrad_mma_ts(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    const int NKT = (lo.K + 1 + 7) / 8;
    constexpr int BPC = WM_ROWS / TS;
    const int KP = NKT * 8;
    const size_t stage_floats = (size_t)WM_ROWS * WM_COLS + (size_t)BPC * pad4(lo.Hd * TS) + (size_t)BPC * lo.du_floats +
                                (size_t)pad4(BPC * lo.K * TS);
    const size_t smem = 2 * stage_floats * sizeof(float) + (size_t)(2 * WM_ROWS * WM_AS + 2 * KP * WM_HS) * sizeof(__half);
    if (smem > 227 * 1024) {
        set_error("ncde_backward: wgrad kernel needs %zu bytes of shared memory", smem);
        return ANCDE_EUNSUPPORTED;
    }
    const unsigned* gmax = reinterpret_cast<const unsigned*>(a.wpack + lo.off_gscale);   // written by ncde_bwd_kernel
    const int threads = 256;
    const int gx = (lo.NCP + WM_COLS - 1) / WM_COLS;
    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + BPC - 1) / BPC;
    ANCDE_CUDA(cudaFuncSetAttribute(ncde_wgrad_mma_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    ANCDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ncde_wgrad_mma_kernel<TS>, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int num_sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // exactly one wave: as many row slices as fit next to each other, never more
    int64_t gy = ((int64_t)num_sms * per_sm) / gx;
    if (gy > nchunks) gy = nchunks;
    if (gy < 1) gy = 1;
    const int per = (int)((nchunks + gy - 1) / gy);
    gy = (nchunks + per - 1) / per;
    static const char* wnames[3] = {"ncde_wgrad_plain", "ncde_wgrad_bottom", "ncde_wgrad_top"};
    prof_begin(wnames[lo.mode], stream);
    ncde_wgrad_mma_kernel<TS><<<dim3(gx, (unsigned)gy), threads, smem, stream>>>(a, gW, gB, per, NKT, gmax);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

int launch_wgrad_mma(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream)
{
    switch (a.lo.TS) {
        case 8: return launch_wgrad_mma_ts<8>(a, gW, gB, stream);
        case 4: return launch_wgrad_mma_ts<4>(a, gW, gB, stream);
        case 2: return launch_wgrad_mma_ts<2>(a, gW, gB, stream);
        default: return launch_wgrad_mma_ts<1>(a, gW, gB, stream);
    }
}

}  // namespace ancde
