// Output-layer weight gradient of the fused NCDE solve: the non-recurrent third of the backward.
//
//   dW_out[(i,j)][k] = sum over every (stage n, sample b) of  pre_bar[n,b,(i,j)] * h_last[n,b,k]
//   pre_bar[n,b,(i,j)] = gbar[n,b,i] * dU[n,b,j] * (1 - G[n,b,i,j]^2)          (tanh' of the saved output)
//
// i.e. a GEMM whose reduction dimension is (4*steps*B) rows long (290 816 for the Sepsis batch) and whose
// left operand is produced on the fly from the tanh outputs G the forward saved in HBM (read exactly
// once: this kernel is the main consumer of HBM bandwidth in the backward).  The stage cotangent gbar
// was written over the saved stage input by ncde_bwd_kernel, which therefore runs first.
//
// Tiling: a CTA owns CBLK*128 columns and ALL k; warp = (k group of 8, column block), lane = one
// column group of 4, so a thread accumulates a 4 x 8 register tile.  Rows arrive in chunks of 32
// through a two-stage cp.async (LDGSTS) pipeline; the bias gradient is the extra k column h = 1.
#include "ncde_common.cuh"

namespace ancde {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;     // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int WG_ROWS = 32;   // rows (stage, sample) per pipeline chunk

template <int TS>
__global__ void __launch_bounds__(256) ncde_wgrad_kernel(const __grid_constant__ SolveArgs p, float* __restrict__ gW,
                                                         float* __restrict__ gB, int chunks_per_cta, int NKG, int CBLK)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    constexpr int BPC = WG_ROWS / TS;          // (stage, tile) blocks per chunk
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int kg = warp % NKG, cb = warp / NKG;
    const int Hd = lo.Hd, C4 = lo.C4, NCH = lo.NCH, NCG = lo.NCG, K = lo.K, C = lo.C;
    const int KP = NKG * 8;
    const int CW = CBLK * 32;                  // column groups per CTA
    const int cgl = cb * 32 + lane;            // column group inside the CTA tile
    const int cg = blockIdx.x * CW + cgl;
    const bool cg_ok = cg < NCG;
    const int i_row = cg_ok ? cg / NCH : 0;
    const int chunk = cg_ok ? cg - i_row * NCH : 0;

    // shared memory: two pipeline stages of { G/pre_bar [32][CW*4], gbar [BPC][Hd*TS], dU [BPC][C4*TS], h [BPC][KP*TS] }
    extern __shared__ __align__(16) float smem[];
    const int szP = WG_ROWS * CW * 4, szG = BPC * pad4(Hd * TS), szD = BPC * lo.du_floats, szH = BPC * KP * TS;
    const int stage_floats = szP + szG + szD + szH;

    // rows k >= K of h never get loaded: row K is the bias column (1.0), the rest is padding (0.0)
    for (int s = 0; s < 2; ++s) {
        float* sH = smem + s * stage_floats + szP + szG + szD;
        for (int idx = tid; idx < szH; idx += NT) {
            const int kk = (idx % (KP * TS)) / TS;
            sH[idx] = (kk == K) ? 1.0f : 0.0f;
        }
    }
    __syncthreads();

    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + BPC - 1) / BPC;
    const int64_t c0 = (int64_t)blockIdx.y * chunks_per_cta;
    int64_t c1 = c0 + chunks_per_cta;
    if (c1 > nchunks) c1 = nchunks;
    const int off_hl = lo.act_off_h[lo.n_hidden - 1];

    auto issue = [&](int64_t ch, int s) {
        float* sP = smem + s * stage_floats;
        float* sG = sP + szP;
        float* sD = sG + szG;
        float* sH = sD + szD;
        const int64_t blk0 = ch * BPC;
        // tanh outputs: 32 rows x CW float4, coalesced 16-byte async copies
        for (int piece = tid; piece < WG_ROWS * CW; piece += NT) {
            const int r = piece / CW, c = piece - r * CW;
            const int64_t blk = blk0 + r / TS;
            const int gcg = blockIdx.x * CW + c;
            const bool ok = (blk < nblocks) && (gcg < NCG);
            const float4* src = reinterpret_cast<const float4*>(p.save_G) + ((ok ? blk : 0) * TS + r % TS) * NCG + (ok ? gcg : 0);
            cp_async16(sP + (size_t)piece * 4, src, ok);
        }
        // per-block vectors: stage cotangent (over the saved stage input), dU, last hidden activation
        for (int bl = 0; bl < BPC; ++bl) {
            const int64_t blk = blk0 + bl;
            const bool ok = blk < nblocks;
            const float* act = p.save_act + (ok ? blk : 0) * lo.act_floats;
            if (VW == 4) {
                for (int i = tid; i < Hd * TS / 4; i += NT) cp_async16(sG + bl * pad4(Hd * TS) + i * 4, act + i * 4, ok);
                for (int i = tid; i < lo.du_floats / 4; i += NT) cp_async16(sD + bl * lo.du_floats + i * 4, act + lo.act_off_du + i * 4, ok);
                for (int i = tid; i < K * TS / 4; i += NT) cp_async16(sH + bl * KP * TS + i * 4, act + off_hl + i * 4, ok);
            } else {
                for (int i = tid; i < Hd * TS; i += NT) cp_async4(sG + bl * pad4(Hd * TS) + i, act + i, ok);
                for (int i = tid; i < lo.du_floats; i += NT) cp_async4(sD + bl * lo.du_floats + i, act + lo.act_off_du + i, ok);
                for (int i = tid; i < K * TS; i += NT) cp_async4(sH + bl * KP * TS + i, act + off_hl + i, ok);
            }
        }
    };

    float acc[4][8];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) acc[c][kk] = 0.f;

    if (c0 < c1) issue(c0, 0);
    cp_async_commit();
    if (c0 + 1 < c1) issue(c0 + 1, 1);
    cp_async_commit();

    for (int64_t ch = c0; ch < c1; ++ch) {
        const int s = (int)((ch - c0) & 1);
        float* sP = smem + s * stage_floats;
        const float* sG = sP + szP;
        const float* sD = sG + szG;
        const float* sH = sD + szD;
        cp_async_wait<1>();
        __syncthreads();
        // G -> pre_bar in place: this thread's column group, a share of the rows
        for (int r = kg; r < WG_ROWS; r += NKG) {
            const int bl = r / TS, b = r - bl * TS;
            float4 g = *reinterpret_cast<const float4*>(sP + ((size_t)r * CW + cgl) * 4);
            const float gb = sG[bl * pad4(Hd * TS) + i_row * TS + b];
            const float* du = sD + bl * lo.du_floats + chunk * lo.DUS + b;
            g.x = gb * du[0 * TS] * fmaf(-g.x, g.x, 1.0f);
            g.y = gb * du[1 * TS] * fmaf(-g.y, g.y, 1.0f);
            g.z = gb * du[2 * TS] * fmaf(-g.z, g.z, 1.0f);
            g.w = gb * du[3 * TS] * fmaf(-g.w, g.w, 1.0f);
            *reinterpret_cast<float4*>(sP + ((size_t)r * CW + cgl) * 4) = g;
        }
        __syncthreads();
        // register-tiled accumulate: acc[c][kk] += sum_b pre_bar[b][c] * h[kk][b]
#pragma unroll 1
        for (int bl = 0; bl < BPC; ++bl) {
            float pv[TS][4];
#pragma unroll
            for (int b = 0; b < TS; ++b) {
                const float4 t = *reinterpret_cast<const float4*>(sP + ((size_t)(bl * TS + b) * CW + cgl) * 4);
                pv[b][0] = t.x; pv[b][1] = t.y; pv[b][2] = t.z; pv[b][3] = t.w;
            }
            const float* hp = sH + bl * KP * TS + (kg * 8) * TS;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                float hv[TS];
#pragma unroll
                for (int q = 0; q < NQ; ++q) Vec<VW>::ld(hp + kk * TS + q * VW, hv + q * VW);
#pragma unroll
                for (int b = 0; b < TS; ++b) {
                    acc[0][kk] = fmaf(pv[b][0], hv[b], acc[0][kk]);
                    acc[1][kk] = fmaf(pv[b][1], hv[b], acc[1][kk]);
                    acc[2][kk] = fmaf(pv[b][2], hv[b], acc[2][kk]);
                    acc[3][kk] = fmaf(pv[b][3], hv[b], acc[3][kk]);
                }
            }
        }
        __syncthreads();
        if (ch + 2 < c1) issue(ch + 2, s);
        cp_async_commit();
    }
    cp_async_wait<0>();

    if (cg_ok) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = chunk * 4 + c;
            if (j >= C) continue;
            const int row = i_row * C + j;           // row of the torch (Hd*C, K) weight
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int kr = kg * 8 + kk;
                if (kr < K) atomicAdd(gW + (int64_t)row * K + kr, acc[c][kk]);
                else if (kr == K) atomicAdd(gB + row, acc[c][kk]);
            }
        }
    }
}

template <int TS>
static int launch_wgrad_ts(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    const int NKG = (lo.K + 1 + 7) / 8;
    if (NKG > 8) {
        set_error("ncde_backward: last hidden width %d too large for the wgrad kernel (max 62)", lo.K);
        return ANCDE_EUNSUPPORTED;
    }
    int CBLK = 8 / NKG;
    while (CBLK > 1 && (CBLK - 1) * 32 >= lo.NCG) --CBLK;      // do not over-tile tiny outputs
    const int KP = NKG * 8, CW = CBLK * 32;
    constexpr int BPC = WG_ROWS / TS;
    const size_t stage_floats = (size_t)WG_ROWS * CW * 4 + (size_t)BPC * pad4(lo.Hd * TS) + (size_t)BPC * lo.du_floats +
                                (size_t)BPC * KP * TS;
    const size_t smem = 2 * stage_floats * sizeof(float);
    if (smem > 227 * 1024) {
        set_error("ncde_backward: wgrad kernel needs %zu bytes of shared memory", smem);
        return ANCDE_EUNSUPPORTED;
    }
    const int threads = 32 * NKG * CBLK;
    const int gx = (lo.NCG + CW - 1) / CW;
    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + BPC - 1) / BPC;
    int per_sm = (int)((227 * 1024) / smem);
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int64_t gy = (148 * per_sm + gx - 1) / gx;
    if (gy > nchunks) gy = nchunks;
    if (gy < 1) gy = 1;
    const int per = (int)((nchunks + gy - 1) / gy);
    gy = (nchunks + per - 1) / per;
    static const char* wnames[3] = {"ncde_wgrad_plain", "ncde_wgrad_bottom", "ncde_wgrad_top"};
    ANCDE_CUDA(cudaFuncSetAttribute(ncde_wgrad_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(wnames[lo.mode], stream);
    ncde_wgrad_kernel<TS><<<dim3(gx, (unsigned)gy), threads, smem, stream>>>(a, gW, gB, per, NKG, CBLK);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

int launch_wgrad(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream)
{
    switch (a.lo.TS) {
        case 8: return launch_wgrad_ts<8>(a, gW, gB, stream);
        case 4: return launch_wgrad_ts<4>(a, gW, gB, stream);
        case 2: return launch_wgrad_ts<2>(a, gW, gB, stream);
        default: return launch_wgrad_ts<1>(a, gW, gB, stream);
    }
}

}  // namespace ancde
