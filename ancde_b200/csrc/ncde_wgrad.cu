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
#include <stdlib.h>

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

    // cp.async issue with everything that does not depend on the chunk hoisted out: NT = CW * NKG, so the pieces
    // of a thread are the rows r0 + m*NKG of one fixed column group
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    const int r0 = tid / CW, cc0 = tid - r0 * CW;
    const int gcg0 = blockIdx.x * CW + cc0;
    const bool col_ok = gcg0 < NCG;
    auto cp16 = [](uint32_t dst, const void* src, bool ok) {
        const int sz = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    };
    auto cp4 = [](uint32_t dst, const void* src, bool ok) {
        const int sz = ok ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    };
    auto issue = [&](int64_t ch, int s) {
        const uint32_t sP = smem_base + (uint32_t)(s * stage_floats) * 4u;
        const uint32_t sG = sP + (uint32_t)szP * 4u;
        const uint32_t sD = sG + (uint32_t)szG * 4u;
        const uint32_t sH = sD + (uint32_t)szD * 4u;
        const int64_t blk0 = ch * BPC;
        // tanh outputs: 32 rows x CW float4, coalesced 16-byte async copies
        const float4* gsrc = reinterpret_cast<const float4*>(p.save_G) + (col_ok ? gcg0 : 0);
        for (int r = r0; r < WG_ROWS; r += NKG) {
            const int64_t blk = blk0 + r / TS;
            const bool ok = col_ok && (blk < nblocks);
            cp16(sP + (uint32_t)(r * CW + cc0) * 16u, gsrc + ((ok ? blk : 0) * TS + (r % TS)) * NCG, ok);
        }
        // per-block vectors: stage cotangent (over the saved stage input), dU, last hidden activation
        for (int bl = 0; bl < BPC; ++bl) {
            const int64_t blk = blk0 + bl;
            const bool ok = blk < nblocks;
            const float* act = p.save_act + (ok ? blk : 0) * lo.act_floats;
            if (VW == 4) {
                for (int i = tid; i < Hd * TS / 4; i += NT) cp16(sG + (uint32_t)(bl * pad4(Hd * TS) + i * 4) * 4u, act + lo.act_off_gbar + i * 4, ok);
                for (int i = tid; i < lo.du_floats / 4; i += NT) cp16(sD + (uint32_t)(bl * lo.du_floats + i * 4) * 4u, act + lo.act_off_du + i * 4, ok);
                for (int i = tid; i < K * TS / 4; i += NT) cp16(sH + (uint32_t)(bl * KP * TS + i * 4) * 4u, act + off_hl + i * 4, ok);
            } else {
                for (int i = tid; i < Hd * TS; i += NT) cp4(sG + (uint32_t)(bl * pad4(Hd * TS) + i) * 4u, act + lo.act_off_gbar + i, ok);
                for (int i = tid; i < lo.du_floats; i += NT) cp4(sD + (uint32_t)(bl * lo.du_floats + i) * 4u, act + lo.act_off_du + i, ok);
                for (int i = tid; i < K * TS; i += NT) cp4(sH + (uint32_t)(bl * KP * TS + i) * 4u, act + off_hl + i, ok);
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
    ANCDE_CUDA(cudaFuncSetAttribute(ncde_wgrad_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    ANCDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ncde_wgrad_kernel<TS>, threads, smem));
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
    ncde_wgrad_kernel<TS><<<dim3(gx, (unsigned)gy), threads, smem, stream>>>(a, gW, gB, per, NKG, CBLK);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

// ---------------------------------------------------------------------------------------------
// Hidden-layer weight gradients: dW_l[o][k] = sum over (stage, sample) of delta_l[o] * in_l[k], where the
// recurrent backward left the masked layer cotangents delta_l in save_act and in_l is the saved input of
// the layer (stage input for l = 0, previous activation otherwise).  One CTA column per layer
// (blockIdx.x = l), rows split over blockIdx.y; thread = output o x a quarter of the inputs, accumulators
// in registers, rows through a two-stage cp.async pipeline.  Results are added to the torch-layout
// gradients directly.
// ---------------------------------------------------------------------------------------------
struct HiddenGradPtrs { float* gW[ANCDE_MAX_LAYERS]; float* gB[ANCDE_MAX_LAYERS]; };
constexpr int HW_KQ_MAX = 32;  // inputs per thread (4 input quarters x 32 = 128 inputs max)
constexpr int HW_BPC = 4;      // (stage, tile) blocks per pipeline chunk

template <int TS, int HW_KQ>
__global__ void __launch_bounds__(256) ncde_hidden_wgrad_kernel(const __grid_constant__ SolveArgs p,
                                                                const __grid_constant__ HiddenGradPtrs hp, int chunks_per_cta)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    const NcdeLayout& lo = p.lo;
    const int l = blockIdx.x;
    const int N = lo.width[l], Kin = lo.in_dim[l];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int o = tid & 63, kq = tid >> 6;
    const int KQ = (Kin + 3) >> 2;
    const int kbeg = kq * KQ;
    const int off_delta = lo.act_off_delta[l];
    const int off_in = (l == 0) ? 0 : lo.act_off_h[l - 1];

    extern __shared__ __align__(16) float smem[];
    const int szD = pad4(N * TS), szI = pad4(Kin * TS);
    const int blk_floats = szD + szI;
    const int stage_floats = HW_BPC * blk_floats;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);

    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + HW_BPC - 1) / HW_BPC;
    const int64_t c0 = (int64_t)blockIdx.y * chunks_per_cta;
    int64_t c1 = c0 + chunks_per_cta;
    if (c1 > nchunks) c1 = nchunks;

    auto issue = [&](int64_t ch, int s) {
        for (int bl = 0; bl < HW_BPC; ++bl) {
            const int64_t blk = ch * HW_BPC + bl;
            const bool ok = blk < nblocks;
            const float* act = p.save_act + (ok ? blk : 0) * lo.act_floats;
            const uint32_t dD = smem_base + (uint32_t)(s * stage_floats + bl * blk_floats) * 4u;
            const uint32_t dI = dD + (uint32_t)szD * 4u;
            const int sz16 = ok ? 16 : 0, sz4 = ok ? 4 : 0;
            if (VW == 4) {
                for (int i = tid; i < N * TS / 4; i += NT)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dD + i * 16u), "l"(act + off_delta + i * 4), "r"(sz16) : "memory");
                for (int i = tid; i < Kin * TS / 4; i += NT)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dI + i * 16u), "l"(act + off_in + i * 4), "r"(sz16) : "memory");
            } else {
                for (int i = tid; i < N * TS; i += NT)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dD + i * 4u), "l"(act + off_delta + i), "r"(sz4) : "memory");
                for (int i = tid; i < Kin * TS; i += NT)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dI + i * 4u), "l"(act + off_in + i), "r"(sz4) : "memory");
            }
        }
    };

    float acc[HW_KQ];
#pragma unroll
    for (int j = 0; j < HW_KQ; ++j) acc[j] = 0.f;
    float bacc = 0.f;

    if (c0 < c1) issue(c0, 0);
    cp_async_commit();
    if (c0 + 1 < c1) issue(c0 + 1, 1);
    cp_async_commit();
    for (int64_t ch = c0; ch < c1; ++ch) {
        const int s = (int)((ch - c0) & 1);
        cp_async_wait<1>();
        __syncthreads();
        if (o < N) {
#pragma unroll 1
            for (int bl = 0; bl < HW_BPC; ++bl) {
                const float* sD = smem + s * stage_floats + bl * blk_floats;
                const float* sI = sD + szD;
                float dv[TS];
#pragma unroll
                for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sD + o * TS + q * VW, dv + q * VW);
                if (kq == 0) {
#pragma unroll
                    for (int b = 0; b < TS; ++b) bacc += dv[b];
                }
#pragma unroll
                for (int j = 0; j < HW_KQ; ++j) {
                    const int kk = kbeg + j;
                    if (j < KQ && kk < Kin) {
                        float iv[TS];
#pragma unroll
                        for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sI + kk * TS + q * VW, iv + q * VW);
#pragma unroll
                        for (int b = 0; b < TS; ++b) acc[j] = fmaf(dv[b], iv[b], acc[j]);
                    }
                }
            }
        }
        __syncthreads();
        if (ch + 2 < c1) issue(ch + 2, s);
        cp_async_commit();
    }
    cp_async_wait<0>();
    if (o < N && c0 < c1) {
#pragma unroll
        for (int j = 0; j < HW_KQ; ++j) {
            const int kk = kbeg + j;
            if (j < KQ && kk < Kin) atomicAdd(hp.gW[l] + (int64_t)o * Kin + kk, acc[j]);
        }
        if (kq == 0) atomicAdd(hp.gB[l] + o, bacc);
    }
}

template <int TS>
static int launch_hidden_wgrad_ts(const SolveArgs& a, const HiddenGradPtrs& hp, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    size_t smem = 0;
    for (int l = 0; l < lo.n_hidden; ++l) {
        if (lo.width[l] > 64 || lo.in_dim[l] > 4 * HW_KQ_MAX) {
            set_error("ncde_backward: hidden layer %d is %dx%d; the hidden weight-gradient kernel handles up to 64 outputs x 128 inputs",
                      l, lo.width[l], lo.in_dim[l]);
            return ANCDE_EUNSUPPORTED;
        }
        const size_t f = 2 * (size_t)HW_BPC * (pad4(lo.width[l] * TS) + pad4(lo.in_dim[l] * TS)) * sizeof(float);
        if (f > smem) smem = f;
    }
    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + HW_BPC - 1) / HW_BPC;
    int64_t gy = (148 * 4 + lo.n_hidden - 1) / lo.n_hidden;
    if (gy > nchunks) gy = nchunks;
    if (gy < 1) gy = 1;
    const int per = (int)((nchunks + gy - 1) / gy);
    gy = (nchunks + per - 1) / per;
    int kmax = 0;
    for (int l = 0; l < lo.n_hidden; ++l) kmax = lo.in_dim[l] > kmax ? lo.in_dim[l] : kmax;
    static const char* names[3] = {"ncde_hidden_wgrad_plain", "ncde_hidden_wgrad_bottom", "ncde_hidden_wgrad_top"};
    if (kmax <= 64) {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_hidden_wgrad_kernel<TS, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_hidden_wgrad_kernel<TS, 16><<<dim3(lo.n_hidden, (unsigned)gy), 256, smem, stream>>>(a, hp, per);
    } else {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_hidden_wgrad_kernel<TS, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_hidden_wgrad_kernel<TS, 32><<<dim3(lo.n_hidden, (unsigned)gy), 256, smem, stream>>>(a, hp, per);
    }
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

int launch_hidden_wgrad(const SolveArgs& a, const HiddenGradPtrs& hp, cudaStream_t stream)
{
    switch (a.lo.TS) {
        case 8: return launch_hidden_wgrad_ts<8>(a, hp, stream);
        case 4: return launch_hidden_wgrad_ts<4>(a, hp, stream);
        case 2: return launch_hidden_wgrad_ts<2>(a, hp, stream);
        default: return launch_hidden_wgrad_ts<1>(a, hp, stream);
    }
}

int launch_wgrad_mma(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream);   // ncde_wgrad_mma.cu

int launch_wgrad(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream)
{
    static int ffma = -1;        // development switch: ANCDE_WGRAD_FFMA=1 selects the FP32-FMA kernel of this file
    if (ffma < 0) {
        const char* e = getenv("ANCDE_WGRAD_FFMA");
        ffma = (e && e[0] == '1') ? 1 : 0;
    }
    if (!ffma) return launch_wgrad_mma(a, gW, gB, stream);
    switch (a.lo.TS) {
        case 8: return launch_wgrad_ts<8>(a, gW, gB, stream);
        case 4: return launch_wgrad_ts<4>(a, gW, gB, stream);
        case 2: return launch_wgrad_ts<2>(a, gW, gB, stream);
        default: return launch_wgrad_ts<1>(a, gW, gB, stream);
    }
}

}  // namespace ancde
