// Backprop through the fused NCDE solve (discrete adjoint of the RK4 3/8 recurrence).
//
// Three kernels:
//  * ncde_bwd_kernel   -- the recurrent part.  One persistent CTA per tile of TS samples walks the
//    steps backwards; per RK4 stage it turns the cotangent of the stage derivative into the
//    cotangent of the stage input (a vector-Jacobian product through tanh, the output layer and
//    the hidden ReLU layers) using the activations and tanh outputs the forward saved in HBM
//    (nothing is recomputed).  Hidden-layer weight gradients accumulate in shared memory; the
//    cotangent of the attention (top NCDE, backprop mode) is reduced in shared memory and added
//    to grad_attention.  The stage cotangent is written back over the saved stage input.
//  * ncde_wgrad_kernel -- the non-recurrent part: dW_out[col][k] = sum over all (stage, sample)
//    of pre_bar[col] * h_last[k] as a register-tiled FP32 GEMM over the saved tanh outputs
//    (read exactly once), with the bias gradient folded in as an extra k column.
//  * unpack_hidden_grads_kernel -- packed (transposed) hidden gradients -> torch layout.
//
// Matches autograd through torchdiffeq's rk4_alt_step_func + the reference vector fields
// (cdeint_module.py:25-32,56-72,103-138; vector_fields.py:205-218,237-250) to FP32 round-off.
#include "ncde_common.cuh"

namespace ancde {

int fill_args(const ancde_ncde_desc* d, const NcdeLayout& lo, SolveArgs* a);
int pack_weights(const ancde_ncde_desc* d, const NcdeLayout& lo, cudaStream_t stream);

// Sum of v[idx] over the 32 lanes of a warp for V values per lane in V/2 + V/4 + ... shuffles:
// afterwards lane l holds the totals of idx = l*(V/32) .. +V/32-1 in v[0 .. V/32-1]  (V >= 32), or
// v[0] = total of idx = l / (32/V) (V < 32).
template <int V>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[V], int lane)
{
    int mask = 16;
#pragma unroll
    for (int half = V / 2; half >= 1; half /= 2) {
        if (mask >= 1) {
            const bool up = (lane & mask) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float send = up ? v[i] : v[i + half];
                const float keep = up ? v[i + half] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
            }
        }
        mask >>= 1;
    }
    if (V < 32) {
#pragma unroll
        for (int m = 16 / V; m >= 1; m >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
    }
}

template <int TS, bool WO_SMEM>
__global__ void __launch_bounds__(512, 1) ncde_bwd_kernel(const __grid_constant__ SolveArgs p)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int b0 = tile * TS;
    const int Hd = lo.Hd, C4 = lo.C4, NCH = lo.NCH, NCG = lo.NCG, NCP = lo.NCP, K = lo.K;
    const int NKG = (K + 7) / 8;
    int NSPLIT = (NT / 32) / NKG;
    if (NSPLIT < 1) NSPLIT = 1;
    const int HT = pad4(Hd * TS);
    const bool need_att = (lo.mode == ANCDE_MODE_TOP) && (p.grad_attention != nullptr);

    extern __shared__ __align__(16) float smem[];
    float* sHW = smem;                                   // hidden weights (packed, transposed)
    float* sGW = sHW + pad4(lo.hidden_floats);           // hidden weight-gradient accumulators
    float* sp = sGW + pad4(lo.hidden_floats);
    float* sWo = sp;
    if (WO_SMEM) sp += (size_t)K * NCP + NCP;
    float* sPre = sp;   sp += (size_t)NCP * TS;          // [q][col][VW]
    float* sAct = sp;   sp += lo.act_floats;             // saved activations of the stage
    float* sG = sp;     sp += HT;                        // cotangent of the stage derivative
    float* sLam = sp;   sp += HT;                        // cotangent of z at the end of the step
    float* sMu = sp;    sp += HT;                        // pending cotangent of z at the start
    float* sV4 = sp;    sp += HT;
    float* sV3 = sp;    sp += HT;
    float* sV2 = sp;    sp += HT;
    float* sHb0 = sp;   sp += pad4(lo.Hmax * TS);
    float* sHb1 = sp;   sp += pad4(lo.Hmax * TS);
    float* sHbPart = sp; sp += (size_t)NSPLIT * NKG * 8 * TS;
    float* sTmp = sp;                                    // [C4][TS]

    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack);
        float4* dst = reinterpret_cast<float4*>(sHW);
        float4* dz = reinterpret_cast<float4*>(sGW);
        for (int i = tid; i < pad4(lo.hidden_floats) / 4; i += NT) {
            dst[i] = __ldg(src + i);
            dz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (WO_SMEM) {
            const float4* s2 = reinterpret_cast<const float4*>(p.wpack + lo.off_wo);
            float4* d2 = reinterpret_cast<float4*>(sWo);
            for (int i = tid; i < (K * NCP + NCP) / 4; i += NT) d2[i] = __ldg(s2 + i);
        }
        for (int item = tid; item < Hd * TS; item += NT) { sLam[item] = 0.f; sMu[item] = 0.f; }
    }
    __syncthreads();
    const float* Wo = WO_SMEM ? sWo : (p.wpack + lo.off_wo);

    int jb = lo.n_out - 1;   // last output time whose cotangent has not been consumed yet

    for (int k = lo.n_steps - 1; k >= 0; --k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step);
        const float dt = __fsub_rn(t1, t0);
        // ---- cotangents of the outputs emitted during this step: t_out[j] in (t0, t1]
        {
            int j = jb;
            for (int item = tid; item < Hd * TS; item += NT) {
                const int i = item / TS, b = item - i * TS;
                const int bg = b0 + b;
                float lam = sLam[item], mu = sMu[item];
                for (j = jb; j >= 1; --j) {
                    const float tj = __ldg(p.t_out + j);
                    if (!(tj > t0)) break;
                    const float g = (bg < lo.B) ? __ldg(p.grad_z_out + ((int64_t)j * lo.B + bg) * Hd + i) : 0.f;
                    if (tj == t1) lam += g;
                    else {
                        const float th = (tj - t0) / (t1 - t0);
                        lam += th * g;
                        mu += (1.0f - th) * g;
                    }
                }
                sLam[item] = lam;
                sMu[item] = mu;
            }
            while (jb >= 1 && __ldg(p.t_out + jb) > t0) --jb;
        }
        __syncthreads();

#pragma unroll 1
        for (int s = 3; s >= 0; --s) {
            const int n = 4 * k + s;
            float* act = p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats;
            // ---- cotangent of the stage derivative k_{s+1}  (transpose of rk4_alt_step_func)
            for (int item = tid; item < Hd * TS; item += NT) {
                const float lam = sLam[item];
                float g;
                if (s == 3) g = dt * 0.125f * lam;
                else if (s == 2) g = dt * (0.375f * lam + sV4[item]);
                else if (s == 1) g = dt * (0.375f * lam - sV4[item] + sV3[item]);
                else g = dt * (0.125f * lam + sV4[item] + (sV2[item] - sV3[item]) * (1.0f / 3.0f));
                sG[item] = g;
            }
            // ---- saved activations of this stage -> shared memory
            {
                const float4* src = reinterpret_cast<const float4*>(act);
                float4* dst = reinterpret_cast<float4*>(sAct);
                for (int i = tid; i < lo.act_floats / 4; i += NT) dst[i] = __ldg(src + i);
            }
            __syncthreads();
            // the stage cotangent replaces the saved stage input in HBM (read by the wgrad kernel);
            // the copy in sAct stays valid for the first hidden layer's weight gradient
            for (int item = tid; item < Hd * TS; item += NT) act[item] = sG[item];

            const float* sDU = sAct + lo.act_off_du;
            // ---- phase 1: tanh outputs from HBM, pre-activation cotangent into shared memory
            const float4* gbase = reinterpret_cast<const float4*>(p.save_G) + (((int64_t)n * lo.ntiles + tile) * TS) * NCG;
            if (need_att) {
                // cotangent of dU: dUbar[j][b] = sum_i gbar[i][b] * G[i][j][b]; the products go through
                // shared memory (same layout as pre_bar) and are reduced over the rows i
                for (int cg = tid; cg < NCG; cg += NT) {
                    const int i = cg / NCH;
                    float gb[TS];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sG + i * TS + q * VW, gb + q * VW);
                    float g[TS][4];
#pragma unroll
                    for (int b = 0; b < TS; ++b) {
                        const float4 t = __ldg(gbase + (int64_t)b * NCG + cg);
                        g[b][0] = t.x; g[b][1] = t.y; g[b][2] = t.z; g[b][3] = t.w;
                    }
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float v[VW];
#pragma unroll
                            for (int e = 0; e < VW; ++e) v[e] = gb[q * VW + e] * g[q * VW + e][c];
                            Vec<VW>::st(sPre + ((size_t)q * NCP + cg * 4 + c) * VW, v);
                        }
                }
                __syncthreads();
                for (int item = tid; item < C4 * NQ; item += NT) {
                    const int q = item / C4, j = item - q * C4;
                    float acc[VW];
#pragma unroll
                    for (int e = 0; e < VW; ++e) acc[e] = 0.f;
                    for (int ii = 0; ii < Hd; ++ii) {
                        float v[VW];
                        Vec<VW>::ld(sPre + ((size_t)q * NCP + ii * C4 + j) * VW, v);
#pragma unroll
                        for (int e = 0; e < VW; ++e) acc[e] += v[e];
                    }
                    float du[VW];
                    Vec<VW>::ld(sAct + lo.act_off_duda + j * TS + q * VW, du);
#pragma unroll
                    for (int e = 0; e < VW; ++e) acc[e] *= du[e];
                    Vec<VW>::st(sTmp + j * TS + q * VW, acc);
                }
                __syncthreads();
            }
            for (int cg = tid; cg < NCG; cg += NT) {
                const int i = cg / NCH, chunk = cg - i * NCH;
                float gb[TS];
#pragma unroll
                for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sG + i * TS + q * VW, gb + q * VW);
                float g[TS][4];
#pragma unroll
                for (int b = 0; b < TS; ++b) {
                    const float4 t = __ldg(gbase + (int64_t)b * NCG + cg);
                    g[b][0] = t.x; g[b][1] = t.y; g[b][2] = t.z; g[b][3] = t.w;
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float du[VW], v[VW];
                        Vec<VW>::ld(sDU + (chunk * 4 + c) * TS + q * VW, du);
#pragma unroll
                        for (int e = 0; e < VW; ++e) {
                            const float gg = g[q * VW + e][c];
                            v[e] = gb[q * VW + e] * du[e] * fmaf(-gg, gg, 1.0f);
                        }
                        Vec<VW>::st(sPre + ((size_t)q * NCP + cg * 4 + c) * VW, v);
                    }
            }
            __syncthreads();
            // ---- cotangent of the attention: a[q] += dUbar * dU/da   
            if (need_att) {
                const float ts = stage_time(t0, dt, s);
                int qa = (int)floorf(ts) - 1;
                if (qa < 0) qa += lo.att_L;
                qa = qa < 0 ? 0 : (qa >= lo.att_L ? lo.att_L - 1 : qa);
                if (lo.att_C == 1) {
                    if (tid < TS && b0 + tid < lo.B) {
                        float acc = 0.f;
                        for (int j = 0; j < lo.C; ++j) acc += sTmp[j * TS + tid];
                        atomicAdd(p.grad_attention + (int64_t)qa * lo.B + b0 + tid, acc);
                    }
                } else {
                    for (int item = tid; item < lo.C * TS; item += NT) {
                        const int b = item / lo.C, j = item - b * lo.C;
                        if (b0 + b < lo.B)
                            atomicAdd(p.grad_attention + ((int64_t)qa * lo.B + b0 + b) * lo.C + j, sTmp[j * TS + b]);
                    }
                }
            }
            // ---- phase 2: hbar[k][b] = sum_col pre_bar[col][b] * Wt_out[k][col]
            if (warp < NKG * NSPLIT) {
                const int kg = warp % NKG, split = warp / NKG;
                float acc[8 * TS];
#pragma unroll
                for (int i = 0; i < 8 * TS; ++i) acc[i] = 0.f;
                for (int col = split * 32 + lane; col < NCP; col += 32 * NSPLIT) {
                    float pv[TS];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sPre + ((size_t)q * NCP + col) * VW, pv + q * VW);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const int kr = kg * 8 + kk;
                        float w = 0.f;
                        if (kr < K) w = WO_SMEM ? Wo[(size_t)kr * NCP + col] : __ldg(Wo + (size_t)kr * NCP + col);
#pragma unroll
                        for (int b = 0; b < TS; ++b) acc[kk * TS + b] = fmaf(w, pv[b], acc[kk * TS + b]);
                    }
                }
                warp_reduce_scatter<8 * TS>(acc, lane);
                constexpr int V = 8 * TS;
                if constexpr (V >= 32) {
#pragma unroll
                    for (int r = 0; r < V / 32; ++r) {
                        const int idx = lane * (V / 32) + r;
                        sHbPart[((size_t)split * NKG + kg) * V + idx] = acc[r];
                    }
                } else {
                    constexpr int per = 32 / V;
                    if (lane % per == 0) sHbPart[((size_t)split * NKG + kg) * V + lane / per] = acc[0];
                }
            }
            __syncthreads();
            // ---- phase 3: hidden layers backwards
            float* hb = sHb0;
            float* hb_next = sHb1;
            for (int l = lo.n_hidden - 1; l >= 0; --l) {
                const int N = lo.width[l], Kin = lo.in_dim[l], S = lo.stride[l];
                const float* hl = sAct + lo.act_off_h[l];
                // delta = hbar * relu'(h)
                for (int item = tid; item < N * TS; item += NT) {
                    float v;
                    if (l == lo.n_hidden - 1) {
                        const int o = item / TS, b = item - o * TS;
                        v = 0.f;
                        for (int sp2 = 0; sp2 < NSPLIT; ++sp2)
                            v += sHbPart[((size_t)sp2 * NKG + o / 8) * (8 * TS) + (o % 8) * TS + b];
                    } else {
                        v = hb[item];
                    }
                    hb[item] = (hl[item] > 0.f) ? v : 0.f;
                }
                __syncthreads();
                const float* inl = (l == 0) ? sAct : sAct + lo.act_off_h[l - 1];
                // weight / bias gradient accumulators: dWt[k][o] += sum_b delta[o][b] * in[k][b]
                for (int item = tid; item < Kin * N; item += NT) {
                    const int kk = item / N, o = item - kk * N;
                    float acc = 0.f;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        float dv[VW], iv[VW];
                        Vec<VW>::ld(hb + o * TS + q * VW, dv);
                        Vec<VW>::ld(inl + kk * TS + q * VW, iv);
#pragma unroll
                        for (int e = 0; e < VW; ++e) acc = fmaf(dv[e], iv[e], acc);
                    }
                    sGW[lo.off_w[l] + kk * S + o] += acc;
                }
                for (int o = tid; o < N; o += NT) {
                    float acc = 0.f;
#pragma unroll
                    for (int b = 0; b < TS; ++b) acc += hb[o * TS + b];
                    sGW[lo.off_b[l] + o] += acc;
                }
                // cotangent of the layer input: hb_next[k][b] = sum_o Wt[k][o] * delta[o][b]
                const float* Wt = sHW + lo.off_w[l];
                for (int item = tid; item < Kin * NQ; item += NT) {
                    const int q = item / Kin, kk = item - q * Kin;
                    float acc[VW];
#pragma unroll
                    for (int e = 0; e < VW; ++e) acc[e] = 0.f;
                    const float* wp = Wt + kk * S;
#pragma unroll 4
                    for (int o = 0; o < N; ++o) {
                        const float w = wp[o];
                        float dv[VW];
                        Vec<VW>::ld(hb + o * TS + q * VW, dv);
#pragma unroll
                        for (int e = 0; e < VW; ++e) acc[e] = fmaf(w, dv[e], acc[e]);
                    }
                    Vec<VW>::st(hb_next + kk * TS + q * VW, acc);
                }
                __syncthreads();
                float* t = hb; hb = hb_next; hb_next = t;
            }
            // hb now holds the cotangent of the stage input
            for (int item = tid; item < Hd * TS; item += NT) {
                const float v = hb[item];
                if (s == 3) sV4[item] = v;
                else if (s == 2) sV3[item] = v;
                else if (s == 1) sV2[item] = v;
                else {
                    sLam[item] = sLam[item] + sV4[item] + sV3[item] + sV2[item] + v + sMu[item];
                    sMu[item] = 0.f;
                }
            }
            __syncthreads();
        }
    }
    // ---- results
    for (int item = tid; item < Hd * TS; item += NT) {
        const int i = item / TS, b = item - i * TS;
        const int bg = b0 + b;
        if (bg < lo.B) p.grad_z0[(int64_t)bg * Hd + i] = sLam[item] + __ldg(p.grad_z_out + (int64_t)bg * Hd + i);
    }
    for (int i = tid; i < lo.hidden_floats; i += NT) {
        const float v = sGW[i];
        if (v != 0.f) atomicAdd(p.grad_hidden + i, v);
    }
}

// ---------------------------------------------------------------------------------------------
// Output-layer weight gradient: a GEMM whose reduction runs over every (stage, tile, sample) row.
// CTA = 32 column groups (128 columns) x all k; thread = 4 columns x 8 k; rows are staged through
// shared memory in chunks of one (stage, tile) block at a time.
// ---------------------------------------------------------------------------------------------
constexpr int WG_CG = 32;        // column groups per CTA
constexpr int WG_MAXKG = 16;     // k groups of 8 (K + 1 bias column <= 128)

template <int TS>
__global__ void __launch_bounds__(512) ncde_wgrad_kernel(const __grid_constant__ SolveArgs p, float* gW, float* gB,
                                                         int blocks_per_cta)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, kg = tid >> 5;
    const int Hd = lo.Hd, C4 = lo.C4, NCH = lo.NCH, NCG = lo.NCG, K = lo.K, C = lo.C;
    const int NKG = (K + 1 + 7) / 8;        // + the bias column
    const int KP = NKG * 8;
    const int cg0 = blockIdx.x * WG_CG;
    const int cg = cg0 + lane;
    const bool cg_ok = cg < NCG;
    const int i_row = cg_ok ? cg / NCH : 0;
    const int chunk = cg_ok ? cg - i_row * NCH : 0;

    extern __shared__ __align__(16) float smem[];
    float* sP = smem;                         // [TS][WG_CG*4] pre_bar of this block
    float* sH = sP + TS * WG_CG * 4;          // [TS][KP] last hidden activation (+1.0 bias column)
    float* sGb = sH + TS * KP;                // [Hd][TS] stage cotangent
    float* sDU = sGb + pad4(Hd * TS);         // [C4][TS]

    float acc[4][8];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) acc[c][kk] = 0.f;

    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t blk0 = (int64_t)blockIdx.y * blocks_per_cta;
    int64_t blk1 = blk0 + blocks_per_cta;
    if (blk1 > nblocks) blk1 = nblocks;
    const float* hl_off = nullptr;
    for (int64_t blk = blk0; blk < blk1; ++blk) {
        const float* act = p.save_act + blk * lo.act_floats;
        // stage cotangent, dU, last hidden activation
        for (int item = tid; item < Hd * TS; item += NT) sGb[item] = __ldg(act + item);
        for (int item = tid; item < C4 * TS; item += NT) sDU[item] = __ldg(act + lo.act_off_du + item);
        hl_off = act + lo.act_off_h[lo.n_hidden - 1];
        for (int item = tid; item < TS * KP; item += NT) {
            const int b = item / KP, kk = item - b * KP;
            float v = 0.f;
            if (kk < K) v = __ldg(hl_off + kk * TS + b);
            else if (kk == K) v = 1.0f;
            sH[item] = v;
        }
        __syncthreads();
        // pre_bar for this CTA's 128 columns (warp 0..: every warp computes a share of the samples)
        for (int b = kg; b < TS; b += NT / 32) {
            float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cg_ok) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(p.save_G) + (blk * TS + b) * NCG + cg);
                const float gb = sGb[i_row * TS + b];
                pv.x = gb * sDU[(chunk * 4 + 0) * TS + b] * fmaf(-g.x, g.x, 1.0f);
                pv.y = gb * sDU[(chunk * 4 + 1) * TS + b] * fmaf(-g.y, g.y, 1.0f);
                pv.z = gb * sDU[(chunk * 4 + 2) * TS + b] * fmaf(-g.z, g.z, 1.0f);
                pv.w = gb * sDU[(chunk * 4 + 3) * TS + b] * fmaf(-g.w, g.w, 1.0f);
            }
            *reinterpret_cast<float4*>(sP + (b * WG_CG + lane) * 4) = pv;
        }
        __syncthreads();
        if (kg < NKG) {
#pragma unroll
            for (int b = 0; b < TS; ++b) {
                const float4 pv = *reinterpret_cast<const float4*>(sP + (b * WG_CG + lane) * 4);
                const float4 h0 = *reinterpret_cast<const float4*>(sH + b * KP + kg * 8);
                const float4 h1 = *reinterpret_cast<const float4*>(sH + b * KP + kg * 8 + 4);
                const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                const float pc[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) acc[c][kk] = fmaf(pc[c], hv[kk], acc[c][kk]);
            }
        }
        __syncthreads();
    }
    (void)NQ;
    if (cg_ok && kg < NKG) {
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

struct UnpackArgs {
    NcdeLayout lo;
    const float* packed;
    float* gW[ANCDE_MAX_LAYERS];
    float* gB[ANCDE_MAX_LAYERS];
};

__global__ void unpack_hidden_grads_kernel(const __grid_constant__ UnpackArgs a)
{
    const NcdeLayout& lo = a.lo;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int l = 0; l < lo.n_hidden; ++l) {
        const int in = lo.in_dim[l], out = lo.width[l], S = lo.stride[l];
        for (int idx = gtid; idx < in * out; idx += gsz) {
            const int o = idx / in, k = idx - o * in;
            a.gW[l][idx] += a.packed[lo.off_w[l] + k * S + o];
        }
        for (int o = gtid; o < out; o += gsz) a.gB[l][o] += a.packed[lo.off_b[l] + o];
    }
}

static int bwd_threads(const NcdeLayout& lo)
{
    int per = (lo.NCG + 511) / 512;
    int nt = round_up((lo.NCG + per - 1) / per, 32);
    const int NKG = (lo.K + 7) / 8;
    if (nt < NKG * 32) nt = NKG * 32;
    if (nt < 64) nt = 64;
    return nt;
}

size_t bwd_smem_floats(const NcdeLayout& lo, bool wo_smem, int nt)
{
    const int NKG = (lo.K + 7) / 8;
    int nsplit = (nt / 32) / NKG;
    if (nsplit < 1) nsplit = 1;
    size_t f = 2 * (size_t)pad4(lo.hidden_floats);
    if (wo_smem) f += (size_t)lo.K * lo.NCP + lo.NCP;
    f += (size_t)lo.NCP * lo.TS + lo.act_floats + 6 * (size_t)pad4(lo.Hd * lo.TS) + 2 * (size_t)pad4(lo.Hmax * lo.TS) +
         (size_t)nsplit * NKG * 8 * lo.TS + (size_t)lo.C4 * lo.TS;
    return f;
}

// Used by make_layout's tile-size choice: does the backward fit with this many samples per CTA?
bool bwd_fits(const NcdeLayout& lo)
{
    const int nt = bwd_threads(lo);
    return nt <= 512 && bwd_smem_floats(lo, false, nt) * 4 <= 227 * 1024;
}

template <int TS>
static int launch_bwd(const ancde_ncde_desc* d, SolveArgs& a, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    const int nt = bwd_threads(lo);
    const size_t max_smem = 227 * 1024;
    if (nt > 512) {
        set_error("ncde_backward: needs %d threads per CTA (max 512)", nt);
        return ANCDE_EUNSUPPORTED;
    }
    const bool wo_smem = bwd_smem_floats(lo, true, nt) * 4 <= max_smem;
    const size_t smem = bwd_smem_floats(lo, wo_smem, nt) * 4;
    if (smem > max_smem) {
        set_error("ncde_backward: needs %zu bytes of shared memory per CTA (max %zu)", smem, max_smem);
        return ANCDE_EUNSUPPORTED;
    }
    // hidden-gradient scratch lives behind the packed weights
    a.grad_hidden = d->wpack + lo.wpack_floats;
    ANCDE_CUDA(cudaMemsetAsync(a.grad_hidden, 0, sizeof(float) * pad4(lo.hidden_floats), stream));
    static const char* names[3] = {"ncde_bwd_plain", "ncde_bwd_bottom", "ncde_bwd_top"};
    static const char* wnames[3] = {"ncde_wgrad_plain", "ncde_wgrad_bottom", "ncde_wgrad_top"};
    if (wo_smem) {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_bwd_kernel<TS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_bwd_kernel<TS, true><<<lo.ntiles, nt, smem, stream>>>(a);
    } else {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_bwd_kernel<TS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_bwd_kernel<TS, false><<<lo.ntiles, nt, smem, stream>>>(a);
    }
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());

    // output-layer weight gradient
    {
        const int NKG = (lo.K + 1 + 7) / 8;
        if (NKG > WG_MAXKG) {
            set_error("ncde_backward: last hidden width %d too large for the wgrad kernel", lo.K);
            return ANCDE_EUNSUPPORTED;
        }
        const int KP = NKG * 8;
        const int nthreads = NKG * 32 < 64 ? 64 : NKG * 32;
        const int gx = (lo.NCG + WG_CG - 1) / WG_CG;
        const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
        int64_t gy = (148 * 4 + gx - 1) / gx;
        if (gy > nblocks) gy = nblocks;
        const int per = (int)((nblocks + gy - 1) / gy);
        gy = (nblocks + per - 1) / per;
        const size_t sm = sizeof(float) * ((size_t)TS * WG_CG * 4 + (size_t)TS * KP + pad4(lo.Hd * TS) + (size_t)lo.C4 * TS);
        prof_begin(wnames[lo.mode], stream);
        ncde_wgrad_kernel<TS><<<dim3(gx, (unsigned)gy), nthreads, sm, stream>>>(a, d->grad_W[lo.n_hidden],
                                                                               d->grad_bias[lo.n_hidden], per);
        prof_end(stream);
        count_launch();
        ANCDE_CUDA(cudaGetLastError());
    }
    {
        UnpackArgs ua;
        ua.lo = lo;
        ua.packed = a.grad_hidden;
        for (int l = 0; l < lo.n_hidden; ++l) { ua.gW[l] = d->grad_W[l]; ua.gB[l] = d->grad_bias[l]; }
        prof_begin("unpack_hidden_grads", stream);
        unpack_hidden_grads_kernel<<<32, 256, 0, stream>>>(ua);
        prof_end(stream);
        count_launch();
        ANCDE_CUDA(cudaGetLastError());
    }
    return ANCDE_OK;
}

}  // namespace ancde

using namespace ancde;

extern "C" int ancde_ncde_backward(const ancde_ncde_desc* d, void* stream_)
{
    reset_launch_count();
    cudaStream_t stream = (cudaStream_t)stream_;
    NcdeLayout lo;
    int rc = make_layout(d, &lo);
    if (rc != ANCDE_OK) return rc;
    ANCDE_CHECK_ARG(d->save_act && d->save_G, "ncde_backward: needs the save_act / save_G buffers of the forward");
    ANCDE_CHECK_ARG(d->grad_z_out && d->grad_z0, "ncde_backward: null grad_z_out / grad_z0");
    for (int l = 0; l <= d->n_hidden; ++l)
        ANCDE_CHECK_ARG(d->grad_W[l] && d->grad_bias[l], "ncde_backward: null gradient buffer for layer %d", l);
    SolveArgs a;
    rc = fill_args(d, lo, &a);
    if (rc != ANCDE_OK) return rc;
    rc = pack_weights(d, lo, stream);   // the packed weights of the forward may have been reused
    if (rc != ANCDE_OK) return rc;
    switch (lo.TS) {
        case 8: return launch_bwd<8>(d, a, stream);
        case 4: return launch_bwd<4>(d, a, stream);
        case 2: return launch_bwd<2>(d, a, stream);
        default: return launch_bwd<1>(d, a, stream);
    }
}
