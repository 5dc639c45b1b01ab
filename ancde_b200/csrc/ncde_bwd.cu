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
//    Data movement: the saved activations of the next stage arrive by cp.async into a second
//    buffer, the next tile of tanh outputs is pulled into L2 with a bulk prefetch, and an output
//    layer that does not fit in shared memory is streamed through a TMA ring (k-contiguous copy).
//  * ncde_wgrad_kernel (ncde_wgrad.cu) -- the non-recurrent output-layer weight gradient.
//  * unpack_hidden_grads_kernel -- packed (transposed) hidden gradients -> torch layout.
//
// Matches autograd through torchdiffeq's rk4_alt_step_func + the reference vector fields
// (cdeint_module.py:25-32,56-72,103-138; vector_fields.py:205-218,237-250) to FP32 round-off.
#include "ncde_common.cuh"
#include "pipeline.cuh"

namespace ancde {

int fill_args(const ancde_ncde_desc* d, const NcdeLayout& lo, SolveArgs* a);
int pack_weights(const ancde_ncde_desc* d, const NcdeLayout& lo, cudaStream_t stream);
int launch_wgrad(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream);   // ncde_wgrad.cu
struct HiddenGradPtrs { float* gW[ANCDE_MAX_LAYERS]; float* gB[ANCDE_MAX_LAYERS]; };
int launch_hidden_wgrad(const SolveArgs& a, const HiddenGradPtrs& hp, cudaStream_t stream);

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

__device__ __forceinline__ void cp_async16_b(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// WO_SMEM: the k-contiguous output-layer copy W_out[col][KPS] is resident in shared memory; otherwise it is
// streamed from L2 through a ring of slots (32*NSPLIT columns each) by a producer warp.
template <int TS, bool WO_SMEM>
__global__ void __launch_bounds__(512, 1) ncde_bwd_kernel(const __grid_constant__ SolveArgs p)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = p.ntc;
    const int NW = NT / 32;
    const int tile = blockIdx.x;
    const int b0 = tile * TS;
    const int Hd = lo.Hd, C4 = lo.C4, NCH = lo.NCH, NCG = lo.NCG, NCP = lo.NCP, K = lo.K, KPS = lo.KPS, DUS = lo.DUS;
    const int NKG = (K + 7) / 8;
    int NSPLIT = NW / NKG;
    if (NSPLIT < 1) NSPLIT = 1;
    const int CS = p.ring_kc;                          // column steps (of 32*NSPLIT columns) per phase-2 chunk
    const int CC = 32 * NSPLIT * CS;                   // columns per phase-2 chunk
    const int nchunks = (NCP + CC - 1) / CC;
    const int NSLOT = p.ring_slots;
    const int HT = pad4(Hd * TS);
    const bool need_att = (lo.mode == ANCDE_MODE_TOP) && (p.grad_attention != nullptr);
    const int64_t g_tile_f4 = (int64_t)TS * NCG;        // float4 per (stage, tile) block of saved tanh outputs

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* bar_empty = bar_full + 8;
    float* smem = reinterpret_cast<float*>(smem_raw + 128);
    float* sHW = smem;                                   // hidden weights (packed, transposed)
    float* sp = sHW + pad4(lo.hidden_floats);
    float* sWoT = sp;                                    // resident W_out[col][KPS], or the ring
    sp += WO_SMEM ? (size_t)NCP * KPS : (size_t)NSLOT * CC * KPS;
    float* sPre = sp;   sp += (size_t)NCP * TS;          // [q][col][VW]
    float* sActBuf = sp; sp += 2 * (size_t)lo.act_fwd_floats;   // saved activations, double buffered
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

    if (!WO_SMEM) {
        if (tid == 0) {
            for (int s = 0; s < NSLOT; ++s) {
                mbar_init(bar_full + s, 1);
                mbar_init(bar_empty + s, NKG * NSPLIT);  // the warps that take part in phase 2
            }
            mbar_fence_init();
        }
        __syncthreads();
        if (tid >= NT) {
            // ---- producer warp: W_out[col][KPS] chunks through the ring + L2 prefetch of the tanh outputs
            if (lane == 0) {
                const float* Wg = p.wpack + lo.off_wot;
                int slot = 0;
                uint32_t phase = 1;
                for (int n = lo.n_stages - 1; n >= 0; --n) {
                    if (n >= 1)
                        bulk_prefetch_l2(reinterpret_cast<const float4*>(p.save_G) + ((int64_t)(n - 1) * lo.ntiles + tile) * g_tile_f4,
                                         (uint32_t)(g_tile_f4 * 16));
                    for (int c = 0; c < nchunks; ++c) {
                        mbar_wait(bar_empty + slot, phase);
                        const int cols = (NCP - c * CC < CC) ? NCP - c * CC : CC;
                        const uint32_t bytes = (uint32_t)(cols * KPS * 4);
                        mbar_arrive_expect_tx(bar_full + slot, bytes);
                        bulk_g2s(sWoT + (size_t)slot * CC * KPS, Wg + (size_t)c * CC * KPS, bytes, bar_full + slot);
                        if (++slot == NSLOT) { slot = 0; phase ^= 1; }
                    }
                }
            }
            return;
        }
    }
    auto cta_sync = [&]() { named_bar_sync(1, NT); };
    auto act_block = [&](int n) { return p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats; };
    auto issue_act = [&](int n, float* dst) {
        const float* src = act_block(n);
        for (int i = tid; i < lo.act_fwd_floats / 4; i += NT) cp_async16_b(dst + i * 4, src + i * 4);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack);
        float4* dst = reinterpret_cast<float4*>(sHW);
        for (int i = tid; i < pad4(lo.hidden_floats) / 4; i += NT) dst[i] = __ldg(src + i);
        if (WO_SMEM) {
            const float4* s2 = reinterpret_cast<const float4*>(p.wpack + lo.off_wot);
            float4* d2 = reinterpret_cast<float4*>(sWoT);
            for (int i = tid; i < NCP * KPS / 4; i += NT) d2[i] = __ldg(s2 + i);
        }
        for (int item = tid; item < Hd * TS; item += NT) { sLam[item] = 0.f; sMu[item] = 0.f; }
        issue_act(lo.n_stages - 1, sActBuf);
    }
    cta_sync();

    int jb = lo.n_out - 1;   // last output time whose cotangent has not been consumed yet
    int rslot = 0;           // ring slot / phase parity of the next fill to consume
    uint32_t rphase = 0;
    int abuf = 0;            // which half of sActBuf holds the current stage
    float gmax = 0.f;        // largest |stage cotangent| this thread wrote (scale of the FP16 weight-gradient operands)

    for (int k = lo.n_steps - 1; k >= 0; --k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step);
        const float dt = __fsub_rn(t1, t0);
        // ---- cotangents of the outputs emitted during this step: t_out[j] in (t0, t1]
        {
            for (int item = tid; item < Hd * TS; item += NT) {
                const int i = item / TS, b = item - i * TS;
                const int bg = b0 + b;
                float lam = sLam[item], mu = sMu[item];
                for (int j = jb; j >= 1; --j) {
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
        cta_sync();

#pragma unroll 1
        for (int s = 3; s >= 0; --s) {
            const int n = 4 * k + s;
            float* act = act_block(n);
            float* sAct = sActBuf + (size_t)abuf * lo.act_fwd_floats;
            PHASE_T0();
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
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");     // saved activations of this stage have landed
            cta_sync();
            // prefetch the next stage's activations (this kernel walks n downwards); pull its tanh outputs into L2
            if (n >= 1) issue_act(n - 1, sActBuf + (size_t)(abuf ^ 1) * lo.act_fwd_floats);
            if (WO_SMEM && tid == 0 && n >= 1)
                bulk_prefetch_l2(reinterpret_cast<const float4*>(p.save_G) + ((int64_t)(n - 1) * lo.ntiles + tile) * g_tile_f4,
                                 (uint32_t)(g_tile_f4 * 16));
            // the stage cotangent goes to HBM for the weight-gradient kernels
            for (int item = tid; item < Hd * TS; item += NT) {
                const float gv = sG[item];
                act[lo.act_off_gbar + item] = gv;
                gmax = fmaxf(gmax, fabsf(gv));
            }

            const float* sDU = sAct + lo.act_off_du;
            PHASE_MARK(8);    // stage cotangent + saved activations
            // ---- phase 1: tanh outputs, pre-activation cotangent into shared memory
            const float4* gbase = reinterpret_cast<const float4*>(p.save_G) + ((int64_t)n * lo.ntiles + tile) * g_tile_f4;
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
                cta_sync();
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
                    Vec<VW>::ld(sAct + lo.act_off_duda + du_index(j, q * VW, TS, DUS), du);
#pragma unroll
                    for (int e = 0; e < VW; ++e) acc[e] *= du[e];
                    Vec<VW>::st(sTmp + j * TS + q * VW, acc);
                }
                cta_sync();
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
                        Vec<VW>::ld(sDU + chunk * DUS + c * TS + q * VW, du);
#pragma unroll
                        for (int e = 0; e < VW; ++e) {
                            const float gg = g[q * VW + e][c];
                            v[e] = gb[q * VW + e] * du[e] * fmaf(-gg, gg, 1.0f);
                        }
                        Vec<VW>::st(sPre + ((size_t)q * NCP + cg * 4 + c) * VW, v);
                    }
            }
            cta_sync();
            PHASE_MARK(9);    // phase 1: G -> pre_bar (+ attention products)
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
            PHASE_MARK(10);   // attention atomics
            // ---- phase 2: hbar[k][b] = sum_col pre_bar[col][b] * W_out[col][k]; warp = (k group, column split)
            if (warp < NKG * NSPLIT) {
                const int kg = warp % NKG, split = warp / NKG;
                float acc[8 * TS];
#pragma unroll
                for (int i = 0; i < 8 * TS; ++i) acc[i] = 0.f;
                for (int c = 0; c < nchunks; ++c) {
                    int slot = 0;
                    const float* wbase;
                    if (WO_SMEM) {
                        wbase = sWoT + (size_t)c * CC * KPS;
                    } else {
                        slot = rslot;
                        mbar_wait(bar_full + slot, rphase);
                        wbase = sWoT + (size_t)slot * CC * KPS;
                    }
                    for (int cs = 0; cs < CS; ++cs) {
                        const int lcol = (cs * NSPLIT + split) * 32 + lane;
                        const int col = c * CC + lcol;
                        if (col < NCP) {
                            const float* wrow = wbase + (size_t)lcol * KPS + kg * 8;
                            const float4 w0 = *reinterpret_cast<const float4*>(wrow);
                            const float4 w1 = *reinterpret_cast<const float4*>(wrow + 4);
                            float pv[TS];
#pragma unroll
                            for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sPre + ((size_t)q * NCP + col) * VW, pv + q * VW);
                            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                                for (int b = 0; b < TS; ++b) acc[kk * TS + b] = fmaf(wv[kk], pv[b], acc[kk * TS + b]);
                        }
                    }
                    if (!WO_SMEM) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_empty + slot);
                        if (++rslot == NSLOT) { rslot = 0; rphase ^= 1; }
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
            } else if (!WO_SMEM) {
                // warps without a phase-2 role still track the ring position
                for (int c = 0; c < nchunks; ++c)
                    if (++rslot == NSLOT) { rslot = 0; rphase ^= 1; }
            }
            cta_sync();
            PHASE_MARK(11);   // phase 2: pre_bar x W_out + reduce + barrier
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
                    v = (hl[item] > 0.f) ? v : 0.f;
                    hb[item] = v;
                    act[lo.act_off_delta[l] + item] = v;     // for the hidden weight-gradient kernel
                }
                cta_sync();
                {
                    // cotangent of the layer input: hb_next[k][b] = sum_o Wt[k][o] * delta[o][b]; KS adjacent lanes
                    // split the sum over o and combine with shuffles
                    const int nitems = Kin * NQ;
                    int KS = 4;
                    while (KS > 1 && nitems * KS > NT) KS >>= 1;
                    const int ks = tid & (KS - 1);
                    const int per = NT / KS;
                    const int OL = (N + KS - 1) / KS;
                    const int obeg = ks * OL;
                    int ocnt = N - obeg;
                    ocnt = ocnt > OL ? OL : ocnt;
                    const float* Wt = sHW + lo.off_w[l];
                    for (int base = 0; base < nitems; base += per) {
                        const int item = base + tid / KS;
                        const bool valid = item < nitems;
                        int q = 0, kk = valid ? item : 0;
                        if (NQ > 1 && kk >= Kin) { q = 1; kk -= Kin; }
                        float acc[VW];
#pragma unroll
                        for (int e = 0; e < VW; ++e) acc[e] = 0.f;
                        if (valid) {
                            const float* wp = Wt + kk * S + obeg;
                            const float* dp = hb + obeg * TS + q * VW;
#pragma unroll 4
                            for (int o = 0; o < ocnt; ++o) {
                                const float w = *wp;
                                float dv[VW];
                                Vec<VW>::ld(dp, dv);
                                ++wp;
                                dp += TS;
#pragma unroll
                                for (int e = 0; e < VW; ++e) acc[e] = fmaf(w, dv[e], acc[e]);
                            }
                        }
                        for (int m = KS >> 1; m >= 1; m >>= 1)
#pragma unroll
                            for (int e = 0; e < VW; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], m);
                        if (valid && ks == 0) Vec<VW>::st(hb_next + kk * TS + q * VW, acc);
                    }
                }
                cta_sync();
                float* t = hb; hb = hb_next; hb_next = t;
            }
            PHASE_MARK(12);   // phase 3: hidden layers
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
            abuf ^= 1;
            cta_sync();
            PHASE_MARK(13);   // RK4 adjoint combine + barrier
        }
    }
    // ---- results
    for (int o = 16; o >= 1; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0 && gmax > 0.f && gmax < 3.0e38f)
        atomicMax(reinterpret_cast<unsigned*>(const_cast<float*>(p.wpack) + lo.off_gscale), __float_as_uint(gmax));
    for (int item = tid; item < Hd * TS; item += NT) {
        const int i = item / TS, b = item - i * TS;
        const int bg = b0 + b;
        if (bg < lo.B) p.grad_z0[(int64_t)bg * Hd + i] = sLam[item] + __ldg(p.grad_z_out + (int64_t)bg * Hd + i);
    }
}

static int bwd_threads(const NcdeLayout& lo)
{
    int per = (lo.NCG + 447) / 448;
    int nt = round_up((lo.NCG + per - 1) / per, 32);
    const int NKG = (lo.K + 7) / 8;
    if (nt < NKG * 32) nt = NKG * 32;
    if (nt < 64) nt = 64;
    return nt;
}

// shared memory (bytes) of the backward without the output-layer weights / ring
static size_t bwd_smem_base(const NcdeLayout& lo, int nt)
{
    const int NKG = (lo.K + 7) / 8;
    int nsplit = (nt / 32) / NKG;
    if (nsplit < 1) nsplit = 1;
    size_t f = (size_t)pad4(lo.hidden_floats);
    f += (size_t)lo.NCP * lo.TS + 2 * (size_t)lo.act_fwd_floats + 6 * (size_t)pad4(lo.Hd * lo.TS) +
         2 * (size_t)pad4(lo.Hmax * lo.TS) + (size_t)nsplit * NKG * 8 * lo.TS + (size_t)lo.C4 * lo.TS;
    return f * 4 + 128;
}

static size_t bwd_ring_bytes(const NcdeLayout& lo, int nt, int slots, int cs)
{
    const int NKG = (lo.K + 7) / 8;
    int nsplit = (nt / 32) / NKG;
    if (nsplit < 1) nsplit = 1;
    return (size_t)slots * 32 * nsplit * cs * lo.KPS * 4;
}

// Used by make_layout's tile-size choice: does the backward fit with this many samples per CTA?
bool bwd_fits(const NcdeLayout& lo)
{
    const int nt = bwd_threads(lo);
    return nt <= 480 && bwd_smem_base(lo, nt) + bwd_ring_bytes(lo, nt, 2, 1) <= 227 * 1024;
}

template <int TS>
static int launch_bwd(const ancde_ncde_desc* d, SolveArgs& a, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    const int nt = bwd_threads(lo);
    const size_t max_smem = 227 * 1024;
    if (nt > 480) {
        set_error("ncde_backward: needs %d threads per CTA (max 480)", nt);
        return ANCDE_EUNSUPPORTED;
    }
    const size_t base = bwd_smem_base(lo, nt);
    const bool wo_smem = base + (size_t)lo.NCP * lo.KPS * 4 <= max_smem;
    size_t smem;
    int threads = nt;
    a.ntc = nt;
    a.ring_kc = 1; a.ring_pw = 0;
    if (wo_smem) {
        a.ring_slots = 1;
        smem = base + (size_t)lo.NCP * lo.KPS * 4;
    } else {
        // the deepest ring that fits: prefer two column steps per slot (fewer barrier waits), at least 3 slots
        int slots = 4, cs = 2;
        while (base + bwd_ring_bytes(lo, nt, slots, cs) > max_smem) {
            if (slots > 3) --slots;
            else if (cs > 1) { cs = 1; slots = 4; }
            else if (slots > 2) --slots;
            else break;
        }
        smem = base + bwd_ring_bytes(lo, nt, slots, cs);
        if (smem > max_smem) {
            set_error("ncde_backward: needs %zu bytes of shared memory per CTA (max %zu)", smem, max_smem);
            return ANCDE_EUNSUPPORTED;
        }
        a.ring_slots = slots;
        a.ring_kc = cs;
        threads = nt + 32;
    }
    static const char* names[3] = {"ncde_bwd_plain", "ncde_bwd_bottom", "ncde_bwd_top"};
    if (wo_smem) {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_bwd_kernel<TS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_bwd_kernel<TS, true><<<lo.ntiles, threads, smem, stream>>>(a);
    } else {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_bwd_kernel<TS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_bwd_kernel<TS, false><<<lo.ntiles, threads, smem, stream>>>(a);
    }
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());

    // weight gradients (read the stage / layer cotangents the kernel above left in save_act)
    {
        int rc = launch_wgrad(a, d->grad_W[lo.n_hidden], d->grad_bias[lo.n_hidden], stream);
        if (rc != ANCDE_OK) return rc;
        HiddenGradPtrs hp;
        for (int l = 0; l < lo.n_hidden; ++l) { hp.gW[l] = d->grad_W[l]; hp.gB[l] = d->grad_bias[l]; }
        rc = launch_hidden_wgrad(a, hp, stream);
        if (rc != ANCDE_OK) return rc;
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
