// Fused neural-CDE forward solve (fixed-grid RK4, 3/8 rule) for a tile of TS samples per CTA.
//
// One persistent CTA integrates its TS samples over ALL time steps: the state z, the RK4 stage
// derivatives and every MLP activation stay in shared memory / registers; per stage the CTA
//   1. evaluates the control increment dU (spline derivative lookup, and for the top NCDE the
//      attention product dX*a + a(1-a)*dX*h'),                [cdeint_module.py:58-60,103-129]
//   2. runs the hidden ReLU layers of the vector-field MLP,   [vector_fields.py:205-216,237-247]
//   3. runs the output layer in register tiles (4 columns x TS samples per thread), applies
//      tanh and contracts the (Hd x C) field with dU,          [vector_fields.py:216-218, cdeint_module.py:61,133]
//   4. combines the stage into the RK4 recurrence              [torchdiffeq rk4_alt_step_func]
// Hidden-layer weights are packed once per call (transposed, padded) and stay in shared memory.
// The output-layer weights stay there too when they fit; the Sepsis-size layers (1715x49 / 3381x49
// floats > 227 KB) are streamed from L2 instead: a producer warp keeps a ring of shared-memory
// slots full with TMA bulk copies (cp.async.bulk + mbarrier) that run ahead of the consumers, so
// the FMA loop only ever reads shared memory.  Plain FP32 FMA (north-star tolerance 1e-4).
#include "ncde_common.cuh"
#include "fwd_hidden.cuh"
#include "pipeline.cuh"

namespace ancde {

constexpr int CTL_ITEMS = 2;   // control elements (j, b) per thread held in registers

struct CtlRegs {
    float b[CTL_ITEMS], c2[CTL_ITEMS], d3[CTL_ITEMS], a[CTL_ITEMS], hp[CTL_ITEMS];
    float frac;
};

// What does not change from stage to stage for the control elements (j, b) a thread owns: worked out once per kernel
// (an integer division and 64-bit index arithmetic per element and stage otherwise).
struct CtlItems {
    int obase[CTL_ITEMS];   // (bg * (L-1)) * C + j: coefficient offset of piece 0, or -1 when the element does not exist
    int abase[CTL_ITEMS];   // bg * att_C + (elementwise attention: j)
    int hbase[CTL_ITEMS];   // bg * C + j
    int di[CTL_ITEMS];      // du_index(j, b), or -1 past the C4 * TS elements
};

template <int TS>
__device__ __forceinline__ void ctl_setup(const NcdeLayout& lo, int b0, int tid, int NT, CtlItems& ci)
{
#pragma unroll
    for (int it = 0; it < CTL_ITEMS; ++it) {
        const int item = tid + it * NT;
        ci.obase[it] = -1; ci.abase[it] = 0; ci.hbase[it] = 0; ci.di[it] = -1;
        if (item < lo.C4 * TS) {
            const int b = item / lo.C4, j = item - b * lo.C4;
            const int bg = b0 + b;
            ci.di[it] = du_index(j, b, TS, lo.DUS);
            if (j < lo.C && bg < lo.B) {
                ci.obase[it] = bg * (lo.L - 1) * lo.C + j;
                ci.abase[it] = bg * lo.att_C + (lo.att_C == 1 ? 0 : j);
                ci.hbase[it] = bg * lo.C + j;
            }
        }
    }
}

// Issues the global loads of the control of stage (k, s); nothing is consumed here so the
// latency overlaps with the MLP of the current stage.
template <int TS>
__device__ __forceinline__ void ctl_issue(const SolveArgs& p, const float* sTimes, int& cnt, int k, int s,
                                          const CtlItems& ci, CtlRegs& r)
{
    const NcdeLayout& lo = p.lo;
    const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
    const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
    const float ts = stage_time(t0, __fsub_rn(t1, t0), s);
    // index = clamp(#(times < t) - 1, 0, L-2): the LEFT piece at a knot (interpolate.py:261-267)
    while (cnt < lo.L && sTimes[cnt] < ts) ++cnt;
    int idx = cnt - 1;
    idx = idx < 0 ? 0 : (idx > lo.L - 2 ? lo.L - 2 : idx);
    r.frac = __fsub_rn(ts, sTimes[idx]);
    const int n = 4 * k + s;
    int q = (int)floorf(ts) - 1;                      // attention[int(floor(t)) - 1], python wrap
    if (q < 0) q += lo.att_L;
    q = q < 0 ? 0 : (q >= lo.att_L ? lo.att_L - 1 : q);
    const bool top = lo.mode == ANCDE_MODE_TOP;
    const int m = top ? n % (lo.n_hp - 1) : 0;        // call counter + reset rule (:134-137)
    const float* cb = p.cb + (int64_t)idx * lo.C;
    const float* cc2 = p.cc2 + (int64_t)idx * lo.C;
    const float* cd3 = p.cd3 + (int64_t)idx * lo.C;
    const float* att = top ? p.attention + (int64_t)q * lo.B * lo.att_C : nullptr;
    const float* hpr = top ? p.h_prime + (int64_t)m * lo.B * lo.C : nullptr;
#pragma unroll
    for (int it = 0; it < CTL_ITEMS; ++it) {
        r.b[it] = r.c2[it] = r.d3[it] = 0.f;
        r.a[it] = r.hp[it] = 0.f;
        const int o = ci.obase[it];
        if (o >= 0) {
            r.b[it] = __ldg(cb + o);
            r.c2[it] = __ldg(cc2 + o);
            r.d3[it] = __ldg(cd3 + o);
            if (top) {
                r.a[it] = __ldg(att + ci.abase[it]);
                r.hp[it] = __ldg(hpr + ci.hbase[it]);
            }
        }
    }
}

// Turns the loaded coefficients into dU (and dU/da for the backward) and stores them [j][TS].
template <int TS>
__device__ __forceinline__ void ctl_finish(const SolveArgs& p, const CtlRegs& r, const CtlItems& ci, float* sDU, float* act)
{
    const NcdeLayout& lo = p.lo;
#pragma unroll
    for (int it = 0; it < CTL_ITEMS; ++it) {
        const int di = ci.di[it];
        if (di >= 0) {
            // derivative = b + (two_c + three_d * frac) * frac   (interpolate.py:298-303)
            float dX = __fadd_rn(r.b[it], __fmul_rn(__fadd_rn(r.c2[it], __fmul_rn(r.d3[it], r.frac)), r.frac));
            float dU = dX, duda = 0.f;
            if (lo.mode == ANCDE_MODE_TOP) {
                const float a = r.a[it], hp = r.hp[it];
                // dY = dX*a + ((a*(1-a))*dX)*h'   (cdeint_module.py:113-129; "Xt" is dX/dt, Q4)
                dU = __fadd_rn(__fmul_rn(dX, a),
                               __fmul_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.0f, a)), dX), hp));
                duda = dX * (1.0f + (1.0f - 2.0f * a) * hp);
            }
            sDU[di] = dU;
            if (act) {
                act[lo.act_off_du + di] = dU;
                if (lo.mode == ANCDE_MODE_TOP) act[lo.act_off_duda + di] = duda;
            }
        }
    }
}

// Hidden layers wider than 32 units (one CTA barrier per layer): out[o][.] = relu(bias[o] + sum_k W[o][k] * in[k][.])
// for all TS samples, activations [unit][TS].  One (output, group of VW samples) item per thread pair -- the two lanes
// of a pair split k and add with one shuffle; the weight row is read in 128-bit pieces.
template <int TS>
__device__ __forceinline__ void dense_relu(const float* __restrict__ Wr, const float* __restrict__ bias, int nk4,
                                           int N, int S, const float* __restrict__ in, float* __restrict__ out,
                                           float* __restrict__ gsave, int tid, int NT)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    const int ks = tid & 1, slot = tid >> 1, nslots = NT >> 1;
    const int h4 = (nk4 + 1) >> 1;
    const int qbeg = ks * h4;
    const int qcnt = ks ? nk4 - h4 : h4;
    const int nitems = N * NQ;
    for (int base = 0; base < nitems; base += nslots) {
        const int item = base + slot;
        const bool valid = item < nitems;
        int q = 0, o = valid ? item : 0;
        if (NQ > 1 && o >= N) { q = 1; o -= N; }
        float acc[VW];
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[e] = 0.f;
        if (valid) {
            const float4* wp = reinterpret_cast<const float4*>(Wr + o * S) + qbeg;
            const float* ip = in + (qbeg * 4) * TS + q * VW;
#pragma unroll 2
            for (int j = 0; j < qcnt; ++j) {
                const float4 w = wp[j];
                float v0[VW], v1[VW], v2[VW], v3[VW];
                Vec<VW>::ld(ip, v0);
                Vec<VW>::ld(ip + TS, v1);
                Vec<VW>::ld(ip + 2 * TS, v2);
                Vec<VW>::ld(ip + 3 * TS, v3);
                ip += 4 * TS;
#pragma unroll
                for (int e = 0; e < VW; ++e) {
                    acc[e] = fmaf(w.x, v0[e], acc[e]);
                    acc[e] = fmaf(w.y, v1[e], acc[e]);
                    acc[e] = fmaf(w.z, v2[e], acc[e]);
                    acc[e] = fmaf(w.w, v3[e], acc[e]);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 1);
        if (valid && ks == 0) {
            const float bv = bias[o];
#pragma unroll
            for (int e = 0; e < VW; ++e) acc[e] = fmaxf(acc[e] + bv, 0.f);
            Vec<VW>::st(out + o * TS + q * VW, acc);
            if (gsave) Vec<VW>::st(gsave + o * TS + q * VW, acc);
        }
    }
}

// WO_SMEM: output-layer weights resident in shared memory; otherwise streamed through the TMA ring.
// CHAIN: every hidden layer has at most 32 units -> warp-synchronous hidden_chain; otherwise dense_relu per layer.
template <int TS, bool WO_SMEM, bool CHAIN>
__global__ void __launch_bounds__(512, 1) ncde_fwd_kernel(const __grid_constant__ SolveArgs p)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, lane = tid & 31;
    const int NT = p.ntc;                         // consumer threads (the producer warp, if any, sits behind them)
    const int tile = blockIdx.x;
    const int b0 = tile * TS;
    const int Hd = lo.Hd, NCH = lo.NCH, NCG = lo.NCG, NCP = lo.NCP, K = lo.K;
    const int KC = p.ring_kc, NSLOT = p.ring_slots, PW = p.ring_pw;
    const int npass = (NCG + NT - 1) / NT;        // column groups are walked NT at a time
    const int nchunks = (K + KC - 1) / KC;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw);      // [NSLOT]
    uint64_t* bar_empty = bar_full + 8;                              // [NSLOT]
    float* smem = reinterpret_cast<float*>(smem_raw + 128);
    float* sHW = smem;                                   // hidden weights + biases (packed)
    float* sp = sHW + pad4(lo.hidden_floats);
    float* sBo = sp;    sp += NCP;                       // output-layer bias
    float* sWo = sp;                                     // output-layer weights, or the streaming ring
    sp += WO_SMEM ? (size_t)K * NCP : (size_t)NSLOT * KC * PW;
    float* sTimes = sp; sp += pad4(lo.L);
    const int HP = pad4(lo.Hmax);                        // floats per sample row of the hidden chain's buffers
    // CHAIN: rows per sample [sample][unit]; otherwise rows per unit [unit][TS] (both TS * HP floats, padding stays finite)
    float* sInT = sp;   sp += TS * HP;                   // stage input z_s
    float* sT0 = sp;    sp += TS * HP;                   // hidden activations ping
    float* sT1 = sp;    sp += TS * HP;                   // hidden activations pong
    float* sLast = sp;  sp += pad4(lo.Hmax * TS);        // last hidden layer    [unit][TS]
    int* sLp = reinterpret_cast<int*>(sp); sp += 8 * ANCDE_MAX_LAYERS;   // per-layer parameters of hidden_chain
    float* sY = sp;     sp += pad4(Hd * TS);             // z at the start of the step
    float* sK1 = sp;    sp += pad4(Hd * TS);
    float* sK2 = sp;    sp += pad4(Hd * TS);
    float* sK3 = sp;    sp += pad4(Hd * TS);
    float* sDU = sp;    sp += 2 * lo.du_floats;          // control increment, double buffered [chunk][4][TS]
    float* sPart = sp;                                   // per column-group partial dot products

    if (!WO_SMEM) {
        if (tid == 0) {
            for (int s = 0; s < NSLOT; ++s) {
                mbar_init(bar_full + s, 1);              // one arrive.expect_tx per fill
                mbar_init(bar_empty + s, NT / 32);       // one arrive per consumer warp
            }
            mbar_fence_init();
        }
        __syncthreads();
        if (tid >= NT) {
            // ---- producer warp: streams Wt_out[k][pass columns] through the ring, always ahead
            if (lane == 0) {
                const float* Wg = p.wpack + lo.off_wo;
                int slot = 0;
                uint32_t phase = 1;
                for (int n = 0; n < lo.n_stages; ++n)
                    for (int pass = 0; pass < npass; ++pass) {
                        const int col0 = pass * NT * 4;
                        const int pw = (NCP - col0 < PW) ? NCP - col0 : PW;
                        for (int c = 0; c < nchunks; ++c) {
                            mbar_wait(bar_empty + slot, phase);
                            const int rows = (K - c * KC < KC) ? K - c * KC : KC;
#ifdef ANCDE_EXPERIMENT_HALF_STREAM
                            const int pwx = (pw / 8) * 4;     // timing experiment only: stream half of every row
#else
                            const int pwx = pw;
#endif
                            mbar_arrive_expect_tx(bar_full + slot, (uint32_t)(rows * pwx * 4));
                            float* dst = sWo + (size_t)slot * KC * PW;
                            for (int r = 0; r < rows; ++r)
                                bulk_g2s(dst + (size_t)r * PW, Wg + (size_t)(c * KC + r) * NCP + col0, (uint32_t)(pwx * 4),
                                         bar_full + slot);
                            if (++slot == NSLOT) { slot = 0; phase ^= 1; }
                        }
                    }
            }
            return;
        }
    }
    auto cta_sync = [&]() { named_bar_sync(1, NT); };

    // ---- prologue: weights, knot times, initial state
    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack);
        float4* dst = reinterpret_cast<float4*>(sHW);
        for (int i = tid; i < pad4(lo.hidden_floats) / 4; i += NT) dst[i] = __ldg(src + i);
        const float4* sb = reinterpret_cast<const float4*>(p.wpack + lo.off_bo);
        for (int i = tid; i < NCP / 4; i += NT) reinterpret_cast<float4*>(sBo)[i] = __ldg(sb + i);
        if (WO_SMEM) {
            const float4* s2 = reinterpret_cast<const float4*>(p.wpack + lo.off_wo);
            float4* d2 = reinterpret_cast<float4*>(sWo);
            for (int i = tid; i < K * NCP / 4; i += NT) d2[i] = __ldg(s2 + i);
        }
        for (int i = tid; i < lo.L; i += NT) sTimes[i] = p.times[i];
        if (tid < lo.n_hidden) {
            int* q = sLp + 8 * tid;
            q[0] = (lo.in_dim[tid] + 3) / 4; q[1] = lo.width[tid]; q[2] = lo.stride[tid]; q[3] = lo.off_w[tid];
            q[4] = lo.off_b[tid]; q[5] = lo.act_off_h[tid]; q[6] = 0; q[7] = 0;
        }
        for (int i = tid; i < 3 * TS * HP; i += NT) sInT[i] = 0.f;      // sInT, sT0, sT1 (contiguous)
        cta_sync();
        for (int item = tid; item < Hd * TS; item += NT) {
            const int i = item / TS, b = item - i * TS;
            const int bg = b0 + b;
            const float v = (bg < lo.B) ? p.z0[(int64_t)bg * Hd + i] : 0.f;
            sInT[CHAIN ? b * HP + i : item] = v;
            sY[item] = v;
            if (bg < lo.B) p.z_out[(int64_t)bg * Hd + i] = v;          // solution[0] = y0
            if (p.save_act) p.save_act[(int64_t)tile * lo.act_floats + item] = v;
        }
    }
    cta_sync();

    int cnt = 0;        // knots strictly below the current stage time (monotone)
    int jout = 1;       // next requested output time
    int rslot = 0;      // ring slot / phase parity of the next fill to consume
    uint32_t rphase = 0;
    CtlItems ci;
    ctl_setup<TS>(lo, b0, tid, NT, ci);
    {
        CtlRegs r;
        ctl_issue<TS>(p, sTimes, cnt, 0, 0, ci, r);
        ctl_finish<TS>(p, r, ci, sDU, p.save_act ? p.save_act + (int64_t)tile * lo.act_floats : nullptr);
    }
    cta_sync();

    for (int k = 0; k < lo.n_steps; ++k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float dt = __fsub_rn(t1, t0);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const int n = 4 * k + s;
            if (n >= lo.n_stages) break;             // single-stage mode: the field at (t_start, z0) only
            float* act = p.save_act ? p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats : nullptr;
            float* act_next = (p.save_act && n + 1 < lo.n_stages)
                                  ? p.save_act + ((int64_t)(n + 1) * lo.ntiles + tile) * lo.act_floats : nullptr;
            const float* sDUcur = sDU + (n & 1) * lo.du_floats;
            float* sDUnext = sDU + ((n + 1) & 1) * lo.du_floats;

            PHASE_T0();
            // (1) control of the NEXT stage: loads in flight while this stage's MLP runs
            CtlRegs ctl;
            const bool have_next = (n + 1 < lo.n_stages);
            if (have_next) ctl_issue<TS>(p, sTimes, cnt, s == 3 ? k + 1 : k, s == 3 ? 0 : s + 1, ci, ctl);
            PHASE_MARK(5);   // control issue

            // (2) hidden layers: warp-synchronous chain, one CTA barrier at the end
            const float* in = sLast;
            if constexpr (CHAIN) {
                hidden_chain<TS>(sLp, lo.n_hidden, sHW, sInT, sT0, sT1, sLast, act, HP, tid >> 5, NT >> 5, lane);
                cta_sync();
            } else {
                in = sInT;
                float* outb = sT0;
                for (int l = 0; l < lo.n_hidden; ++l) {
                    const int4 q0 = *reinterpret_cast<const int4*>(sLp + 8 * l);
                    const int2 q1 = *reinterpret_cast<const int2*>(sLp + 8 * l + 4);
                    dense_relu<TS>(sHW + q0.w, sHW + q1.x, q0.x, q0.y, q0.z, in, outb, act ? act + q1.y : nullptr, tid, NT);
                    cta_sync();
                    in = outb;
                    outb = (outb == sT0) ? sT1 : sT0;
                }
            }
            PHASE_MARK(0);   // ctl issue + hidden layers

            // (3) output layer, tanh, contraction with dU: thread = 4 columns x TS samples
            for (int pass = 0; pass < npass; ++pass) {
                const int cg = pass * NT + tid;
                const bool active = cg < NCG;
                float acc[TS][4];
                {
                    const float4 bv = *reinterpret_cast<const float4*>(sBo + (active ? cg : 0) * 4);
#pragma unroll
                    for (int b = 0; b < TS; ++b) { acc[b][0] = bv.x; acc[b][1] = bv.y; acc[b][2] = bv.z; acc[b][3] = bv.w; }
                }
                for (int c = 0; c < nchunks; ++c) {
                    const float* wp;
                    int wstride;
                    int slot = 0;
                    if (WO_SMEM) {
                        wp = sWo + (size_t)(c * KC) * NCP + (active ? cg : 0) * 4;
                        wstride = NCP;
                    } else {
                        slot = rslot;
                        mbar_wait(bar_full + slot, rphase);
                        wp = sWo + (size_t)slot * KC * PW + tid * 4;
                        wstride = PW;
                    }
                    const int rows = (K - c * KC < KC) ? K - c * KC : KC;
                    if (active) {
                        const float* ip = in + (c * KC) * TS;
#pragma unroll 7
                        for (int kk = 0; kk < rows; ++kk) {
                            const float4 w = *reinterpret_cast<const float4*>(wp);
                            wp += wstride;
#pragma unroll
                            for (int q = 0; q < NQ; ++q) {
                                float v[VW];
                                Vec<VW>::ld(ip + q * VW, v);
#pragma unroll
                                for (int e = 0; e < VW; ++e) {
                                    acc[q * VW + e][0] = fmaf(v[e], w.x, acc[q * VW + e][0]);
                                    acc[q * VW + e][1] = fmaf(v[e], w.y, acc[q * VW + e][1]);
                                    acc[q * VW + e][2] = fmaf(v[e], w.z, acc[q * VW + e][2]);
                                    acc[q * VW + e][3] = fmaf(v[e], w.w, acc[q * VW + e][3]);
                                }
                            }
                            ip += TS;
                        }
                    }
                    if (!WO_SMEM) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_empty + slot);
                        if (++rslot == NSLOT) { rslot = 0; rphase ^= 1; }
                    }
                }
                PHASE_MARK(1);   // output-layer FMA loop
                if (active) {
                    const int chunk = cg % NCH;
                    float part[TS];
#pragma unroll
                    for (int b = 0; b < TS; ++b) part[b] = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float dy[TS];
#pragma unroll
                        for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sDUcur + chunk * lo.DUS + c * TS + q * VW, dy + q * VW);
#pragma unroll
                        for (int b = 0; b < TS; ++b) {
                            const float gt = tanh_mufu(acc[b][c]);
                            acc[b][c] = gt;
                            part[b] = fmaf(gt, dy[b], part[b]);
                        }
                    }
                    if (p.save_G) {
                        float4* gp = reinterpret_cast<float4*>(p.save_G) +
                                     (((int64_t)n * lo.ntiles + tile) * TS) * NCG + cg;
#pragma unroll
                        for (int b = 0; b < TS; ++b)
                            gp[(int64_t)b * NCG] = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
                    }
#pragma unroll
                    for (int q = 0; q < NQ; ++q) Vec<VW>::st(sPart + cg * TS + q * VW, part + q * VW);
                }
            }
            PHASE_MARK(2);   // tanh, contraction, G store
            if (have_next) ctl_finish<TS>(p, ctl, ci, sDUnext, act_next);
            cta_sync();
            PHASE_MARK(3);   // control finish + barrier

            // (4) finish the contraction and advance the RK4 recurrence (rk4_alt_step_func)
            for (int item = tid; item < Hd * TS; item += NT) {
                const int i = item / TS, b = item - i * TS;
                const int bg = b0 + b;
                float o = 0.f;
                for (int ch = 0; ch < NCH; ++ch) o += sPart[(i * NCH + ch) * TS + b];
                if (p.k_out && bg < lo.B) p.k_out[((int64_t)n * lo.B + bg) * Hd + i] = o;
                const float y = sY[item];
                float zin;
                if (s == 0) {
                    sK1[item] = o;
                    zin = __fadd_rn(y, __fdiv_rn(__fmul_rn(dt, o), 3.0f));                       // y + dt*k1/3
                } else if (s == 1) {
                    sK2[item] = o;
                    zin = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(o, __fdiv_rn(sK1[item], 3.0f)))); // y + dt*(k2 - k1/3)
                } else if (s == 2) {
                    sK3[item] = o;
                    zin = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(sK1[item], sK2[item]), o)));  // y + dt*(k1-k2+k3)
                } else {
                    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(sK1[item], __fmul_rn(3.0f, sK2[item])),
                                                          __fmul_rn(3.0f, sK3[item])), o);
                    const float y1 = __fadd_rn(y, __fmul_rn(sum, __fdiv_rn(dt, 8.0f)));
                    // emit every requested time in (t0, t1] (FixedGridODESolver.integrate + _linear_interp)
                    int j = jout;
                    while (j < lo.n_out) {
                        const float tj = __ldg(p.t_out + j);
                        if (!(t1 >= tj)) break;
                        float v;
                        if (tj == t0) v = y;
                        else if (tj == t1) v = y1;
                        else v = __fadd_rn(y, __fmul_rn(__fdiv_rn(__fsub_rn(y1, y), __fsub_rn(t1, t0)), __fsub_rn(tj, t0)));
                        if (bg < lo.B) p.z_out[((int64_t)j * lo.B + bg) * Hd + i] = v;
                        ++j;
                    }
                    sY[item] = y1;
                    zin = y1;
                }
                sInT[CHAIN ? b * HP + i : item] = zin;
                if (act_next) act_next[item] = zin;
            }
            if (s == 3) {
                while (jout < lo.n_out && t1 >= __ldg(p.t_out + jout)) ++jout;
            }
            cta_sync();
            PHASE_MARK(4);   // RK4 combine + barrier
        }
    }
}

// ---- launch -----------------------------------------------------------------------------------
static int fwd_threads(const NcdeLayout& lo)
{
    int per = (lo.NCG + 447) / 448;
    int nt = round_up((lo.NCG + per - 1) / per, 32);
    int need = (lo.C4 * lo.TS + CTL_ITEMS - 1) / CTL_ITEMS;
    if (nt < need) nt = round_up(need, 32);
    if (nt < 64) nt = 64;
    return nt;
}

// shared memory (bytes) without the output-layer weights / ring
static size_t fwd_smem_base(const NcdeLayout& lo)
{
    size_t f = pad4(lo.hidden_floats) + lo.NCP;
    f += pad4(lo.L) + 3 * (size_t)lo.TS * pad4(lo.Hmax) + pad4(lo.Hmax * lo.TS) + 8 * ANCDE_MAX_LAYERS +
         4 * pad4(lo.Hd * lo.TS) + 2 * (size_t)lo.du_floats + (size_t)lo.NCG * lo.TS;
    return f * 4 + 128;
}

template <int TS>
static int launch_fwd_ts(SolveArgs a, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    if ((int64_t)lo.B * (lo.L - 1) * lo.C >= ((int64_t)1 << 31) - lo.C) {
        set_error("ncde_forward: B * (L-1) * C = %lld does not fit the 32-bit coefficient offsets of the control lookup",
                  (long long)lo.B * (lo.L - 1) * lo.C);
        return ANCDE_EUNSUPPORTED;
    }
    const int nt = fwd_threads(lo);
    if (nt > 480) {
        set_error("ncde_forward: C=%d with %d samples per CTA needs %d threads (max 480)", lo.C, lo.TS, nt);
        return ANCDE_EUNSUPPORTED;
    }
    const size_t max_smem = 227 * 1024;
    const size_t base = fwd_smem_base(lo);
    if (base > max_smem) {
        set_error("ncde_forward: hidden layers need %zu bytes of shared memory (max %zu)", base, max_smem);
        return ANCDE_EUNSUPPORTED;
    }
    const bool wo_smem = base + (size_t)lo.K * lo.NCP * 4 <= max_smem;
    a.ntc = nt;
    size_t smem;
    int threads = nt;
    if (wo_smem) {
        a.ring_kc = lo.K; a.ring_slots = 1; a.ring_pw = lo.NCP;
        smem = base + (size_t)lo.K * lo.NCP * 4;
    } else {
        // ring of NSLOT slots, each KC rows of the pass width; as deep as the remaining shared memory allows
        const int pw = (lo.NCP < nt * 4) ? lo.NCP : nt * 4;
        const size_t avail = max_smem - base;
        int slots = 3;
        int kc = (int)(avail / ((size_t)slots * pw * 4));
        if (kc < 1) { slots = 2; kc = (int)(avail / ((size_t)slots * pw * 4)); }
        if (kc < 1) {
            set_error("ncde_forward: no shared memory left for the weight ring (%zu bytes free)", avail);
            return ANCDE_EUNSUPPORTED;
        }
        if (kc > 8) kc = 8;
        if (kc > lo.K) kc = lo.K;
        a.ring_kc = kc; a.ring_slots = slots; a.ring_pw = pw;
        smem = base + (size_t)slots * kc * pw * 4;
        threads = nt + 32;
    }
    static const char* names[3] = {"ncde_fwd_plain", "ncde_fwd_bottom", "ncde_fwd_top"};
    bool chain = true;
    // one or two samples per CTA (small batches): the chain's single warp beats a CTA barrier per layer up to 64 units
    for (int l = 0; l < lo.n_hidden; ++l) chain = chain && lo.width[l] <= (TS <= 2 ? 64 : 32);
    void (*kern)(SolveArgs) = wo_smem ? (chain ? ncde_fwd_kernel<TS, true, true> : ncde_fwd_kernel<TS, true, false>)
                                      : (chain ? ncde_fwd_kernel<TS, false, true> : ncde_fwd_kernel<TS, false, false>);
    ANCDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(names[lo.mode], stream);
    kern<<<lo.ntiles, threads, smem, stream>>>(a);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

int launch_fwd(const SolveArgs& a, cudaStream_t stream)
{
    switch (a.lo.TS) {
        case 8: return launch_fwd_ts<8>(a, stream);
        case 4: return launch_fwd_ts<4>(a, stream);
        case 2: return launch_fwd_ts<2>(a, stream);
        default: return launch_fwd_ts<1>(a, stream);
    }
}

}  // namespace ancde
