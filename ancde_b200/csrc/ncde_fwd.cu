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
// Weights are packed once per call (transposed, padded) and kept in shared memory; only an
// output layer too large for the 227 KB of one SM (Sepsis: 1715x49, 3381x49) is streamed from
// L2 with coalesced 128-bit loads.  Plain FP32 FMA throughout (north-star tolerance 1e-4).
#include "ncde_common.cuh"

namespace ancde {

constexpr int CTL_ITEMS = 2;   // control elements (j, b) per thread held in registers

struct CtlRegs {
    float b[CTL_ITEMS], c2[CTL_ITEMS], d3[CTL_ITEMS], a[CTL_ITEMS], hp[CTL_ITEMS];
    float frac;
};

// Issues the global loads of the control of stage (k, s); nothing is consumed here so the
// latency overlaps with the MLP of the current stage.
template <int TS>
__device__ __forceinline__ void ctl_issue(const SolveArgs& p, const float* sTimes, int& cnt, int k, int s,
                                          int b0, int tid, int NT, CtlRegs& r)
{
    const NcdeLayout& lo = p.lo;
    const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step);
    const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step);
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
    const int m = n % (lo.n_hp - 1);                  // call counter with the reset rule (:134-137)
#pragma unroll
    for (int it = 0; it < CTL_ITEMS; ++it) {
        const int item = tid + it * NT;
        r.b[it] = r.c2[it] = r.d3[it] = 0.f;
        r.a[it] = r.hp[it] = 0.f;
        if (item < lo.C4 * TS) {
            const int b = item / lo.C4, j = item - b * lo.C4;
            const int bg = b0 + b;
            if (j < lo.C && bg < lo.B) {
                const int64_t o = ((int64_t)bg * (lo.L - 1) + idx) * lo.C + j;
                r.b[it] = __ldg(p.cb + o);
                r.c2[it] = __ldg(p.cc2 + o);
                r.d3[it] = __ldg(p.cd3 + o);
                if (lo.mode == ANCDE_MODE_TOP) {
                    r.a[it] = __ldg(p.attention + ((int64_t)q * lo.B + bg) * lo.att_C + (lo.att_C == 1 ? 0 : j));
                    r.hp[it] = __ldg(p.h_prime + ((int64_t)m * lo.B + bg) * lo.C + j);
                }
            }
        }
    }
}

// Turns the loaded coefficients into dU (and dU/da for the backward) and stores them [j][TS].
template <int TS>
__device__ __forceinline__ void ctl_finish(const SolveArgs& p, const CtlRegs& r, float* sDU, float* act,
                                           int tid, int NT)
{
    const NcdeLayout& lo = p.lo;
#pragma unroll
    for (int it = 0; it < CTL_ITEMS; ++it) {
        const int item = tid + it * NT;
        if (item < lo.C4 * TS) {
            const int b = item / lo.C4, j = item - b * lo.C4;
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
            sDU[j * TS + b] = dU;
            if (act) {
                act[lo.act_off_du + j * TS + b] = dU;
                if (lo.mode == ANCDE_MODE_TOP) act[lo.act_off_duda + j * TS + b] = duda;
            }
        }
    }
}

// out[o][.] = relu(bias[o] + sum_k Wt[k][o] * in[k][.]) for all TS samples.
template <int TS>
__device__ __forceinline__ void dense_relu(const float* __restrict__ Wt, const float* __restrict__ bias, int K,
                                           int N, int S, const float* __restrict__ in, float* __restrict__ out,
                                           float* __restrict__ gsave, int tid, int NT)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    for (int item = tid; item < N * NQ; item += NT) {
        const int q = item / N, o = item - q * N;
        float acc[VW];
        const float bv = bias[o];
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[e] = bv;
        const float* wp = Wt + o;
        const float* ip = in + q * VW;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float w = wp[k * S];
            float v[VW];
            Vec<VW>::ld(ip + k * TS, v);
#pragma unroll
            for (int e = 0; e < VW; ++e) acc[e] = fmaf(w, v[e], acc[e]);
        }
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[e] = fmaxf(acc[e], 0.f);
        Vec<VW>::st(out + o * TS + q * VW, acc);
        if (gsave) Vec<VW>::st(gsave + o * TS + q * VW, acc);
    }
}

template <int TS, bool WO_SMEM>
__global__ void __launch_bounds__(512, 1) ncde_fwd_kernel(const __grid_constant__ SolveArgs p)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int tile = blockIdx.x;
    const int b0 = tile * TS;
    const int Hd = lo.Hd, C4 = lo.C4, NCH = lo.NCH, NCG = lo.NCG, NCP = lo.NCP, K = lo.K;

    extern __shared__ __align__(16) float smem[];
    float* sHW = smem;                                   // hidden weights + biases (packed)
    float* sp = sHW + pad4(lo.hidden_floats);
    float* sWo = sp;                                     // output layer (only if it fits)
    if (WO_SMEM) sp += (size_t)K * NCP + NCP;
    float* sTimes = sp; sp += pad4(lo.L);
    float* sIn = sp;    sp += pad4(Hd * TS);             // stage input z_s      [i][TS]
    float* sA0 = sp;    sp += pad4(lo.Hmax * TS);        // activations ping
    float* sA1 = sp;    sp += pad4(lo.Hmax * TS);        // activations pong
    float* sY = sp;     sp += pad4(Hd * TS);             // z at the start of the step
    float* sK1 = sp;    sp += pad4(Hd * TS);
    float* sK2 = sp;    sp += pad4(Hd * TS);
    float* sK3 = sp;    sp += pad4(Hd * TS);
    float* sDU = sp;    sp += 2 * pad4(C4 * TS);         // control increment, double buffered [j][TS]
    float* sPart = sp;                                   // per column-group partial dot products

    // ---- prologue: weights, knot times, initial state
    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack);
        float4* dst = reinterpret_cast<float4*>(sHW);
        for (int i = tid; i < pad4(lo.hidden_floats) / 4; i += NT) dst[i] = __ldg(src + i);
        if (WO_SMEM) {
            const float4* s2 = reinterpret_cast<const float4*>(p.wpack + lo.off_wo);
            float4* d2 = reinterpret_cast<float4*>(sWo);
            for (int i = tid; i < (K * NCP + NCP) / 4; i += NT) d2[i] = __ldg(s2 + i);
        }
        for (int i = tid; i < lo.L; i += NT) sTimes[i] = p.times[i];
        for (int item = tid; item < Hd * TS; item += NT) {
            const int i = item / TS, b = item - i * TS;
            const int bg = b0 + b;
            const float v = (bg < lo.B) ? p.z0[(int64_t)bg * Hd + i] : 0.f;
            sIn[item] = v;
            sY[item] = v;
            if (bg < lo.B) p.z_out[(int64_t)bg * Hd + i] = v;          // solution[0] = y0
            if (p.save_act) p.save_act[(int64_t)tile * lo.act_floats + item] = v;
        }
    }
    __syncthreads();
    const float* Wo = WO_SMEM ? sWo : (p.wpack + lo.off_wo);
    const float* bo = Wo + (size_t)K * NCP;

    int cnt = 0;        // knots strictly below the current stage time (monotone)
    int jout = 1;       // next requested output time
    {
        CtlRegs r;
        ctl_issue<TS>(p, sTimes, cnt, 0, 0, b0, tid, NT, r);
        ctl_finish<TS>(p, r, sDU, p.save_act ? p.save_act + (int64_t)tile * lo.act_floats : nullptr, tid, NT);
    }
    __syncthreads();

    for (int k = 0; k < lo.n_steps; ++k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step);
        const float dt = __fsub_rn(t1, t0);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const int n = 4 * k + s;
            float* act = p.save_act ? p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats : nullptr;
            float* act_next = (p.save_act && n + 1 < lo.n_stages)
                                  ? p.save_act + ((int64_t)(n + 1) * lo.ntiles + tile) * lo.act_floats : nullptr;
            const float* sDUcur = sDU + (n & 1) * pad4(C4 * TS);
            float* sDUnext = sDU + ((n + 1) & 1) * pad4(C4 * TS);

            // (1) control of the NEXT stage: loads in flight while this stage's MLP runs
            CtlRegs ctl;
            const bool have_next = (n + 1 < lo.n_stages);
            if (have_next) ctl_issue<TS>(p, sTimes, cnt, s == 3 ? k + 1 : k, s == 3 ? 0 : s + 1, b0, tid, NT, ctl);

            // (2) hidden layers
            const float* in = sIn;
            float* outb = sA0;
            for (int l = 0; l < lo.n_hidden; ++l) {
                dense_relu<TS>(sHW + lo.off_w[l], sHW + lo.off_b[l], lo.in_dim[l], lo.width[l], lo.stride[l], in,
                               outb, act ? act + lo.act_off_h[l] : nullptr, tid, NT);
                __syncthreads();
                in = outb;
                outb = (outb == sA0) ? sA1 : sA0;
            }

            // (3) output layer, tanh, contraction with dU: thread = 4 columns x TS samples
            for (int cg = tid; cg < NCG; cg += NT) {
                float acc[TS][4];
                {
                    const float4 bv = WO_SMEM ? *reinterpret_cast<const float4*>(bo + cg * 4)
                                              : __ldg(reinterpret_cast<const float4*>(bo + cg * 4));
#pragma unroll
                    for (int b = 0; b < TS; ++b) { acc[b][0] = bv.x; acc[b][1] = bv.y; acc[b][2] = bv.z; acc[b][3] = bv.w; }
                }
                const float* wp = Wo + cg * 4;
#pragma unroll 7
                for (int kk = 0; kk < K; ++kk) {
                    const float4 w = WO_SMEM ? *reinterpret_cast<const float4*>(wp + (size_t)kk * NCP)
                                             : __ldg(reinterpret_cast<const float4*>(wp + (size_t)kk * NCP));
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        float v[VW];
                        Vec<VW>::ld(in + kk * TS + q * VW, v);
#pragma unroll
                        for (int e = 0; e < VW; ++e) {
                            acc[q * VW + e][0] = fmaf(v[e], w.x, acc[q * VW + e][0]);
                            acc[q * VW + e][1] = fmaf(v[e], w.y, acc[q * VW + e][1]);
                            acc[q * VW + e][2] = fmaf(v[e], w.z, acc[q * VW + e][2]);
                            acc[q * VW + e][3] = fmaf(v[e], w.w, acc[q * VW + e][3]);
                        }
                    }
                }
                const int chunk = cg % NCH;
                float part[TS];
#pragma unroll
                for (int b = 0; b < TS; ++b) part[b] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float dy[TS];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sDUcur + (chunk * 4 + c) * TS + q * VW, dy + q * VW);
#pragma unroll
                    for (int b = 0; b < TS; ++b) {
                        const float g = tanh_mufu(acc[b][c]);
                        acc[b][c] = g;
                        part[b] = fmaf(g, dy[b], part[b]);
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
            if (have_next) ctl_finish<TS>(p, ctl, sDUnext, act_next, tid, NT);
            __syncthreads();

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
                sIn[item] = zin;
                if (act_next) act_next[item] = zin;
            }
            if (s == 3) {
                while (jout < lo.n_out && t1 >= __ldg(p.t_out + jout)) ++jout;
            }
            __syncthreads();
        }
    }
}

// ---- weight packing: torch (out, in) row-major -> transposed, padded layout of NcdeLayout
struct PackArgs {
    NcdeLayout lo;
    const float* W[ANCDE_MAX_LAYERS + 1];
    const float* bias[ANCDE_MAX_LAYERS + 1];
    float* wpack;
};

__global__ void pack_weights_kernel(const __grid_constant__ PackArgs a)
{
    const NcdeLayout& lo = a.lo;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int l = 0; l < lo.n_hidden; ++l) {
        const int in = lo.in_dim[l], out = lo.width[l], S = lo.stride[l];
        for (int idx = gtid; idx < in * S; idx += gsz) {
            const int k = idx / S, o = idx - k * S;
            a.wpack[lo.off_w[l] + idx] = (o < out) ? a.W[l][(int64_t)o * in + k] : 0.f;
        }
        for (int o = gtid; o < pad4(out); o += gsz) a.wpack[lo.off_b[l] + o] = (o < out) ? a.bias[l][o] : 0.f;
    }
    const int K = lo.K, NCP = lo.NCP, C4 = lo.C4, C = lo.C;
    const float* Wout = a.W[lo.n_hidden];
    const float* bout = a.bias[lo.n_hidden];
    for (int64_t idx = gtid; idx < (int64_t)K * NCP; idx += gsz) {
        const int k = (int)(idx / NCP), col = (int)(idx - (int64_t)k * NCP);
        const int i = col / C4, j = col - i * C4;
        a.wpack[lo.off_wo + idx] = (j < C) ? Wout[((int64_t)i * C + j) * K + k] : 0.f;
    }
    for (int col = gtid; col < NCP; col += gsz) {
        const int i = col / C4, j = col - i * C4;
        a.wpack[lo.off_bo + col] = (j < C) ? bout[i * C + j] : 0.f;
    }
}

// ---- host side -------------------------------------------------------------------------------
bool bwd_fits(const NcdeLayout& lo);

static void fill_layout(const ancde_ncde_desc* d, int TS, NcdeLayout* lo)
{
    *lo = NcdeLayout();
    lo->B = d->B; lo->L = d->L; lo->C = d->C; lo->Hd = d->Hd; lo->n_hidden = d->n_hidden; lo->mode = d->mode;
    lo->att_C = d->att_C; lo->att_L = d->att_L; lo->n_hp = d->n_hp; lo->n_steps = d->n_steps; lo->n_out = d->n_out;
    lo->n_stages = 4 * d->n_steps;
    lo->Hmax = d->Hd;
    for (int l = 0; l < d->n_hidden; ++l) {
        lo->width[l] = d->width[l];
        lo->in_dim[l] = (l == 0) ? d->Hd : d->width[l - 1];
        lo->stride[l] = d->width[l] | 1;
        if (d->width[l] > lo->Hmax) lo->Hmax = d->width[l];
    }
    lo->TS = TS;
    lo->ntiles = (d->B + TS - 1) / TS;
    lo->C4 = pad4(d->C);
    lo->NCH = lo->C4 / 4;
    lo->NCG = d->Hd * lo->NCH;
    lo->NCP = d->Hd * lo->C4;
    lo->K = d->width[d->n_hidden - 1];
    int off = 0;
    for (int l = 0; l < d->n_hidden; ++l) {
        lo->off_w[l] = off; off += pad4(lo->in_dim[l] * lo->stride[l]);
        lo->off_b[l] = off; off += pad4(lo->width[l]);
    }
    lo->hidden_floats = off;
    lo->off_wo = off; off += lo->K * lo->NCP;
    lo->off_bo = off; off += lo->NCP;
    lo->wpack_floats = off;
    int a = d->Hd * TS;
    for (int l = 0; l < d->n_hidden; ++l) { lo->act_off_h[l] = a; a += d->width[l] * TS; }
    lo->act_off_du = a; a += lo->C4 * TS;
    lo->act_off_duda = a;
    if (d->mode == ANCDE_MODE_TOP) a += lo->C4 * TS;
    lo->act_floats = pad4(a);
}

int make_layout(const ancde_ncde_desc* d, NcdeLayout* lo)
{
    ANCDE_CHECK_ARG(d != nullptr, "ncde: null descriptor");
    ANCDE_CHECK_ARG(d->B >= 1 && d->L >= 2 && d->C >= 1 && d->Hd >= 1, "ncde: bad shape B=%d L=%d C=%d Hd=%d",
                    d->B, d->L, d->C, d->Hd);
    ANCDE_CHECK_ARG(d->n_hidden >= 1 && d->n_hidden <= ANCDE_MAX_LAYERS, "ncde: n_hidden=%d out of range",
                    d->n_hidden);
    ANCDE_CHECK_ARG(d->mode >= ANCDE_MODE_PLAIN && d->mode <= ANCDE_MODE_TOP, "ncde: bad mode %d", d->mode);
    ANCDE_CHECK_ARG(d->n_steps >= 1 && d->n_out >= 1, "ncde: n_steps=%d n_out=%d", d->n_steps, d->n_out);
    ANCDE_CHECK_ARG(d->step > 0.f, "ncde: step must be positive");
    if (d->mode == ANCDE_MODE_TOP) {
        ANCDE_CHECK_ARG(d->att_C == 1 || d->att_C == d->C, "ncde: attention channels %d (need 1 or C=%d)", d->att_C, d->C);
        ANCDE_CHECK_ARG(d->att_L >= 1 && d->n_hp >= 2, "ncde: att_L=%d n_hp=%d", d->att_L, d->n_hp);
    }
    for (int l = 0; l < d->n_hidden; ++l) ANCDE_CHECK_ARG(d->width[l] >= 1, "ncde: width[%d]=%d", l, d->width[l]);
    int TS = d->samples_per_cta;
    if (TS == 0) {
        // the largest tile that still gives every SM work (148 SMs); weight reuse grows with TS.
        // The backward's shared-memory footprint grows with TS too, so shrink until it fits.
        TS = 8;
        while (TS > 1 && (d->B + TS - 1) / TS < 120) TS /= 2;
        for (;;) {
            fill_layout(d, TS, lo);
            if (TS == 1 || bwd_fits(*lo)) break;
            TS /= 2;
        }
    }
    ANCDE_CHECK_ARG(TS == 8 || TS == 4 || TS == 2 || TS == 1, "ncde: samples_per_cta=%d (need 8, 4, 2 or 1)", TS);
    fill_layout(d, TS, lo);
    return ANCDE_OK;
}

static int fwd_threads(const NcdeLayout& lo)
{
    int per = (lo.NCG + 511) / 512;
    int nt = round_up((lo.NCG + per - 1) / per, 32);
    int need = (lo.C4 * lo.TS + CTL_ITEMS - 1) / CTL_ITEMS;
    if (nt < need) nt = round_up(need, 32);
    if (nt < 64) nt = 64;
    return nt;
}

static size_t fwd_smem_floats(const NcdeLayout& lo, bool wo_smem)
{
    size_t f = pad4(lo.hidden_floats);
    if (wo_smem) f += (size_t)lo.K * lo.NCP + lo.NCP;
    f += pad4(lo.L) + pad4(lo.Hd * lo.TS) + 2 * pad4(lo.Hmax * lo.TS) + 4 * pad4(lo.Hd * lo.TS) +
         2 * pad4(lo.C4 * lo.TS) + (size_t)lo.NCG * lo.TS;
    return f;
}

template <int TS>
static int launch_fwd(const SolveArgs& a, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    const int nt = fwd_threads(lo);
    if (nt > 512) {
        set_error("ncde_forward: C=%d with %d samples per CTA needs %d threads (max 512)", lo.C, lo.TS, nt);
        return ANCDE_EUNSUPPORTED;
    }
    const size_t max_smem = 227 * 1024;
    bool wo_smem = fwd_smem_floats(lo, true) * 4 <= max_smem;
    const size_t smem = fwd_smem_floats(lo, wo_smem) * 4;
    if (smem > max_smem) {
        set_error("ncde_forward: hidden layers need %zu bytes of shared memory (max %zu)", smem, max_smem);
        return ANCDE_EUNSUPPORTED;
    }
    static const char* names[3] = {"ncde_fwd_plain", "ncde_fwd_bottom", "ncde_fwd_top"};
    if (wo_smem) {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_fwd_kernel<TS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_fwd_kernel<TS, true><<<lo.ntiles, nt, smem, stream>>>(a);
    } else {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_fwd_kernel<TS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_fwd_kernel<TS, false><<<lo.ntiles, nt, smem, stream>>>(a);
    }
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

int pack_weights(const ancde_ncde_desc* d, const NcdeLayout& lo, cudaStream_t stream)
{
    PackArgs pa;
    pa.lo = lo;
    for (int l = 0; l <= d->n_hidden; ++l) {
        ANCDE_CHECK_ARG(d->W[l] && d->bias[l], "ncde: null weight/bias pointer for layer %d", l);
        pa.W[l] = d->W[l];
        pa.bias[l] = d->bias[l];
    }
    pa.wpack = d->wpack;
    prof_begin("pack_weights", stream);
    pack_weights_kernel<<<64, 256, 0, stream>>>(pa);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

int fill_args(const ancde_ncde_desc* d, const NcdeLayout& lo, SolveArgs* a)
{
    ANCDE_CHECK_ARG(d->times && d->t_out && d->coeff_b && d->coeff_c2 && d->coeff_d3 && d->z0 && d->z_out && d->wpack,
                    "ncde: null input/output pointer");
    if (d->mode == ANCDE_MODE_TOP) ANCDE_CHECK_ARG(d->attention && d->h_prime, "ncde: TOP mode needs attention and h_prime");
    *a = SolveArgs();
    a->lo = lo;
    a->t_start = d->t_start; a->t_end = d->t_end; a->step = d->step;
    a->times = d->times; a->t_out = d->t_out; a->cb = d->coeff_b; a->cc2 = d->coeff_c2; a->cd3 = d->coeff_d3;
    a->z0 = d->z0; a->attention = d->attention; a->h_prime = d->h_prime; a->wpack = d->wpack;
    a->z_out = d->z_out; a->k_out = d->k_out; a->save_act = d->save_act; a->save_G = d->save_G;
    a->grad_z_out = d->grad_z_out; a->grad_z0 = d->grad_z0; a->grad_attention = d->grad_attention;
    return ANCDE_OK;
}

}  // namespace ancde

using namespace ancde;

extern "C" int64_t ancde_ncde_wpack_floats(const ancde_ncde_desc* d)
{
    NcdeLayout lo;
    if (make_layout(d, &lo) != ANCDE_OK) return -1;
    return lo.wpack_floats + pad4(lo.hidden_floats);   // + hidden-gradient scratch used by the backward
}

extern "C" int64_t ancde_ncde_save_act_floats(const ancde_ncde_desc* d)
{
    NcdeLayout lo;
    if (make_layout(d, &lo) != ANCDE_OK) return -1;
    return (int64_t)lo.n_stages * lo.ntiles * lo.act_floats;
}

extern "C" int64_t ancde_ncde_save_G_floats(const ancde_ncde_desc* d)
{
    NcdeLayout lo;
    if (make_layout(d, &lo) != ANCDE_OK) return -1;
    return (int64_t)lo.n_stages * lo.ntiles * lo.TS * lo.NCP;
}

extern "C" int ancde_ncde_forward(const ancde_ncde_desc* d, void* stream_)
{
    reset_launch_count();
    cudaStream_t stream = (cudaStream_t)stream_;
    NcdeLayout lo;
    int rc = make_layout(d, &lo);
    if (rc != ANCDE_OK) return rc;
    ANCDE_CHECK_ARG((d->save_act == nullptr) == (d->save_G == nullptr), "ncde_forward: save_act and save_G go together");
    SolveArgs a;
    rc = fill_args(d, lo, &a);
    if (rc != ANCDE_OK) return rc;
    rc = pack_weights(d, lo, stream);
    if (rc != ANCDE_OK) return rc;
    switch (lo.TS) {
        case 8: return launch_fwd<8>(a, stream);
        case 4: return launch_fwd<4>(a, stream);
        case 2: return launch_fwd<2>(a, stream);
        default: return launch_fwd<1>(a, stream);
    }
}
