// The lane engine (ANCDE_ENGINE_LANE): the latency regime of the fused NCDE solve -- small batches of small networks
// (UEA CharacterTrajectories: 32 samples, 724 dependent RK4 stages; Google Stock: 200 samples).
//
// The streamed engine gives a tile of samples a whole CTA and pays, per stage, two or three CTA barriers, a hidden-layer
// chain whose every 128-bit shared load is waited for, and index arithmetic spread over a few dozen threads: 5-7 k cycles
// per stage for a network that needs ~10 k FMAs.  With 32 samples there is nothing to overlap that latency with, so here
// ONE WARP owns one sample from the first to the last stage and nothing but __syncwarp separates the phases:
//   * control (spline derivative, attention product): lane = control channel j = lane % C4, loads issued a stage ahead
//     [cdeint_module.py:58-60,103-129; interpolate.py:261-303]
//   * hidden ReLU layers: lane = unit (two units for layers wider than 32); the layer input is read from shared memory
//     as 128-bit broadcasts and the lane's weight row as 128-bit pieces, ALL loads of a dot product issued before its
//     first FMA (fully unrolled, four accumulators)          [vector_fields.py:205-216,237-247]
//   * output layer: lane owns the columns lane, lane + 32, ... of the (Hd x C4) field = rows i0 + (32/C4) m at ITS channel
//     j, so tanh(.) * dU needs no data exchange and the contraction over j is log2(C4) shuffles inside groups of C4 lanes
//     [vector_fields.py:216-218, cdeint_module.py:61,133]
//   * the RK4 recurrence (3/8 rule) of those rows lives in registers, replicated over the C4 lanes of a group
//     [torchdiffeq rk4_alt_step_func]
// The backward walks the stages in reverse with the same ownership: saved activations and tanh outputs of a stage arrive
// by two TMA bulk copies on a per-warp mbarrier three stages ahead; the output layer's vector-Jacobian product is 4Q
// accumulators per lane over its columns followed by a shuffle reduce-scatter; hidden layers backwards lane = input unit.
// Both kernels read / write the streamed engine's TS = 1 layout (NcdeLayout), so the weight-gradient kernels and the
// packed weights are shared.  C4 must divide 32 (C <= 32, C4 in {4, 8, 16, 32}), state and hidden widths <= 64.
#include "ncde_lane.cuh"
#include "pipeline.cuh"

namespace ancde {

namespace {

constexpr int LANE_NBUF = 4;       // backward: stages of saved activations in flight per warp

// pieces of 4 inputs of the layer input (zero beyond nk4: the matching weights are never loaded)
template <int Q>
__device__ __forceinline__ void lane_load_x(const float* __restrict__ in, int nk4, float4 (&x)[Q])
{
#pragma unroll
    for (int q = 0; q < Q; ++q)
        x[q] = (q < nk4) ? reinterpret_cast<const float4*>(in)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// init + sum_k wrow[k] * x[k]: every load is issued before the first FMA, four independent accumulators
template <int Q>
__device__ __forceinline__ float lane_dot(const float* __restrict__ wrow, int nk4, const float4 (&x)[Q], float init)
{
    float4 w[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q)
        w[q] = (q < nk4) ? reinterpret_cast<const float4*>(wrow)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
    float a0 = init, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        a0 = fmaf(w[q].x, x[q].x, a0);
        a1 = fmaf(w[q].y, x[q].y, a1);
        a2 = fmaf(w[q].z, x[q].z, a2);
        a3 = fmaf(w[q].w, x[q].w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// Sum of v[idx] over the 32 lanes.  V = 32: lane l ends with the total of idx l in v[0]; V = 8: lane l ends with the total
// of idx l / 4 in v[0].
template <int V>
__device__ __forceinline__ void lane_reduce_scatter(float (&v)[V], int lane)
{
    int mask = 16;
#pragma unroll
    for (int half = V / 2; half >= 1; half /= 2) {
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
        mask >>= 1;
    }
    if (V < 32) {
#pragma unroll
        for (int m = 16 / V; m >= 1; m >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
    }
}

struct LaneCtl { float b, c2, d3, a, hp, frac; };

// Global loads of the control of stage n = (k, s) for this lane's channel j; consumed a stage later.
__device__ __forceinline__ void lane_ctl_issue(const SolveArgs& p, const float* sTimes, int& cnt, int k, int s, bool jvalid,
                                               int obase, int abase, int hbase, LaneCtl& r)
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
    r.b = r.c2 = r.d3 = r.a = r.hp = 0.f;
    if (jvalid) {
        const int o = obase + idx * lo.C;
        r.b = __ldg(p.cb + o);
        r.c2 = __ldg(p.cc2 + o);
        r.d3 = __ldg(p.cd3 + o);
        if (lo.mode == ANCDE_MODE_TOP) {
            int q = (int)floorf(ts) - 1;                      // attention[int(floor(t)) - 1], python wrap (Q3)
            if (q < 0) q += lo.att_L;
            q = q < 0 ? 0 : (q >= lo.att_L ? lo.att_L - 1 : q);
            const int m = (4 * k + s) % (lo.n_hp - 1);        // call counter + reset rule (cdeint_module.py:134-137)
            r.a = __ldg(p.attention + (int64_t)q * lo.B * lo.att_C + abase);
            r.hp = __ldg(p.h_prime + (int64_t)m * lo.B * lo.C + hbase);
        }
    }
}

// dU (and dU/da) of this lane's channel from the loaded coefficients (ctl_finish of the streamed engine)
__device__ __forceinline__ void lane_ctl_finish(int mode, const LaneCtl& r, float& dU, float& duda)
{
    // derivative = b + (two_c + three_d * frac) * frac   (interpolate.py:298-303)
    const float dX = __fadd_rn(r.b, __fmul_rn(__fadd_rn(r.c2, __fmul_rn(r.d3, r.frac)), r.frac));
    dU = dX;
    duda = 0.f;
    if (mode == ANCDE_MODE_TOP) {
        // dY = dX*a + ((a*(1-a))*dX)*h'   (cdeint_module.py:113-129)
        dU = __fadd_rn(__fmul_rn(dX, r.a), __fmul_rn(__fmul_rn(__fmul_rn(r.a, __fsub_rn(1.0f, r.a)), dX), r.hp));
        duda = dX * (1.0f + (1.0f - 2.0f * r.a) * r.hp);
    }
}

// ---------------------------------------------------------------------------------------------------------- forward
template <int MC, int Q>
__global__ void __launch_bounds__(256, 1) ncde_lane_fwd_kernel(const __grid_constant__ SolveArgs p)
{
    constexpr int U = (Q > 8) ? 2 : 1;
    constexpr int HP = 4 * Q;
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, WPC = NT >> 5;
    const int bg = blockIdx.x * WPC + warp;                 // this warp's sample = its tile (TS = 1)
    const int Hd = lo.Hd, C4 = lo.C4, NCP = lo.NCP, KPS = lo.KPS;
    const int sh = 31 - __clz(C4);                          // log2(C4)
    const int RP = 32 >> sh;                                // field rows per pass of 32 columns
    const int j = lane & (C4 - 1), i0 = lane >> sh;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sp = reinterpret_cast<float*>(smem_raw);
    const float* sHW = sp;  sp += pad4(lo.hidden_floats);   // hidden weights + biases (packed rows W[o][stride])
    const float* sBo = sp;  sp += NCP;
    const float* sWot = sp; sp += (size_t)NCP * KPS;        // output layer W_out[col][KPS] (k contiguous)
    float* sTimes = sp;     sp += pad4(lo.L);
    int* sLp = reinterpret_cast<int*>(sp); sp += 8 * ANCDE_MAX_LAYERS;
    float* sIn = sp + (size_t)warp * 3 * HP;                // per warp: stage input, activations ping / pong
    float* sT0 = sIn + HP;
    float* sT1 = sT0 + HP;

    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack);
        float4* dst = reinterpret_cast<float4*>(const_cast<float*>(sHW));
        for (int i = tid; i < pad4(lo.hidden_floats) / 4; i += NT) dst[i] = __ldg(src + i);
        const float4* sb = reinterpret_cast<const float4*>(p.wpack + lo.off_bo);
        for (int i = tid; i < NCP / 4; i += NT) reinterpret_cast<float4*>(const_cast<float*>(sBo))[i] = __ldg(sb + i);
        const float4* sw = reinterpret_cast<const float4*>(p.wpack + lo.off_wot);
        for (int i = tid; i < NCP * KPS / 4; i += NT) reinterpret_cast<float4*>(const_cast<float*>(sWot))[i] = __ldg(sw + i);
        for (int i = tid; i < lo.L; i += NT) sTimes[i] = p.times[i];
        if (tid < lo.n_hidden) {
            int* q = sLp + 8 * tid;
            q[0] = (lo.in_dim[tid] + 3) / 4; q[1] = lo.width[tid]; q[2] = lo.stride[tid]; q[3] = lo.off_w[tid];
            q[4] = lo.off_b[tid]; q[5] = lo.act_off_h[tid]; q[6] = 0; q[7] = 0;
        }
        for (int i = tid; i < WPC * 3 * HP; i += NT) (sIn - (size_t)warp * 3 * HP)[i] = 0.f;
    }
    __syncthreads();
    if (bg >= lo.B) return;                                  // no CTA barrier below

    const int tile = bg;
    const bool jvalid = j < lo.C;
    const int obase = bg * (lo.L - 1) * lo.C + j;
    const int abase = bg * lo.att_C + (lo.att_C == 1 ? 0 : j);
    const int hbase = bg * lo.C + j;
    const int duidx = (j >> 2) * lo.DUS + (j & 3);           // du_index(j, 0, 1, DUS)
    const int nk4o = (lo.K + 3) >> 2;
    const bool saving = p.save_act != nullptr;

    float y[MC], k1[MC], k2[MC], k3[MC];
#pragma unroll
    for (int m = 0; m < MC; ++m) {
        const int i = i0 + RP * m;
        const bool v = i < Hd;
        y[m] = v ? p.z0[(int64_t)bg * Hd + i] : 0.f;
        k1[m] = k2[m] = k3[m] = 0.f;
        if (v && j == 0) {
            sIn[i] = y[m];
            p.z_out[(int64_t)bg * Hd + i] = y[m];            // solution[0] = y0
            if (saving) p.save_act[(int64_t)tile * lo.act_floats + i] = y[m];
        }
    }
    int cnt = 0;        // knots strictly below the current stage time (monotone)
    int jout = 1;       // next requested output time
    LaneCtl ctl;
    lane_ctl_issue(p, sTimes, cnt, 0, 0, jvalid, obase, abase, hbase, ctl);
    __syncwarp();

    for (int k = 0; k < lo.n_steps; ++k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float dt = __fsub_rn(t1, t0);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const int n = 4 * k + s;
            float* act = saving ? p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats : nullptr;
            float* act_next = (saving && n + 1 < lo.n_stages)
                                  ? p.save_act + ((int64_t)(n + 1) * lo.ntiles + tile) * lo.act_floats : nullptr;
            // (1) control of this stage from the registers loaded a stage ago; the loads of the next stage go out now
            float dU, duda;
            lane_ctl_finish(lo.mode, ctl, dU, duda);
            if (n + 1 < lo.n_stages)
                lane_ctl_issue(p, sTimes, cnt, s == 3 ? k + 1 : k, s == 3 ? 0 : s + 1, jvalid, obase, abase, hbase, ctl);
            if (act && lane < C4) {
                act[lo.act_off_du + duidx] = dU;
                if (lo.mode == ANCDE_MODE_TOP) act[lo.act_off_duda + duidx] = duda;
            }

            // (2) hidden layers
            const float* in = sIn;
            for (int l = 0; l < lo.n_hidden; ++l) {
                const int4 q0 = *reinterpret_cast<const int4*>(sLp + 8 * l);
                const int2 q1 = *reinterpret_cast<const int2*>(sLp + 8 * l + 4);
                const int nk4 = q0.x, N = q0.y;
                float* out = (l & 1) ? sT1 : sT0;
                float4 x[Q];
                lane_load_x<Q>(in, nk4, x);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (u == 1 && N <= 32) break;
                    const int o = lane + 32 * u;
                    const bool valid = o < N;
                    const int oc = valid ? o : N - 1;
                    float v = lane_dot<Q>(sHW + q0.w + oc * q0.z, nk4, x, sHW[q1.x + oc]);
                    v = valid ? fmaxf(v, 0.f) : 0.f;
                    if (o < HP) out[o] = v;
                    if (valid && act) act[q1.y + o] = v;
                }
                __syncwarp();
                in = out;
            }

            // (3) output layer, tanh, contraction with dU over the C4 lanes of a group
            float ko[MC];
            {
                float4 x[Q];
                lane_load_x<Q>(in, nk4o, x);
                float* gsave = p.save_G ? p.save_G + ((int64_t)n * lo.ntiles + tile) * NCP : nullptr;
#pragma unroll
                for (int m = 0; m < MC; ++m) {
                    ko[m] = 0.f;
                    if (32 * m < NCP) {
                        const int col = lane + 32 * m;
                        const bool cv = col < NCP;
                        const int cc = cv ? col : NCP - 1;
                        const float pre = lane_dot<Q>(sWot + (size_t)cc * KPS, nk4o, x, sBo[cc]);
                        const float g = cv ? tanh_mufu(pre) : 0.f;
                        if (gsave && cv) gsave[col] = g;
                        float part = g * dU;
                        for (int o = 1; o < C4; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                        ko[m] = part;
                    }
                }
            }

            // (4) the RK4 recurrence of this lane's rows (rk4_alt_step_func)
#pragma unroll
            for (int m = 0; m < MC; ++m) {
                const int i = i0 + RP * m;
                const bool v = i < Hd;
                const bool wr = v && j == 0;
                const float o = ko[m];
                if (wr && p.k_out) p.k_out[((int64_t)n * lo.B + bg) * Hd + i] = o;
                const float yy = y[m];
                float zin;
                if (s == 0) {
                    k1[m] = o;
                    zin = __fadd_rn(yy, __fdiv_rn(__fmul_rn(dt, o), 3.0f));                       // y + dt*k1/3
                } else if (s == 1) {
                    k2[m] = o;
                    zin = __fadd_rn(yy, __fmul_rn(dt, __fsub_rn(o, __fdiv_rn(k1[m], 3.0f))));     // y + dt*(k2 - k1/3)
                } else if (s == 2) {
                    k3[m] = o;
                    zin = __fadd_rn(yy, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[m], k2[m]), o)));    // y + dt*(k1-k2+k3)
                } else {
                    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(k1[m], __fmul_rn(3.0f, k2[m])), __fmul_rn(3.0f, k3[m])), o);
                    const float y1 = __fadd_rn(yy, __fmul_rn(sum, __fdiv_rn(dt, 8.0f)));
                    // emit every requested time in (t0, t1] (FixedGridODESolver.integrate + _linear_interp)
                    if (wr) {
                        int jj = jout;
                        while (jj < lo.n_out) {
                            const float tj = __ldg(p.t_out + jj);
                            if (!(t1 >= tj)) break;
                            float val;
                            if (tj == t0) val = yy;
                            else if (tj == t1) val = y1;
                            else val = __fadd_rn(yy, __fmul_rn(__fdiv_rn(__fsub_rn(y1, yy), __fsub_rn(t1, t0)), __fsub_rn(tj, t0)));
                            p.z_out[((int64_t)jj * lo.B + bg) * Hd + i] = val;
                            ++jj;
                        }
                    }
                    y[m] = y1;
                    zin = y1;
                }
                if (wr) {
                    sIn[i] = zin;
                    if (act_next) act_next[i] = zin;
                }
            }
            if (s == 3) {
                while (jout < lo.n_out && t1 >= __ldg(p.t_out + jout)) ++jout;
            }
            __syncwarp();
        }
    }
}

// --------------------------------------------------------------------------------------------------------- backward
template <int MC, int Q>
__global__ void __launch_bounds__(256, 1) ncde_lane_bwd_kernel(const __grid_constant__ SolveArgs p)
{
    constexpr int U = (Q > 8) ? 2 : 1;
    constexpr int HP = 4 * Q;
    constexpr int VB = 4 * Q - 32;                           // accumulators of the hidden units 32.. (0, 8 or 32)
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, WPC = NT >> 5;
    const int bg = blockIdx.x * WPC + warp;
    const int Hd = lo.Hd, C4 = lo.C4, NCP = lo.NCP, KPS = lo.KPS, K = lo.K;
    const int sh = 31 - __clz(C4);
    const int RP = 32 >> sh;
    const int j = lane & (C4 - 1), i0 = lane >> sh;
    const int AF = lo.act_fwd_floats;
    const bool need_att = (lo.mode == ANCDE_MODE_TOP) && (p.grad_attention != nullptr);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * LANE_NBUF;       // [WPC][LANE_NBUF]
    float* sp = reinterpret_cast<float*>(smem_raw + (size_t)WPC * LANE_NBUF * 8);
    const float* sWtb = sp; sp += pad4(lo.hiddenb_floats);   // transposed hidden weights Wt_l[k][tstride]
    const float* sWot = sp; sp += (size_t)NCP * KPS;
    int* sLp = reinterpret_cast<int*>(sp); sp += 8 * ANCDE_MAX_LAYERS;
    float* wbase = sp + (size_t)warp * (LANE_NBUF * (AF + NCP) + 3 * HP);
    float* sActBuf = wbase;                                  // [LANE_NBUF][AF]
    float* sGBuf = sActBuf + (size_t)LANE_NBUF * AF;         // [LANE_NBUF][NCP]
    float* sD0 = sGBuf + (size_t)LANE_NBUF * NCP;
    float* sD1 = sD0 + HP;
    float* sZb = sD1 + HP;

    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack + lo.off_hiddenb);
        float4* dst = reinterpret_cast<float4*>(const_cast<float*>(sWtb));
        for (int i = tid; i < pad4(lo.hiddenb_floats) / 4; i += NT) dst[i] = __ldg(src + i);
        const float4* sw = reinterpret_cast<const float4*>(p.wpack + lo.off_wot);
        for (int i = tid; i < NCP * KPS / 4; i += NT) reinterpret_cast<float4*>(const_cast<float*>(sWot))[i] = __ldg(sw + i);
        if (tid < lo.n_hidden) {
            const int l = tid;
            int* q = sLp + 8 * l;
            q[0] = (lo.width[l] + 3) / 4; q[1] = lo.in_dim[l]; q[2] = lo.tstride[l]; q[3] = lo.off_wt[l] - lo.off_hiddenb;
            q[4] = l > 0 ? lo.act_off_h[l - 1] : 0; q[5] = l > 0 ? lo.act_off_delta[l - 1] : 0; q[6] = 0; q[7] = 0;
        }
        for (int i = lane; i < 3 * HP; i += 32) sD0[i] = 0.f;
        if (lane == 0) {
            for (int s = 0; s < LANE_NBUF; ++s) mbar_init(bars + s, 1);
            mbar_fence_init();
        }
    }
    __syncthreads();
    if (bg >= lo.B) return;                                  // no CTA barrier below

    const int tile = bg;
    const int duidx = (j >> 2) * lo.DUS + (j & 3);
    const int nk4o = (K + 3) >> 2;
    const int NS = lo.n_stages;
    // saved activations (forward part of the block) and tanh outputs of reverse stage r = NS - 1 - n -> ring slot r % NBUF
    auto issue = [&](int r) {
        if (lane == 0 && r < NS) {
            const int n = NS - 1 - r, slot = r % LANE_NBUF;
            mbar_arrive_expect_tx(bars + slot, (uint32_t)((AF + NCP) * 4));
            bulk_g2s(sActBuf + (size_t)slot * AF, p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats, (uint32_t)(AF * 4), bars + slot);
            bulk_g2s(sGBuf + (size_t)slot * NCP, p.save_G + ((int64_t)n * lo.ntiles + tile) * NCP, (uint32_t)(NCP * 4), bars + slot);
        }
    };
    for (int r = 0; r < LANE_NBUF - 1; ++r) issue(r);

    float lam[MC], mu[MC], V4[MC], V3[MC], V2[MC];
#pragma unroll
    for (int m = 0; m < MC; ++m) lam[m] = mu[m] = V4[m] = V3[m] = V2[m] = 0.f;
    int jb = lo.n_out - 1;   // last output time whose cotangent has not been consumed yet
    float gmax = 0.f;
    int r = 0;

    for (int k = lo.n_steps - 1; k >= 0; --k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float dt = __fsub_rn(t1, t0);
        // ---- cotangents of the outputs emitted during this step: t_out[jj] in (t0, t1]
        for (int jj = jb; jj >= 1; --jj) {
            const float tj = __ldg(p.t_out + jj);
            if (!(tj > t0)) break;
            const float th = (tj - t0) / (t1 - t0);
#pragma unroll
            for (int m = 0; m < MC; ++m) {
                const int i = i0 + RP * m;
                const float gz = (i < Hd) ? __ldg(p.grad_z_out + ((int64_t)jj * lo.B + bg) * Hd + i) : 0.f;
                if (tj == t1) lam[m] += gz;
                else {
                    lam[m] += th * gz;
                    mu[m] += (1.0f - th) * gz;
                }
            }
        }
        while (jb >= 1 && __ldg(p.t_out + jb) > t0) --jb;

#pragma unroll 1
        for (int s = 3; s >= 0; --s, ++r) {
            const int n = 4 * k + s;
            float* act = p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats;
            const int slot = r % LANE_NBUF;
            issue(r + LANE_NBUF - 1);                        // its slot was read last in stage r - 1 (behind a __syncwarp)
            // ---- cotangent of the stage derivative k_{s+1} (transpose of rk4_alt_step_func)
            float gv[MC];
#pragma unroll
            for (int m = 0; m < MC; ++m) {
                const int i = i0 + RP * m;
                float g;
                if (s == 3) g = dt * 0.125f * lam[m];
                else if (s == 2) g = dt * (0.375f * lam[m] + V4[m]);
                else if (s == 1) g = dt * (0.375f * lam[m] - V4[m] + V3[m]);
                else g = dt * (0.125f * lam[m] + V4[m] + (V2[m] - V3[m]) * (1.0f / 3.0f));
                g = (i < Hd) ? g : 0.f;
                gv[m] = g;
                if (i < Hd && j == 0) act[lo.act_off_gbar + i] = g;      // for the weight-gradient kernels
                gmax = fmaxf(gmax, fabsf(g));
            }
            mbar_wait(bars + slot, (uint32_t)((r / LANE_NBUF) & 1));
            const float* sAct = sActBuf + (size_t)slot * AF;
            const float* sG = sGBuf + (size_t)slot * NCP;
            const float dU = sAct[lo.act_off_du + duidx];

            // ---- pre_bar = gbar * dU * (1 - G^2); cotangent of dU for the attention
            float pre[MC];
            float e = 0.f;
#pragma unroll
            for (int m = 0; m < MC; ++m) {
                const int col = lane + 32 * m;
                const float g = (col < NCP) ? sG[col] : 0.f;
                pre[m] = gv[m] * dU * fmaf(-g, g, 1.0f);
                e = fmaf(gv[m], g, e);
            }
            if (need_att) {
                // dUbar[j] = sum_i gbar[i] G[i][j]: this lane's rows, then the lanes of the other row groups
                const float ts = stage_time(t0, dt, s);
                int qa = (int)floorf(ts) - 1;
                if (qa < 0) qa += lo.att_L;
                qa = qa < 0 ? 0 : (qa >= lo.att_L ? lo.att_L - 1 : qa);
                float a = e * sAct[lo.act_off_duda + duidx];
                for (int o = C4; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lo.att_C == 1) {
                    for (int o = 1; o < C4; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (lane == 0) atomicAdd(p.grad_attention + (int64_t)qa * lo.B + bg, a);
                } else if (lane < lo.C) {
                    atomicAdd(p.grad_attention + ((int64_t)qa * lo.B + bg) * lo.C + lane, a);
                }
            }

            // ---- hbar[k] = sum_col W_out[col][k] pre_bar[col]: 4Q accumulators over this lane's columns, then a shuffle
            // reduce-scatter: unit o < 32 ends in lane o, unit 32 + o in lane o (VB = 32) or lanes 4o .. 4o+3 (VB = 8)
            float hb[4 * Q];
#pragma unroll
            for (int i = 0; i < 4 * Q; ++i) hb[i] = 0.f;
#pragma unroll
            for (int m = 0; m < MC; ++m) {
                if (32 * m < NCP) {
                    const int col = lane + 32 * m;
                    const int cc = col < NCP ? col : NCP - 1;       // pre[m] = 0 beyond NCP
                    const float4* wr = reinterpret_cast<const float4*>(sWot + (size_t)cc * KPS);
                    float4 w[Q];
#pragma unroll
                    for (int q = 0; q < Q; ++q) w[q] = (q < nk4o) ? wr[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float pm = pre[m];
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        hb[4 * q + 0] = fmaf(w[q].x, pm, hb[4 * q + 0]);
                        hb[4 * q + 1] = fmaf(w[q].y, pm, hb[4 * q + 1]);
                        hb[4 * q + 2] = fmaf(w[q].z, pm, hb[4 * q + 2]);
                        hb[4 * q + 3] = fmaf(w[q].w, pm, hb[4 * q + 3]);
                    }
                }
            }
            float ha[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) ha[i] = hb[i];
            lane_reduce_scatter<32>(ha, lane);
            {
                // delta of the last hidden layer: hbar * relu'(h)
                const int l = lo.n_hidden - 1;
                const float* hl = sAct + lo.act_off_h[l];
                float* dsave = act + lo.act_off_delta[l];
                {
                    const int o = lane;
                    const bool valid = o < K;
                    const float v = (valid && hl[valid ? o : 0] > 0.f) ? ha[0] : 0.f;
                    sD0[o] = v;
                    if (valid) dsave[o] = v;
                }
                if constexpr (VB == 32) {
                    float hc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) hc[i] = hb[32 + i];
                    lane_reduce_scatter<32>(hc, lane);
                    const int o = 32 + lane;
                    const bool valid = o < K;
                    const float v = (valid && hl[valid ? o : 0] > 0.f) ? hc[0] : 0.f;
                    sD0[o] = v;
                    if (valid) dsave[o] = v;
                } else if constexpr (VB == 8) {
                    float hc[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) hc[i] = hb[32 + i];
                    lane_reduce_scatter<8>(hc, lane);
                    const int o = 32 + (lane >> 2);
                    const bool valid = o < K;
                    const float v = (valid && hl[valid ? o : 0] > 0.f) ? hc[0] : 0.f;
                    if ((lane & 3) == 0) {
                        sD0[o] = v;
                        if (valid) dsave[o] = v;
                    }
                }
            }
            __syncwarp();

            // ---- hidden layers backwards: delta_{l-1} = (W_l^T delta_l) * relu'(h_{l-1}); the stage-input cotangent is W_0^T delta_0
            {
                const float* in = sD0;
                int pp = 1;
                for (int l = lo.n_hidden - 1; l >= 0; --l) {
                    const int4 q0 = *reinterpret_cast<const int4*>(sLp + 8 * l);
                    const int2 q1 = *reinterpret_cast<const int2*>(sLp + 8 * l + 4);
                    const int nk4 = q0.x, N = q0.y;           // pieces of the layer's outputs, inputs of the layer
                    float* out = pp ? sD1 : sD0;
                    float4 x[Q];
                    lane_load_x<Q>(in, nk4, x);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (u == 1 && N <= 32) break;
                        const int kin = lane + 32 * u;
                        const bool valid = kin < N;
                        const int kc = valid ? kin : N - 1;
                        float v = lane_dot<Q>(sWtb + q0.w + kc * q0.z, nk4, x, 0.f);
                        if (l > 0) {
                            v = (valid && sAct[q1.x + kc] > 0.f) ? v : 0.f;
                            if (kin < HP) out[kin] = v;
                            if (valid) act[q1.y + kin] = v;   // for the hidden weight-gradient kernel
                        } else if (valid) {
                            sZb[kin] = v;                     // cotangent of the stage input
                        }
                    }
                    __syncwarp();
                    in = out;
                    pp ^= 1;
                }
            }

            // ---- RK4 adjoint recurrence of this lane's rows
#pragma unroll
            for (int m = 0; m < MC; ++m) {
                const int i = i0 + RP * m;
                const float v = (i < Hd) ? sZb[i] : 0.f;
                if (s == 3) V4[m] = v;
                else if (s == 2) V3[m] = v;
                else if (s == 1) V2[m] = v;
                else {
                    lam[m] = lam[m] + V4[m] + V3[m] + V2[m] + v + mu[m];
                    mu[m] = 0.f;
                }
            }
            __syncwarp();
        }
    }
    // ---- results
    for (int o = 16; o >= 1; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0 && gmax > 0.f && gmax < 3.0e38f)
        atomicMax(reinterpret_cast<unsigned*>(const_cast<float*>(p.wpack) + lo.off_gscale), __float_as_uint(gmax));
#pragma unroll
    for (int m = 0; m < MC; ++m) {
        const int i = i0 + RP * m;
        if (i < Hd && j == 0) p.grad_z0[(int64_t)bg * Hd + i] = lam[m] + __ldg(p.grad_z_out + (int64_t)bg * Hd + i);
    }
}

// ------------------------------------------------------------------------------------------------------------ launch
int lane_q(int maxdim) { return maxdim <= 32 ? 8 : (maxdim <= 40 ? 10 : 16); }
int lane_mc(int ncp)
{
    const int m = (ncp + 31) / 32;
    return m <= 1 ? 1 : (m <= 3 ? 3 : (m <= 5 ? 5 : 8));
}

size_t lane_fwd_smem(const NcdeLayout& lo, int wpc, int Q)
{
    size_t f = pad4(lo.hidden_floats) + lo.NCP + (size_t)lo.NCP * lo.KPS + pad4(lo.L) + 8 * ANCDE_MAX_LAYERS + (size_t)wpc * 3 * 4 * Q;
    return f * 4;
}
size_t lane_bwd_smem(const NcdeLayout& lo, int wpc, int Q)
{
    size_t f = pad4(lo.hiddenb_floats) + (size_t)lo.NCP * lo.KPS + 8 * ANCDE_MAX_LAYERS +
               (size_t)wpc * (LANE_NBUF * ((size_t)lo.act_fwd_floats + lo.NCP) + 3 * 4 * Q);
    return f * 4 + (size_t)wpc * LANE_NBUF * 8;
}

int lane_maxdim(const NcdeLayout& lo)
{
    int m = lo.Hd;
    for (int l = 0; l < lo.n_hidden; ++l) m = lo.width[l] > m ? lo.width[l] : m;
    return m;
}

// warps (= samples) per CTA: one while every sample can have an SM sub-partition to itself, more for larger batches
int lane_wpc(int B)
{
    int w = 1;
    while (w < 8 && (B + w - 1) / w > 2 * 148) w *= 2;
    return w;
}

typedef void (*LaneKernel)(SolveArgs);

template <int MC>
LaneKernel pick_q(bool bwd, int Q)
{
    if (Q == 8) return bwd ? ncde_lane_bwd_kernel<MC, 8> : ncde_lane_fwd_kernel<MC, 8>;
    if (Q == 10) return bwd ? ncde_lane_bwd_kernel<MC, 10> : ncde_lane_fwd_kernel<MC, 10>;
    return bwd ? ncde_lane_bwd_kernel<MC, 16> : ncde_lane_fwd_kernel<MC, 16>;
}
LaneKernel pick(bool bwd, int MC, int Q)
{
    switch (MC) {
        case 1: return pick_q<1>(bwd, Q);
        case 3: return pick_q<3>(bwd, Q);
        case 5: return pick_q<5>(bwd, Q);
        default: return pick_q<8>(bwd, Q);
    }
}

int launch_lane(SolveArgs& a, bool bwd, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    if (lo.TS != 1) {
        set_error("lane engine: needs one sample per tile (got %d)", lo.TS);
        return ANCDE_EINVAL;
    }
    if ((int64_t)lo.B * (lo.L - 1) * lo.C >= ((int64_t)1 << 31) - lo.C) {
        set_error("lane engine: B * (L-1) * C does not fit the 32-bit coefficient offsets of the control lookup");
        return ANCDE_EUNSUPPORTED;
    }
    const int Q = lane_q(lane_maxdim(lo)), MC = lane_mc(lo.NCP);
    const int wpc = lane_wpc(lo.B);
    const size_t smem = bwd ? lane_bwd_smem(lo, wpc, Q) : lane_fwd_smem(lo, wpc, Q);
    if (smem > 227 * 1024) {
        set_error("lane engine: needs %zu bytes of shared memory per CTA", smem);
        return ANCDE_EUNSUPPORTED;
    }
    LaneKernel kern = pick(bwd, MC, Q);
    ANCDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const char* fnames[3] = {"ncde_fwd_plain", "ncde_fwd_bottom", "ncde_fwd_top"};
    static const char* bnames[3] = {"ncde_bwd_plain", "ncde_bwd_bottom", "ncde_bwd_top"};
    prof_begin(bwd ? bnames[lo.mode] : fnames[lo.mode], stream);
    kern<<<(lo.B + wpc - 1) / wpc, 32 * wpc, smem, stream>>>(a);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

}  // namespace

bool lane_shape_ok(const ancde_ncde_desc* d)
{
    if (!d || d->single_stage) return false;
    if (d->B < 1 || d->L < 2 || d->C < 1 || d->Hd < 1 || d->n_hidden < 1 || d->n_hidden > ANCDE_MAX_LAYERS) return false;
    const int C4 = pad4(d->C);
    if (C4 != 4 && C4 != 8 && C4 != 16 && C4 != 32) return false;
    if (d->Hd > 64 || d->Hd * C4 > 256) return false;
    int64_t hidden = 0;
    for (int l = 0; l < d->n_hidden; ++l) {
        if (d->width[l] < 1 || d->width[l] > 64) return false;
        hidden += (int64_t)(d->width[l] + 4) * 72 * 2;
    }
    // generous bound on the shared memory of one CTA (the launchers check the exact figure)
    const int K = d->width[d->n_hidden - 1];
    const int64_t bytes = 4 * (hidden + (int64_t)d->Hd * C4 * (8 * ((K + 7) / 8) + 5) + d->L + 4096);
    return bytes <= 200 * 1024;
}

int launch_lane_fwd(SolveArgs a, cudaStream_t stream) { return launch_lane(a, false, stream); }
int launch_lane_bwd(SolveArgs a, cudaStream_t stream) { return launch_lane(a, true, stream); }

}  // namespace ancde
