// The lane engine (ANCDE_ENGINE_LANE): the latency regime of the fused NCDE solve -- small batches of small networks
// (UEA CharacterTrajectories: 32 samples, 724 dependent RK4 stages per network; Google Stock: 200 samples).
//
// The streamed engine gives a tile of samples a CTA built for throughput and pays, per stage, a hidden-layer chain on ONE
// warp whose every shared load is waited for, index arithmetic spread over a few dozen threads and two or three CTA
// barriers: 5-7 k cycles per stage for a network of ~10 k FMAs.  With 32 samples there is nothing to overlap that
// latency with.  What ncu showed on B200 (UEA shape; every state of this file was profiled): a warp that is alone on
// its scheduler issues one instruction every 4-7 cycles WHATEVER it does (fixed-latency dependencies, shared-memory
// round trips), so the time of a stage is the per-warp INSTRUCTION COUNT of its critical warp:
//   * one warp per sample, every dot product unrolled with predicates (first version): ~2000 instructions per stage,
//     no faster than the streamed kernel (UEA step 8.7 ms against 10.7);
//   * a CTA per sample with layer tables in shared memory and run-time loop bounds: ~800 per warp and stage (8.4 ms);
//   * the network STRUCTURE as compile-time constants (LaneStatic): 400-550 (5.2 ms);
//   * shared-memory base addresses pinned in registers (the compiler rematerialised them from kernel parameters at
//     every use: a third of the instructions) and units split over threads only above 4 pieces per thread: 4.65 ms.
//
//   ONE SMALL CTA PER SAMPLE, ONE THREAD PER OUTPUT-LAYER COLUMN (T = 32 * ceil(max(Hd * C4, widest layer) / 32) threads:
//   1..8 warps on the SM's four schedulers), every phase spread over the threads it can use:
//   * control (spline derivative, attention product): thread = column (i, j) evaluates ITS channel j from coefficients
//     loaded a stage ahead                                    [cdeint_module.py:58-60,103-129; interpolate.py:261-303]
//   * hidden ReLU layers: thread = unit, or KP threads per unit when a dot product has more than LANE_SPLIT 128-bit
//     pieces per thread (weight row and layer input in shared memory, log2(KP) shuffles)
//                                                             [vector_fields.py:205-216,237-247]
//   * output layer: thread = column: full dot product, tanh, times dU of its channel; the contraction over j is
//     log2(C4) shuffles inside groups of C4 lanes              [vector_fields.py:216-218, cdeint_module.py:61,133]
//   * the RK4 recurrence (3/8 rule) of row i in registers, replicated over the C4 threads of the row
//     [torchdiffeq rk4_alt_step_func]
//   One CTA barrier per layer and one per stage (a __syncwarp when the CTA is a single warp: UEA's attention network).
// The backward walks the stages in reverse with the same ownership: saved activations and tanh outputs of a stage arrive
// by two TMA bulk copies on an mbarrier three stages ahead; the output layer's vector-Jacobian product runs with
// KP threads per hidden unit over a padded copy of W_out^T [k][columns] (column-contiguous, conflict free) and
// the pre-activation cotangent broadcast from shared memory; hidden layers backwards as in the forward.
// Both kernels read / write the streamed engine's saved layout (NcdeLayout: tiles of 8 samples -- what the tcgen05 weight-gradient
// kernels take -- or of 1 sample on request; a CTA owns one slot of a tile, slots without a sample are zero-filled), so the
// weight-gradient kernels and the packed weights are shared.  For large batches (B >= 240: several CTAs per SM) idle warps skip the
// layer bodies and only warp 0 evaluates the control (template parameter SKIP).  C4 must divide 32 (C <= 32, C4 in {4, 8, 16, 32}), Hd * C4 <= 256, widths <= 64.
#include "ncde_lane.cuh"
#include "pipeline.cuh"

#include <cstdlib>

namespace ancde {

namespace {

#ifndef LANE_SPLIT
#define LANE_SPLIT 4
#endif
constexpr int LANE_NBUF = 4;       // backward: stages of saved activations in flight
constexpr int LANE_ML = ANCDE_MAX_LAYERS;

// What a thread does in one dense layer, worked out ONCE before the stage loop and kept in registers (the layer loops are
// unrolled at compile time): measured with ncu on the UEA shape, recomputing unit / part / row addresses / loop bounds from
// shared-memory layer tables every stage made 850 warp instructions per stage out of a network that needs 80 FMAs per
// thread -- and a warp that is alone on its scheduler issues one instruction every ~5 cycles.
struct LaneLayer {
    int wofs;      // float offset of this thread's first 128-bit weight piece in the layer's weight array
    int xofs;      // float offset of its first piece of the layer input (4 * part)
    int step;      // floats between its pieces (4 * threads per unit)
    int np;        // its pieces (threads beyond the last unit recompute the last unit and do not write)
    int npmin, npmax;   // bounds of np over the threads of a unit (compile-time constants for a static shape)
    int out;       // unit index when this thread writes the unit's result (part 0), else -1
    bool warp_on;  // this thread's WARP holds at least one unit of the layer (warp-uniform): the others only meet at the barrier --
                   // with several CTAs per SM (large batches) idle warps' instructions are not free
    int aofs;      // where that result goes in the saved-activation block
    float bias;
};

__host__ __device__ constexpr int lane_pad4(int x) { return (x + 3) & ~3; }
__host__ __device__ constexpr int lane_pad4odd(int x) { return 4 * (((x + 3) >> 2) | 1); }
__host__ __device__ constexpr int lane_max(int a, int b) { return a > b ? a : b; }
// threads of a CTA: one per output-layer column and at least one per unit of the widest layer -- a single warp (and
// __syncwarp instead of CTA barriers) for the smallest networks
__host__ __device__ constexpr int lane_threads_of(int NCP, int Hmax) { return 32 * ((lane_max(NCP, Hmax) + 31) / 32); }

// The STRUCTURE of a network (layer count and widths, state size, padded channel count) as compile-time constants: every
// loop bound, pieces-per-thread count and shuffle depth of the kernels folds, the layer loops unroll to exactly NL layers
// and the dot products to straight-line code.  Instantiated for the networks of the BASELINE configurations the engine
// serves (UEA, Google Stock, Mujoco); any other supported shape runs the same code with run-time bounds (LaneDyn).
// MB: CTAs per SM the large-batch kernels are compiled for (a register cap; 1 = none)
template <int HD, int CC4, int NL_, int W0, int W1, int W2, int W3, int MB = 1>
struct LaneStatic {
    static constexpr bool kStatic = true;
    static constexpr int kMinBlocksBig = MB;
    static constexpr int Hd = HD, C4 = CC4, NL = NL_, NCP = HD * CC4;
    __host__ __device__ static constexpr int width(int l) { return l == 0 ? W0 : (l == 1 ? W1 : (l == 2 ? W2 : W3)); }
    static constexpr int K = width(NL_ - 1);
    static constexpr int Hmax = lane_max(lane_max(HD, W0), lane_max(NL_ > 1 ? W1 : 0, lane_max(NL_ > 2 ? W2 : 0, NL_ > 3 ? W3 : 0)));
    static constexpr int T = lane_threads_of(NCP, Hmax);
};
struct LaneDyn {
    static constexpr bool kStatic = false;
    static constexpr int kMinBlocksBig = 1;
    static constexpr int Hd = 0, C4 = 0, NL = LANE_ML, NCP = 0, K = 0, Hmax = 0, T = 0;
    __host__ __device__ static constexpr int width(int) { return 0; }
};

// sum over this thread's pieces of w . x (128-bit pieces, four accumulators; chunks of four pieces with every load of a
// chunk issued before its first FMA).  npmin <= np <= npmax: with compile-time bounds the loop is straight-line code and
// only the pieces between npmin and npmax keep a predicate.
__device__ __forceinline__ float lane_dot(const float* __restrict__ w, const float* __restrict__ x, int step, int np, int npmin, int npmax)
{
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int q0 = 0; q0 < npmax; q0 += 4) {
        float4 ww[4], xx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (q0 + u < npmin || q0 + u < np) {
                ww[u] = *reinterpret_cast<const float4*>(w + (q0 + u) * step);
                xx[u] = *reinterpret_cast<const float4*>(x + (q0 + u) * step);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (q0 + u < npmin || q0 + u < np) {
                a0 = fmaf(ww[u].x, xx[u].x, a0);
                a1 = fmaf(ww[u].y, xx[u].y, a1);
                a2 = fmaf(ww[u].z, xx[u].z, a2);
                a3 = fmaf(ww[u].w, xx[u].w, a3);
            }
        }
    }
    return (a0 + a1) + (a2 + a3);
}

// log2 of the threads that share one unit of a layer of N units whose dot products have nk4 128-bit pieces, in a CTA of T
// threads.  A unit is split (one shuffle step per doubling) while a thread would have more than LANE_SPLIT pieces: measured on
// B200, UEA shape (10 pieces per unit in the forward, 40 in the backward's output layer): thresholds 12 / 4 / 2 give
// 5.05 / 4.67 / 4.65 ms per training step.
__host__ __device__ constexpr int lane_kshift(int T, int N, int nk4)
{
    int s = 0;
    while (s < 3 && (N << (s + 1)) <= T && (nk4 >> s) > LANE_SPLIT) ++s;
    return s;
}

// this thread's share of a layer of N units whose dot products have nk4 pieces; rows of `stride` floats
__device__ __forceinline__ LaneLayer lane_layer(int tid, int T, int N, int nk4, int stride)
{
    LaneLayer L;
    const int ksh = lane_kshift(T, N, nk4), KP = 1 << ksh;
    const int o = tid >> ksh, part = tid & (KP - 1);
    const int oc = o < N ? o : N - 1;
    L.wofs = oc * stride + 4 * part;
    L.xofs = 4 * part;
    L.step = 4 * KP;
    L.npmin = nk4 >> ksh;
    L.npmax = (nk4 + KP - 1) >> ksh;
    L.np = part < nk4 ? (nk4 - part + KP - 1) >> ksh : 0;
    L.out = (o < N && part == 0) ? o : -1;
    L.warp_on = (tid & ~31) < (N << ksh);
    L.aofs = 0;
    L.bias = 0.f;
    return L;
}

// row stride (floats) of the backward's copy of W_out^T [k][columns]: a quarter warp = 8 / KP units x KP column pieces hits
// 8 distinct 16-byte bank groups when the stride is 4 * KP modulo 32
__host__ __device__ inline int lane_wt_stride(int NCP, int kshift)
{
    const int want = (4 << kshift) & 31;
    return NCP + ((want - NCP % 32) + 32) % 32;
}

// A shared-memory pointer the compiler must keep in a register: measured with ncu, a third of the instructions of a stage
// were smem base addresses REMATERIALISED from kernel parameters at every use (the compiler considers that free).
template <typename P>
__device__ __forceinline__ P* lane_pin(P* p)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return reinterpret_cast<P*>(__cvta_shared_to_generic(a));
}

// CTA-wide synchronisation: a __syncwarp when the CTA is one warp (T folds to a constant for the static shapes)
__device__ __forceinline__ void lane_sync(int T)
{
    if (T == 32) __syncwarp();
    else __syncthreads();
}

struct LaneCtl { float b, c2, d3, a, hp, frac; };

// Global loads of the control of the stage at time ts (call counter m) for this thread's channel; consumed a stage later.
// tknot = times[cnt] (+inf past the last knot) is carried in a register: the loop compares against it and only reloads when the
// count advances -- a shared-memory round trip per comparison sat on every warp's critical path otherwise.
__device__ __forceinline__ void lane_ctl_issue(const SolveArgs& p, const float* sTimes, int& cnt, float& tknot, float ts, int m, bool jvalid,
                                               int obase, int abase, int hbase, LaneCtl& r)
{
    const NcdeLayout& lo = p.lo;
    // index = clamp(#(times < t) - 1, 0, L-2): the LEFT piece at a knot (interpolate.py:261-267)
    while (tknot < ts) {
        ++cnt;
        tknot = (cnt < lo.L) ? sTimes[cnt] : __int_as_float(0x7f800000);
    }
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
            r.a = __ldg(p.attention + (int64_t)q * lo.B * lo.att_C + abase);
            r.hp = __ldg(p.h_prime + (int64_t)m * lo.B * lo.C + hbase);      // m: call counter + reset rule (cdeint_module.py:134-137)
        }
    }
}

// dU (and dU/da) of this thread's channel from the loaded coefficients (ctl_finish of the streamed engine)
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
// SAVE: training (save_act / save_G written) or inference -- a template parameter so that no store carries a null check.
// SKIP: warps that hold no unit of a layer skip its body (large batches: several CTAs per SM share the issue slots -- Mujoco at B = 1024
// 1.15 -> 1.07 ms per step); in the latency regime the extra branches cost more than they save (UEA 4.56 -> 4.79 ms), so the
// tiles-of-one-sample kernels are compiled without them.
template <class S, bool SAVE, bool SKIP>
__global__ void __launch_bounds__(256, SKIP ? S::kMinBlocksBig : 1) ncde_lane_fwd_kernel(const __grid_constant__ SolveArgs p)
{
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, T = S::kStatic ? S::T : (int)blockDim.x;
    const int bg = blockIdx.x;                              // this CTA's sample (slot bg % TS of tile bg / TS of the saved layout)
    const int NL = S::kStatic ? S::NL : lo.n_hidden;
    const int Hd = S::kStatic ? S::Hd : lo.Hd, C4 = S::kStatic ? S::C4 : lo.C4, NCP = S::kStatic ? S::NCP : lo.NCP;
    const int K = S::kStatic ? S::K : lo.K;
    const int KPS = 8 * ((K + 7) / 8) + 4;                  // = lo.KPS
    const int sh = 31 - __clz(C4);                          // log2(C4)
    const int col = tid, i = tid >> sh, j = tid & (C4 - 1); // this thread's column (i, j) of the field
    const bool cv = col < NCP;                              // <=> i < Hd
    const int HP = pad4(S::kStatic ? S::Hmax : lo.Hmax);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sp = reinterpret_cast<float*>(smem_raw);
    float* sHW = sp;   sp += pad4(lo.hidden_floats);        // hidden weights + biases (packed rows W[o][stride])
    float* sBo = sp;   sp += NCP;
    float* sWot = sp;  sp += (size_t)NCP * KPS;             // output layer W_out[col][KPS] (k contiguous)
    float* sTimes = sp; sp += pad4(lo.L);
    float* sIn = sp;   sp += HP;                            // stage input
    float* sT0 = sp;   sp += HP;                            // activations ping / pong
    float* sT1 = sp;   sp += HP;
    float* sDU = sp;                                        // SKIP: the control increment of the stage, from warp 0 (32 floats)

    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack);
        for (int q = tid; q < pad4(lo.hidden_floats) / 4; q += T) reinterpret_cast<float4*>(sHW)[q] = __ldg(src + q);
        const float4* sb = reinterpret_cast<const float4*>(p.wpack + lo.off_bo);
        for (int q = tid; q < NCP / 4; q += T) reinterpret_cast<float4*>(sBo)[q] = __ldg(sb + q);
        const float4* sw = reinterpret_cast<const float4*>(p.wpack + lo.off_wot);
        for (int q = tid; q < NCP * KPS / 4; q += T) reinterpret_cast<float4*>(sWot)[q] = __ldg(sw + q);
        for (int q = tid; q < lo.L; q += T) sTimes[q] = p.times[q];
        for (int q = tid; q < 3 * HP; q += T) sIn[q] = 0.f;
    }
    __syncthreads();

    // the saved layout: tiles of TS samples (1, or 8 for large batches: the tcgen05 weight-gradient kernels' layout), every
    // sub-block [feature][TS]; this CTA's sample is slot sb of tile `tile`
    const int TS = lo.TS, tile = bg / TS, sb = bg - tile * TS;
    if (bg >= lo.B) {
        // a slot of the last tile without a sample: the weight-gradient kernels read whole tiles, so it holds zeros
        if (SAVE) {
            for (int n = 0; n < lo.n_stages; ++n) {
                float* a = p.save_act + ((size_t)n * lo.ntiles + tile) * lo.act_floats;
                for (int q = tid; q < Hd; q += T) a[q * TS + sb] = 0.f;
                for (int l = 0; l < lo.n_hidden; ++l)
                    for (int q = tid; q < lo.width[l]; q += T) a[lo.act_off_h[l] + q * TS + sb] = 0.f;
                for (int q = tid; q < C4; q += T) {
                    a[lo.act_off_du + (q >> 2) * lo.DUS + (q & 3) * TS + sb] = 0.f;
                    if (lo.mode == ANCDE_MODE_TOP) a[lo.act_off_duda + (q >> 2) * lo.DUS + (q & 3) * TS + sb] = 0.f;
                }
                float* g = p.save_G + (((size_t)n * lo.ntiles + tile) * TS + sb) * NCP;
                for (int q = tid; q < NCP; q += T) g[q] = 0.f;
            }
        }
        return;
    }

    // this thread's share of every hidden layer
    LaneLayer ly[LANE_ML];
#pragma unroll
    for (int l = 0; l < S::NL; ++l) {
        ly[l] = LaneLayer();
        if (l < NL) {
            const int N = S::kStatic ? S::width(l) : lo.width[l];
            const int IN = S::kStatic ? (l == 0 ? S::Hd : S::width(l > 0 ? l - 1 : 0)) : lo.in_dim[l];
            ly[l] = lane_layer(tid, T, N, (IN + 3) >> 2, S::kStatic ? lane_pad4odd(IN) : lo.stride[l]);
            ly[l].wofs += lo.off_w[l];
            if (ly[l].out >= 0) {
                ly[l].bias = sHW[lo.off_b[l] + ly[l].out];
                ly[l].aofs = lo.act_off_h[l] + ly[l].out * TS + sb;
            }
        }
    }

    const float* pw[LANE_ML];
    const float* px[LANE_ML];
    float* po[LANE_ML];
#pragma unroll
    for (int l = 0; l < S::NL; ++l) {
        pw[l] = px[l] = nullptr; po[l] = nullptr;
        if (l < NL) {
            pw[l] = lane_pin(sHW + ly[l].wofs);
            px[l] = lane_pin(((l == 0) ? sIn : ((l & 1) ? sT0 : sT1)) + ly[l].xofs);
            po[l] = lane_pin(((l & 1) ? sT1 : sT0) + (ly[l].out >= 0 ? ly[l].out : 0));
        }
    }

    const bool jvalid = j < lo.C;
    const bool wr = cv && j == 0;                            // the thread that writes row i
    const int obase = bg * (lo.L - 1) * lo.C + j;
    const int abase = bg * lo.att_C + (lo.att_C == 1 ? 0 : j);
    const int hbase = bg * lo.C + j;
    const int duidx = (j >> 2) * lo.DUS + (j & 3) * TS + sb;   // du_index(j, b, TS, DUS)
    const int nk4o = (K + 3) >> 2;
    const bool top = lo.mode == ANCDE_MODE_TOP;
    const float* lastbuf = lane_pin(((NL - 1) & 1) ? sT1 : sT0);
    const float* wo_row = lane_pin(sWot + (size_t)(cv ? col : 0) * KPS);
    float* const sIn_i = lane_pin(sIn + (cv ? i : 0));
    const float* const sTimesP = lane_pin(sTimes);
    const float bo = sBo[cv ? col : 0];
    // per-stage strides of the output pointers
    const size_t act_stride = (size_t)lo.ntiles * lo.act_floats, g_stride = (size_t)lo.ntiles * TS * NCP, k_stride = (size_t)lo.B * Hd;
    // running pointers (never dereferenced when the buffer is absent)
    float* act = p.save_act + (size_t)tile * lo.act_floats;
    float* gsave = p.save_G + ((size_t)tile * TS + sb) * NCP + (cv ? col : 0);
    float* kout = p.k_out + (size_t)bg * Hd + (cv ? i : 0);
    const bool have_kout = p.k_out != nullptr;

    float y = cv ? p.z0[(int64_t)bg * Hd + i] : 0.f;
    float k1 = 0.f, k2 = 0.f, k3 = 0.f;
    if (wr) {
        sIn[i] = y;
        p.z_out[(int64_t)bg * Hd + i] = y;                   // solution[0] = y0
        if (SAVE) act[i * TS + sb] = y;
    }
    int cnt = 0;        // knots strictly below the current stage time (monotone)
    float tknot = sTimesP[0];
    int jout = 1;       // next requested output time
    float tout = (lo.n_out > 1) ? __ldg(p.t_out + 1) : 0.f;
    int mcall = 0;      // call counter of the stage whose control is being loaded, modulo n_hp - 1
    LaneCtl ctl;
    float t0 = grid_time(0, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
    float t1 = grid_time(1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
    lane_ctl_issue(p, sTimesP, cnt, tknot, t0, 0, jvalid, obase, abase, hbase, ctl);
    lane_sync(T);

    for (int k = 0; k < lo.n_steps; ++k) {
        const float dt = __fsub_rn(t1, t0);
        const float t2 = (k + 1 < lo.n_steps) ? grid_time(k + 2, lo.n_steps, p.t_start, p.t_end, p.step, p.grid) : t1;
        const float ts1 = stage_time(t0, dt, 1), ts2 = stage_time(t0, dt, 2), ts3 = stage_time(t0, dt, 3);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const bool have_next = 4 * k + s + 1 < lo.n_stages;
            float* act_next = act + act_stride;
            // (1) control of this stage from the registers loaded a stage ago; the loads of the next stage go out now
            // (SKIP, large batches: only warp 0 evaluates it -- the C4 channels are all it takes -- and hands dU over through
            // shared memory; the hidden layers' barriers lie between its store and the other warps' load)
            float dU = 0.f, duda = 0.f;
            if (!SKIP || tid < 32) {
                lane_ctl_finish(lo.mode, ctl, dU, duda);
                if (have_next) {
                    if (top && ++mcall == lo.n_hp - 1) mcall = 0;
                    const float tsn = (s == 0) ? ts1 : (s == 1 ? ts2 : (s == 2 ? ts3 : t1));      // stage 0 of the next step is at its t0 = t1
                    lane_ctl_issue(p, sTimesP, cnt, tknot, tsn, mcall, jvalid, obase, abase, hbase, ctl);
                }
                if (tid < C4) {
                    if (SKIP) sDU[tid] = dU;
                    if (SAVE) {
                        act[lo.act_off_du + duidx] = dU;
                        if (top) act[lo.act_off_duda + duidx] = duda;
                    }
                }
            }

            // (2) hidden layers: KP threads per unit, one CTA barrier per layer
#pragma unroll
            for (int l = 0; l < S::NL; ++l) {
                if (l < NL) {
                    if (!SKIP || ly[l].warp_on) {
                        float acc = lane_dot(pw[l], px[l], ly[l].step, ly[l].np, ly[l].npmin, ly[l].npmax);
                        const int kp = S::kStatic ? (1 << lane_kshift(S::T, S::width(l), ((l == 0 ? S::Hd : S::width(l > 0 ? l - 1 : 0)) + 3) >> 2))
                                                  : (ly[l].step >> 2);
                        for (int m = 1; m < kp; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
                        if (ly[l].out >= 0) {
                            const float v = fmaxf(acc + ly[l].bias, 0.f);
                            *po[l] = v;
                            if (SAVE) act[ly[l].aofs] = v;
                        }
                    }
                    lane_sync(T);
                }
            }

            // (3) output layer, tanh, contraction with dU over the C4 threads of the row
            float ko;
            {
                float g = 0.f;
                if (cv) {
                    g = tanh_mufu(lane_dot(wo_row, lastbuf, 4, nk4o, nk4o, nk4o) + bo);
                    if (SAVE) *gsave = g;
                }
                if (SKIP) dU = sDU[j];
                ko = g * dU;
                for (int m = 1; m < C4; m <<= 1) ko += __shfl_xor_sync(0xffffffffu, ko, m);
            }

            // (4) the RK4 recurrence of row i (rk4_alt_step_func)
            {
                if (wr && have_kout) *kout = ko;
                float zin;
                if (s == 0) {
                    k1 = ko;
                    zin = __fadd_rn(y, __fdiv_rn(__fmul_rn(dt, ko), 3.0f));                       // y + dt*k1/3
                } else if (s == 1) {
                    k2 = ko;
                    zin = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(ko, __fdiv_rn(k1, 3.0f))));        // y + dt*(k2 - k1/3)
                } else if (s == 2) {
                    k3 = ko;
                    zin = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1, k2), ko)));          // y + dt*(k1-k2+k3)
                } else {
                    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(k1, __fmul_rn(3.0f, k2)), __fmul_rn(3.0f, k3)), ko);
                    const float y1 = __fadd_rn(y, __fmul_rn(sum, __fdiv_rn(dt, 8.0f)));
                    // emit every requested time in (t0, t1] (FixedGridODESolver.integrate + _linear_interp); the next requested
                    // time is loaded one emission ahead
                    while (jout < lo.n_out && t1 >= tout) {
                        if (wr) {
                            float val;
                            if (tout == t0) val = y;
                            else if (tout == t1) val = y1;
                            else val = __fadd_rn(y, __fmul_rn(__fdiv_rn(__fsub_rn(y1, y), __fsub_rn(t1, t0)), __fsub_rn(tout, t0)));
                            p.z_out[((int64_t)jout * lo.B + bg) * Hd + i] = val;
                        }
                        ++jout;
                        if (jout < lo.n_out) tout = __ldg(p.t_out + jout);
                    }
                    y = y1;
                    zin = y1;
                }
                if (wr) {
                    *sIn_i = zin;
                    if (SAVE && have_next) act_next[i * TS + sb] = zin;
                }
            }
            act = act_next;
            gsave += g_stride;
            kout += k_stride;
            lane_sync(T);
        }
        t0 = t1;
        t1 = t2;
    }
}

// --------------------------------------------------------------------------------------------------------- backward
template <class S, bool SKIP>
__global__ void __launch_bounds__(256, SKIP ? S::kMinBlocksBig : 1) ncde_lane_bwd_kernel(const __grid_constant__ SolveArgs p)
{
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, T = S::kStatic ? S::T : (int)blockDim.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const int bg = blockIdx.x;
    const int NL = S::kStatic ? S::NL : lo.n_hidden;
    const int Hd = S::kStatic ? S::Hd : lo.Hd, C4 = S::kStatic ? S::C4 : lo.C4, NCP = S::kStatic ? S::NCP : lo.NCP;
    const int K = S::kStatic ? S::K : lo.K;
    const int sh = 31 - __clz(C4);
    const int col = tid, i = tid >> sh, j = tid & (C4 - 1);
    const bool cv = col < NCP;
    const int HP = pad4(S::kStatic ? S::Hmax : lo.Hmax);
    const int AF = lo.act_fwd_floats;
    const bool need_att = (lo.mode == ANCDE_MODE_TOP) && (p.grad_attention != nullptr);
    const int kso = lane_kshift(T, K, NCP >> 2);             // output layer backwards: 2^kso threads per hidden unit
    const int SW = lane_wt_stride(NCP, kso);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // [LANE_NBUF]
    float* sp = reinterpret_cast<float*>(smem_raw + LANE_NBUF * 8);
    float* sWtb = sp;  sp += pad4(lo.hiddenb_floats);        // transposed hidden weights Wt_l[k][tstride]
    float* sWt = sp;   sp += (size_t)K * SW;                 // W_out^T [k][SW] (columns contiguous)
    float* sActBuf = sp; sp += (size_t)LANE_NBUF * AF;       // [LANE_NBUF][AF]
    float* sGBuf = sp; sp += (size_t)LANE_NBUF * NCP;        // [LANE_NBUF][NCP]
    float* sPre = sp;  sp += NCP;                            // cotangent of the output layer's pre-activation
    float* sD0 = sp;   sp += HP;
    float* sD1 = sp;   sp += HP;
    float* sZb = sp;   sp += HP;
    float* sAtt = sp;                                        // [W][C4] partial sums of the attention cotangent

    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack + lo.off_hiddenb);
        for (int q = tid; q < pad4(lo.hiddenb_floats) / 4; q += T) reinterpret_cast<float4*>(sWtb)[q] = __ldg(src + q);
        const int NQ = NCP / 4;
        const float4* sw = reinterpret_cast<const float4*>(p.wpack + lo.off_wo);      // Wt_out[k][NCP]
        for (int q = tid; q < K * NQ; q += T) {
            const int k = q / NQ, c = q - k * NQ;
            *reinterpret_cast<float4*>(sWt + (size_t)k * SW + 4 * c) = __ldg(sw + q);
        }
        for (int q = tid; q < NCP + 3 * HP; q += T) sPre[q] = 0.f;
        if (tid == 0) {
            for (int s = 0; s < LANE_NBUF; ++s) mbar_init(bars + s, 1);
            mbar_fence_init();
        }
    }
    __syncthreads();

    const int TS = lo.TS, tile = bg / TS, sb = bg - tile * TS;      // the saved layout, as in the forward
    if (bg >= lo.B) {
        // a slot of the last tile without a sample: zero cotangents for the weight-gradient kernels
        for (int n = 0; n < lo.n_stages; ++n) {
            float* a = p.save_act + ((size_t)n * lo.ntiles + tile) * lo.act_floats;
            for (int q = tid; q < Hd; q += T) a[lo.act_off_gbar + q * TS + sb] = 0.f;
            for (int l = 0; l < lo.n_hidden; ++l)
                for (int q = tid; q < lo.width[l]; q += T) a[lo.act_off_delta[l] + q * TS + sb] = 0.f;
        }
        return;
    }

    // this thread's share of the output layer (backwards: unit k of the last hidden layer <- all columns) and of every hidden layer
    // (backwards: input unit kin of layer l <- the layer's outputs); layer l reads its cotangent from D[(l+1) & 1], writes D[l & 1]
    LaneLayer lyo = lane_layer(tid, T, K, NCP >> 2, SW);
    float* const sDlast = (NL & 1) ? sD1 : sD0;
    int lyo_mask = 0;
    if (lyo.out >= 0) {
        lyo_mask = lo.act_off_h[NL - 1] + lyo.out * TS + sb;
        lyo.aofs = lo.act_off_delta[NL - 1] + lyo.out * TS + sb;
    }
    LaneLayer ly[LANE_ML];
    int lmask[LANE_ML];
#pragma unroll
    for (int l = 0; l < S::NL; ++l) {
        ly[l] = LaneLayer();
        lmask[l] = 0;
        if (l < NL) {
            const int N = S::kStatic ? S::width(l) : lo.width[l];
            const int IN = S::kStatic ? (l == 0 ? S::Hd : S::width(l > 0 ? l - 1 : 0)) : lo.in_dim[l];
            ly[l] = lane_layer(tid, T, IN, (N + 3) >> 2, S::kStatic ? lane_pad4odd(N) : lo.tstride[l]);
            ly[l].wofs += lo.off_wt[l] - lo.off_hiddenb;
            if (ly[l].out >= 0 && l > 0) {
                lmask[l] = lo.act_off_h[l - 1] + ly[l].out * TS + sb;
                ly[l].aofs = lo.act_off_delta[l - 1] + ly[l].out * TS + sb;
            }
        }
    }

    const float* pw[LANE_ML];
    const float* px[LANE_ML];
    float* po[LANE_ML];
#pragma unroll
    for (int l = 0; l < S::NL; ++l) {
        pw[l] = px[l] = nullptr; po[l] = nullptr;
        if (l < NL) {
            pw[l] = lane_pin(sWtb + ly[l].wofs);
            px[l] = lane_pin((((l + 1) & 1) ? sD1 : sD0) + ly[l].xofs);
            po[l] = lane_pin((l == 0 ? sZb : ((l & 1) ? sD1 : sD0)) + (ly[l].out >= 0 ? ly[l].out : 0));
        }
    }
    const float* const pwo = lane_pin(sWt + lyo.wofs);
    const float* const pxo = lane_pin(sPre + lyo.xofs);
    float* const poo = lane_pin(sDlast + (lyo.out >= 0 ? lyo.out : 0));
    float* const sPre_c = lane_pin(sPre + (cv ? col : 0));
    const float* const sZb_i = lane_pin(sZb + (cv ? i : 0));
    const float* const sActP = lane_pin(sActBuf);
    const float* const sGP = lane_pin(sGBuf + (cv ? col : 0));

    const bool wr = cv && j == 0;
    const int duidx = (j >> 2) * lo.DUS + (j & 3) * TS + sb;
    const int NS = lo.n_stages;
    const size_t act_stride = (size_t)lo.ntiles * lo.act_floats;
    // saved activations (forward part of the block) and tanh outputs of reverse stage r = NS - 1 - n -> ring slot r % NBUF
    auto issue = [&](int r) {
        if (tid == 0 && r < NS) {
            const int n = NS - 1 - r, slot = r % LANE_NBUF;
            mbar_arrive_expect_tx(bars + slot, (uint32_t)((AF + NCP) * 4));
            bulk_g2s(sActBuf + (size_t)slot * AF, p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats, (uint32_t)(AF * 4), bars + slot);
            bulk_g2s(sGBuf + (size_t)slot * NCP, p.save_G + (((int64_t)n * lo.ntiles + tile) * TS + sb) * NCP, (uint32_t)(NCP * 4), bars + slot);
        }
    };
    for (int r = 0; r < LANE_NBUF - 1; ++r) issue(r);

    float lam = 0.f, mu = 0.f, V4 = 0.f, V3 = 0.f, V2 = 0.f;
    // the next output time whose cotangent has not been consumed yet, and that cotangent, loaded a step ahead
    int jb = lo.n_out - 1;
    float tjb = (jb >= 1) ? __ldg(p.t_out + jb) : 0.f;
    float gzb = (jb >= 1 && cv) ? __ldg(p.grad_z_out + ((int64_t)jb * lo.B + bg) * Hd + i) : 0.f;
    float gmax = 0.f;
    int r = 0;
    float* act = p.save_act + ((int64_t)(NS - 1) * lo.ntiles + tile) * lo.act_floats;

    for (int k = lo.n_steps - 1; k >= 0; --k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float dt = __fsub_rn(t1, t0);
        // ---- cotangents of the outputs emitted during this step: t_out[jb] in (t0, t1]
        while (jb >= 1 && tjb > t0) {
            if (tjb == t1) lam += gzb;
            else {
                const float th = (tjb - t0) / (t1 - t0);
                lam += th * gzb;
                mu += (1.0f - th) * gzb;
            }
            --jb;
            if (jb >= 1) {
                tjb = __ldg(p.t_out + jb);
                gzb = cv ? __ldg(p.grad_z_out + ((int64_t)jb * lo.B + bg) * Hd + i) : 0.f;
            }
        }

#pragma unroll 1
        for (int s = 3; s >= 0; --s, ++r) {
            const int slot = r & (LANE_NBUF - 1);
            issue(r + LANE_NBUF - 1);                        // its slot was read last in stage r - 1, behind a barrier
            // ---- cotangent of the stage derivative k_{s+1} (transpose of rk4_alt_step_func)
            float gv;
            if (s == 3) gv = dt * 0.125f * lam;
            else if (s == 2) gv = dt * (0.375f * lam + V4);
            else if (s == 1) gv = dt * (0.375f * lam - V4 + V3);
            else gv = dt * (0.125f * lam + V4 + (V2 - V3) * (1.0f / 3.0f));
            gv = cv ? gv : 0.f;
            if (wr) act[lo.act_off_gbar + i * TS + sb] = gv;    // for the weight-gradient kernels
            gmax = fmaxf(gmax, fabsf(gv));
            mbar_wait(bars + slot, (uint32_t)((r >> 2) & 1));
            static_assert(LANE_NBUF == 4, "slot / parity arithmetic");
            const float* sAct = sActP + slot * AF;

            // ---- pre_bar = gbar * dU * (1 - G^2) of this thread's column; cotangent of dU for the attention
            {
                const float g = cv ? sGP[slot * NCP] : 0.f;
                const float dU = sAct[lo.act_off_du + duidx];
                if (cv) *sPre_c = gv * dU * fmaf(-g, g, 1.0f);
                if (need_att) {
                    // dUbar[j] = sum_i gbar[i] G[i][j]: the rows of this warp by shuffles, the warps through shared memory
                    float a = gv * g * sAct[lo.act_off_duda + duidx];
                    for (int m = C4; m < 32; m <<= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
                    if (lane < C4) sAtt[warp * C4 + lane] = a;
                }
            }
            lane_sync(T);
            if (need_att && tid < C4) {
                float a = 0.f;
                for (int w = 0; w < W; ++w) a += sAtt[w * C4 + tid];
                const float ts = stage_time(t0, dt, s);
                int qa = (int)floorf(ts) - 1;
                if (qa < 0) qa += lo.att_L;
                qa = qa < 0 ? 0 : (qa >= lo.att_L ? lo.att_L - 1 : qa);
                if (lo.att_C == 1) {
                    for (int m = 1; m < C4; m <<= 1) a += __shfl_xor_sync(C4 == 32 ? 0xffffffffu : ((1u << C4) - 1u), a, m);
                    if (tid == 0) atomicAdd(p.grad_attention + (int64_t)qa * lo.B + bg, a);
                } else if (tid < lo.C) {
                    atomicAdd(p.grad_attention + ((int64_t)qa * lo.B + bg) * lo.C + tid, a);
                }
            }

            // ---- hbar[k] = sum_col W_out[col][k] pre_bar[col]: 2^kso threads per unit, then delta = hbar * relu'(h)
            {
                if (!SKIP || lyo.warp_on) {
                    float acc = lane_dot(pwo, pxo, lyo.step, lyo.np, lyo.npmin, lyo.npmax);
                    for (int m = 1; m < (1 << kso); m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
                    if (lyo.out >= 0) {
                        const float v = (sAct[lyo_mask] > 0.f) ? acc : 0.f;
                        *poo = v;
                        act[lyo.aofs] = v;
                    }
                }
            }
            lane_sync(T);

            // ---- hidden layers backwards: delta_{l-1} = (W_l^T delta_l) * relu'(h_{l-1}); the stage-input cotangent is W_0^T delta_0
#pragma unroll
            for (int l = S::NL - 1; l >= 0; --l) {
                if (l < NL) {
                    if (!SKIP || ly[l].warp_on) {
                        float acc = lane_dot(pw[l], px[l], ly[l].step, ly[l].np, ly[l].npmin, ly[l].npmax);
                        const int kp = S::kStatic ? (1 << lane_kshift(S::T, l == 0 ? S::Hd : S::width(l > 0 ? l - 1 : 0), (S::width(l) + 3) >> 2))
                                                  : (ly[l].step >> 2);
                        for (int m = 1; m < kp; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
                        if (ly[l].out >= 0) {
                            if (l > 0) {
                                const float v = (sAct[lmask[l]] > 0.f) ? acc : 0.f;
                                *po[l] = v;
                                act[ly[l].aofs] = v;              // for the hidden weight-gradient kernel
                            } else {
                                *po[l] = acc;                     // cotangent of the stage input (sZb)
                            }
                        }
                    }
                    lane_sync(T);
                }
            }

            // ---- RK4 adjoint recurrence of row i
            {
                const float v = cv ? *sZb_i : 0.f;
                if (s == 3) V4 = v;
                else if (s == 2) V3 = v;
                else if (s == 1) V2 = v;
                else {
                    lam = lam + V4 + V3 + V2 + v + mu;
                    mu = 0.f;
                }
            }
            act -= act_stride;
        }
    }
    // ---- results
    for (int o = 16; o >= 1; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0 && gmax > 0.f && gmax < 3.0e38f)
        atomicMax(reinterpret_cast<unsigned*>(const_cast<float*>(p.wpack) + lo.off_gscale), __float_as_uint(gmax));
    if (wr) p.grad_z0[(int64_t)bg * Hd + i] = lam + __ldg(p.grad_z_out + (int64_t)bg * Hd + i);
}

// ------------------------------------------------------------------------------------------------------------ launch
int lane_threads(const NcdeLayout& lo) { return lane_threads_of(lo.NCP, lo.Hmax); }

size_t lane_fwd_smem(const NcdeLayout& lo)
{
    size_t f = pad4(lo.hidden_floats) + lo.NCP + (size_t)lo.NCP * lo.KPS + pad4(lo.L) + 3 * (size_t)pad4(lo.Hmax) + 32;
    return f * 4;
}
size_t lane_bwd_smem(const NcdeLayout& lo, int T)
{
    const int SW = lane_wt_stride(lo.NCP, lane_kshift(T, lo.K, lo.NCP >> 2));
    size_t f = pad4(lo.hiddenb_floats) + (size_t)lo.K * SW + LANE_NBUF * ((size_t)lo.act_fwd_floats + lo.NCP) +
               lo.NCP + 3 * (size_t)pad4(lo.Hmax) + (size_t)(T / 32) * lo.C4;
    return f * 4 + LANE_NBUF * 8;
}

typedef void (*LaneKernel)(SolveArgs);

template <class S>
bool lane_matches(const NcdeLayout& lo)
{
    if (lo.Hd != S::Hd || lo.C4 != S::C4 || lo.n_hidden != S::NL) return false;
    for (int l = 0; l < S::NL; ++l)
        if (lo.width[l] != S::width(l)) return false;
    return lane_threads(lo) == S::T;
}

// the networks of the BASELINE configurations (experiments/models of the reference: FinalTanh g / FinalTanh_ff6 f)
typedef LaneStatic<40, 4, 3, 40, 40, 40, 0> LaneUeaTop;        // UEA: h = hh = 40, 3 layers, 3 channels + time
typedef LaneStatic<4, 4, 4, 10, 20, 20, 20> LaneUeaBottom;     //      attention 20 / 10
typedef LaneStatic<12, 8, 2, 12, 12, 0, 0> LaneStockTop;       // Google Stock: h = hh = 12, 2 layers, 6 channels
typedef LaneStatic<6, 8, 3, 8, 4, 4, 0> LaneStockBottom;       //      attention 4 / 8
// Mujoco (14 channels) runs at B = 1024: 64 registers = four CTAs per SM put its attention network's forward in two waves of CTAs
// instead of three (1.05 -> 1.02 ms per step); the same cap costs the UEA networks 27 % in spills at that batch size
typedef LaneStatic<12, 16, 2, 12, 12, 0, 0, 4> LaneMujocoTop;
typedef LaneStatic<14, 16, 3, 8, 4, 4, 0, 4> LaneMujocoBottom;

LaneKernel lane_kernel(const NcdeLayout& lo, bool bwd, bool save)
{
    const bool big = lo.B >= 240;         // several CTAs per SM: idle warps skip the layer bodies, control from warp 0
    const bool dyn_only = getenv("ANCDE_LANE_DYNAMIC") && getenv("ANCDE_LANE_DYNAMIC")[0] == '1';   // tests: the run-time-bounds code
#define LANE_PICK(S) (bwd ? (big ? ncde_lane_bwd_kernel<S, true> : ncde_lane_bwd_kernel<S, false>)                                   \
                          : (save ? (big ? ncde_lane_fwd_kernel<S, true, true> : ncde_lane_fwd_kernel<S, true, false>)              \
                                  : (big ? ncde_lane_fwd_kernel<S, false, true> : ncde_lane_fwd_kernel<S, false, false>)))
#define LANE_TRY(S) if (!dyn_only && lane_matches<S>(lo)) return LANE_PICK(S)
    LANE_TRY(LaneUeaTop);
    LANE_TRY(LaneUeaBottom);
    LANE_TRY(LaneStockTop);
    LANE_TRY(LaneStockBottom);
    LANE_TRY(LaneMujocoTop);
    LANE_TRY(LaneMujocoBottom);
#undef LANE_TRY
    return LANE_PICK(LaneDyn);
}

int launch_lane(SolveArgs& a, bool bwd, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    if (lo.TS != 1 && lo.TS != 8) {
        set_error("lane engine: saves in tiles of 1 or 8 samples (got %d)", lo.TS);
        return ANCDE_EINVAL;
    }
    if ((int64_t)lo.B * (lo.L - 1) * lo.C >= ((int64_t)1 << 31) - lo.C) {
        set_error("lane engine: B * (L-1) * C does not fit the 32-bit coefficient offsets of the control lookup");
        return ANCDE_EUNSUPPORTED;
    }
    const int T = lane_threads(lo);
    const size_t smem = bwd ? lane_bwd_smem(lo, T) : lane_fwd_smem(lo);
    if (T > 256 || smem > 227 * 1024) {
        set_error("lane engine: needs %d threads and %zu bytes of shared memory per CTA", T, smem);
        return ANCDE_EUNSUPPORTED;
    }
    void (*kern)(SolveArgs) = lane_kernel(lo, bwd, a.save_act != nullptr);
    ANCDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const char* fnames[3] = {"ncde_fwd_plain", "ncde_fwd_bottom", "ncde_fwd_top"};
    static const char* bnames[3] = {"ncde_bwd_plain", "ncde_bwd_bottom", "ncde_bwd_top"};
    prof_begin(bwd ? bnames[lo.mode] : fnames[lo.mode], stream);
    kern<<<lo.ntiles * lo.TS, T, smem, stream>>>(a);      // one CTA per slot of the saved layout (slots past B hold zeros)
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
    const int64_t bytes = 4 * (hidden + (int64_t)d->Hd * C4 * (8 * ((K + 7) / 8) + 37) + d->L + 4096);
    return bytes <= 200 * 1024;
}

// Samples per tile of the saved layout when the lane engine runs this descriptor (1 or 8), 0 when it does not.
//   * Tiles of 8 samples -- the layout of the tcgen05 weight-gradient kernels -- unless the caller asks for tiles of one: measured
//     on B200 with the same recurrent kernels, UEA (B = 32) 4.69 -> 4.42 ms and Google Stock (B = 200) 0.64 -> 0.57 ms per step
//     (the mma.sync weight-gradient kernels of the one-sample layout take 50-100 us each for this little work); at B = 1024 the
//     one-sample layout costs UEA 9.6 ms of weight gradients.
//   * Small batches (B < 240, the streamed engine's own "one sample per CTA" regime): always the lane engine.
//   * Larger batches of the small networks: measured at B = 1024 (recurrent kernels, lane against streamed): Google Stock 0.44
//     against 1.22 ms, UEA 8.8 against 15.6 ms, Mujoco 0.92 against 1.12 ms -- a CTA per sample keeps 4-7 CTAs resident per SM
//     where the streamed engine has one latency-bound CTA of 8 samples.  Restricted to the measured range: output layer <= 8192
//     weights, <= 80 KB of shared memory per CTA of the forward.
int lane_choice(const ancde_ncde_desc* d)
{
    if (!lane_shape_ok(d)) return 0;
    const int spc = d->samples_per_cta;
    if (spc != 0 && spc != 1 && spc != 8) return 0;
    const int ts = spc ? spc : 8;
    if (d->engine == ANCDE_ENGINE_LANE) return ts;
    if (d->engine != 0) return 0;
    if (d->B < 240) return ts;
    if (spc != 1) {
        const int C4 = pad4(d->C), K = d->width[d->n_hidden - 1];
        int64_t f = (int64_t)d->Hd * C4 * (8 * ((K + 7) / 8) + 5) + d->L + 256;
        for (int l = 0; l < d->n_hidden; ++l) f += (int64_t)d->width[l] * (pad4odd(l == 0 ? d->Hd : d->width[l - 1]) + 1);
        if ((int64_t)d->Hd * C4 * K <= 8192 && 4 * f <= 80 * 1024) return 8;
    }
    return 0;
}

int launch_lane_fwd(SolveArgs a, cudaStream_t stream) { return launch_lane(a, false, stream); }
int launch_lane_bwd(SolveArgs a, cudaStream_t stream) { return launch_lane(a, true, stream); }

}  // namespace ancde
