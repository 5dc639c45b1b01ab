// Tensor-core fused NCDE forward solve (fixed-grid RK4, 3/8 rule): one thread-block cluster integrates NS samples
// over all time steps.  See ncde2_common.cuh for the decomposition.  Per RK4 stage a CTA
//   1. runs the hidden ReLU layers for all NS samples on the tensor cores (FP16 hi/lo split, FP32 accumulate);
//      activations ping-pong between two shared-memory buffers ([n][k] FP16 hi/lo matrices, written with
//      stmatrix.trans straight from the accumulator fragments, read with ldmatrix),               [vector_fields.py:205-216,237-247]
//   2. multiplies its resident slice of the output layer with the last hidden activation (B fragments held in
//      registers for the whole tile loop), applies tanh in the accumulator registers, contracts with dU (spline
//      derivative x attention) and reduces over the control channels with shuffles,                [vector_fields.py:216-218; cdeint_module.py:61,133]
//   3. scatters its rows of the stage derivative to every CTA of the cluster (distributed shared memory) and
//   4. advances the RK4 recurrence for all NS samples; the state lives in registers in accumulator-fragment
//      order, which is also the order every saved tensor is written in (128-bit coalesced stores).  [torchdiffeq rk4_alt_step_func]
// The control of the next stage (spline derivative lookup, attention product) is fetched while the cluster exchange
// is in flight.                                                                                    [cdeint_module.py:58-60,103-129; interpolate.py:261-303]
#include <cuda_runtime.h>
#include <stdlib.h>

#include "ncde2_common.cuh"

namespace ancde {

constexpr int V2_CTL_ITEMS = 6;
constexpr int V2_RK_TILES = 2;
constexpr int V2_FWD_THREADS = 384;

struct FwdSmem {
    size_t wo, wh, bias, act, part, nsl, k, dy, times, total;
    int ntl_max, nipl_max, dys, kp, tpw, slots;
};

__host__ __device__ static inline FwdSmem fwd_smem(const V2Layout& lo, const V2Launch& la)
{
    FwdSmem s;
    int nip = 0;
    for (int r = 0; r < la.CS; ++r) {
        const int n = la.ip_begin[r + 1] - la.ip_begin[r];
        nip = n > nip ? n : nip;
    }
    const int nw = la.threads / 32;
    s.nipl_max = nip;
    s.ntl_max = nip * lo.NJC;
    s.tpw = (s.ntl_max + nw - 1) / nw;
    s.slots = (lo.NJC + s.tpw - 1) / s.tpw + 1;
    s.dys = lo.NS + 8;
    s.kp = 16 * lo.KSB + 8;
    size_t o = 0;
    s.wo = o;    o += (size_t)s.ntl_max * lo.of.stride * 16;
    s.wh = o;    o += (size_t)lo.hidden_u4 * 16;
    s.bias = o;
    int nb = s.ntl_max * 16;
    for (int l = 0; l < lo.n_hidden; ++l) nb += lo.hf[l].mtiles * 16;
    o += (size_t)nb * 4;
    s.act = o;   o += (size_t)4 * lo.NS * s.kp * 2;            // hi0, lo0, hi1, lo1
    s.part = o;  o += (size_t)nip * s.slots * 2 * lo.NS * 4;
    s.nsl = o;   o += (size_t)pad4(nip) * 4;
    s.k = o;     o += (size_t)2 * lo.MTz * 16 * lo.NS * 4;
    s.dy = o;    o += (size_t)lo.C8 * s.dys * 4;
    s.times = o; o += (size_t)pad4(lo.L) * 4;
    s.total = (o + 15) & ~(size_t)15;
    return s;
}

// tanh(x) = 1 - 2 / (exp(2x) + 1), x = acc / weight-scale; |abs err| < 4e-7 (MUFU ex2 + rcp)
__device__ __forceinline__ float tanh_scaled(float acc, float c)
{
    float ex, r;
    const float a = acc * c;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(a));
    const float dd = ex + 1.0f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dd));
    return fmaf(-2.0f, r, 1.0f);
}

template <int NT, int MAXTHR>
__global__ void __launch_bounds__(MAXTHR, 1) ncde2_fwd_kernel(const __grid_constant__ V2Args p)
{
    constexpr int NS = 8 * NT;
    const V2Layout& lo = p.lo;
    const V2Launch& la = p.la;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NTHR = blockDim.x, NW = NTHR >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int CS = la.CS;
    const int rank = (int)cluster_ctarank();
    const int cl = blockIdx.x / CS;
    const int b0 = cl * NS;
    const int Hd = lo.Hd, C = lo.C, NJC = lo.NJC, B = lo.B;
    const int ip0 = la.ip_begin[rank], nipl = la.ip_begin[rank + 1] - ip0;
    const int ntl = nipl * NJC, tile0 = ip0 * NJC;
    const bool training = p.save_act != nullptr;
    const bool top = lo.mode == ANCDE_MODE_TOP;
    const int64_t act_stage = (int64_t)lo.nclusters * lo.act_block;          // floats per stage of the saved activations
    const int64_t g_stage = (int64_t)lo.nclusters * lo.ntiles * NT * 128;    // floats per stage of the saved tanh outputs

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FwdSmem sm = fwd_smem(lo, la);
    const uint4* sWo = reinterpret_cast<const uint4*>(smem_raw + sm.wo);
    const uint4* sWh = reinterpret_cast<const uint4*>(smem_raw + sm.wh);
    float* sBias = reinterpret_cast<float*>(smem_raw + sm.bias);     // [ntl_max*16 output bias | hidden biases]
    float* sPart = reinterpret_cast<float*>(smem_raw + sm.part);
    int* sNsl = reinterpret_cast<int*>(smem_raw + sm.nsl);
    float* sK = reinterpret_cast<float*>(smem_raw + sm.k);
    float* sDY = reinterpret_cast<float*>(smem_raw + sm.dy);
    float* sTimes = reinterpret_cast<float*>(smem_raw + sm.times);
    const int DYS = sm.dys, KP = sm.kp, TPW = sm.tpw, SLOTS = sm.slots;
    const int KROWS = lo.MTz * 16;
    const uint32_t act_bytes = (uint32_t)(NS * KP * 2);              // one FP16 matrix
    const uint32_t act_base = smem_u32(smem_raw + sm.act);           // buffer b: hi at act_base + 2*b*act_bytes, lo right behind
    // this lane's row address inside an x4 ldmatrix/stmatrix access, relative to the hi matrix of a buffer
    const uint32_t lane_off = (uint32_t)(((lane & 7) * KP + ((lane >> 3) & 1) * 8) * 2) + ((lane >> 4) ? act_bytes : 0u);
    auto act_addr = [&](int buf, int ks, int nt) -> uint32_t {
        return act_base + (uint32_t)buf * 2u * act_bytes + lane_off + (uint32_t)((nt * 8 * KP + ks * 16) * 2);
    };

    // ---- prologue: resident weights, knot times, zeroed operand buffers
    {
        uint4* dWo = reinterpret_cast<uint4*>(smem_raw + sm.wo);
        const uint4* src = p.wpack + lo.of.off + (size_t)tile0 * lo.of.stride;
        for (int i = tid; i < ntl * lo.of.stride; i += NTHR) dWo[i] = __ldg(src + i);
        uint4* dWh = reinterpret_cast<uint4*>(smem_raw + sm.wh);
        const uint4* srch = p.wpack + lo.hf[0].off;
        for (int i = tid; i < lo.hidden_u4; i += NTHR) dWh[i] = __ldg(srch + i);
        const float* wf = reinterpret_cast<const float*>(p.wpack);
        for (int i = tid; i < ntl * 16; i += NTHR) sBias[i] = __ldg(wf + lo.off_obias + tile0 * 16 + i);
        int nb = sm.ntl_max * 16;
        for (int l = 0; l < lo.n_hidden; ++l) {
            for (int i = tid; i < lo.hf[l].mtiles * 16; i += NTHR) sBias[nb + i] = __ldg(wf + lo.off_hbias[l] + i);
            nb += lo.hf[l].mtiles * 16;
        }
        uint4* za = reinterpret_cast<uint4*>(smem_raw + sm.act);
        for (int i = tid; i < (int)(4 * act_bytes / 16); i += NTHR) za[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < 2 * KROWS * NS; i += NTHR) sK[i] = 0.f;
        for (int i = tid; i < lo.C8 * DYS; i += NTHR) sDY[i] = 0.f;
        for (int i = tid; i < lo.L; i += NTHR) sTimes[i] = p.times[i];
        // number of warps that contribute partial sums to i-pair ipl (tiles are dealt to warps in runs of TPW)
        for (int i = tid; i < nipl; i += NTHR) sNsl[i] = ((i + 1) * NJC - 1) / TPW - (i * NJC) / TPW + 1;
    }
    __syncthreads();

    // ---- RK4 state in accumulator-fragment order: state tile = (m-tile of the state, n-tile)
    float ry[V2_RK_TILES][4], rk1[V2_RK_TILES][4], rk2[V2_RK_TILES][4], rk3[V2_RK_TILES][4];
    const int n_rkt = lo.MTz * NT;
#pragma unroll
    for (int r = 0; r < V2_RK_TILES; ++r) {
        const int tile = warp + r * NW;
#pragma unroll
        for (int e = 0; e < 4; ++e) ry[r][e] = rk1[r][e] = rk2[r][e] = rk3[r][e] = 0.f;
        if (tile < n_rkt) {
            const int mt = tile / NT, nt = tile - mt * NT;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = mt * 16 + g + (e >> 1) * 8, n = nt * 8 + 2 * t + (e & 1);
                const int bg = b0 + n;
                const float v = (i < Hd && bg < B) ? __ldg(p.z0 + (int64_t)bg * Hd + i) : 0.f;
                ry[r][e] = v;
                if (i < Hd && bg < B && (nt % CS) == rank) p.z_out[(int64_t)bg * Hd + i] = v;      // solution[0] = y0
            }
            store_act_tile(act_addr(0, mt, nt), ry[r]);
            if (training && (nt % CS) == rank)
                *reinterpret_cast<float4*>(p.save_act + (int64_t)cl * lo.act_block + lo.fo_z + (tile * 32 + lane) * 4) =
                    make_float4(ry[r][0], ry[r][1], ry[r][2], ry[r][3]);
        }
    }

    // ---- control: dU[j][n] of a stage (spline derivative, attention product); item = n*C + j
    int cnt = 0;        // knots strictly below the current stage time (monotone)
    float cb[V2_CTL_ITEMS], cc2[V2_CTL_ITEMS], cd3[V2_CTL_ITEMS], ca[V2_CTL_ITEMS], chp[V2_CTL_ITEMS];
    float cfrac = 0.f;
    const int n_ctl = C * NS;
    const int n_ctl_it = (n_ctl + NTHR - 1) / NTHR;
    auto ctl_issue = [&](int k, int s) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step);
        const float ts = stage_time(t0, __fsub_rn(t1, t0), s);
        // index = clamp(#(times < t) - 1, 0, L-2): the LEFT piece at a knot (interpolate.py:261-267)
        while (cnt < lo.L && sTimes[cnt] < ts) ++cnt;
        int idx = cnt - 1;
        idx = idx < 0 ? 0 : (idx > lo.L - 2 ? lo.L - 2 : idx);
        cfrac = __fsub_rn(ts, sTimes[idx]);
        int q = (int)floorf(ts) - 1;                  // attention[int(floor(t)) - 1], python wrap
        if (q < 0) q += lo.att_L;
        q = q < 0 ? 0 : (q >= lo.att_L ? lo.att_L - 1 : q);
        const int m = top ? (4 * k + s) % (lo.n_hp - 1) : 0;   // call counter + reset rule (cdeint_module.py:134-137)
        const float* pb = p.cb + idx * C;
        const float* pc = p.cc2 + idx * C;
        const float* pd = p.cd3 + idx * C;
        const float* pa = top ? p.attention + (int64_t)q * B * lo.att_C : nullptr;
        const float* ph = top ? p.h_prime + (int64_t)m * B * C : nullptr;
        const int rowlen = (lo.L - 1) * C;
#pragma unroll
        for (int it = 0; it < V2_CTL_ITEMS; ++it) {
            if (it < n_ctl_it) {
                cb[it] = cc2[it] = cd3[it] = 0.f;
                ca[it] = chp[it] = 0.f;
                const int item = tid + it * NTHR;
                const int n = (C == 1) ? item : (int)__umulhi((unsigned)item, lo.magicC), j = item - n * C;
                const int bg = b0 + n;
                if (item < n_ctl && bg < B) {
                    const int o = bg * rowlen + j;
                    cb[it] = __ldg(pb + o);
                    cc2[it] = __ldg(pc + o);
                    cd3[it] = __ldg(pd + o);
                    if (top) {
                        ca[it] = __ldg(pa + bg * lo.att_C + (lo.att_C == 1 ? 0 : j));
                        chp[it] = __ldg(ph + bg * C + j);
                    }
                }
            }
        }
    };
    auto ctl_finish = [&](int stage) {
        float* act = (training && rank == 0) ? p.save_act + (int64_t)stage * act_stage + (int64_t)cl * lo.act_block : nullptr;
#pragma unroll
        for (int it = 0; it < V2_CTL_ITEMS; ++it) {
            const int item = tid + it * NTHR;
            if (it < n_ctl_it && item < n_ctl) {
                const int n = (C == 1) ? item : (int)__umulhi((unsigned)item, lo.magicC), j = item - n * C;
                // derivative = b + (two_c + three_d * frac) * frac   (interpolate.py:298-303)
                const float dX = __fadd_rn(cb[it], __fmul_rn(__fadd_rn(cc2[it], __fmul_rn(cd3[it], cfrac)), cfrac));
                float dU = dX, duda = 0.f;
                if (top) {
                    const float a = ca[it], hp = chp[it];
                    // dY = dX*a + ((a*(1-a))*dX)*h'   (cdeint_module.py:113-129; "Xt" is dX/dt)
                    dU = __fadd_rn(__fmul_rn(dX, a), __fmul_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.0f, a)), dX), hp));
                    duda = dX * (1.0f + (1.0f - 2.0f * a) * hp);
                }
                sDY[j * DYS + n] = dU;
                if (act) {
                    act[lo.fo_du + item] = dU;
                    if (top) act[lo.fo_duda + item] = duda;
                }
            }
        }
    };
    ctl_issue(0, 0);
    ctl_finish(0);

    // ---- per-warp run of output-layer tiles and the slot its partial sums go to
    const int tl_begin = warp * TPW;
    const int tl_end = (tl_begin + TPW < ntl) ? tl_begin + TPW : ntl;
    const int ipl_first = tl_begin / NJC, jc_first = tl_begin - ipl_first * NJC;
    const int slot_first = warp - (ipl_first * NJC) / TPW;
    uint32_t sK_rank[V2_MAX_CS];
#pragma unroll
    for (int r = 0; r < V2_MAX_CS; ++r) sK_rank[r] = map_to_rank(smem_u32(sK), (uint32_t)(r < CS ? r : 0));

    cluster_arrive();     // every CTA of the cluster has initialised its shared memory before anyone writes into it
    cluster_wait();

    int jout = 1;         // next requested output time
    PH2_DECL();
    const float tanh_c = 2.885390081777927f * V2_WSCALE_INV;      // 2*log2(e) / weight scale
    const int nst = lo.of.full + lo.of.tail8;

    for (int k = 0; k < lo.n_steps; ++k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step);
        const float dt = __fsub_rn(t1, t0);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const int stage = 4 * k + s;
            float* act = training ? p.save_act + (int64_t)stage * act_stage + (int64_t)cl * lo.act_block : nullptr;
            const bool have_next = stage + 1 < lo.n_stages;
            PH2_T0();

            // ---- (1) hidden layers; buffer 0 holds the stage input
            int cur = 0;
            int boff = sm.ntl_max * 16;
#pragma unroll 1
            for (int l = 0; l < lo.n_hidden; ++l) {
                const int mtiles = lo.hf[l].mtiles, full = lo.hf[l].full, tail8 = lo.hf[l].tail8, stride = lo.hf[l].stride;
                const uint4* Wl = sWh + (lo.hf[l].off - lo.hf[0].off);
                const float* bl = sBias + boff;
                for (int item = warp; item < mtiles * NT; item += NW) {
                    const int mt = item / NT, nt = item - mt * NT;
                    const uint4* A = Wl + mt * stride + lane;
                    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4];
                    a2[0] = a2[1] = bl[mt * 16 + g];
                    a2[2] = a2[3] = bl[mt * 16 + g + 8];
                    uint32_t baddr = act_addr(cur, 0, nt);
#pragma unroll 2
                    for (int ks = 0; ks < full; ++ks) {
                        const uint4 ah = A[ks * 64], al = A[ks * 64 + 32];
                        uint32_t bh0, bh1, bl0, bl1;
                        ldmatrix_x4(baddr, bh0, bh1, bl0, bl1);
                        baddr += 32;
                        mma16(a0, al.x, al.y, al.z, al.w, bh0, bh1);
                        mma16(a1, ah.x, ah.y, ah.z, ah.w, bl0, bl1);
                        mma16(a2, ah.x, ah.y, ah.z, ah.w, bh0, bh1);
                    }
                    if (tail8) {
                        const uint4 a = A[full * 64];
                        uint32_t bh0, bh1, bl0, bl1;
                        ldmatrix_x4(baddr, bh0, bh1, bl0, bl1);
                        mma8(a0, a.z, a.w, bh0);
                        mma8(a1, a.x, a.y, bl0);
                        mma8(a2, a.x, a.y, bh0);
                    }
                    float h[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) h[e] = fmaxf((a0[e] + a1[e] + a2[e]) * V2_WSCALE_INV, 0.f);
                    store_act_tile(act_addr(cur ^ 1, mt, nt), h);
                    if (act && (nt % CS) == rank)
                        *reinterpret_cast<float4*>(act + lo.fo_h[l] + (item * 32 + lane) * 4) = make_float4(h[0], h[1], h[2], h[3]);
                }
                __syncthreads();
                cur ^= 1;
                boff += mtiles * 16;
            }
            PH2_MARK(0);

            // ---- (2) output layer slice x all samples, tanh, contraction with dU
            {
                uint32_t bh[V2_MAX_KSTEPS][NT][2], bl[V2_MAX_KSTEPS][NT][2];
#pragma unroll
                for (int ks = 0; ks < V2_MAX_KSTEPS; ++ks)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        bh[ks][nt][0] = bh[ks][nt][1] = bl[ks][nt][0] = bl[ks][nt][1] = 0u;
                        if (ks < nst) ldmatrix_x4(act_addr(cur, ks, nt), bh[ks][nt][0], bh[ks][nt][1], bl[ks][nt][0], bl[ks][nt][1]);
                    }
                float part[4 * NT];
#pragma unroll
                for (int e = 0; e < 4 * NT; ++e) part[e] = 0.f;
                int ipl = ipl_first, jc = jc_first, slot = slot_first;
                float4* gsave = training ? reinterpret_cast<float4*>(p.save_G + (int64_t)stage * g_stage) +
                                               ((int64_t)cl * lo.ntiles + tile0 + tl_begin) * NT * 32 + lane
                                         : nullptr;
                auto flush = [&]() {
                    reduce_scatter_g<4 * NT>(part, lane);
                    constexpr int V = 4 * NT;
                    constexpr int PER = (V >= 8) ? V / 8 : 1;
                    int base;
                    bool writer = true;
                    if (V >= 8) base = (g >> 2) * (V / 2) + ((g >> 1) & 1) * (V / 4) + (g & 1) * (V / 8);
                    else { base = (g >> 2) * 2 + ((g >> 1) & 1); writer = (g & 1) == 0; }
                    if (writer) {
#pragma unroll
                        for (int e = 0; e < PER; ++e) {
                            const int v = base + e;
                            const int nt = v >> 2, c = v & 3;
                            sPart[((ipl * SLOTS + slot) * 2 + (c >> 1)) * NS + nt * 8 + 2 * t + (c & 1)] = part[e];
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4 * NT; ++e) part[e] = 0.f;
                };
#pragma unroll 1
                for (int tl = tl_begin; tl < tl_end; ++tl) {
                    const uint4* A = sWo + tl * lo.of.stride + lane;
                    float acc[NT][4];
                    {
                        const float bia0 = sBias[tl * 16 + g], bia1 = sBias[tl * 16 + g + 8];
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = bia0; acc[nt][2] = acc[nt][3] = bia1; }
                    }
#pragma unroll
                    for (int ks = 0; ks < V2_MAX_KSTEPS; ++ks) {
                        if (ks < lo.of.full) {
                            const uint4 ah = A[ks * 64], al = A[ks * 64 + 32];
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt) mma16x3(acc[nt], ah, al, bh[ks][nt][0], bh[ks][nt][1], bl[ks][nt][0], bl[ks][nt][1]);
                        } else if (ks == lo.of.full && lo.of.tail8) {
                            const uint4 a = A[ks * 64];
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt) mma8x3(acc[nt], a, bh[ks][nt][0], bl[ks][nt][0]);
                        }
                    }
                    // epilogue: rows g / g+8 of the tile are (i0, j) and (i0+1, j); columns 2t, 2t+1 are samples
                    const float* dyp = sDY + (8 * jc + g) * DYS + 2 * t;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const float2 dy = *reinterpret_cast<const float2*>(dyp + nt * 8);
                        float gv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) gv[e] = tanh_scaled(acc[nt][e], tanh_c);
                        part[nt * 4 + 0] = fmaf(gv[0], dy.x, part[nt * 4 + 0]);
                        part[nt * 4 + 1] = fmaf(gv[1], dy.y, part[nt * 4 + 1]);
                        part[nt * 4 + 2] = fmaf(gv[2], dy.x, part[nt * 4 + 2]);
                        part[nt * 4 + 3] = fmaf(gv[3], dy.y, part[nt * 4 + 3]);
                        if (training) gsave[nt * 32] = make_float4(gv[0], gv[1], gv[2], gv[3]);
                    }
                    if (training) gsave += NT * 32;
                    if (++jc == NJC) {       // last control chunk of this i-pair: its partial sums are complete
                        flush();
                        jc = 0; ++ipl; slot = 0;
                    }
                }
                if (jc != 0 && tl_begin < tl_end) flush();
            }
            PH2_MARK(1);
            __syncthreads();
            PH2_MARK(2);

            // ---- (3) control of the next stage (loads in flight during the exchange) + scatter of the stage derivative
            const int kn = (s == 3) ? k + 1 : k, sn = (s == 3) ? 0 : s + 1;
            if (have_next) ctl_issue(kn, sn);
            const int buf = stage & 1;
            for (int item = tid; item < 2 * nipl * NS; item += NTHR) {
                const int il = item / NS, n = item - il * NS;
                const int i = 2 * ip0 + il;
                if (i < Hd) {
                    const int ipl = il >> 1;
                    const float* pp = sPart + ((ipl * SLOTS) * 2 + (il & 1)) * NS + n;
                    const int ns = sNsl[ipl];
                    float o = pp[0];
                    for (int sl = 1; sl < ns; ++sl) o += pp[sl * 2 * NS];
                    const uint32_t off = (uint32_t)(((buf * KROWS + i) * NS + n) * 4);
#pragma unroll
                    for (int r = 0; r < V2_MAX_CS; ++r)
                        if (r < CS) st_cluster_f32(sK_rank[r] + off, o);
                    const int bg = b0 + n;
                    if (p.k_out && bg < B) p.k_out[((int64_t)stage * B + bg) * Hd + i] = o;
                }
            }
            PH2_MARK(3);
            cluster_arrive();
            cluster_wait();
            PH2_MARK(4);
            if (have_next) ctl_finish(stage + 1);
            PH2_MARK(5);

            // ---- (4) RK4 recurrence (rk4_alt_step_func), next stage input into buffer 0
            float* act_next = (training && have_next) ? p.save_act + (int64_t)(stage + 1) * act_stage + (int64_t)cl * lo.act_block : nullptr;
            const float* kk = sK + buf * KROWS * NS;
            const float third = 1.0f / 3.0f;
#pragma unroll
            for (int r = 0; r < V2_RK_TILES; ++r) {
                const int tile = warp + r * NW;
                if (tile < n_rkt) {
                    const int mt = tile / NT, nt = tile - mt * NT;
                    const float2 k_lo = *reinterpret_cast<const float2*>(kk + (mt * 16 + g) * NS + nt * 8 + 2 * t);
                    const float2 k_hi = *reinterpret_cast<const float2*>(kk + (mt * 16 + g + 8) * NS + nt * 8 + 2 * t);
                    const float kv[4] = {k_lo.x, k_lo.y, k_hi.x, k_hi.y};
                    float zin[4];
                    if (s == 0) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { rk1[r][e] = kv[e]; zin[e] = fmaf(dt * third, kv[e], ry[r][e]); }                  // y + dt*k1/3
                    } else if (s == 1) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { rk2[r][e] = kv[e]; zin[e] = fmaf(dt, kv[e] - rk1[r][e] * third, ry[r][e]); }       // y + dt*(k2 - k1/3)
                    } else if (s == 2) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { rk3[r][e] = kv[e]; zin[e] = fmaf(dt, (rk1[r][e] - rk2[r][e]) + kv[e], ry[r][e]); } // y + dt*(k1-k2+k3)
                    } else {
                        float y1[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            y1[e] = fmaf(((rk1[r][e] + 3.0f * rk2[r][e]) + 3.0f * rk3[r][e]) + kv[e], dt * 0.125f, ry[r][e]);
                        // emit every requested time in (t0, t1] (FixedGridODESolver.integrate + _linear_interp)
                        if ((nt % CS) == rank) {
                            int jj = jout;
                            while (jj < lo.n_out) {
                                const float tj = __ldg(p.t_out + jj);
                                if (!(t1 >= tj)) break;
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int i = mt * 16 + g + (e >> 1) * 8, bg = b0 + nt * 8 + 2 * t + (e & 1);
                                    float v;
                                    if (tj == t0) v = ry[r][e];
                                    else if (tj == t1) v = y1[e];
                                    else v = ry[r][e] + (y1[e] - ry[r][e]) / (t1 - t0) * (tj - t0);
                                    if (i < Hd && bg < B) p.z_out[((int64_t)jj * B + bg) * Hd + i] = v;
                                }
                                ++jj;
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) { ry[r][e] = y1[e]; zin[e] = y1[e]; }
                    }
                    store_act_tile(act_addr(0, mt, nt), zin);
                    if (act_next && (nt % CS) == rank)
                        *reinterpret_cast<float4*>(act_next + lo.fo_z + (tile * 32 + lane) * 4) = make_float4(zin[0], zin[1], zin[2], zin[3]);
                }
            }
            if (s == 3) {
                while (jout < lo.n_out && t1 >= __ldg(p.t_out + jout)) ++jout;
            }
            __syncthreads();
            PH2_MARK(6);
        }
    }
    PH2_FLUSH(16 + (top ? 8 : 0));
    cluster_arrive();     // nobody leaves while a peer may still address its shared memory
    cluster_wait();
}

// ---- launch -----------------------------------------------------------------------------------------------
// Cluster size: the smallest power of two whose slice of the output layer fits in shared memory next to the hidden
// layers; samples per cluster: 8 per CTA (so that every SM owns 8 samples), at most 32.
static int fwd_threads_default(int NT)
{
    static int forced = -1;      // development override: ANCDE_V2_THREADS=256|384
    if (forced < 0) {
        const char* e = getenv("ANCDE_V2_THREADS");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 256 || forced == 384) return forced;
    return NT >= 4 ? 256 : 384;
}

int choose_launch2_fwd(V2Layout* lo, V2Launch* la)
{
    const size_t max_smem = 227 * 1024;
    static const int cand[][2] = {{1, 1}, {2, 2}, {4, 4}, {8, 4}, {8, 2}, {8, 1}};
    for (const auto& c : cand) {
        const int cs = c[0];
        la->CS = cs;
        lo->NT = c[1];
        lo->NS = 8 * lo->NT;
        lo->nclusters = (lo->B + lo->NS - 1) / lo->NS;
        la->threads = fwd_threads_default(lo->NT);
        const int base = lo->NIP / cs, rem = lo->NIP % cs;
        la->ip_begin[0] = 0;
        for (int r = 0; r < cs; ++r) la->ip_begin[r + 1] = la->ip_begin[r] + base + (r < rem ? 1 : 0);
        for (int r = cs; r < V2_MAX_CS; ++r) la->ip_begin[r + 1] = la->ip_begin[cs];
        if (lo->NIP < cs) continue;
        la->smem = fwd_smem(*lo, *la).total;
        const int nw = la->threads / 32;
        if (la->smem <= max_smem && lo->MTz * lo->NT <= V2_RK_TILES * nw && lo->C * lo->NS <= V2_CTL_ITEMS * la->threads &&
            (int64_t)lo->B * (lo->L - 1) * lo->C < ((int64_t)1 << 31))
            return ANCDE_OK;
    }
    set_error("ncde_forward: the output layer (%d x %d) does not fit the shared memory of a cluster of %d CTAs", lo->Hd * lo->C, lo->K,
              V2_MAX_CS);
    return ANCDE_EUNSUPPORTED;
}

template <int NT, int MAXTHR>
static int launch_fwd2_nt(const V2Args& a, cudaStream_t stream)
{
    const V2Launch& la = a.la;
    ANCDE_CUDA(cudaFuncSetAttribute(ncde2_fwd_kernel<NT, MAXTHR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)la.smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.lo.nclusters * la.CS));
    cfg.blockDim = dim3((unsigned)la.threads);
    cfg.dynamicSmemBytes = la.smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)la.CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static const char* names[3] = {"ncde_fwd_plain", "ncde_fwd_bottom", "ncde_fwd_top"};
    prof_begin(names[a.lo.mode], stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, ncde2_fwd_kernel<NT, MAXTHR>, a);
    prof_end(stream);
    count_launch();
    if (e != cudaSuccess) {
        set_error("ncde_forward: launch failed: %s (cluster %d, %zu bytes of shared memory)", cudaGetErrorString(e), la.CS, la.smem);
        return ANCDE_ECUDA;
    }
    return ANCDE_OK;
}

int launch_fwd2(V2Args& a, cudaStream_t stream)
{
    int rc = choose_launch2_fwd(&a.lo, &a.la);
    if (rc != ANCDE_OK) return rc;
    const bool wide = a.la.threads > 256;
    switch (a.lo.NT) {
        case 4: return wide ? launch_fwd2_nt<4, 384>(a, stream) : launch_fwd2_nt<4, 256>(a, stream);
        case 2: return wide ? launch_fwd2_nt<2, 384>(a, stream) : launch_fwd2_nt<2, 256>(a, stream);
        default: return wide ? launch_fwd2_nt<1, 384>(a, stream) : launch_fwd2_nt<1, 256>(a, stream);
    }
}

}  // namespace ancde
