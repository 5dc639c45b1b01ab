// Tensor-core fused NCDE forward solve (fixed-grid RK4, 3/8 rule): one thread-block cluster integrates NS samples
// over all time steps.  See ncde2_common.cuh for the decomposition.  Per RK4 stage a CTA
//   1. runs the hidden ReLU layers for all NS samples on the tensor cores (FP16 hi/lo split, FP32 accumulate),
//      activations ping-pong between two shared-memory buffers kept in B-fragment order,          [vector_fields.py:205-216,237-247]
//   2. multiplies its resident slice of the output layer with the last hidden activation (held in registers as
//      B fragments for the whole tile loop), applies tanh in the accumulator registers, contracts with dU
//      (spline derivative x attention) and reduces over the control channels with shuffles,       [vector_fields.py:216-218; cdeint_module.py:61,133]
//   3. scatters its rows of the stage derivative to every CTA of the cluster (distributed shared memory) and
//   4. advances the RK4 recurrence for all NS samples (state in registers).                        [torchdiffeq rk4_alt_step_func]
// The control of the next stage (spline derivative lookup, attention product) is fetched while the cluster exchange
// is in flight.                                                                                    [cdeint_module.py:58-60,103-129; interpolate.py:261-303]
#include <cuda_runtime.h>

#include "ncde2_common.cuh"

namespace ancde {

constexpr int V2_CTL_ITEMS = 6;
constexpr int V2_FWD_THREADS = 384;

struct FwdSmem {
    size_t wo, wh, bias, bf, part, k, dy, times, total;
    int ntl_max, dys;
};

__host__ __device__ static inline FwdSmem fwd_smem(const V2Layout& lo, const V2Launch& la)
{
    FwdSmem s;
    int ntl = 0;
    for (int r = 0; r < la.CS; ++r) {
        const int n = (la.ip_begin[r + 1] - la.ip_begin[r]) * lo.NJC;
        ntl = n > ntl ? n : ntl;
    }
    s.ntl_max = ntl;
    s.dys = la.NS + 8;
    size_t o = 0;
    s.wo = o;    o += (size_t)ntl * lo.of.stride * 16;
    s.wh = o;    o += (size_t)lo.hidden_u4 * 16;
    s.bias = o;
    int nb = ntl * 16;
    for (int l = 0; l < lo.n_hidden; ++l) nb += lo.hf[l].mtiles * 16;
    o += (size_t)nb * 4;
    s.bf = o;    o += (size_t)2 * lo.KSB * la.NT * 32 * 16;
    s.part = o;  o += (size_t)ntl * 2 * la.NS * 4;
    s.k = o;     o += (size_t)2 * lo.Hd * la.NS * 4;
    s.dy = o;    o += (size_t)lo.C8 * s.dys * 4;
    s.times = o; o += (size_t)pad4(lo.L) * 4;
    s.total = (o + 15) & ~(size_t)15;
    return s;
}

template <int NT>
__global__ void __launch_bounds__(V2_FWD_THREADS, 1) ncde2_fwd_kernel(const __grid_constant__ V2Args p)
{
    constexpr int NS = 8 * NT;
    const V2Layout& lo = p.lo;
    const V2Launch& la = p.la;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NTHR = blockDim.x, NW = NTHR >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int CS = la.CS;
    const int rank = (int)cluster_ctarank();
    const int b0 = (blockIdx.x / CS) * NS;
    const int Hd = lo.Hd, C = lo.C, C8 = lo.C8, NJC = lo.NJC, B = lo.B;
    const int ip0 = la.ip_begin[rank], nipl = la.ip_begin[rank + 1] - ip0;
    const int ntl = nipl * NJC, tile0 = ip0 * NJC;
    const bool training = p.save_act != nullptr;
    const int64_t act_stage = (int64_t)lo.act_rows * B;     // floats per stage of the saved activations

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FwdSmem sm = fwd_smem(lo, la);
    uint4* sWo = reinterpret_cast<uint4*>(smem_raw + sm.wo);
    uint4* sWh = reinterpret_cast<uint4*>(smem_raw + sm.wh);
    float* sBias = reinterpret_cast<float*>(smem_raw + sm.bias);     // [ntl_max*16 output bias | hidden biases]
    uint4* sBf = reinterpret_cast<uint4*>(smem_raw + sm.bf);
    float* sPart = reinterpret_cast<float*>(smem_raw + sm.part);
    float* sK = reinterpret_cast<float*>(smem_raw + sm.k);
    float* sDY = reinterpret_cast<float*>(smem_raw + sm.dy);
    float* sTimes = reinterpret_cast<float*>(smem_raw + sm.times);
    const int DYS = sm.dys;
    const int BFN = lo.KSB * NT * 32;        // uint4 per activation buffer
    const uint32_t sK_addr = smem_u32(sK);

    // ---- prologue: resident weights, knot times, zeroed operand buffers
    {
        const uint4* src = p.wpack + lo.of.off + (size_t)tile0 * lo.of.stride;
        for (int i = tid; i < ntl * lo.of.stride; i += NTHR) sWo[i] = __ldg(src + i);
        const uint4* srch = p.wpack + lo.hf[0].off;
        for (int i = tid; i < lo.hidden_u4; i += NTHR) sWh[i] = __ldg(srch + i);
        const float* wf = reinterpret_cast<const float*>(p.wpack);
        for (int i = tid; i < ntl * 16; i += NTHR) sBias[i] = __ldg(wf + lo.off_obias + tile0 * 16 + i);
        int nb = sm.ntl_max * 16;
        for (int l = 0; l < lo.n_hidden; ++l) {
            for (int i = tid; i < lo.hf[l].mtiles * 16; i += NTHR) sBias[nb + i] = __ldg(wf + lo.off_hbias[l] + i);
            nb += lo.hf[l].mtiles * 16;
        }
        for (int i = tid; i < 2 * BFN; i += NTHR) sBf[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < C8 * DYS; i += NTHR) sDY[i] = 0.f;
        for (int i = tid; i < lo.L; i += NTHR) sTimes[i] = p.times[i];
    }
    __syncthreads();

    // ---- RK4 state: item = (state element i, sample n), n fastest
    float ry[V2_RK_ITEMS], rk1[V2_RK_ITEMS], rk2[V2_RK_ITEMS], rk3[V2_RK_ITEMS];
    const int n_items = Hd * NS;
#pragma unroll
    for (int it = 0; it < V2_RK_ITEMS; ++it) {
        const int item = tid + it * NTHR;
        ry[it] = rk1[it] = rk2[it] = rk3[it] = 0.f;
        if (item < n_items) {
            const int i = item / NS, n = item - i * NS;
            const int bg = b0 + n;
            const float v = (bg < B) ? __ldg(p.z0 + (int64_t)bg * Hd + i) : 0.f;
            ry[it] = v;
            bfrag_store(sBf, NT, i, n, v);
            if (bg < B && ((n >> 3) % CS) == rank) {
                p.z_out[(int64_t)bg * Hd + i] = v;                                   // solution[0] = y0
                if (training) p.save_act[(int64_t)(lo.row_z + i) * B + bg] = v;
            }
        }
    }

    // ---- control: dU[j][n] of a stage (spline derivative, attention product)
    int cnt = 0;        // knots strictly below the current stage time (monotone)
    float cb[V2_CTL_ITEMS], cc2[V2_CTL_ITEMS], cd3[V2_CTL_ITEMS], ca[V2_CTL_ITEMS], chp[V2_CTL_ITEMS];
    float cfrac = 0.f;
    const int n_ctl = C * NS;
    auto ctl_index = [&](int k, int s, int& idx, int& q, int& m) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step);
        const float ts = stage_time(t0, __fsub_rn(t1, t0), s);
        // index = clamp(#(times < t) - 1, 0, L-2): the LEFT piece at a knot (interpolate.py:261-267)
        while (cnt < lo.L && sTimes[cnt] < ts) ++cnt;
        idx = cnt - 1;
        idx = idx < 0 ? 0 : (idx > lo.L - 2 ? lo.L - 2 : idx);
        cfrac = __fsub_rn(ts, sTimes[idx]);
        q = (int)floorf(ts) - 1;                      // attention[int(floor(t)) - 1], python wrap
        if (q < 0) q += lo.att_L;
        q = q < 0 ? 0 : (q >= lo.att_L ? lo.att_L - 1 : q);
        const int n = 4 * k + s;
        m = (lo.mode == ANCDE_MODE_TOP) ? n % (lo.n_hp - 1) : 0;   // call counter + reset rule (cdeint_module.py:134-137)
    };
    auto ctl_issue = [&](int k, int s) {
        int idx, q, m;
        ctl_index(k, s, idx, q, m);
#pragma unroll
        for (int it = 0; it < V2_CTL_ITEMS; ++it) {
            const int item = tid + it * NTHR;
            cb[it] = cc2[it] = cd3[it] = 0.f;
            ca[it] = chp[it] = 0.f;
            if (item < n_ctl) {
                const int n = item / C, j = item - n * C;
                const int bg = b0 + n;
                if (bg < B) {
                    const int64_t o = ((int64_t)bg * (lo.L - 1) + idx) * C + j;
                    cb[it] = __ldg(p.cb + o);
                    cc2[it] = __ldg(p.cc2 + o);
                    cd3[it] = __ldg(p.cd3 + o);
                    if (lo.mode == ANCDE_MODE_TOP) {
                        ca[it] = __ldg(p.attention + ((int64_t)q * B + bg) * lo.att_C + (lo.att_C == 1 ? 0 : j));
                        chp[it] = __ldg(p.h_prime + ((int64_t)m * B + bg) * C + j);
                    }
                }
            }
        }
    };
    auto ctl_finish = [&](int stage) {
        float* act = training ? p.save_act + (int64_t)stage * act_stage : nullptr;
#pragma unroll
        for (int it = 0; it < V2_CTL_ITEMS; ++it) {
            const int item = tid + it * NTHR;
            if (item < n_ctl) {
                const int n = item / C, j = item - n * C;
                // derivative = b + (two_c + three_d * frac) * frac   (interpolate.py:298-303)
                const float dX = __fadd_rn(cb[it], __fmul_rn(__fadd_rn(cc2[it], __fmul_rn(cd3[it], cfrac)), cfrac));
                float dU = dX, duda = 0.f;
                if (lo.mode == ANCDE_MODE_TOP) {
                    const float a = ca[it], hp = chp[it];
                    // dY = dX*a + ((a*(1-a))*dX)*h'   (cdeint_module.py:113-129; "Xt" is dX/dt)
                    dU = __fadd_rn(__fmul_rn(dX, a), __fmul_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.0f, a)), dX), hp));
                    duda = dX * (1.0f + (1.0f - 2.0f * a) * hp);
                }
                sDY[j * DYS + n] = dU;
                const int bg = b0 + n;
                if (act && bg < B && ((n >> 3) % CS) == rank) {
                    act[(int64_t)(lo.row_du + j) * B + bg] = dU;
                    if (lo.mode == ANCDE_MODE_TOP) act[(int64_t)(lo.row_duda + j) * B + bg] = duda;
                }
            }
        }
    };
    ctl_issue(0, 0);
    ctl_finish(0);
    cluster_arrive();     // every CTA of the cluster has initialised its shared memory before anyone writes into it
    cluster_wait();

    int jout = 1;         // next requested output time
    const float tanh_c = 2.885390081777927f * V2_WSCALE_INV;      // 2*log2(e) / weight scale

    for (int k = 0; k < lo.n_steps; ++k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step);
        const float dt = __fsub_rn(t1, t0);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const int stage = 4 * k + s;
            float* act = training ? p.save_act + (int64_t)stage * act_stage : nullptr;
            const bool have_next = stage + 1 < lo.n_stages;

            // ---- (1) hidden layers; buffer 0 holds the stage input
            int cur = 0;
            int boff = sm.ntl_max * 16;
            for (int l = 0; l < lo.n_hidden; ++l) {
                const APack ap = lo.hf[l];
                const uint4* Bin = sBf + cur * BFN;
                uint4* Bout = sBf + (cur ^ 1) * BFN;
                const uint4* Wl = sWh + (ap.off - lo.hf[0].off);
                const float* bl = sBias + boff;
                for (int item = warp; item < ap.mtiles * NT; item += NW) {
                    const int mt = item / NT, nt = item - mt * NT;
                    const uint4* A = Wl + mt * ap.stride;
                    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4];
                    a2[0] = a2[1] = bl[mt * 16 + g];
                    a2[2] = a2[3] = bl[mt * 16 + g + 8];
                    for (int ks = 0; ks < ap.full; ++ks) {
                        const uint4 ah = A[ks * 64 + lane], al = A[ks * 64 + 32 + lane];
                        const uint4 bb = Bin[(ks * NT + nt) * 32 + lane];
                        mma16(a0, al.x, al.y, al.z, al.w, bb.x, bb.y);
                        mma16(a1, ah.x, ah.y, ah.z, ah.w, bb.z, bb.w);
                        mma16(a2, ah.x, ah.y, ah.z, ah.w, bb.x, bb.y);
                    }
                    if (ap.tail8) {
                        const uint4 a = A[ap.full * 64 + lane];
                        const uint4 bb = Bin[(ap.full * NT + nt) * 32 + lane];
                        mma8(a0, a.z, a.w, bb.x);
                        mma8(a1, a.x, a.y, bb.z);
                        mma8(a2, a.x, a.y, bb.x);
                    }
                    float h[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) h[e] = fmaxf((a0[e] + a1[e] + a2[e]) * V2_WSCALE_INV, 0.f);
                    const int o0 = mt * 16 + g, o1 = o0 + 8, n0 = nt * 8 + 2 * t;
                    if (o0 < ap.M) { bfrag_store(Bout, NT, o0, n0, h[0]); bfrag_store(Bout, NT, o0, n0 + 1, h[1]); }
                    if (o1 < ap.M) { bfrag_store(Bout, NT, o1, n0, h[2]); bfrag_store(Bout, NT, o1, n0 + 1, h[3]); }
                    if (act && (nt % CS) == rank) {
                        const int bg = b0 + n0;
                        float* r0 = act + (int64_t)(lo.row_h[l] + o0) * B + bg;
                        float* r1 = act + (int64_t)(lo.row_h[l] + o1) * B + bg;
                        if (o0 < ap.M) { if (bg < B) r0[0] = h[0]; if (bg + 1 < B) r0[1] = h[1]; }
                        if (o1 < ap.M) { if (bg < B) r1[0] = h[2]; if (bg + 1 < B) r1[1] = h[3]; }
                    }
                }
                __syncthreads();
                cur ^= 1;
                boff += ap.mtiles * 16;
            }

            // ---- (2) output layer slice x all samples, tanh, contraction with dU
            {
                uint32_t bh[V2_MAX_KSTEPS][NT][2], bl[V2_MAX_KSTEPS][NT][2];
                const uint4* Bin = sBf + cur * BFN;
                const int nst = lo.of.full + lo.of.tail8;
#pragma unroll
                for (int ks = 0; ks < V2_MAX_KSTEPS; ++ks)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        uint4 bb = make_uint4(0u, 0u, 0u, 0u);
                        if (ks < nst) bb = Bin[(ks * NT + nt) * 32 + lane];
                        bh[ks][nt][0] = bb.x; bh[ks][nt][1] = bb.y;
                        bl[ks][nt][0] = bb.z; bl[ks][nt][1] = bb.w;
                    }
                for (int tl = warp; tl < ntl; tl += NW) {
                    const uint4* A = sWo + tl * lo.of.stride;
                    float acc[NT][4];
                    {
                        const float bia0 = sBias[tl * 16 + g], bia1 = sBias[tl * 16 + g + 8];
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = bia0; acc[nt][2] = acc[nt][3] = bia1; }
                    }
#pragma unroll
                    for (int ks = 0; ks < V2_MAX_KSTEPS; ++ks) {
                        if (ks < lo.of.full) {
                            const uint4 ah = A[ks * 64 + lane], al = A[ks * 64 + 32 + lane];
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt) mma16x3(acc[nt], ah, al, bh[ks][nt][0], bh[ks][nt][1], bl[ks][nt][0], bl[ks][nt][1]);
                        } else if (ks == lo.of.full && lo.of.tail8) {
                            const uint4 a = A[ks * 64 + lane];
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt) mma8x3(acc[nt], a, bh[ks][nt][0], bl[ks][nt][0]);
                        }
                    }
                    // epilogue: rows g / g+8 of the tile are (i0, j) and (i0+1, j); columns 2t, 2t+1 are samples
                    const int tg = tile0 + tl;
                    const int ip = tg / NJC, jc = tg - ip * NJC;
                    const int i0 = 2 * ip, j = 8 * jc + g;
                    float part[4 * NT];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const float2 dy = *reinterpret_cast<const float2*>(sDY + j * DYS + nt * 8 + 2 * t);
                        float gv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float ex, r;
                            const float a = acc[nt][e] * tanh_c;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(a));
                            const float dd = ex + 1.0f;
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dd));
                            gv[e] = fmaf(-2.0f, r, 1.0f);
                        }
                        part[nt * 4 + 0] = gv[0] * dy.x;
                        part[nt * 4 + 1] = gv[1] * dy.y;
                        part[nt * 4 + 2] = gv[2] * dy.x;
                        part[nt * 4 + 3] = gv[3] * dy.y;
                        if (training) {
                            const int bg = b0 + nt * 8 + 2 * t;
                            float* gp = p.save_G + (((int64_t)stage * B + bg) * Hd + i0) * C8 + j;
                            const int64_t sstride = (int64_t)Hd * C8;
                            if (bg < B) { gp[0] = gv[0]; if (i0 + 1 < Hd) gp[C8] = gv[2]; }
                            if (bg + 1 < B) { gp[sstride] = gv[1]; if (i0 + 1 < Hd) gp[sstride + C8] = gv[3]; }
                        }
                    }
                    reduce_scatter_g<4 * NT>(part, lane);
                    {
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
                                sPart[(tl * 2 + (c >> 1)) * NS + nt * 8 + 2 * t + (c & 1)] = part[e];
                            }
                        }
                    }
                }
            }
            __syncthreads();

            // ---- (3) control of the next stage (loads in flight during the exchange) + scatter of the stage derivative
            const int kn = (s == 3) ? k + 1 : k, sn = (s == 3) ? 0 : s + 1;
            if (have_next) ctl_issue(kn, sn);
            const int buf = stage & 1;
            for (int item = tid; item < 2 * nipl * NS; item += NTHR) {
                const int il = item / NS, n = item - il * NS;
                const int i = 2 * ip0 + il;
                if (i < Hd) {
                    const float* pp = sPart + (((il >> 1) * NJC) * 2 + (il & 1)) * NS + n;
                    float o = 0.f;
                    for (int jc = 0; jc < NJC; ++jc) o += pp[jc * 2 * NS];
                    const uint32_t a = sK_addr + (uint32_t)(((buf * Hd + i) * NS + n) * 4);
                    for (int r = 0; r < CS; ++r) st_cluster_f32(map_to_rank(a, (uint32_t)r), o);
                    const int bg = b0 + n;
                    if (p.k_out && bg < B) p.k_out[((int64_t)stage * B + bg) * Hd + i] = o;
                }
            }
            cluster_arrive();
            cluster_wait();
            if (have_next) ctl_finish(stage + 1);

            // ---- (4) RK4 recurrence (rk4_alt_step_func), next stage input into buffer 0
            float* act_next = (training && have_next) ? p.save_act + (int64_t)(stage + 1) * act_stage : nullptr;
            const float* kk = sK + buf * Hd * NS;
#pragma unroll
            for (int it = 0; it < V2_RK_ITEMS; ++it) {
                const int item = tid + it * NTHR;
                if (item < n_items) {
                    const int i = item / NS, n = item - i * NS;
                    const int bg = b0 + n;
                    const bool owner = bg < B && ((n >> 3) % CS) == rank;
                    const float o = kk[item];
                    const float y = ry[it];
                    float zin;
                    if (s == 0) {
                        rk1[it] = o;
                        zin = __fadd_rn(y, __fdiv_rn(__fmul_rn(dt, o), 3.0f));                         // y + dt*k1/3
                    } else if (s == 1) {
                        rk2[it] = o;
                        zin = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(o, __fdiv_rn(rk1[it], 3.0f))));     // y + dt*(k2 - k1/3)
                    } else if (s == 2) {
                        rk3[it] = o;
                        zin = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(rk1[it], rk2[it]), o)));  // y + dt*(k1-k2+k3)
                    } else {
                        const float sum = __fadd_rn(__fadd_rn(__fadd_rn(rk1[it], __fmul_rn(3.0f, rk2[it])), __fmul_rn(3.0f, rk3[it])), o);
                        const float y1 = __fadd_rn(y, __fmul_rn(sum, __fdiv_rn(dt, 8.0f)));
                        // emit every requested time in (t0, t1] (FixedGridODESolver.integrate + _linear_interp)
                        int jj = jout;
                        while (jj < lo.n_out) {
                            const float tj = __ldg(p.t_out + jj);
                            if (!(t1 >= tj)) break;
                            float v;
                            if (tj == t0) v = y;
                            else if (tj == t1) v = y1;
                            else v = __fadd_rn(y, __fmul_rn(__fdiv_rn(__fsub_rn(y1, y), __fsub_rn(t1, t0)), __fsub_rn(tj, t0)));
                            if (owner) p.z_out[((int64_t)jj * B + bg) * Hd + i] = v;
                            ++jj;
                        }
                        ry[it] = y1;
                        zin = y1;
                    }
                    bfrag_store(sBf, NT, i, n, zin);
                    if (act_next && owner) act_next[(int64_t)(lo.row_z + i) * B + bg] = zin;
                }
            }
            if (s == 3) {
                while (jout < lo.n_out && t1 >= __ldg(p.t_out + jout)) ++jout;
            }
            __syncthreads();
        }
    }
    cluster_arrive();     // nobody leaves while a peer may still address its shared memory
    cluster_wait();
}

// ---- launch -----------------------------------------------------------------------------------------------
// Cluster size: the smallest power of two whose slice of the output layer fits in shared memory next to the hidden
// layers; samples per cluster: 8 per CTA (so that every SM owns 8 samples), at most 32.
int choose_launch2(const V2Layout& lo, bool backward, V2Launch* la);
size_t bwd2_smem_bytes(const V2Layout& lo, const V2Launch& la);

int choose_launch2_fwd(const V2Layout& lo, V2Launch* la)
{
    const size_t max_smem = 227 * 1024;
    static const int cand[][2] = {{1, 1}, {2, 2}, {4, 4}, {8, 4}, {8, 2}, {8, 1}};
    for (const auto& c : cand) {
        const int cs = c[0];
        la->CS = cs;
        la->NT = c[1];
        la->NS = 8 * la->NT;
        la->nclusters = (lo.B + la->NS - 1) / la->NS;
        la->threads = V2_FWD_THREADS;
        const int base = lo.NIP / cs, rem = lo.NIP % cs;
        la->ip_begin[0] = 0;
        for (int r = 0; r < cs; ++r) la->ip_begin[r + 1] = la->ip_begin[r] + base + (r < rem ? 1 : 0);
        for (int r = cs; r < V2_MAX_CS; ++r) la->ip_begin[r + 1] = la->ip_begin[cs];
        if (lo.NIP < cs) continue;
        la->smem = fwd_smem(lo, *la).total;
        if (la->smem <= max_smem && lo.Hd * la->NS <= V2_RK_ITEMS * la->threads && lo.C * la->NS <= V2_CTL_ITEMS * la->threads)
            return ANCDE_OK;
    }
    set_error("ncde_forward: the output layer (%d x %d) does not fit the shared memory of a cluster of %d CTAs", lo.Hd * lo.C, lo.K,
              V2_MAX_CS);
    return ANCDE_EUNSUPPORTED;
}

template <int NT>
static int launch_fwd2_nt(const V2Args& a, cudaStream_t stream)
{
    const V2Launch& la = a.la;
    ANCDE_CUDA(cudaFuncSetAttribute(ncde2_fwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)la.smem));
    if (la.CS > 8) ANCDE_CUDA(cudaFuncSetAttribute(ncde2_fwd_kernel<NT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(la.nclusters * la.CS));
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, ncde2_fwd_kernel<NT>, a);
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
    int rc = choose_launch2_fwd(a.lo, &a.la);
    if (rc != ANCDE_OK) return rc;
    switch (a.la.NT) {
        case 4: return launch_fwd2_nt<4>(a, stream);
        case 2: return launch_fwd2_nt<2>(a, stream);
        default: return launch_fwd2_nt<1>(a, stream);
    }
}

}  // namespace ancde
