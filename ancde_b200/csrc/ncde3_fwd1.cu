// UMMA forward solve, single-CTA variant: networks with NARROW hidden layers (<= 32 units; the attention / bottom network).
//
// The cluster kernel (ncde3_fwd.cu) pays four tensor-core round trips per stage for the hidden layers (~1.5 k cycles each,
// whatever their size) and a DSMEM exchange of the stage derivative; the streamed FP32 engine (ncde_fwd.cu) runs such
// layers as a warp-synchronous chain (2.8 k cycles for four layers of 20 units) but spends 3.4 k cycles per stage on the
// output layer in FP32 (bound by the shared-memory return path) and 2.2 k on tanh / contraction / stores.  This kernel takes
// the best of both for the shapes where ALL of W_out fits one SM's tensor memory next to the accumulators (bottom network
// of the Sepsis shape: 35*36 rows x 21 -> 10 tiles; 320 + 160 of the 512 columns): one CTA integrates 8 samples,
//   1. hidden ReLU layers: FP32 warp chain (hidden_chain, 4 warps x 2 samples) -> last activation as the FP16 hi/lo
//      B operand [h_hi ; h_lo] (N = 16); meanwhile the other warps evaluate the control of the next stage and issue the
//      TMA row stores of the previous stage's tanh outputs;                       [vector_fields.py:205-216]
//   2. per 128-row tile of W_out (A operand resident in TENSOR MEMORY, FP16 hi and lo tiles) four tcgen05.mma
//      (W_hi + W_lo) x [h_hi ; h_lo] (M = 128, N = 16, K = 16), issued by one warp per tile; the warps read the accumulators
//      back (tcgen05.ld), add the two halves, tanh, and park the rows in shared memory [sample][i*C4 + j];
//                                                                                   [vector_fields.py:216-218]
//   3. a thread per (sample, field row) contracts its row with dU (128-bit shared loads) and advances the RK4 recurrence of
//      ITS state element in registers -- no exchange of the stage derivative at all.  [cdeint_module.py:61; rk4_alt_step_func]
// Three CTA barriers per stage.  Saved tensors: the streamed engine's layout (tiles of 8 samples).
#include <cuda_runtime.h>
#include <stdlib.h>

#include "fwd_hidden.cuh"
#include "mma_common.cuh"
#include "ncde3_common.cuh"
#include "umma.cuh"

namespace ancde {

__global__ void __launch_bounds__(V3_THREADS, 1) ncde3_fwd1_kernel(const __grid_constant__ V3Args p)
{
    const V3Layout& lo = p.lo;
    constexpr int NTHR = V3_THREADS, NW = NTHR / 32, TS = 8;
    constexpr int SPWC = 1;                             // samples per chain warp
    constexpr int CHAIN_WARPS = TS / SPWC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int wq = warp & 3, wg = warp >> 2;
    const int b0 = blockIdx.x * TS;
    const int Hd = lo.Hd, C = lo.C, B = lo.B, C4 = lo.C4, K = lo.K, HP = lo.c_HP;
    const bool training = p.save_act != nullptr;
    const bool top = lo.mode == ANCDE_MODE_TOP;
    const int my_tile = blockIdx.x;
    const bool save_mine = training && my_tile < lo.sv_ntiles;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const V3S1Smem sm = v3_s1_smem(lo);
    float* sHW = reinterpret_cast<float*>(smem_raw + sm.hw);
    int* sLp = reinterpret_cast<int*>(smem_raw + sm.lp);
    float* sInT = reinterpret_cast<float*>(smem_raw + sm.z);                   // stage input [sample][HP]
    float* sT0 = sInT + TS * HP;
    float* sT1 = sT0 + TS * HP;
    float* sLast = reinterpret_cast<float*>(smem_raw + sm.last);              // last hidden activation [unit][8]
    unsigned char* sActHi = smem_raw + sm.act;
    unsigned char* sActLo = sActHi + lo.SBOa;
    float* sDU = reinterpret_cast<float*>(smem_raw + sm.du);
    float* sG = reinterpret_cast<float*>(smem_raw + sm.sg);                    // staged tanh outputs [sample][sg_str]
    float* sTimes = reinterpret_cast<float*>(smem_raw + sm.times);
    float* sTout = reinterpret_cast<float*>(smem_raw + sm.tout);
    uint64_t* bar_t = reinterpret_cast<uint64_t*>(smem_raw + sm.bars);        // [V3_MAX_TILES] W_out tile t ready
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar_t + V3_MAX_TILES);
    // B operand, MN-major: the 8 samples of one feature k are 16 contiguous bytes
    auto act_off = [&](int n, int k) -> uint32_t { return (uint32_t)((k >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2); };

    // ---- prologue
    {
        const float4* src = reinterpret_cast<const float4*>(p.wpack);
        float4* dst = reinterpret_cast<float4*>(sHW);
        for (int i = tid; i < pad4(lo.c_hidden_floats) / 4; i += NTHR) dst[i] = __ldg(src + i);
        if (tid < lo.n_hidden) {
            int* q = sLp + 8 * tid;
            q[0] = (lo.in_dim[tid] + 3) / 4; q[1] = lo.width[tid]; q[2] = lo.c_stride[tid]; q[3] = lo.c_off_w[tid];
            q[4] = lo.c_off_b[tid]; q[5] = lo.sv_off_h[tid]; q[6] = 0; q[7] = 0;
        }
        for (int i = tid; i < 3 * TS * HP; i += NTHR) sInT[i] = 0.f;
        for (int i = tid; i < TS * lo.sg_str; i += NTHR) sG[i] = 0.f;         // padding channels stay zero
        for (int i = tid; i < 2 * TS * C4; i += NTHR) sDU[i] = 0.f;
        uint4* za = reinterpret_cast<uint4*>(sActHi);
        for (int i = tid; i < 2 * lo.SBOa / 16; i += NTHR) za[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < lo.L; i += NTHR) sTimes[i] = p.times[i];
        for (int i = tid; i < lo.n_out; i += NTHR) sTout[i] = p.t_out[i];
        if (tid == 0) {
            for (int t = 0; t < V3_MAX_TILES; ++t) mbar_init(bar_t + t, 1);
            mbar_fence_init();
        }
        if (warp == 0) tmem_alloc(tslot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const uint32_t tlane = tbase + ((uint32_t)(32 * wq) << 16);     // this warp's lane quarter
    {
        // W_out as the RESIDENT A operand in tensor memory: row m of a tile = TMEM lane m, 16 FP16 of the row (one k step) per
        // 8 columns; read from the canonical K-major image (two 16-byte core-matrix rows per k step)
        const int kst = lo.KPo >> 4, items = 2 * lo.mtiles * kst;
        for (int it = wg; it < items; it += 4) {
            const int mt = it / kst, ks = it - mt * kst;            // mt = matrix (hi, lo) * mtiles + tile
            const int r = 128 * (mt % lo.mtiles) + 32 * wq + lane;
            const unsigned char* q = p.img + (size_t)(mt / lo.mtiles) * lo.wo_mat_bytes + (size_t)(r >> 3) * lo.SBOo + (size_t)ks * 256 + (r & 7) * 16;
            const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(q)), x1 = __ldg(reinterpret_cast<const uint4*>(q + 128));
            const uint32_t rr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            tmem_st8(tlane + (uint32_t)(lo.a_col + mt * (lo.KPo >> 1) + 8 * ks), rr);
        }
        tmem_wait_st();
        // the bias column of the operand (k = K): ones in the hi half; never rewritten
        if (tid < TS) *reinterpret_cast<__half*>(sActHi + act_off(tid, K)) = __float2half_rn(1.0f);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    // ---- control: dU[n][j] of a stage (spline derivative, attention product), into `du`
    int cnt = 0;        // knots strictly below the current stage time (monotone)
    const int n_ctl = C4 * TS;
    const unsigned magicC4 = (unsigned)((((uint64_t)1 << 32) + C4 - 1) / C4);      // item / C4 == umulhi(item, magicC4) for item < 2^16
    auto control = [&](int k, int s, float* du, int ctid, int cthreads) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float ts = stage_time(t0, __fsub_rn(t1, t0), s);
        // index = clamp(#(times < t) - 1, 0, L-2): the LEFT piece at a knot (interpolate.py:261-267)
        while (cnt < lo.L && sTimes[cnt] < ts) ++cnt;
        int idx = cnt - 1;
        idx = idx < 0 ? 0 : (idx > lo.L - 2 ? lo.L - 2 : idx);
        const float frac = __fsub_rn(ts, sTimes[idx]);
        int q = (int)floorf(ts) - 1;                  // attention[int(floor(t)) - 1], python wrap
        if (q < 0) q += lo.att_L;
        q = q < 0 ? 0 : (q >= lo.att_L ? lo.att_L - 1 : q);
        const int stage = 4 * k + s;
        const int m = top ? stage % (lo.n_hp - 1) : 0;     // call counter + reset rule (cdeint_module.py:134-137)
        const float* pb = p.cb + idx * C;
        const float* pc = p.cc2 + idx * C;
        const float* pd = p.cd3 + idx * C;
        const float* pa = top ? p.attention + (int64_t)q * B * lo.att_C : nullptr;
        const float* ph = top ? p.h_prime + (int64_t)m * B * C : nullptr;
        const int rowlen = (lo.L - 1) * C;
        float* act = save_mine ? p.save_act + ((int64_t)stage * lo.sv_ntiles + my_tile) * lo.sv_act_floats : nullptr;
        for (int base = ctid; base < n_ctl; base += CTL_BATCH * cthreads) {
            float vb[CTL_BATCH], vc[CTL_BATCH], vd[CTL_BATCH], va[CTL_BATCH], vh[CTL_BATCH];
#pragma unroll
            for (int u = 0; u < CTL_BATCH; ++u) {
                const int item = base + u * cthreads;
                vb[u] = vc[u] = vd[u] = va[u] = vh[u] = 0.f;
                const int n = (int)__umulhi((unsigned)item, magicC4), j = item - n * C4;    // item = n*C4 + j: lanes walk the channels
                const int bg = b0 + n;
                if (item < n_ctl && j < C && bg < B) {
                    const int o = bg * rowlen + j;
                    vb[u] = __ldg(pb + o);
                    vc[u] = __ldg(pc + o);
                    vd[u] = __ldg(pd + o);
                    if (top) {
                        va[u] = __ldg(pa + bg * lo.att_C + (lo.att_C == 1 ? 0 : j));
                        vh[u] = __ldg(ph + bg * C + j);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < CTL_BATCH; ++u) {
                const int item = base + u * cthreads;
                const int n = (int)__umulhi((unsigned)item, magicC4), j = item - n * C4;
                if (item < n_ctl) {
                    // derivative = b + (two_c + three_d * frac) * frac   (interpolate.py:298-303)
                    const float dX = __fadd_rn(vb[u], __fmul_rn(__fadd_rn(vc[u], __fmul_rn(vd[u], frac)), frac));
                    float dU = dX, duda = 0.f;
                    if (top) {
                        const float a = va[u], hp = vh[u];
                        // dY = dX*a + ((a*(1-a))*dX)*h'   (cdeint_module.py:113-129; "Xt" is dX/dt)
                        dU = __fadd_rn(__fmul_rn(dX, a), __fmul_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.0f, a)), dX), hp));
                        duda = dX * (1.0f + (1.0f - 2.0f * a) * hp);
                    }
                    if (j < C) du[n * C4 + j] = dU;
                    if (act) {
                        const int di = (j >> 2) * lo.sv_DUS + (j & 3) * 8 + n;
                        act[lo.sv_off_du + di] = dU;
                        if (top) act[lo.sv_off_duda + di] = duda;
                    }
                }
            }
        }
    };

    // ---- RK4 state in registers: thread (sample c_n, field row c_il) owns ONE state element; it is also the thread that
    // contracts the tanh rows of that element with dU
    const int c_il = tid & ((1 << lo.nip_log2) - 1), c_n = tid >> lo.nip_log2;
    const bool c_ok = c_il < Hd && c_n < TS;
    const int c_bg = b0 + c_n;
    const float4* c_g4 = reinterpret_cast<const float4*>(sG + (c_ok ? c_n * lo.sg_str + c_il * C4 : 0));
    const int c_d = c_ok ? c_n * C4 : 0;
    float ry = 0.f, rk1 = 0.f, rk2 = 0.f, rk3 = 0.f;
    const int64_t act_stride = (int64_t)lo.sv_ntiles * lo.sv_act_floats;
    float* act_run = save_mine ? p.save_act + (int64_t)my_tile * lo.sv_act_floats : nullptr;
    // (addresses that only depend on the stage number are carried along, not rebuilt on the critical path)
    const int64_t kout_stride = (int64_t)B * Hd;
    float* kout_run = (p.k_out && c_ok && c_bg < B) ? p.k_out + (int64_t)c_bg * Hd + c_il : nullptr;
    float* const zin_dst = sInT + c_n * HP + c_il;
    const int act_idx = c_il * 8 + c_n;
    if (c_ok) {
        ry = (c_bg < B) ? __ldg(p.z0 + (int64_t)c_bg * Hd + c_il) : 0.f;
        if (c_bg < B) p.z_out[(int64_t)c_bg * Hd + c_il] = ry;                  // solution[0] = y0
        sInT[c_n * HP + c_il] = ry;
        if (act_run) act_run[c_il * 8 + c_n] = ry;
    }
    control(0, 0, sDU, tid, NTHR);
    fence_async_smem();
    __syncthreads();

    // ---- constants
    const uint32_t idesc = umma_idesc_f16(128, 2 * TS, true);
    const uint64_t dB = umma_smem_desc(smem_u32(sActHi), 128, (uint32_t)lo.SBOa);      // [act_hi ; act_lo]: 16 rows
    const float tanh_c = 2.885390081777927f * V3_WSCALE_INV;      // 2*log2(e) / weight scale
    const int nblk = Hd * lo.nfull + (lo.rem ? (Hd + lo.ipb - 1) / lo.ipb : 0);
    const int ksteps = lo.KPo >> 4;
    const int64_t g_rows = (int64_t)lo.sv_ntiles * 8;
    uint32_t ph_t = 0;
    int jout = 1;                     // next requested output time
    // this lane's rows in the tiles its warp reads back (tile wg + 4u, lane quarter wq): offset of the row in a staged sample
    // row (field row il, control channel j), -1 where the lane holds no row; computed once
    constexpr int E_MAX = V3_MAX_TILES / 4;
    int e_off[E_MAX], e_n = 0;
    unsigned e_valid = 0;
#pragma unroll
    for (int u = 0; u < E_MAX; ++u) {
        const int t = wg + 4 * u, blk = 4 * t + wq;
        e_off[u] = -1;
        if (t < lo.mtiles && blk < nblk) {
            e_n = u + 1;
            int il, j;
            bool valid;
            if (blk < Hd * lo.nfull) {
                il = blk / lo.nfull;
                j = 32 * (blk - il * lo.nfull) + lane;
                valid = true;
            } else {
                const int q = blk - Hd * lo.nfull, jj = lane & (lo.seg - 1);
                il = q * lo.ipb + (lane >> lo.seg_log2);
                j = 32 * lo.nfull + jj;
                valid = jj < lo.rem && il < Hd;
            }
            if (il < Hd && j < C4) e_off[u] = il * C4 + j;
            if (valid) e_valid |= 1u << u;
        }
    }
    PH2_DECL();

    for (int k = 0; k < lo.n_steps; ++k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float dt = __fsub_rn(t1, t0);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const int stage = 4 * k + s;
            const bool have_next = stage + 1 < lo.n_stages;
            float* act = act_run;                          // this stage's saved-activation block (or null)
            if (save_mine) act_run += act_stride;
            const float* du_cur = sDU + (stage & 1) * (TS * C4);
            float* du_next = sDU + ((stage + 1) & 1) * (TS * C4);
            PH2_T0();

            // ---- (1) hidden layers: FP32 warp chain; the last activation becomes the FP16 hi/lo B operand
            if (warp < CHAIN_WARPS) {
                hidden_chain<TS, SPWC>(sLp, lo.n_hidden, sHW, sInT, sT0, sT1, sLast, act, HP, warp, CHAIN_WARPS, lane);
                __syncwarp();
                const int s0 = SPWC * warp;
                for (int o = lane; o < K; o += 32) {
                    if (SPWC == 2) {
                        const float2 v = *reinterpret_cast<const float2*>(sLast + o * TS + s0);
                        uint32_t hh, hl;
                        split_pair(v.x, v.y, hh, hl);
                        *reinterpret_cast<uint32_t*>(sActHi + act_off(s0, o)) = hh;
                        *reinterpret_cast<uint32_t*>(sActLo + act_off(s0, o)) = hl;
                    } else {
                        __half hh, hl;
                        split_half(sLast[o * TS + s0], hh, hl);
                        *reinterpret_cast<__half*>(sActHi + act_off(s0, o)) = hh;
                        *reinterpret_cast<__half*>(sActLo + act_off(s0, o)) = hl;
                    }
                }
                fence_async_smem();
                PH2_MARK(0);
            } else {
                // meanwhile: the previous stage's staged tanh rows leave through one TMA row copy per sample, and the control of the
                // next stage is evaluated
                if (training && warp == NW - 1 && stage > 0) {
                    if (lane < TS && b0 + lane < g_rows)
                        bulk_s2g(p.save_G + ((int64_t)(stage - 1) * g_rows + b0 + lane) * lo.sv_NCP, sG + lane * lo.sg_str, (uint32_t)(Hd * C4 * 4));
                    bulk_commit();
                    bulk_wait_read0();
                }
                if (have_next) control(s == 3 ? k + 1 : k, s == 3 ? 0 : s + 1, du_next, tid - 32 * CHAIN_WARPS, NTHR - 32 * CHAIN_WARPS);
            }
            __syncthreads();
            PH2_MARK(1);

            // ---- (2) W_out x [h_hi ; h_lo] per 128-row tile, tanh, rows staged in shared memory
            for (int t = warp_u; t < lo.mtiles; t += NW) {
                if (elect_one()) {
                    tc_fence_after();
                    const uint32_t d = tbase + (uint32_t)(16 * t);
                    const uint32_t tAhi = tbase + (uint32_t)(lo.a_col + t * (lo.KPo >> 1));
                    const uint32_t tAlo = tAhi + (uint32_t)(lo.mtiles * (lo.KPo >> 1));
#pragma unroll 2
                    for (int ks = 0; ks < ksteps; ++ks) umma_f16_ts(d, tAhi + (uint32_t)(ks * 8), dB + (uint64_t)(ks * 16), idesc, ks != 0);
#pragma unroll 2
                    for (int ks = 0; ks < ksteps; ++ks) umma_f16_ts(d, tAlo + (uint32_t)(ks * 8), dB + (uint64_t)(ks * 16), idesc, true);
                    umma_commit(bar_t + t);
                }
                __syncwarp();
            }
            PH2_MARK(2);
            // This warp's tiles wg, wg + 4, ...: the tcgen05.ld of the next tile is in flight while the current one goes through tanh.
            // (Reading all tiles back behind ONE tcgen05.wait::ld was measured slower, 1.41 -> 1.51 ms: the first tile's tanh no
            // longer overlaps the products of the later ones.)
            if (e_n > 0) {
                uint32_t ra[16], rb[16];
                mbar_wait(bar_t + wg, ph_t);
                tc_fence_after();
                tmem_ld16_issue(tlane + (uint32_t)(16 * wg), ra);
#pragma unroll
                for (int u = 0; u < E_MAX; ++u) {
                    if (u >= e_n) break;
                    uint32_t (&cur)[16] = (u & 1) ? rb : ra;
                    uint32_t (&nxt)[16] = (u & 1) ? ra : rb;
                    tmem_ld16_wait(cur);
                    if (u + 1 < e_n) {
                        mbar_wait(bar_t + wg + 4 * (u + 1), ph_t);
                        tc_fence_after();
                        tmem_ld16_issue(tlane + (uint32_t)(16 * (wg + 4 * (u + 1))), nxt);
                    }
                    if (e_off[u] >= 0) {
                        float* gs = sG + e_off[u];
                        const bool valid = (e_valid >> u) & 1;
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            gs[e * lo.sg_str] = valid ? tanh_scaled3(__uint_as_float(cur[e]) + __uint_as_float(cur[8 + e]), tanh_c) : 0.f;
                    }
                }
            }
            ph_t ^= 1;
            PH2_MARK(3);
            tc_fence_before();
            fence_async_smem();
            __syncthreads();
            PH2_MARK(4);

            // ---- (3) contraction with dU from the staged rows, RK4 recurrence of this thread's state element
            if (c_ok) {
                const float4* d4 = reinterpret_cast<const float4*>(du_cur + c_d);
                float o0 = 0.f, o1 = 0.f;
#pragma unroll 3
                for (int q = 0; q < (C4 >> 2); ++q) {
                    const float4 a = c_g4[q], b = d4[q];
                    o0 = fmaf(a.x, b.x, o0); o1 = fmaf(a.y, b.y, o1);
                    o0 = fmaf(a.z, b.z, o0); o1 = fmaf(a.w, b.w, o1);
                }
                const float o = o0 + o1;
                if (kout_run) { *kout_run = o; kout_run += kout_stride; }
                const float third = 1.0f / 3.0f;
                float zin;
                if (s == 0) { rk1 = o; zin = fmaf(dt * third, o, ry); }                        // y + dt*k1/3
                else if (s == 1) { rk2 = o; zin = fmaf(dt, o - rk1 * third, ry); }             // y + dt*(k2 - k1/3)
                else if (s == 2) { rk3 = o; zin = fmaf(dt, (rk1 - rk2) + o, ry); }             // y + dt*(k1-k2+k3)
                else {
                    const float y1 = fmaf(((rk1 + 3.0f * rk2) + 3.0f * rk3) + o, dt * 0.125f, ry);
                    // emit every requested time in (t0, t1] (FixedGridODESolver.integrate + _linear_interp)
                    int jj = jout;
                    while (jj < lo.n_out) {
                        const float tj = sTout[jj];
                        if (!(t1 >= tj)) break;
                        float vv;
                        if (tj == t0) vv = ry;
                        else if (tj == t1) vv = y1;
                        else vv = ry + (y1 - ry) / (t1 - t0) * (tj - t0);
                        if (c_bg < B) p.z_out[((int64_t)jj * B + c_bg) * Hd + c_il] = vv;
                        ++jj;
                    }
                    ry = y1;
                    zin = y1;
                }
                *zin_dst = zin;
                if (have_next && act_run) act_run[act_idx] = zin;
            }
            if (s == 3) {
                while (jout < lo.n_out && t1 >= sTout[jout]) ++jout;
            }
            PH2_MARK(5);
            __syncthreads();
            PH2_MARK(6);
        }
    }
    PH2_FLUSH(16 + (top ? 8 : 0));
    // the last stage's tanh rows
    if (training && warp == NW - 1) {
        if (lane < TS && b0 + lane < g_rows)
            bulk_s2g(p.save_G + ((int64_t)(lo.n_stages - 1) * g_rows + b0 + lane) * lo.sv_NCP, sG + lane * lo.sg_str, (uint32_t)(Hd * C4 * 4));
        bulk_commit();
        bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int launch_fwd3_single(const V3Args& a, cudaStream_t stream)
{
    const V3Layout& lo = a.lo;
    const size_t smem = v3_s1_smem(lo).total;
    ANCDE_CUDA(cudaFuncSetAttribute(ncde3_fwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const char* names[3] = {"ncde_fwd_plain", "ncde_fwd_bottom", "ncde_fwd_top"};
    prof_begin(names[lo.mode], stream);
    ncde3_fwd1_kernel<<<lo.nclusters, V3_THREADS, smem, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    prof_end(stream);
    count_launch();
    if (e != cudaSuccess) {
        set_error("ncde_forward: launch failed: %s (%zu bytes of shared memory)", cudaGetErrorString(e), smem);
        return ANCDE_ECUDA;
    }
    return ANCDE_OK;
}

}  // namespace ancde
