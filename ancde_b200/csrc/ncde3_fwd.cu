// UMMA fused NCDE forward solve (fixed-grid RK4, 3/8 rule) on the tcgen05 tensor cores: one thread-block cluster
// integrates NS = 8*CS samples over all time steps.  See ncde3_common.cuh for the decomposition.  Per RK4 stage a CTA
//   1. runs the hidden ReLU layers for all NS samples: one thread issues tcgen05.mma (A = 64*W_l hi/lo from shared
//      memory, B = activations hi/lo, D in tensor memory), two to eight warps read D back (tcgen05.ld), apply ReLU,
//      split into FP16 hi/lo and overwrite the activation operand in place;                       [vector_fields.py:205-216,237-247]
//      meanwhile the other warps evaluate the control of the NEXT stage (spline derivative, attention product);
//                                                                                                  [cdeint_module.py:58-60,103-129; interpolate.py:261-303]
//   2. multiplies its resident slice of the output layer with the last hidden activation (tcgen05.mma, one commit per
//      128-row tile so that the epilogue of tile t overlaps the products of tile t+1), applies tanh to the
//      accumulators, multiplies with dU and reduces over the control channel with a shuffle butterfly;
//                                                                                                  [vector_fields.py:216-218; cdeint_module.py:61,133]
//   3. scatters its rows of the stage derivative to every CTA of the cluster (distributed shared memory) and
//   4. advances the RK4 recurrence for all NS samples (state in registers) and writes the next stage input as the
//      B operand of the first hidden layer.                                                        [torchdiffeq rk4_alt_step_func]
// In training mode the stage inputs, hidden activations, controls and tanh outputs are saved in the streamed engine's
// layout (tiles of 8 samples), so the backward kernels of ncde_bwd.cu / ncde_wgrad_mma.cu run unchanged.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "mma_common.cuh"
#include "ncde3_common.cuh"
#include "umma.cuh"

namespace ancde {


// TSM: the W_out slice lives in tensor memory, the tanh outputs of a stage are staged in shared memory [NS][sg_str] (one TMA
// row copy per sample writes them out in training mode) and the contraction with dU reads them back from there.
template <bool TSM>
__global__ void __launch_bounds__(V3_THREADS, 1) ncde3_fwd_kernel(const __grid_constant__ V3Args p)
{
    const V3Layout& lo = p.lo;
    constexpr int NTHR = V3_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // the same value, provably warp-uniform (uniform datapath for the MMA issue)
    const int wq = warp & 3, wg = warp >> 2;          // TMEM lane quarter this warp may read; warp group
    const int CS = lo.CS, NS = lo.NS, NSL = lo.ns_log2;
    const int rank = (int)cluster_ctarank();
    const int cl = blockIdx.x / CS;
    const int b0 = cl * NS;
    const int Hd = lo.Hd, C = lo.C, B = lo.B;
    const int i0 = lo.i_begin[rank], ni = lo.i_begin[rank + 1] - i0;
    const bool training = p.save_act != nullptr;
    const bool top = lo.mode == ANCDE_MODE_TOP;
    const int my_tile = cl * CS + rank;               // tile of 8 samples this CTA writes the saved tensors of
    const bool save_mine = training && my_tile < lo.sv_ntiles;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const V3Smem sm = v3_smem(lo);
    unsigned char* sImg = smem_raw + sm.img;
    unsigned char* sActHi = smem_raw + sm.act;
    unsigned char* sActLo = sActHi + (size_t)(NS / 8) * lo.SBOa;
    float* sDU = reinterpret_cast<float*>(smem_raw + sm.du);
    float* sK = reinterpret_cast<float*>(smem_raw + sm.k);
    float* sKpart = reinterpret_cast<float*>(smem_raw + sm.kpart);
    float* sTimes = reinterpret_cast<float*>(smem_raw + sm.times);
    float* sTout = reinterpret_cast<float*>(smem_raw + sm.tout);    // requested output times (a global load per look-up cost ~600 cycles)
    uint64_t* bar_h = reinterpret_cast<uint64_t*>(smem_raw + sm.bars);          // hidden layers' accumulator ready
    uint64_t* bar_t = bar_h + 1;                                                // [V3_MAX_TILES] W_out tile t ready
    uint64_t* bar_k = bar_t + V3_MAX_TILES;                                     // [2] stage derivative of buffer b complete (transaction bytes)
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar_k + 2);
    float* sG = reinterpret_cast<float*>(smem_raw + sm.sg);          // staged tanh outputs [NS][sg_str] (lo.g_stage)
    const int SBOa = lo.SBOa;
    const int KSTR = lo.NS + 4;                       // row stride of the stage-derivative buffers (conflict-free 128-bit reads)
    // activation operand (B of every product): MN-major, i.e. 8 consecutive SAMPLES of one feature k are contiguous, so
    // that a thread that owns one feature of 8 samples (accumulator row) writes them with one 128-bit store
    auto act_off = [&](int n, int k) -> uint32_t { return (uint32_t)((n >> 3) * SBOa + (k >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2); };

    // ---- prologue: resident operand image, zeroed buffers, barriers, tensor memory
    {
        const uint4* srcH = reinterpret_cast<const uint4*>(p.img);
        uint4* dst = reinterpret_cast<uint4*>(sImg);
        for (int i = tid; i < lo.hidden_bytes / 16; i += NTHR) dst[i] = __ldg(srcH + i);
        if constexpr (!TSM) {
            const uint4* srcW = reinterpret_cast<const uint4*>(p.img + lo.hidden_bytes + (size_t)rank * 2 * lo.wo_mat_bytes);
            uint4* dstW = reinterpret_cast<uint4*>(sImg + lo.hidden_bytes);
            for (int i = tid; i < 2 * lo.wo_mat_bytes / 16; i += NTHR) dstW[i] = __ldg(srcW + i);
        }
        if constexpr (TSM)      // padding channels of the staged rows are never written again: they stay zero
            for (int i = tid; i < NS * lo.sg_str; i += NTHR) sG[i] = 0.f;
        uint4* za = reinterpret_cast<uint4*>(sActHi);
        for (int i = tid; i < 2 * (NS / 8) * SBOa / 16; i += NTHR) za[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < 2 * lo.du_floats; i += NTHR) sDU[i] = 0.f;
        for (int i = tid; i < 2 * Hd * KSTR; i += NTHR) sK[i] = 0.f;
        for (int i = tid; i < lo.L; i += NTHR) sTimes[i] = p.times[i];
        for (int i = tid; i < lo.n_out; i += NTHR) sTout[i] = p.t_out[i];
        if (tid == 0) {
            mbar_init(bar_h, 1);
            for (int t = 0; t < V3_MAX_TILES; ++t) mbar_init(bar_t + t, 1);
            mbar_init(bar_k + 0, 1);
            mbar_init(bar_k + 1, 1);
            mbar_fence_init();
        }
        if (warp == 0) tmem_alloc(tslot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const uint32_t tlane = tbase + ((uint32_t)(32 * wq) << 16);     // this warp's lane quarter
    if constexpr (TSM) {
        // the W_out slice as the RESIDENT A operand in tensor memory: row m of a tile = TMEM lane m, 16 FP16 of the row (one k
        // step) per 8 columns; read from the canonical K-major image (two 16-byte core-matrix rows per k step)
        const int kst = lo.KPo >> 4, items = 2 * lo.mtiles * kst;
        const unsigned char* srcW = p.img + lo.hidden_bytes + (size_t)rank * 2 * lo.wo_mat_bytes;
        for (int it = wg; it < items; it += 4) {
            const int mt = it / kst, ks = it - mt * kst;            // mt = matrix (hi, lo) * mtiles + tile
            const int r = 128 * (mt % lo.mtiles) + 32 * wq + lane;
            const unsigned char* q = srcW + (size_t)(mt / lo.mtiles) * lo.wo_mat_bytes + (size_t)(r >> 3) * lo.SBOo + (size_t)ks * 256 + (r & 7) * 16;
            const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(q)), x1 = __ldg(reinterpret_cast<const uint4*>(q + 128));
            const uint32_t rr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            tmem_st8(tlane + (uint32_t)(lo.a_col + mt * (lo.KPo >> 1) + 8 * ks), rr);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    // ---- control: dU[j][n] of a stage (spline derivative, attention product) for all NS samples, into `du`
    int cnt = 0;        // knots strictly below the current stage time (monotone)
    const int C4 = training ? lo.sv_C4 : C;
    const int n_ctl = C4 << NSL;
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
        // all loads of a batch are issued before the first use: one memory round trip per CTL_BATCH items of a thread
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
                    if (j < C) du[TSM ? n * lo.DUSTR + j : j * lo.DUSTR + n] = dU;
                    if (act && (n >> 3) == rank) {
                        const int di = (j >> 2) * lo.sv_DUS + (j & 3) * 8 + (n & 7);
                        act[lo.sv_off_du + di] = dU;
                        if (top) act[lo.sv_off_duda + di] = duda;
                    }
                }
            }
        }
    };

    // ---- RK4 state in registers: thread tid < Hd*CS owns field row si = tid % Hd of the 8 samples of chunk sc = tid / Hd
    float ry[8], rk1[8], rk2[8], rk3[8];
    const int sc = tid / Hd, si = tid - sc * Hd;
    const bool state_thr = tid < Hd * CS;
    const uint32_t state_off = act_off(8 * sc, si);            // one 128-bit store = the 8 samples of feature si
    auto store_state = [&](const float (&z)[8]) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_pair(z[2 * e], z[2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4*>(sActHi + state_off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(sActLo + state_off) = make_uint4(l[0], l[1], l[2], l[3]);
    };
    {
        float* act0 = save_mine ? p.save_act + (int64_t)my_tile * lo.sv_act_floats : nullptr;
#pragma unroll
        for (int e = 0; e < 8; ++e) ry[e] = rk1[e] = rk2[e] = rk3[e] = 0.f;
        if (state_thr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int bg = b0 + 8 * sc + e;
                ry[e] = (bg < B) ? __ldg(p.z0 + (int64_t)bg * Hd + si) : 0.f;
                if (sc == rank && bg < B) p.z_out[(int64_t)bg * Hd + si] = ry[e];          // solution[0] = y0
            }
            store_state(ry);
            if (act0 && sc == rank) {
                reinterpret_cast<float4*>(act0 + si * 8)[0] = make_float4(ry[0], ry[1], ry[2], ry[3]);
                reinterpret_cast<float4*>(act0 + si * 8)[1] = make_float4(ry[4], ry[5], ry[6], ry[7]);
            }
        }
    }
    // the columns behind the state: a one for the bias column, zeros up to the padded k of the first layer
    auto write_const_cols = [&]() {
        const int ncol = lo.KPh[0] - Hd;
        for (int item = tid; item < (ncol << NSL); item += NTHR) {
            const int k = Hd + (item >> NSL), n = item & (NS - 1);
            *reinterpret_cast<__half*>(sActHi + act_off(n, k)) = __float2half_rn(k == Hd ? 1.0f : 0.0f);
            *reinterpret_cast<__half*>(sActLo + act_off(n, k)) = __float2half_rn(0.0f);
        }
    };
    write_const_cols();
    // the hidden layers overwrite the operand in place; their last write leaves exactly this pattern behind when the last
    // hidden width equals the state size -- then the constant columns never need rewriting
    const bool const_cols_stable = lo.width[lo.n_hidden - 1] == Hd && lo.KPo >= lo.KPh[0];
    control(0, 0, sDU, tid, NTHR);

    uint32_t sK_rank[V3_MAX_CS];
#pragma unroll
    for (int r = 0; r < V3_MAX_CS; ++r) sK_rank[r] = map_to_rank(smem_u32(sK), (uint32_t)(r < CS ? r : 0));
    // the peers' copies of bar_k sit at the same distance from their sK
    const uint32_t bark_delta = smem_u32(bar_k) - smem_u32(sK);

    fence_async_smem();
    cluster_arrive();     // every CTA of the cluster has initialised its shared memory before anyone writes into it
    cluster_wait();

    // ---- constants of the epilogues
    const uint32_t idesc = umma_idesc_f16(128, NS, true), idesc2 = umma_idesc_f16(128, 2 * NS, true);
    const uint32_t aHi = smem_u32(sActHi), aLo = smem_u32(sActLo);
    const uint32_t wHi = smem_u32(sImg) + lo.hidden_bytes, wLo = wHi + lo.wo_mat_bytes;
    // operand descriptors: the activations (B), and the W_out tile this warp's lane 0 issues the products of
    const uint64_t dBhi = umma_smem_desc(aHi, 128, SBOa), dBlo = umma_smem_desc(aLo, 128, SBOa);
    const uint64_t dWhi = umma_smem_desc(wHi + (warp_u < lo.mtiles ? warp_u : 0) * 16 * lo.SBOo, 128, lo.SBOo);
    const uint64_t dWlo = umma_smem_desc(wLo + (warp_u < lo.mtiles ? warp_u : 0) * 16 * lo.SBOo, 128, lo.SBOo);
    const float tanh_c = 2.885390081777927f * V3_WSCALE_INV;      // 2*log2(e) / weight scale
    const int nslab = NS >> 4;                                    // 16-column slabs of an accumulator tile
    const int nblk = ni * lo.nfull + (lo.rem ? (ni + lo.ipb - 1) / lo.ipb : 0);
    const int nbase4 = ((lane & 1) << 3) | ((lane & 2) << 1) | ((lane & 4) >> 1) | ((lane & 8) >> 3);   // sample a lane ends up with after 4 butterfly rounds
    int hk = lo.KPh[0] >> 4;          // k steps and A descriptors of the next hidden-layer product
    uint64_t hHi = umma_smem_desc(smem_u32(sImg) + lo.off_hhi[0], 128, (uint32_t)lo.SBOh[0]);
    uint64_t hLo = umma_smem_desc(smem_u32(sImg) + lo.off_hlo[0], 128, (uint32_t)lo.SBOh[0]);
    const bool hidden_warp = wq < 2 && wg < nslab;                // warps that read the hidden layers' accumulator (warp 0 issues)
    uint32_t ph_h = 0, ph_t = 0;      // phase parities of bar_h / bar_t
    // warps that neither read the hidden layers' accumulator, nor evaluate the control, nor hold RK4 state: they issue the TMA
    // stores of the staged tanh outputs (every warp, when there is no such warp)
    int tma_idx = -1, n_tma = 0;
    for (int w = 0; w < NTHR / 32; ++w)
        if ((w & 3) < 2 && (w >> 2) >= nslab && w * 32 >= Hd * CS) {
            if (w == warp) tma_idx = n_tma;
            ++n_tma;
        }
    if (n_tma == 0) { tma_idx = warp; n_tma = NTHR / 32; }
    int jout = 1;                     // next requested output time
    // addresses that only depend on the stage number are carried along instead of being rebuilt on the critical path (every
    // phase of a stage is a chain of dependent instructions of ONE warp at ~4 cycles each)
    const int64_t act_stride = (int64_t)lo.sv_ntiles * lo.sv_act_floats;
    float* act_run = save_mine ? p.save_act + (int64_t)my_tile * lo.sv_act_floats : nullptr;
    const float* kk0 = sK + si * KSTR + 8 * sc;
    // first (usually only) item of this thread in the contraction: field row, sample, offsets
    const int c_il = (tid >> 2) & ((1 << lo.nip_log2) - 1), c_n = ((tid >> (2 + lo.nip_log2)) << 2) | (tid & 3);
    const bool c_ok = c_il < ni && c_n < NS;
    const int c_g = c_ok ? c_n * lo.sg_str + c_il * lo.C4 : 0, c_d = c_ok ? c_n * lo.DUSTR : 0;
    const uint32_t c_off = (uint32_t)((((i0 + c_il) * KSTR) + c_n) * 4);
    PH2_DECL();
#ifdef ANCDE_PHASE_TIMING
    long long phx_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, phx_t = 0;
#define PHX_T0() phx_t = clock64()
#define PHX_MARK(i) do { const long long now__ = clock64(); phx_acc[i] += now__ - phx_t; phx_t = now__; } while (0)
#else
#define PHX_T0() do { } while (0)
#define PHX_MARK(i) do { } while (0)
#endif

    for (int k = 0; k < lo.n_steps; ++k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float dt = __fsub_rn(t1, t0);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const int stage = 4 * k + s;
            const bool have_next = stage + 1 < lo.n_stages;
            float* act = act_run;                          // this stage's saved-activation block of my tile (or null)
            if (save_mine) act_run += act_stride;
            const float* du_cur = sDU + (stage & 1) * lo.du_floats;
            float* du_next = sDU + ((stage + 1) & 1) * lo.du_floats;
            PH2_T0();

            // ---- (1) hidden layers on the tensor cores; the activation operand is overwritten in place
#pragma unroll 1
            for (int l = 0; l < lo.n_hidden; ++l) {
                if (!hidden_warp) break;                  // the other warps: control of the next stage, below
                if (warp_u == 0) {
                    if (elect_one()) {
                        tc_fence_after();
                        const uint32_t d = tbase + lo.tcol_h;
                        // W_hi x [act_hi ; act_lo] (the two activation matrices are adjacent: one B operand of 2*NS rows) fills
                        // columns [0, NS) and [NS, 2*NS); then W_lo x act_hi is added to the first NS columns.
                        // One k step = two core matrices = 256 bytes = 16 descriptor units.
#pragma unroll 4
                        for (int ks = 0; ks < hk; ++ks) umma_f16(d, hHi + (uint64_t)(ks * 16), dBhi + (uint64_t)(ks * 16), idesc2, ks != 0);
#pragma unroll 4
                        for (int ks = 0; ks < hk; ++ks) umma_f16(d, hLo + (uint64_t)(ks * 16), dBhi + (uint64_t)(ks * 16), idesc, true);
                        umma_commit(bar_h);
                    }
                    // the descriptors of the next product (next layer, or layer 0 of the next stage) while this one runs
                    const int ln = (l + 1 < lo.n_hidden) ? l + 1 : 0;
                    hk = lo.KPh[ln] >> 4;
                    hHi = umma_smem_desc(smem_u32(sImg) + lo.off_hhi[ln], 128, (uint32_t)lo.SBOh[ln]);
                    hLo = umma_smem_desc(smem_u32(sImg) + lo.off_hlo[ln], 128, (uint32_t)lo.SBOh[ln]);
                }
                __syncwarp();
                PH2_MARK(0);
                const int KPn = (l + 1 < lo.n_hidden) ? lo.KPh[l + 1] : lo.KPo;      // k extent the consumer reads
                if (32 * wq < KPn) {
                    mbar_spin(bar_h, ph_h);
                    tc_fence_after();
                    PH2_MARK(1);
                    const int m = 32 * wq + lane, n0 = 16 * wg, width = lo.width[l];
                    float v[16], v2[16];
                    tmem_ld16x2(tlane + lo.tcol_h + n0, tlane + lo.tcol_h + NS + n0, v, v2);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] += v2[e];
                    if (m < KPn) {
                        uint32_t hh[8], hl[8];
#pragma unroll
                        for (int e = 0; e < 16; e += 2) {
                            // rows behind the layer's units: the bias column of the next layer, then zeros
                            v[e] = (m < width) ? fmaxf(v[e] * V3_WSCALE_INV, 0.f) : (m == width ? 1.0f : 0.0f);
                            v[e + 1] = (m < width) ? fmaxf(v[e + 1] * V3_WSCALE_INV, 0.f) : (m == width ? 1.0f : 0.0f);
                            split_pair(v[e], v[e + 1], hh[e >> 1], hl[e >> 1]);
                        }
                        const uint32_t o = act_off(n0, m);
                        *reinterpret_cast<uint4*>(sActHi + o) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                        *reinterpret_cast<uint4*>(sActHi + o + SBOa) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
                        *reinterpret_cast<uint4*>(sActLo + o) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
                        *reinterpret_cast<uint4*>(sActLo + o + SBOa) = make_uint4(hl[4], hl[5], hl[6], hl[7]);
                        if (act && m < width && (rank >> 1) == wg) {
                            float4* dst = reinterpret_cast<float4*>(act + lo.sv_off_h[l] + m * 8);
                            const bool up = rank & 1;
                            dst[0] = make_float4(up ? v[8] : v[0], up ? v[9] : v[1], up ? v[10] : v[2], up ? v[11] : v[3]);
                            dst[1] = make_float4(up ? v[12] : v[4], up ? v[13] : v[5], up ? v[14] : v[6], up ? v[15] : v[7]);
                        }
                    }
                    PH2_MARK(2);
                    fence_async_smem();
                    tc_fence_before();
                }
                ph_h ^= 1;
                if (l + 1 < lo.n_hidden) named_bar_sync(1, 64 * nslab);      // only the warps of the hidden layers
                PH2_MARK(3);
            }
            if (!hidden_warp && (lo.n_hidden & 1)) ph_h ^= 1;      // (keeps the parity in step; these warps never wait on bar_h)
#ifdef ANCDE_PHASE_TIMING
            const long long ctl_t0 = clock64();
#endif
            if (!hidden_warp && wq >= 2 && have_next)     // meanwhile: control of the next stage
                control(s == 3 ? k + 1 : k, s == 3 ? 0 : s + 1, du_next, ((wg << 1) + (wq - 2)) * 32 + lane, NTHR / 2);
#ifdef ANCDE_PHASE_TIMING
            if (tid == 64 && blockIdx.x == 0 && p.dbg) p.dbg[top ? 1 : 0] += clock64() - ctl_t0;
#endif
            if (TSM && training && tma_idx >= 0) bulk_wait_read0();     // the staged rows of the previous stage have left
            __syncthreads();

            // ---- (2) output layer slice x all samples (one commit per 128-row tile), tanh, contraction with dU
            if (warp_u < lo.mtiles) {                  // one issuing warp (one elected lane) per 128-row tile
                if (elect_one()) {
                    tc_fence_after();
                    const int ksteps = lo.KPo >> 4;
                    const uint32_t d = tbase + warp_u * 2 * NS;
                    if constexpr (TSM) {
                        const uint32_t tAhi = tbase + (uint32_t)(lo.a_col + warp_u * (lo.KPo >> 1));
                        const uint32_t tAlo = tAhi + (uint32_t)(lo.mtiles * (lo.KPo >> 1));
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ++ks) umma_f16_ts(d, tAhi + (uint32_t)(ks * 8), dBhi + (uint64_t)(ks * 16), idesc2, ks != 0);
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ++ks) umma_f16_ts(d, tAlo + (uint32_t)(ks * 8), dBhi + (uint64_t)(ks * 16), idesc, true);
                    } else {
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ++ks) umma_f16(d, dWhi + (uint64_t)(ks * 16), dBhi + (uint64_t)(ks * 16), idesc2, ks != 0);
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ++ks) umma_f16(d, dWlo + (uint64_t)(ks * 16), dBhi + (uint64_t)(ks * 16), idesc, true);
                    }
                    umma_commit(bar_t + warp_u);
                }
            }
            __syncwarp();
            PH2_MARK(4);
            for (int t = wg; t < lo.mtiles; t += 4) {
                const int blk = 4 * t + wq;
                if (blk >= nblk) continue;
                // identity of this lane's row: field row i0 + il, control channel j
                int il, j, chunk, R;
                bool valid;
                if (blk < ni * lo.nfull) {
                    il = blk / lo.nfull;
                    chunk = blk - il * lo.nfull;
                    j = 32 * chunk + lane;
                    valid = true;
                    R = 5;
                } else {
                    const int q = blk - ni * lo.nfull, jj = lane & (lo.seg - 1);
                    il = q * lo.ipb + (lane >> lo.seg_log2);
                    j = 32 * lo.nfull + jj;
                    valid = jj < lo.rem && il < ni;
                    chunk = lo.nfull;
                    R = lo.seg_log2;
                }
                const int jc = valid ? j : 0;
                const bool vec_g = blk < ni * lo.nfull || lo.seg >= 4;      // 4 consecutive lanes = 4 consecutive channels of one field row
                const int nb = nbase4 & ~((16 >> (R < 4 ? R : 4)) - 1);       // bits fixed by the rounds actually run
                float* kp = sKpart + ((il < ni ? il : 0) * lo.nchunks + chunk) * NS;
                const bool writer = il < ni && (R < 5 || lane < 16);
                const int cntv = R >= 4 ? 1 : (16 >> R);
                PHX_T0();
                mbar_spin(bar_t + t, ph_t);
                tc_fence_after();
                PHX_MARK(0);
                if constexpr (TSM) {
                    // tanh of the accumulator rows into the staged rows: lanes with consecutive channels write consecutive words
                    const bool st_ok = il < ni && j < lo.C4;
                    float* gs = sG + il * lo.C4 + j;
                    for (int sl = 0; sl < nslab; ++sl) {
                        const int n0 = 16 * sl;
                        float v[16], v2[16];
                        tmem_ld16x2(tlane + t * 2 * NS + n0, tlane + t * 2 * NS + NS + n0, v, v2);
                        PHX_MARK(1);
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = valid ? tanh_scaled3(v[e] + v2[e], tanh_c) : 0.f;
                        PHX_MARK(2);
                        if (st_ok) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) gs[(n0 + e) * lo.sg_str] = v[e];
                        }
                        PHX_MARK(3);
                    }
                    continue;
                }
                for (int sl = 0; sl < nslab; ++sl) {
                    const int n0 = 16 * sl;
                    float v[16];
                    {
                        float v2[16];
                        tmem_ld16x2(tlane + t * 2 * NS + n0, tlane + t * 2 * NS + NS + n0, v, v2);
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] += v2[e];
                    }
                    PHX_MARK(1);
                    const float4* dup = reinterpret_cast<const float4*>(du_cur + jc * lo.DUSTR + n0);
                    float du[16];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const float4 x = dup[q4];
                        du[4 * q4] = x.x; du[4 * q4 + 1] = x.y; du[4 * q4 + 2] = x.z; du[4 * q4 + 3] = x.w;
                    }
                    float g[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        g[e] = tanh_scaled3(v[e], tanh_c);
                        v[e] = valid ? g[e] * du[e] : 0.f;
                    }
                    PHX_MARK(2);
                    if (training && vec_g) {
                        // tanh outputs in the streamed engine's layout [stage][sample][i*C4 + j]: 4x4 transposes over the 4 lanes
                        // that own 4 consecutive channels, so that a lane stores 4 channels of one sample (128-bit stores, 4
                        // samples x 128 contiguous bytes per warp instruction; scalar stores cost ~4x the store-issue time)
                        const int r4 = lane & 3;
                        const int64_t rows = (int64_t)lo.sv_ntiles * 8;
                        const int j0 = j & ~3;
                        const bool st_ok = il < ni && j0 < lo.sv_C4;
                        float* gp = p.save_G + ((int64_t)stage * rows + b0 + n0 + r4) * lo.sv_NCP + (i0 + il) * lo.sv_C4 + j0;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            float a0 = valid ? g[4 * q4] : 0.f, a1 = valid ? g[4 * q4 + 1] : 0.f, a2 = valid ? g[4 * q4 + 2] : 0.f,
                                  a3 = valid ? g[4 * q4 + 3] : 0.f;
                            {
                                const bool up = r4 & 1;
                                float t0 = up ? a0 : a1, t1 = up ? a2 : a3;
                                t0 = __shfl_xor_sync(0xffffffffu, t0, 1);
                                t1 = __shfl_xor_sync(0xffffffffu, t1, 1);
                                if (up) { a0 = t0; a2 = t1; } else { a1 = t0; a3 = t1; }
                            }
                            {
                                const bool up = r4 & 2;
                                float t0 = up ? a0 : a2, t1 = up ? a1 : a3;
                                t0 = __shfl_xor_sync(0xffffffffu, t0, 2);
                                t1 = __shfl_xor_sync(0xffffffffu, t1, 2);
                                if (up) { a0 = t0; a1 = t1; } else { a2 = t0; a3 = t1; }
                            }
                            if (st_ok && b0 + n0 + 4 * q4 + r4 < rows)
                                *reinterpret_cast<float4*>(gp + (int64_t)(4 * q4) * lo.sv_NCP) = make_float4(a0, a1, a2, a3);
                        }
                    } else if (training && valid) {
                        // (partial chunks narrower than 4 lanes: scalar stores)
                        const int64_t rows = (int64_t)lo.sv_ntiles * 8;
                        float* gp = p.save_G + ((int64_t)stage * rows + b0 + n0) * lo.sv_NCP + (i0 + il) * lo.sv_C4 + j;
                        const int nrow = (int)((rows - b0 - n0 < 16) ? rows - b0 - n0 : 16);
#pragma unroll
                        for (int e = 0; e < 16; ++e) { if (e < nrow) *gp = g[e]; gp += lo.sv_NCP; }
                        if (j == C - 1 && lo.sv_C4 > C) {          // the padding channels of the row: zeros
                            gp -= (int64_t)16 * lo.sv_NCP;
                            for (int e = 0; e < nrow; ++e)
                                for (int jp = 1; jp <= lo.sv_C4 - C; ++jp) gp[(int64_t)e * lo.sv_NCP + jp] = 0.f;
                        }
                    }
                    PHX_MARK(3);
                    // reduce over the lanes of a segment, halving the live values each round
                    if (R >= 1) {
                        const bool up = lane & 1;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float send = up ? v[e] : v[e + 8], keep = up ? v[e + 8] : v[e];
                            v[e] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                        }
                    }
                    if (R >= 2) {
                        const bool up = lane & 2;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float send = up ? v[e] : v[e + 4], keep = up ? v[e + 4] : v[e];
                            v[e] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                        }
                    }
                    if (R >= 3) {
                        const bool up = lane & 4;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float send = up ? v[e] : v[e + 2], keep = up ? v[e + 2] : v[e];
                            v[e] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
                    }
                    if (R >= 4) {
                        const bool up = lane & 8;
                        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
                        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    if (R >= 5) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
                    PHX_MARK(4);
                    if (writer) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (e < cntv) kp[n0 + nb + e] = v[e];
                    }
                    PHX_MARK(5);
                }
            }
            ph_t ^= 1;
            tc_fence_before();
            PH2_MARK(5);
            if constexpr (TSM) fence_async_smem();
            __syncthreads();
            const int buf = stage & 1;
            if constexpr (TSM) {
                PHX_T0();
                // ---- (3) contraction with dU from the staged rows (a thread per (sample, field row): lanes walk the field rows, 144-byte
                // stride: conflict-free 128-bit reads) and scatter of the stage derivative to every CTA of the cluster
                // (4 consecutive lanes = 4 consecutive samples of one field row: one 128-bit remote store per rank)
                for (int item0 = 0; item0 < (NS << lo.nip_log2); item0 += NTHR) {
                    int il = c_il, n = c_n, goff = c_g, doff = c_d;
                    bool ok = c_ok;
                    uint32_t off = c_off + (uint32_t)(buf * Hd * KSTR * 4);
                    if (item0) {
                        const int item = item0 + tid;
                        il = (item >> 2) & ((1 << lo.nip_log2) - 1);
                        n = ((item >> (2 + lo.nip_log2)) << 2) | (item & 3);
                        ok = il < ni && n < NS;
                        goff = ok ? n * lo.sg_str + il * lo.C4 : 0;
                        doff = ok ? n * lo.DUSTR : 0;
                        off = (uint32_t)(((buf * Hd + i0 + il) * KSTR + n) * 4);
                    }
                    const float4* g4 = reinterpret_cast<const float4*>(sG + goff);
                    const float4* d4 = reinterpret_cast<const float4*>(du_cur + doff);
                    float o0 = 0.f, o1 = 0.f;
#pragma unroll 3
                    for (int q = 0; q < (lo.C4 >> 2); ++q) {
                        const float4 a = g4[q], b = d4[q];
                        o0 = fmaf(a.x, b.x, o0); o1 = fmaf(a.y, b.y, o1);
                        o0 = fmaf(a.z, b.z, o0); o1 = fmaf(a.w, b.w, o1);
                    }
                    const float o = o0 + o1;
                    PHX_MARK(4);
                    const float oa = __shfl_down_sync(0xffffffffu, o, 1), ob = __shfl_down_sync(0xffffffffu, o, 2),
                                oc = __shfl_down_sync(0xffffffffu, o, 3);
                    const int i = i0 + il;
                    if (ok && (lane & 3) == 0) {
#pragma unroll
                        for (int r = 0; r < V3_MAX_CS; ++r)
                            if (r < CS) st_async_v4(sK_rank[r] + off, o, oa, ob, oc, sK_rank[r] + bark_delta + buf * 8);
                    }
                    const int bg = b0 + n;
                    if (ok && p.k_out && bg < B) p.k_out[((int64_t)stage * B + bg) * Hd + i] = o;
                }
                PHX_MARK(5);
                // the staged rows leave through one TMA row copy per sample (this CTA's field rows are contiguous in
                // [stage][sample][i*C4 + j]).  Issuing a bulk copy costs its thread ~100 cycles and more when the queue is
                // full (measured: 1.5 k cycles for two per warp), so warps that have nothing else to do until the next
                // epilogue issue them -- off the critical path of the RK4 recurrence
                if (training && tma_idx >= 0) {
                    const int64_t rows = (int64_t)lo.sv_ntiles * 8;
                    for (int n = tma_idx + n_tma * lane; n < NS; n += 32 * n_tma)
                        if (b0 + n < rows)
                            bulk_s2g(p.save_G + ((int64_t)stage * rows + b0 + n) * lo.sv_NCP + i0 * lo.C4, sG + n * lo.sg_str,
                                     (uint32_t)(ni * lo.C4 * 4));
                    bulk_commit();
                }
            } else {
            // ---- (3) scatter of the stage derivative: this CTA's field rows to every CTA of the cluster
            for (int item = tid; item < (ni << NSL); item += NTHR) {
                const int il = item >> NSL, n = item & (NS - 1);
                const float* pp = sKpart + (il * lo.nchunks) * NS + n;
                float o = pp[0];
                for (int c = 1; c < lo.nchunks; ++c) o += pp[c * NS];
                const int i = i0 + il;
                const uint32_t off = (uint32_t)(((buf * Hd + i) * KSTR + n) * 4);
#pragma unroll
                for (int r = 0; r < V3_MAX_CS; ++r)
                    if (r < CS) st_async_f32(sK_rank[r] + off, o, sK_rank[r] + bark_delta + buf * 8);
                const int bg = b0 + n;
                if (p.k_out && bg < B) p.k_out[((int64_t)stage * B + bg) * Hd + i] = o;
            }
            }
            // Every CTA of the cluster sends its rows to everybody: buffer `buf` is complete when Hd*NS floats have landed.
            // No cluster barrier: a peer can only overwrite this buffer two stages later, after it has received OUR rows of
            // the next stage, which we send after the reads below.
            if (tid == 0) mbar_arrive_expect_tx(bar_k + buf, (uint32_t)(Hd * NS * 4));
            PHX_MARK(6);
            if (state_thr) mbar_wait(bar_k + buf, (uint32_t)((stage >> 1) & 1));
            PHX_MARK(7);
            PH2_MARK(6);

            // ---- (4) RK4 recurrence (rk4_alt_step_func); next stage input -> B operand of the first hidden layer
            float* act_next = have_next ? act_run : nullptr;       // (act_run already points at the next stage's block)
            if (state_thr) {
                const float* kk = kk0 + buf * (Hd * KSTR);
                const float4 k0 = reinterpret_cast<const float4*>(kk)[0], k1 = reinterpret_cast<const float4*>(kk)[1];
                const float o[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
                const float third = 1.0f / 3.0f;
                float zin[8];
                if (s == 0) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) { rk1[e] = o[e]; zin[e] = fmaf(dt * third, o[e], ry[e]); }                      // y + dt*k1/3
                } else if (s == 1) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) { rk2[e] = o[e]; zin[e] = fmaf(dt, o[e] - rk1[e] * third, ry[e]); }            // y + dt*(k2 - k1/3)
                } else if (s == 2) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) { rk3[e] = o[e]; zin[e] = fmaf(dt, (rk1[e] - rk2[e]) + o[e], ry[e]); }        // y + dt*(k1-k2+k3)
                } else {
                    float y1[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) y1[e] = fmaf(((rk1[e] + 3.0f * rk2[e]) + 3.0f * rk3[e]) + o[e], dt * 0.125f, ry[e]);
                    // emit every requested time in (t0, t1] (FixedGridODESolver.integrate + _linear_interp)
                    if (sc == rank) {
                        int jj = jout;
                        while (jj < lo.n_out) {
                            const float tj = sTout[jj];
                            if (!(t1 >= tj)) break;
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int bg = b0 + 8 * sc + e;
                                float vv;
                                if (tj == t0) vv = ry[e];
                                else if (tj == t1) vv = y1[e];
                                else vv = ry[e] + (y1[e] - ry[e]) / (t1 - t0) * (tj - t0);
                                if (bg < B) p.z_out[((int64_t)jj * B + bg) * Hd + si] = vv;
                            }
                            ++jj;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) { ry[e] = y1[e]; zin[e] = y1[e]; }
                }
                store_state(zin);
                if (act_next && sc == rank) {
                    reinterpret_cast<float4*>(act_next + si * 8)[0] = make_float4(zin[0], zin[1], zin[2], zin[3]);
                    reinterpret_cast<float4*>(act_next + si * 8)[1] = make_float4(zin[4], zin[5], zin[6], zin[7]);
                }
            }
            if (!const_cols_stable) write_const_cols();
            if (s == 3) {
                while (jout < lo.n_out && t1 >= sTout[jout]) ++jout;
            }
            fence_async_smem();
            __syncthreads();
            PH2_MARK(7);
        }
    }
    PH2_FLUSH(16 + (top ? 8 : 0));
#ifdef ANCDE_PHASE_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0 && p.dbg)
        for (int i = 0; i < 8; ++i) p.dbg[40 + (top ? 8 : 0) + i] += phx_acc[i];
#endif
    if (TSM && training && tma_idx >= 0) bulk_wait0();
    tc_fence_before();
    cluster_arrive();     // nobody leaves while a peer may still address its shared memory
    cluster_wait();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

static int launch_fwd3_t(const V3Args& a, cudaStream_t stream)
{
    const V3Layout& lo = a.lo;
    if (lo.chain1) return launch_fwd3_single(a, stream);
    const size_t smem = v3_smem(lo).total;
    auto kern = lo.ts ? ncde3_fwd_kernel<true> : ncde3_fwd_kernel<false>;
    ANCDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(lo.nclusters * lo.CS));
    cfg.blockDim = dim3(V3_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)lo.CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static const char* names[3] = {"ncde_fwd_plain", "ncde_fwd_bottom", "ncde_fwd_top"};
    prof_begin(names[lo.mode], stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    prof_end(stream);
    count_launch();
    if (e != cudaSuccess) {
        set_error("ncde_forward: launch failed: %s (cluster %d, %zu bytes of shared memory)", cudaGetErrorString(e), lo.CS, smem);
        return ANCDE_ECUDA;
    }
    return ANCDE_OK;
}

int launch_fwd3(const ancde_ncde_desc* d, const V3Layout& lo, const unsigned char* img, cudaStream_t stream)
{
    ANCDE_CHECK_ARG(d->times && d->t_out && d->coeff_b && d->coeff_c2 && d->coeff_d3 && d->z0 && d->z_out && img,
                    "ncde: null input/output pointer");
    ANCDE_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15) == 0, "ncde: the packed operand image must be 16-byte aligned");
    if (d->mode == ANCDE_MODE_TOP) ANCDE_CHECK_ARG(d->attention && d->h_prime, "ncde: TOP mode needs attention and h_prime");
    V3Args a = V3Args();
    a.lo = lo;
    a.t_start = d->t_start; a.t_end = d->t_end; a.step = d->step; a.grid = d->grid;
    a.times = d->times; a.t_out = d->t_out; a.cb = d->coeff_b; a.cc2 = d->coeff_c2; a.cd3 = d->coeff_d3;
    a.z0 = d->z0; a.attention = d->attention; a.h_prime = d->h_prime;
    a.img = img;
    a.wpack = d->wpack;
    a.z_out = d->z_out; a.k_out = d->k_out; a.save_act = d->save_act; a.save_G = d->save_G;
    a.dbg = debug_buffer();
    return launch_fwd3_t(a, stream);
}

}  // namespace ancde
