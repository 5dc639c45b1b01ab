// Cluster / tcgen05 recurrent backward of the fused NCDE solve (discrete adjoint of the RK4 3/8 recurrence): one
// thread-block cluster walks NS = 8*CS samples backwards with W_out^T resident in shared memory.  See ncde_bwd4.cuh
// for the decomposition.  Per RK4 stage a CTA
//   0. turns its RK4 adjoint state (registers) into the cotangent gbar of the stage derivative of ITS 8 samples, scales
//      it by a power of two (FP16 range) and broadcasts it to the cluster;          [transpose of torchdiffeq rk4_alt_step_func]
//   1. builds prebar = gbar * dU * (1 - G^2) of its slice of the reduction index for all NS samples from the saved tanh
//      outputs (FP16 hi/lo, B operand ring) while one elected thread issues  hbar_partial = W_out^T[slice] prebar  as
//      tcgen05.mma products into tensor memory; the timewise attention cotangent sum_p gbar*G*dU/da rides along in
//      registers;                                                                   [vector_fields.py:216-218,248-250; cdeint_module.py:113-133]
//   2. reads the accumulator back (tcgen05.ld) and reduce-scatters the partial sums over the cluster (st.async);
//   3. sums the CS partial hbar of its own samples, walks the hidden ReLU layers backwards (bwd_hidden.cuh) and
//   4. advances the RK4 adjoint recurrence.
// Matches autograd through torchdiffeq's rk4_alt_step_func + the reference vector fields
// (cdeint_module.py:25-32,56-72,103-138) to FP32 round-off, like the streamed kernel it replaces.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "bwd_hidden.cuh"
#include "mma_common.cuh"
#include "ncde_bwd4.cuh"
#include "umma.cuh"

namespace ancde {

// Waits on an mbarrier: every thread suspends in mbarrier.try_wait.  (Measured alternative: one lane per warp
// busy-polling test_wait while the others sit behind __syncwarp made every stage 4 k cycles SLOWER -- the polling warps
// take issue slots and shared-memory bandwidth from the few warps that work in the latency-bound phases.)
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity, int)
{
    mbar_wait(bar, parity);
}

// Named barrier over the worker warps (the issuer warp never joins).
__device__ __forceinline__ void workers_sync() { named_bar_sync(1, V4_WORKERS); }

// HM: how the hidden layers run backwards -- 0: FP32 warp-synchronous chain (a warp per sample, widths <= 32), 1: mma.sync
// (a warp per 16 inputs), 2: FP32 with one (unit, sample) output per thread over all the warps a layer fills.
template <int HM>
__global__ void __launch_bounds__(V4_THREADS, 1) ncde_bwd4_kernel(const __grid_constant__ Bwd4Args p)
{
    constexpr bool CHAIN = HM != 1;       // FP32 buffers and transposed FP32 weights (modes 0 and 2)
    const NcdeLayout& lo = p.s.lo;            // streamed layout with tiles of 8 samples
    const V4Layout& v = p.v;
    constexpr int TS = 8, NT = V4_WORKERS, NW = NT / 32;      // worker threads; warp NW is the MMA issuer
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);      // provably warp-uniform (uniform datapath for the MMA issue)
    const int wq = warp & 3, wg = warp >> 2;
    const int CS = v.CS, NS = v.NS, CK = v.CK, NSLOT = v.nslots;
    const int rank = (int)cluster_ctarank();
    const int cl = blockIdx.x / CS;
    const int tile = cl * CS + rank;          // the tile of 8 samples this CTA owns
    const bool have_tile = tile < lo.ntiles;
    const int b0 = tile * TS;
    const int Hd = lo.Hd, NCP = lo.NCP, K = lo.K, DUS = lo.DUS;
    const int HT = Hd * TS;
    const bool need_att = v.need_att != 0;
    const int64_t rows = (int64_t)lo.ntiles * TS;               // sample rows of the saved tanh outputs

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar_g = reinterpret_cast<uint64_t*>(smem_raw + v.o_bars);     // gathered stage cotangent complete (tx bytes)
    uint64_t* bar_du = bar_g + 1;                                             // dU blocks of the stage have landed (TMA)
    uint64_t* bar_act = bar_g + 2;                                            // own saved activations have landed (TMA)
    uint64_t* bar_recv = bar_g + 3;                                           // partial sums of my samples complete (tx bytes)
    uint64_t* bar_built = bar_g + 4;                                          // [2] ring chunk written by all worker warps
    uint64_t* bar_free = bar_g + 6;                                           // [2] products of a ring chunk complete
    uint64_t* bar_gin = bar_g + 8;                                            // tanh outputs of the stage have landed (TMA)
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar_g + 9);
    unsigned char* sRing = smem_raw + v.o_ring;
    const float* sWtb = reinterpret_cast<const float*>(smem_raw + v.o_hid);  // CHAIN: transposed hidden weights (FP32)
    const uint4* sWhb = reinterpret_cast<const uint4*>(smem_raw + v.o_hid);  // !CHAIN: W_l^T as FP16 hi/lo A fragments
    float* sRecv = reinterpret_cast<float*>(smem_raw + v.o_recv);            // [src rank][unit][8 samples]
    float* sRecvAtt = reinterpret_cast<float*>(smem_raw + v.o_recvatt);      // [src rank][warp of the tile][8 samples]
    float* sGall = reinterpret_cast<float*>(smem_raw + v.o_gall);            // scaled gbar of the cluster [Hd][GSTR]
    float* sDU = reinterpret_cast<float*>(smem_raw + v.o_du);                // [tile of the cluster][dU | dU/da]
    float* sAct = reinterpret_cast<float*>(smem_raw + v.o_act);              // own saved activations z_s | h_1 .. h_n
    float* sZb = reinterpret_cast<float*>(smem_raw + v.o_zb);                // cotangent of the stage input [Hd][8]
    __half* sBd = reinterpret_cast<__half*>(smem_raw + v.o_bd);              // !CHAIN: two buffers of {hi, lo} [8][KP]
    float* sD0 = reinterpret_cast<float*>(smem_raw + v.o_bd);                // CHAIN: hidden cotangents ping [sample][unit]
    float* sD1 = sD0 + TS * v.HP;                                             //        pong
    float* sMax = reinterpret_cast<float*>(smem_raw + v.o_max);
    int* sLp = reinterpret_cast<int*>(smem_raw + v.o_lp);
    float* sGst = reinterpret_cast<float*>(smem_raw + v.o_gst);              // staged tanh outputs [NS rows][GROW floats]
    float* sTout = reinterpret_cast<float*>(smem_raw + v.o_tout);            // requested output times
    int4* sTab = reinterpret_cast<int4*>(smem_raw + v.o_tab);                // per run of 4 columns: {gbar row, dU index, G column, -}

    const int nks = v.ks_begin[rank + 1] - v.ks_begin[rank];
    const int nchunks = ((nks + v.PASSK - 1) / v.PASSK + v.PPC - 1) / v.PPC;       // CK = PPC * PASSK k steps per chunk
    const int pbase = 16 * v.ks_begin[rank];

    // ---- prologue: resident operands, tables, barriers, tensor memory
    if (tid < NT) {
        for (int i = tid; i < lo.n_out; i += NT) sTout[i] = p.s.t_out[i];
        if constexpr (CHAIN) {
            const float4* hs = reinterpret_cast<const float4*>(p.s.wpack + lo.off_hiddenb);
            float4* hd = reinterpret_cast<float4*>(smem_raw + v.o_hid);
            for (int i = tid; i < lo.hiddenb_floats / 4; i += NT) hd[i] = __ldg(hs + i);
            for (int i = tid; i < 2 * TS * v.HP; i += NT) sD0[i] = 0.f;
        } else {
            const uint4* hs = reinterpret_cast<const uint4*>(p.s.wpack + lo.off_whb);
            uint4* hd = reinterpret_cast<uint4*>(smem_raw + v.o_hid);
            for (int i = tid; i < lo.hb_u4; i += NT) hd[i] = __ldg(hs + i);
            uint32_t* zb = reinterpret_cast<uint32_t*>(sBd);
            for (int i = tid; i < 2 * 8 * v.KP; i += NT) zb[i] = 0u;      // 4 matrices of 8*KP halfs
        }
        // operand rows / columns that are never written stay zero (0 * finite garbage would be fine, 0 * NaN is not)
        uint4* zr = reinterpret_cast<uint4*>(sRing);
        for (int i = tid; i < NSLOT * v.ring_bytes / 16; i += NT) zr[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < NS * v.GROW; i += NT) sGst[i] = 0.f;              // rows behind the batch stay zero
        for (int i = tid; i < (Hd + 1) * v.GSTR; i += NT) sGall[i] = 0.f;      // row Hd stays zero
        for (int i = tid; i < CS * v.dub; i += NT) sDU[i] = 0.f;          // tiles behind the batch stay zero
        for (int i = tid; i < v.act_ld; i += NT) sAct[i] = 0.f;
        for (int i = tid; i < HT; i += NT) sZb[i] = 0.f;
        if (tid < 32) sMax[tid] = 0.f;
        if (tid < lo.n_hidden) {
            const int l = tid;
            int* q = sLp + 8 * l;
            if (CHAIN) {
                q[0] = (lo.width[l] + 3) / 4; q[1] = lo.in_dim[l]; q[2] = lo.tstride[l]; q[3] = lo.off_wt[l] - lo.off_hiddenb;
                q[4] = l > 0 ? lo.act_off_h[l - 1] : 0; q[5] = l > 0 ? lo.act_off_delta[l - 1] : 0; q[6] = 0; q[7] = 0;
            } else {
                q[0] = lo.hbp[l].mtiles; q[1] = lo.hbp[l].full; q[2] = lo.hbp[l].tail8; q[3] = lo.hbp[l].stride;
                q[4] = lo.hbp[l].off; q[5] = lo.hbp[l].M;
                q[6] = l > 0 ? lo.act_off_h[l - 1] : 0; q[7] = l > 0 ? lo.act_off_delta[l - 1] : 0;
            }
        }
        // columns 8*gq .. 8*gq + 7 of this rank's slice = two runs of 4 control channels of one field row each; per run
        // {offset of the field row in the gathered cotangent, float4 index of the 4 channels in the re-laid-out controls,
        // column offset in the saved tanh outputs}.  Columns behind the slice or behind the last field row point at the ZERO
        // row of the gathered cotangent (and at column 0): no predicates in the conversion loop.
        for (int gq = tid; gq < 2 * v.PASSK * v.npass_max; gq += NT) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int pp = pbase + 8 * gq + 4 * h;
                const bool ok = gq < 2 * nks && pp < NCP;
                const int i = pp / lo.C4, j0 = pp - i * lo.C4;
                sTab[2 * gq + h] = ok ? make_int4(i * v.GSTR, (j0 >> 2) * 8, 8 * gq + 4 * h, 0) : make_int4(Hd * v.GSTR, 0, 0, 0);
            }
        }
        if (tid == 0) {
            mbar_init(bar_g, 1);
            mbar_init(bar_du, 1);
            mbar_init(bar_act, 1);
            mbar_init(bar_recv, 1);
            mbar_init(bar_built + 0, NW);
            mbar_init(bar_built + 1, NW);
            mbar_init(bar_free + 0, 1);
            mbar_init(bar_free + 1, 1);
            mbar_init(bar_gin, 1);
            mbar_fence_init();
        }
        if (warp == 0) tmem_alloc(tslot, (uint32_t)v.tmem_cols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const uint32_t tlane = tbase + ((uint32_t)(32 * wq) << 16);
    if (tid < NT) {
        // resident A operand: row m = TMEM lane m (this warp's quarter), 16 FP16 of the row per 8 columns; the four warps of
        // a lane quarter share the k steps
        const int m = 32 * wq + lane;
        const uint4* src = reinterpret_cast<const uint4*>(p.img + (size_t)rank * v.a_bytes + (size_t)m * v.nks_max * 32);
        for (int ks = wg; ks < v.nks_max; ks += 4) {
            const uint4 x0 = __ldg(src + 2 * ks), x1 = __ldg(src + 2 * ks + 1);
            const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            tmem_st8(tlane + (uint32_t)(v.a_col + 8 * ks), r);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    fence_async_smem();
    cluster_arrive();     // every CTA of the cluster has initialised its shared memory and barriers before anyone writes into it
    cluster_wait();

    auto act_block = [&](int n, int tl) { return p.s.save_act + ((int64_t)n * lo.ntiles + tl) * lo.act_floats; };
    // dU (| dU/da) blocks of all tiles of the cluster for stage n: one TMA bulk copy per tile, ONE issuing thread.  They
    // land in the operand buffer, which is idle between the last product of a stage and the operand build of the next,
    // and are re-laid out from there at the start of the stage (below).
    float* sStage = reinterpret_cast<float*>(sRing);
    auto issue_du = [&](int n) {
        uint32_t bytes = 0;
        for (int t = 0; t < CS; ++t)
            if (cl * CS + t < lo.ntiles) bytes += (uint32_t)(v.dub * 4);
        mbar_arrive_expect_tx(bar_du, bytes);
        for (int t = 0; t < CS; ++t)
            if (cl * CS + t < lo.ntiles)
                bulk_g2s(sStage + (size_t)t * v.dub, act_block(n, cl * CS + t) + lo.act_off_du, (uint32_t)(v.dub * 4), bar_du);
    };
    // The slice of the saved tanh outputs this CTA converts in stage n: one TMA bulk copy per sample row (rows padded by 16
    // bytes in shared memory: the 8 rows a quarter warp reads fall into 8 different bank groups).  Called by a whole warp.
    const int g_cols = (pbase + 16 * nks <= NCP) ? 16 * nks : NCP - pbase;          // floats of a row slice
    auto issue_g = [&](int n) {
        const int64_t r0 = (int64_t)cl * NS;
        const int nrows = (rows - r0 < NS) ? (int)(rows - r0) : NS;
        if (lane == 0) mbar_arrive_expect_tx(bar_gin, (uint32_t)(nrows * g_cols * 4));
        __syncwarp();
        for (int r = lane; r < nrows; r += 32)
            bulk_g2s(sGst + (size_t)r * v.GROW, p.s.save_G + ((int64_t)n * rows + r0 + r) * NCP + pbase, (uint32_t)(g_cols * 4), bar_gin);
    };
    auto issue_act = [&](int n) {
        mbar_arrive_expect_tx(bar_act, (uint32_t)(v.act_ld * 4));
        bulk_g2s(sAct, act_block(n, tile), (uint32_t)(v.act_ld * 4), bar_act);
    };

    if (warp_u == NW) {
        // ================================ the issuer warp: products of every operand chunk ================================
        const uint32_t idesc = umma_idesc_f16(128, 2 * NS, false);
        const uint32_t tA = tbase + (uint32_t)v.a_col;
        const uint64_t dB0 = umma_smem_desc(smem_u32(sRing), 128, (uint32_t)v.SBOb);
        const uint64_t dB1 = umma_smem_desc(smem_u32(sRing) + (uint32_t)v.ring_bytes, 128, (uint32_t)v.SBOb);
        uint32_t pb0 = 0, pb1 = 0;
        int use = 0;                   // ring position of the next chunk (all stages)
        issue_g(lo.n_stages - 1);
        for (int n = lo.n_stages - 1; n >= 0; --n) {
            for (int c = 0; c < nchunks; ++c) {
                const int slot = (NSLOT == 2) ? (use & 1) : 0;
                ++use;
                warp_wait(bar_built + slot, slot ? pb1 : pb0, lane);
                if (slot) pb1 ^= 1; else pb0 ^= 1;
                if (elect_one()) {
                    tc_fence_after();
                    const int k0 = CK * c;
                    const int kn = (nks - k0 < CK) ? nks - k0 : CK;
                    const uint64_t dB = slot ? dB1 : dB0;
#pragma unroll 4
                    for (int ks = 0; ks < kn; ++ks)
                        umma_f16_ts(tbase, tA + (uint32_t)((k0 + ks) * 8), dB + (uint64_t)(ks * 16), idesc, (k0 + ks) != 0);
                    umma_commit(bar_free + slot);
                }
                __syncwarp();
                // every worker warp has converted its last item of this stage: the staging buffer takes the next stage's rows
                if (c == nchunks - 1 && n >= 1) issue_g(n - 1);
            }
        }
    } else {
    // ================================================ the worker warps ================================================
    if (tid == 0) {
        issue_du(lo.n_stages - 1);
        if (have_tile) issue_act(lo.n_stages - 1);
    }

    // ---- constants of the operand build: a thread converts (sample nloc of the cluster, 8 columns q of a chunk) items;
    // the sample is the same for all its items: n8 = lane % 8 of tile `tile_t`, and q = 4 * (item warp / CS) + lane / 8
    const int n8 = lane & 7;
    const int cs_log2 = (CS == 8) ? 3 : (CS == 4 ? 2 : 1);
    const int tile_t = warp & (CS - 1);                      // tile of the cluster this thread's sample belongs to
    const int nloc = 8 * tile_t + n8;
    const bool samp_ok = cl * CS + tile_t < lo.ntiles;
    // one PASS of the 512 workers converts 8 * PASSK columns... (PASSK = 32 / CS k steps) for all NS samples: a thread's
    // items of a stage are the same (sample, 8-column group of the pass) in every pass
    const int PASSK = v.PASSK, PPC = v.PPC;
    const int npass = (nks + PASSK - 1) / PASSK;
    const int qpass = 4 * (warp >> cs_log2) + (lane >> 3);   // 0 .. 2*PASSK - 1
    const uint32_t ring_hi = smem_u32(sRing) + (uint32_t)((nloc >> 3) * v.SBOb + (nloc & 7) * 16 + qpass * 128);
    const uint32_t ring_lo_off = (uint32_t)((NS >> 3) * v.SBOb);
    const int NCH = lo.NCH;
    const bool top = lo.mode == ANCDE_MODE_TOP;
    // re-laid-out controls: float4 {channels 4*j4 .. 4*j4 + 3} per (tile, j4, sample); dU/da behind dU
    const float4* sDUt = reinterpret_cast<const float4*>(sDU) + (size_t)tile_t * NCH * 8 + n8;
    const float4* sDAt = sDUt + (size_t)CS * NCH * 8;
    const float* sGn = sGall + nloc;
    uint32_t gall_rank[V4_MAX_CS];
#pragma unroll
    for (int r = 0; r < V4_MAX_CS; ++r) gall_rank[r] = map_to_rank(smem_u32(sGall), (uint32_t)(r < CS ? r : 0));
    const uint32_t recv_att_dst = map_to_rank(smem_u32(sRecv), (uint32_t)tile_t);     // the owner of this thread's sample
    const uint32_t barg_delta = smem_u32(bar_g) - smem_u32(sGall);         // the peers' barriers sit at the same distances
    const uint32_t barr_delta = smem_u32(bar_recv) - smem_u32(sRecv);
    const uint32_t ratt_delta = smem_u32(sRecvAtt) - smem_u32(sRecv);

    // Controls of a stage: the TMA-staged [chunk j4][channel][sample] blocks of every tile -> float4 per (tile, j4, sample).
    // `first`, `count`: the calling threads' index and number (all workers before the first stage, the idle warps during the
    // hidden layers).
    auto stage_in = [&](int n, int first, int count) {
        for (int idx = first; idx < CS * NCH * 8; idx += count) {
            const int t = idx / (NCH * 8), rem = idx - t * NCH * 8;
            const float* src = sStage + (size_t)t * v.dub + (rem >> 3) * DUS + (rem & 7);
            reinterpret_cast<float4*>(sDU)[idx] = make_float4(src[0], src[8], src[16], src[24]);
            if (top) {
                src += lo.du_floats;
                reinterpret_cast<float4*>(sDU)[CS * NCH * 8 + idx] = make_float4(src[0], src[8], src[16], src[24]);
            }
        }
    };
    const int CW = v.chain_warps;       // warps that walk the hidden layers backwards; the others prepare the next stage meanwhile

    // ---- RK4 adjoint state in registers: thread tid < Hd*8 owns element (i = tid / 8, b = tid % 8) of its tile
    const bool st_thr = tid < HT;
    const int si = tid >> 3, sb = tid & 7;
    float lam = 0.f, mu = 0.f, v4 = 0.f, v3 = 0.f, v2 = 0.f;
    float gmax = 0.f;
    int jb = lo.n_out - 1;             // last output time whose cotangent has not been consumed yet
    uint32_t par_g = 0, par_du = 0, par_act = 0, par_recv = 0, par_gin = 0;
    int use = 0;                       // ring position of the next chunk (all stages; the same in every thread)
    PH2_DECL();
#ifdef ANCDE_PHASE_TIMING
    if (tid == 0 && rank == 0 && p.s.dbg && cl < 32 && lo.mode == ANCDE_MODE_TOP) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.s.dbg[64 + cl] = (long long)gt;
    }
#endif
    warp_wait(bar_du, par_du, lane);          // controls of the last stage (the first one this kernel visits)
    par_du ^= 1;
    stage_in(lo.n_stages - 1, tid, NT);

    for (int k = lo.n_steps - 1; k >= 0; --k) {
        const float t0 = grid_time(k, lo.n_steps, p.s.t_start, p.s.t_end, p.s.step, p.s.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.s.t_start, p.s.t_end, p.s.step, p.s.grid);
        const float dt = __fsub_rn(t1, t0);
        // ---- cotangents of the outputs emitted during this step: t_out[j] in (t0, t1]
        if (st_thr) {
            const int bg = b0 + sb;
            for (int j = jb; j >= 1; --j) {
                const float tj = sTout[j];
                if (!(tj > t0)) break;
                const float gz = (have_tile && bg < lo.B) ? __ldg(p.s.grad_z_out + ((int64_t)j * lo.B + bg) * Hd + si) : 0.f;
                if (tj == t1) lam += gz;
                else {
                    const float th = (tj - t0) / (t1 - t0);
                    lam += th * gz;
                    mu += (1.0f - th) * gz;
                }
            }
        }
        while (jb >= 1 && sTout[jb] > t0) --jb;

#pragma unroll 1
        for (int s = 3; s >= 0; --s) {
            const int n = 4 * k + s;
            float* act = have_tile ? act_block(n, tile) : nullptr;
            PH2_T0();
            // ---- (0) cotangent of the stage derivative k_{s+1} (transpose of rk4_alt_step_func), scaled, to the cluster
            float gv = 0.f;
            if (st_thr) {
                if (s == 3) gv = dt * 0.125f * lam;
                else if (s == 2) gv = dt * (0.375f * lam + v4);
                else if (s == 1) gv = dt * (0.375f * lam - v4 + v3);
                else gv = dt * (0.125f * lam + v4 + (v2 - v3) * (1.0f / 3.0f));
                if (act) act[lo.act_off_gbar + tid] = gv;        // the stage cotangent goes to HBM for the weight-gradient kernels
            }
            float lmax = fabsf(gv);
            gmax = fmaxf(gmax, lmax);
            for (int o = 16; o >= 1; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
            if (lane == 0) sMax[warp] = lmax;
            workers_sync();
            float scale = 1.0f;
            {
                const float4 m0 = *reinterpret_cast<const float4*>(sMax), m1 = *reinterpret_cast<const float4*>(sMax + 4);
                const float4 m2 = *reinterpret_cast<const float4*>(sMax + 8), m3 = *reinterpret_cast<const float4*>(sMax + 12);
                const float m = fmaxf(fmaxf(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m0.z, m0.w)), fmaxf(fmaxf(m1.x, m1.y), fmaxf(m1.z, m1.w))),
                                      fmaxf(fmaxf(fmaxf(m2.x, m2.y), fmaxf(m2.z, m2.w)), fmaxf(fmaxf(m3.x, m3.y), fmaxf(m3.z, m3.w))));
                if (m > 1e-30f && m < 1e30f) {
                    const int e = (int)((__float_as_uint(m) >> 23) & 0xffu) - 126;     // m = f * 2^e, f in [0.5, 1)
                    scale = __uint_as_float((uint32_t)(127 + 6 - e) << 23);            // brings the largest cotangent to ~2^6
                }
            }
            const float inv_scale = __uint_as_float((254u << 23) - __float_as_uint(scale));     // power of two: exponent flip
            if (st_thr) {
                const float val = gv * scale;
                const uint32_t off = (uint32_t)((si * v.GSTR + 8 * rank + sb) * 4);
#pragma unroll
                for (int r = 0; r < V4_MAX_CS; ++r)
                    if (r < CS) st_async_f32(gall_rank[r] + off, val, gall_rank[r] + barg_delta);
            }
            if (tid == 0) mbar_arrive_expect_tx(bar_g, (uint32_t)(HT * CS * 4));
            warp_wait(bar_g, par_g, lane);
            par_g ^= 1;
            PH2_MARK(0);

            // ---- (1) operand build, pass by pass; the issuer warp multiplies a chunk as soon as every warp has written it
            float att = 0.f;
            {
                warp_wait(bar_gin, par_gin, lane);      // this stage's tanh outputs (copies issued a whole stage ago)
                par_gin ^= 1;
                const float* grow = sGst + (size_t)nloc * v.GROW;
                const int4* tp = sTab + 2 * qpass;
                const int tstep = 4 * PASSK;                  // table entries per pass
                // two passes ahead: the tanh outputs of the next two items are in flight while one is converted
                float4 gA0, gB0, gA1, gB1;
                int4 tA0, tB0, tA1, tB1;
                auto fetch = [&](int pass, float4& ga, float4& gb, int4& ta, int4& tb) {
                    const int pz = pass < npass ? pass : npass - 1;          // (past the end: a harmless reload)
                    ta = tp[pz * tstep];
                    tb = tp[pz * tstep + 1];
                    ga = *reinterpret_cast<const float4*>(grow + ta.z);
                    gb = *reinterpret_cast<const float4*>(grow + tb.z);
                };
                fetch(0, gA0, gB0, tA0, tB0);
                fetch(1, gA1, gB1, tA1, tB1);
                int pic = 0;                                   // pass inside the current chunk
                int slot = (NSLOT == 2) ? (use & 1) : 0;
                uint32_t so = ring_hi + (uint32_t)(slot * v.ring_bytes);
                auto convert = [&](int pass, const float4& g0, const float4& g1, const int4& ta, const int4& tb) {
                    float pv[8];
                    {
                        const float gbv = sGn[ta.x];
                        const float4 du = sDUt[ta.y];
                        pv[0] = gbv * du.x * fmaf(-g0.x, g0.x, 1.0f);
                        pv[1] = gbv * du.y * fmaf(-g0.y, g0.y, 1.0f);
                        pv[2] = gbv * du.z * fmaf(-g0.z, g0.z, 1.0f);
                        pv[3] = gbv * du.w * fmaf(-g0.w, g0.w, 1.0f);
                        if (need_att) {
                            // sum_j dUbar[j] * dU/da[j] with dUbar[j] = sum_i gbar[i] * G[i][j]   
                            const float4 da = sDAt[ta.y];
                            att = fmaf(gbv, (g0.x * da.x + g0.y * da.y) + (g0.z * da.z + g0.w * da.w), att);
                        }
                    }
                    {
                        const float gbv = sGn[tb.x];
                        const float4 du = sDUt[tb.y];
                        pv[4] = gbv * du.x * fmaf(-g1.x, g1.x, 1.0f);
                        pv[5] = gbv * du.y * fmaf(-g1.y, g1.y, 1.0f);
                        pv[6] = gbv * du.z * fmaf(-g1.z, g1.z, 1.0f);
                        pv[7] = gbv * du.w * fmaf(-g1.w, g1.w, 1.0f);
                        if (need_att) {
                            const float4 da = sDAt[tb.y];
                            att = fmaf(gbv, (g1.x * da.x + g1.y * da.y) + (g1.z * da.z + g1.w * da.w), att);
                        }
                    }
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) split_pair(pv[2 * x], pv[2 * x + 1], h[x], l[x]);
                    if (pic == 0) {
                        // the buffer is free once the products of the chunk that used it last have completed: chunk number
                        // use - NSLOT (0-based over the whole solve) is completion number (use - NSLOT) / NSLOT of its barrier
                        if (use >= NSLOT) warp_wait(bar_free + slot, (uint32_t)(((use - NSLOT) / NSLOT) & 1), lane);
                        ++use;
                    }
                    const uint32_t dst = so + (uint32_t)(pic * 2 * PASSK * 128);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst + ring_lo_off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
                    if (++pic == PPC || pass == npass - 1) {
                        // chunk complete (for this warp)
                        fence_async_smem();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_built + slot);
                        pic = 0;
                        slot = (NSLOT == 2) ? (use & 1) : 0;
                        so = ring_hi + (uint32_t)(slot * v.ring_bytes);
                    }
                };
                for (int pass = 0; pass < npass; pass += 2) {
                    {
                        const float4 g0 = gA0, g1 = gB0;
                        const int4 ta = tA0, tb = tB0;
                        fetch(pass + 2, gA0, gB0, tA0, tB0);
                        convert(pass, g0, g1, ta, tb);
                    }
                    if (pass + 1 < npass) {
                        const float4 g0 = gA1, g1 = gB1;
                        const int4 ta = tA1, tb = tB1;
                        fetch(pass + 3, gA1, gB1, tA1, tB1);
                        convert(pass + 1, g0, g1, ta, tb);
                    }
                }
            }
            PH2_MARK(1);
            // timewise attention: this thread's share for sample nloc -> the sample's owner
            if (need_att) {
                att += __shfl_xor_sync(0xffffffffu, att, 8);
                att += __shfl_xor_sync(0xffffffffu, att, 16);
                if (lane < 8) {
                    const uint32_t base = recv_att_dst;
                    st_async_f32(base + ratt_delta + (uint32_t)(((rank * v.QW + (warp >> cs_log2)) * 8 + lane) * 4), att, base + barr_delta);
                }
            }

            // ---- (2) accumulator -> registers -> reduce-scatter of the partial sums over the cluster
            {
                const int last_use = use - 1;             // the stage's last chunk
                const int last = (NSLOT == 2) ? (last_use & 1) : 0;
                warp_wait(bar_free + last, (uint32_t)((last_use / NSLOT) & 1), lane);
                tc_fence_after();
                // every product of this stage has completed: the operand buffer is idle, the next stage's controls land in it
                if (tid == 0 && n >= 1) issue_du(n - 1);
                PH2_MARK(2);
                if (wg < (NS >> 4)) {
                    float a[16], b2[16];
                    tmem_ld16x2(tlane + 16 * wg, tlane + NS + 16 * wg, a, b2);
                    const bool hi = (lane & 8) == 0;
                    float out[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        // hi rows: W_hi*p_hi + W_hi*p_lo; lo rows: W_lo*p_hi.  The hi lane keeps samples 0..7 of the slab, the
                        // lo lane samples 8..15: one exchange gives every lane 8 finished sums.
                        const float s_lo = hi ? a[e] + b2[e] : a[e];
                        const float s_hi = hi ? a[8 + e] + b2[8 + e] : a[8 + e];
                        const float send = hi ? s_hi : s_lo;
                        const float keep = hi ? s_lo : s_hi;
                        out[e] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    const int u = 16 * wq + 8 * (lane >> 4) + (lane & 7);
                    const int owner = 2 * wg + (hi ? 0 : 1);
                    if (u < K) {
                        const uint32_t base = map_to_rank(smem_u32(sRecv), (uint32_t)owner);
                        const uint32_t dst = base + (uint32_t)(((rank * K + u) * 8) * 4);
                        st_async_v4(dst, out[0], out[1], out[2], out[3], base + barr_delta);
                        st_async_v4(dst + 16, out[4], out[5], out[6], out[7], base + barr_delta);
                    }
                }
                tc_fence_before();
                if (tid == 0) mbar_arrive_expect_tx(bar_recv, (uint32_t)(CS * K * 8 * 4 + (need_att ? CS * v.QW * 8 * 4 : 0)));
            }

            // ---- (3) my samples: sum of the partial hbar, hidden layers backwards
            warp_wait(bar_recv, par_recv, lane);
            par_recv ^= 1;
            if (have_tile) {
                warp_wait(bar_act, par_act, lane);
                par_act ^= 1;
            }
            PH2_MARK(3);
            {
                const int l = lo.n_hidden - 1;
                if (tid < K * TS) {
                    const int o = tid >> 3, b = tid & 7;
                    float x = 0.f;
                    for (int r = 0; r < CS; ++r) x += sRecv[(r * K + o) * 8 + b];
                    const float hv = sAct[lo.act_off_h[l] + tid];
                    if constexpr (CHAIN) {
                        x = (hv > 0.f) ? x * ((1.0f / 64.0f) * inv_scale) : 0.f;
                        if (act) act[lo.act_off_delta[l] + tid] = x;          // for the hidden weight-gradient kernel
                        sD0[b * v.HP + o] = x;
                    } else {
                        x = (hv > 0.f) ? x * (1.0f / 64.0f) : 0.f;            // stays multiplied by `scale` inside the chain
                        if (act) act[lo.act_off_delta[l] + tid] = x * inv_scale;
                        const __half xh = __float2half_rn(x);
                        sBd[b * v.KP + o] = xh;
                        sBd[8 * v.KP + b * v.KP + o] = __float2half_rn(x - __half2float(xh));
                    }
                }
                if (need_att && warp == NW - 1 && lane < 8) {
                    float a = 0.f;
                    for (int r = 0; r < CS * v.QW; ++r) a += sRecvAtt[r * 8 + lane];
                    const float ts = stage_time(t0, dt, s);
                    int qa = (int)floorf(ts) - 1;
                    if (qa < 0) qa += lo.att_L;
                    qa = qa < 0 ? 0 : (qa >= lo.att_L ? lo.att_L - 1 : qa);
                    if (have_tile && b0 + lane < lo.B)
                        atomicAdd(p.s.grad_attention + (int64_t)qa * lo.B + b0 + lane, a * inv_scale);
                }
            }
            workers_sync();
            if (warp < CW) {
                if constexpr (HM == 0) {
                    hidden_chain_bwd<TS, 1>(sLp, lo.n_hidden, sWtb, sD0, sD1, sAct, act, sZb, v.HP, warp, CW, lane);
                } else if constexpr (HM == 2) {
                    const int cthreads = 32 * CW;
                    hidden_wide_bwd<TS>(sLp, lo.n_hidden, sWtb, sD0, sD1, sAct, act, sZb, v.HP, tid, [cthreads] { named_bar_sync(2, cthreads); });
                } else {
                    const int cthreads = 32 * CW;
                    hidden_mma_bwd<TS>(sLp, lo.n_hidden, sWhb, sBd, v.KP, sAct, act, sZb, inv_scale, warp, lane,
                                       [cthreads] { named_bar_sync(2, cthreads); });
                }
            } else if (n >= 1) {
                // meanwhile the other warps bring in the next stage: its controls (TMA issued after this stage's last product)
                // into the layout of the operand build, its tanh outputs towards L2
                warp_wait(bar_du, par_du, lane);
                stage_in(n - 1, tid - 32 * CW, NT - 32 * CW);
            }
            par_du ^= 1;
            workers_sync();
            // the saved activations of this stage are consumed: fetch the next stage's (this kernel walks n downwards)
            if (tid == 0 && n >= 1 && have_tile) issue_act(n - 1);
            PH2_MARK(4);

            // ---- (4) RK4 adjoint recurrence
            if (st_thr) {
                const float x = sZb[tid];
                if (s == 3) v4 = x;
                else if (s == 2) v3 = x;
                else if (s == 1) v2 = x;
                else {
                    lam = lam + v4 + v3 + v2 + x + mu;
                    mu = 0.f;
                }
            }
            PH2_MARK(5);
        }
    }
#ifdef ANCDE_PHASE_TIMING
    if (tid == 0 && blockIdx.x == 0 && p.s.dbg)
        for (int i = 0; i < 8; ++i) p.s.dbg[(lo.mode == ANCDE_MODE_TOP ? 40 : 32) + i] += ph2_acc[i];
    if (tid == 0 && rank == 0 && p.s.dbg && cl < 32 && lo.mode == ANCDE_MODE_TOP) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.s.dbg[96 + cl] = (long long)gt;
    }
#endif
    // ---- results
    for (int o = 16; o >= 1; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0 && have_tile && gmax > 0.f && gmax < 3.0e38f)
        atomicMax(reinterpret_cast<unsigned*>(const_cast<float*>(p.s.wpack) + lo.off_gscale), __float_as_uint(gmax));
    if (st_thr && have_tile) {
        const int bg = b0 + sb;
        if (bg < lo.B) p.s.grad_z0[(int64_t)bg * Hd + si] = lam + __ldg(p.s.grad_z_out + (int64_t)bg * Hd + si);
    }
    }   // worker warps
    tc_fence_before();
    cluster_arrive();     // nobody leaves while a peer may still address its shared memory
    cluster_wait();
    if (warp == 0) tmem_dealloc(tbase, (uint32_t)v.tmem_cols);
}

// ---- host side ------------------------------------------------------------------------------------------------------
static int pow2ceil4(int x)
{
    int p = 32;
    while (p < x) p *= 2;
    return p;
}

static bool chain4(const NcdeLayout& lo)
{
    for (int l = 0; l < lo.n_hidden; ++l)
        if (lo.width[l] > 32) return false;
    return true;
}

int make_layout4(const ancde_ncde_desc* d, const NcdeLayout& st, bool need_att, int force_cs, V4Layout* v)
{
    *v = V4Layout();
    if (st.TS != 8 || d->single_stage || st.n_stages < 4) return ANCDE_EUNSUPPORTED;
    if (st.K > 64 || st.Hd * 8 > V4_WORKERS || st.K * 8 > V4_WORKERS) return ANCDE_EUNSUPPORTED;
    if (need_att && st.att_C != 1) return ANCDE_EUNSUPPORTED;       // elementwise attention cotangents: streamed kernel
    if ((int64_t)st.ntiles * 8 * st.NCP * st.n_stages >= ((int64_t)1 << 40)) return ANCDE_EUNSUPPORTED;
    const bool chain = chain4(st);
    int maxmt = 1;
    for (int l = 0; l < st.n_hidden; ++l) maxmt = st.hbp[l].mtiles > maxmt ? st.hbp[l].mtiles : maxmt;
    if (!chain && maxmt > V4_WORKERS / 32) return ANCDE_EUNSUPPORTED;
    const int nks_total = (st.NCP + 15) / 16;
    static const int cand[3] = {8, 4, 2};
    for (int c = 0; c < 3; ++c) {
        const int CS = cand[c];
        if (force_cs && CS != force_cs) continue;
        if (nks_total < CS) continue;
        v->CS = CS;
        v->NS = 8 * CS;
        v->QW = 16 / CS;
        v->nclusters = (st.ntiles + CS - 1) / CS;
        const int base = nks_total / CS, extra = nks_total % CS;
        v->ks_begin[0] = 0;
        for (int r = 0; r < CS; ++r) v->ks_begin[r + 1] = v->ks_begin[r] + base + (r < extra ? 1 : 0);
        for (int r = CS; r < V4_MAX_CS; ++r) v->ks_begin[r + 1] = v->ks_begin[CS];
        v->nks_max = base + (extra ? 1 : 0);
        v->a_bytes = 128 * 16 * v->nks_max * 2;
        v->img_bytes = (int64_t)CS * v->a_bytes;
        v->a_col = 2 * v->NS;
        v->tmem_cols = pow2ceil4(2 * v->NS + 8 * v->nks_max);
        if (v->tmem_cols > 512) continue;
        v->GSTR = v->NS + 4;
        v->GROW = 16 * v->nks_max + 4;
        v->dub = st.du_floats * (st.mode == ANCDE_MODE_TOP ? 2 : 1);
        v->act_ld = st.act_off_du;
        v->KP = 16 * st.KSB + 8;
        v->HP = pad4(st.Hmax);
        {
            // Hidden layers (measured per stage at the Sepsis shape): layers of <= 32 units in FP32 with one (unit,
            // sample) output per thread (4.0 k cycles; a warp per sample: 4.2 k; mma.sync: 4.9 k); wider layers on mma.sync
            // (49 units: 5.0 k; the FP32 form moves 26 128-bit shared-memory loads per output and is bound by the 128 B/clk
            // of the shared-memory return path: 7.1 k).  ANCDE_BWD4_HID = 0 / 1 / 2 overrides (development).
            const char* env = getenv("ANCDE_BWD4_HID");
            int hm = env ? atoi(env) : (chain ? 2 : 1);
            if (hm == 0 && !chain) hm = 2;
            if (hm < 0 || hm > 2) hm = chain ? 2 : 1;
            int in_max = st.Hd;
            for (int l = 0; l < st.n_hidden; ++l) in_max = st.in_dim[l] > in_max ? st.in_dim[l] : in_max;
            v->chain = hm;
            v->chain_warps = hm == 0 ? 8 : (hm == 1 ? maxmt : (in_max * 8 + 31) / 32);
        }
        v->need_att = need_att ? 1 : 0;
        if ((v->dub * 4) % 16 || (v->act_ld * 4) % 16 || (st.act_off_du * 4) % 16 || (st.act_floats * 4) % 16) continue;
        // Operand chunks: whole passes.  Two chunks in two buffers where they fit (the products of the first half run while
        // the second is converted), else as few chunks as fit: every chunk costs the workers a proxy fence and a barrier
        // round, and the issuer its wake-up.
        v->PASSK = 32 / CS;
        v->npass_max = (v->nks_max + v->PASSK - 1) / v->PASSK;
        const char* env_nch = getenv("ANCDE_BWD4_NCH");        // development: operand chunks per stage
        int want = env_nch ? atoi(env_nch) : 2;
        if (want > v->npass_max) want = v->npass_max;
        if (want < 1) want = 1;
        for (int nch = want, first = 1; nch <= v->npass_max; nch = first ? 1 : nch + 1, first = 0) {
            const int PPC = (v->npass_max + nch - 1) / nch;
            const int nchunks = (v->npass_max + PPC - 1) / PPC;
            if (nchunks != nch) continue;
            for (int nslots = (nchunks > 1 ? 2 : 1); nslots >= (first && nchunks > 1 ? 2 : 1); --nslots) {
                const int CK = PPC * v->PASSK;
                v->PPC = PPC;
                v->CK = CK;
                v->nslots = nslots;
                v->nchunks_max = nchunks;
                v->SBOb = CK * 256;
                v->ring_bytes = (2 * v->NS / 8) * v->SBOb;
                auto up = [](size_t x) { return (x + 127) & ~(size_t)127; };
                size_t o = 0;
                v->o_bars = (int)o;    o = up(o + 128);
                v->o_ring = (int)o;    o = up(o + (size_t)nslots * v->ring_bytes);
                v->o_hid = (int)o;     o = up(o + (v->chain != 1 ? (size_t)pad4(st.hiddenb_floats) * 4 : (size_t)st.hb_u4 * 16));
                v->o_recv = (int)o;    o = up(o + (size_t)CS * st.K * 8 * 4);
                v->o_recvatt = (int)o; o = up(o + (size_t)CS * v->QW * 8 * 4);
                v->o_gall = (int)o;    o = up(o + (size_t)(st.Hd + 1) * v->GSTR * 4);
                v->o_du = (int)o;      o = up(o + (size_t)CS * v->dub * 4);
                v->o_act = (int)o;     o = up(o + (size_t)v->act_ld * 4);
                v->o_zb = (int)o;      o = up(o + (size_t)pad4(st.Hd * 8) * 4);
                v->o_bd = (int)o;      o = up(o + (v->chain != 1 ? (size_t)2 * 8 * v->HP * 4 : (size_t)4 * 8 * v->KP * 2));
                v->o_max = (int)o;     o = up(o + 128);
                v->o_lp = (int)o;      o = up(o + 8 * ANCDE_MAX_LAYERS * 4);
                v->o_tab = (int)o;     o = up(o + (size_t)4 * v->PASSK * v->npass_max * 16);
                v->o_gst = (int)o;     o = up(o + (size_t)v->NS * v->GROW * 4);
                v->o_tout = (int)o;    o = up(o + (size_t)pad4(st.n_out) * 4);
                v->smem_total = (int)o;
                // the controls of the next stage are staged in the operand buffer
                if ((size_t)CS * v->dub * 4 > (size_t)nslots * v->ring_bytes) continue;
                if (o <= 227 * 1024) {
                    if (getenv("ANCDE_DEBUG_LAYOUT"))
                        fprintf(stderr, "bwd4 layout: mode %d CS %d nks_max %d passes %d per chunk %d chunks %d slots %d tmem cols %d ring %d smem %d\n",
                                st.mode, CS, v->nks_max, v->npass_max, PPC, nchunks, nslots, v->tmem_cols, v->ring_bytes, v->smem_total);
                    return ANCDE_OK;
                }
            }
        }
    }
    return ANCDE_EUNSUPPORTED;
}

struct Pack4Args {
    NcdeLayout lo;
    V4Layout v;
    const float* Wout;
    unsigned char* img;
};

// A image of rank r, row-major [128 rows][16*nks_max FP16]: unit u -> rows 16*(u/8) + u%8 (hi half) and + 8 (lo half);
// column pl -> reduction index p = 16*ks_begin[r] + pl = i*C4 + j.  Values are 64 * W_out[(i*C + j)][u]; rows without a
// unit and columns behind the slice are zero.
__global__ void pack_weights4_kernel(const __grid_constant__ Pack4Args a)
{
    const NcdeLayout& lo = a.lo;
    const V4Layout& v = a.v;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const int P = 16 * v.nks_max;
    const int64_t total = (int64_t)v.CS * 64 * P;
    for (int64_t idx = gtid; idx < total; idx += gsz) {
        const int pl = (int)(idx % P);
        const int u = (int)((idx / P) % 64);
        const int r = (int)(idx / ((int64_t)P * 64));
        const int pp = 16 * v.ks_begin[r] + pl;
        float w = 0.f;
        if (u < lo.K && pl < 16 * (v.ks_begin[r + 1] - v.ks_begin[r]) && pp < lo.NCP) {
            const int i = pp / lo.C4, j = pp - i * lo.C4;
            if (j < lo.C) w = 64.0f * a.Wout[((int64_t)i * lo.C + j) * lo.K + u];
        }
        const __half h = __float2half_rn(w);
        const __half l = __float2half_rn(w - __half2float(h));
        const int mh = 16 * (u >> 3) + (u & 7), ml = mh + 8;
        __half* base = reinterpret_cast<__half*>(a.img + (size_t)r * v.a_bytes);
        base[(size_t)mh * P + pl] = h;
        base[(size_t)ml * P + pl] = l;
    }
}

int pack_weights4(const ancde_ncde_desc* d, const NcdeLayout& st, const V4Layout& v, unsigned char* img, cudaStream_t stream)
{
    Pack4Args pa;
    pa.lo = st;
    pa.v = v;
    pa.Wout = d->W[d->n_hidden];
    pa.img = img;
    prof_begin("pack_weights", stream);
    pack_weights4_kernel<<<148, 256, 0, stream>>>(pa);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

static void launch_config4(const V4Layout& v, cudaStream_t stream, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at)
{
    *cfg = cudaLaunchConfig_t();
    cfg->gridDim = dim3((unsigned)(v.nclusters * v.CS));
    cfg->blockDim = dim3(V4_THREADS);
    cfg->dynamicSmemBytes = (size_t)v.smem_total;
    cfg->stream = stream;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)v.CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg->attrs = at;
    cfg->numAttrs = 1;
}

// Clusters of this layout that can be resident at the same time (0 when the query fails).  A cluster of 8 CTAs with
// one CTA per SM fits 15 times on a B200 (measured: the 16th cluster of a 1024-sample batch ran as a second wave and
// doubled the kernel time), a cluster of 4 fits 33 times.
static int max_active_clusters4(const V4Layout& v)
{
    static int cache[3][V4_MAX_CS + 1][8] = {};          // [hidden-layer mode][CS][smem / 32 KB] -> clusters + 1
    const int bucket = v.smem_total / (32 * 1024);
    int& slot = cache[v.chain][v.CS][bucket > 7 ? 7 : bucket];
    if (slot) return slot - 1;
    auto kern = v.chain == 0 ? ncde_bwd4_kernel<0> : (v.chain == 1 ? ncde_bwd4_kernel<1> : ncde_bwd4_kernel<2>);
    int n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess) {
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        V4Layout q = v;
        q.smem_total = (bucket + 1) * 32 * 1024 - 1 > 227 * 1024 ? 227 * 1024 : (bucket + 1) * 32 * 1024 - 1;
        q.nclusters = 64;
        launch_config4(q, nullptr, &cfg, at);
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    slot = n + 1;
    return n;
}

// The cluster size whose clusters run in the fewest waves; ties go to clusters of 4, the size every measurement of
// DESIGN.md was taken with (clusters of 8 were not faster where they fit in one wave).
int choose_layout4(const ancde_ncde_desc* d, const NcdeLayout& st, bool need_att, int force_cs, V4Layout* v)
{
    if (!force_cs) {      // development: cluster size per network from the environment
        const char* env = getenv(d->mode == ANCDE_MODE_TOP ? "ANCDE_BWD4_CS_TOP" : "ANCDE_BWD4_CS_BOTTOM");
        if (env && atoi(env) > 0) force_cs = atoi(env);
    }
    if (force_cs) return make_layout4(d, st, need_att, force_cs, v);
    int best_waves = 0;
    V4Layout cand;
    // with the same number of waves the smaller cluster wins: less exchange (Sepsis top network: 2.11 ms with 4 CTAs per
    // cluster, 2.04 ms with 2; bottom 1.653 / 1.645)
    static const int order[3] = {2, 4, 8};
    for (int o = 0; o < 3; ++o) {
        const int cs = order[o];
        if (make_layout4(d, st, need_att, cs, &cand) != ANCDE_OK) continue;
        const int act = max_active_clusters4(cand);
        const int waves = act > 0 ? (cand.nclusters + act - 1) / act : 1 << 20;
        if (best_waves == 0 || waves < best_waves) {
            best_waves = waves;
            *v = cand;
        }
    }
    return best_waves ? ANCDE_OK : ANCDE_EUNSUPPORTED;
}

int launch_bwd4(const SolveArgs& a, const V4Layout& v, const unsigned char* img, cudaStream_t stream)
{
    Bwd4Args ba;
    ba.s = a;
    ba.v = v;
    ba.img = img;
    auto kern = v.chain == 0 ? ncde_bwd4_kernel<0> : (v.chain == 1 ? ncde_bwd4_kernel<1> : ncde_bwd4_kernel<2>);
    ANCDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem_total));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    launch_config4(v, stream, &cfg, at);
    static const char* names[3] = {"ncde_bwd_plain", "ncde_bwd_bottom", "ncde_bwd_top"};
    prof_begin(names[a.lo.mode], stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ba);
    prof_end(stream);
    count_launch();
    if (e != cudaSuccess) {
        set_error("ncde_backward: cluster launch failed: %s (cluster %d, %d bytes of shared memory)", cudaGetErrorString(e), v.CS, v.smem_total);
        return ANCDE_ECUDA;
    }
    return ANCDE_OK;
}

}  // namespace ancde
