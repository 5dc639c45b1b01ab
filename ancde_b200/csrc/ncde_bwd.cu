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
#include <mutex>

#include "mma_common.cuh"
#include "bwd_hidden.cuh"
#include "ncde_bwd4.cuh"
#include "ncde_lane.cuh"

namespace ancde {

int fill_args(const ancde_ncde_desc* d, const NcdeLayout& lo, SolveArgs* a);
int pack_weights(const ancde_ncde_desc* d, const NcdeLayout& lo, cudaStream_t stream);
int engine_of(const ancde_ncde_desc* d);
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

constexpr int BW_MTK_MAX = 4;      // output-layer reduction size K <= 64

// The output layer's vector-Jacobian product runs on the tensor cores: per TILE of 16 output rows (2 field rows x 8
// control channels, see NcdeLayout) a warp builds the pre-activation cotangent of its 8 samples directly in
// B-fragment registers from the saved tanh outputs, splits it into FP16 hi/lo halves (scaled by a power of two taken
// from the largest stage cotangent of the tile) and multiplies it with the W_out^T fragments that a producer warp
// streams from L2 through a ring of shared-memory slots (TMA bulk copies).  CT tiles per slot, one per warp of a
// group of CT warps; two groups work on alternate slots.
// CHAIN: every hidden layer has at most 32 units -> FP32 warp-synchronous hidden_chain_bwd; wider layers go through
// the tensor cores (mma.sync, FP16 hi/lo x3), one CTA barrier per layer.
template <int TS, bool CHAIN>
__global__ void __launch_bounds__(512, 1) ncde_bwd_kernel(const __grid_constant__ SolveArgs p)
{
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int NT = p.ntc;
    const int NW = NT / 32;
    const int tile = blockIdx.x;
    const int b0 = tile * TS;
    const int Hd = lo.Hd, C4 = lo.C4, NCP = lo.NCP, DUS = lo.DUS, MTK = lo.MTK, NJC = lo.NJC;
    const int CT = p.ring_kc;                          // tiles per ring slot = warps per group (NW = 2 * CT)
    const int NSLOT = p.ring_slots;
    const int ntl = lo.n_wtiles;
    const int nchunks = (ntl + CT - 1) / CT;
    const int HT = pad4(Hd * TS);
    const bool need_att = (lo.mode == ANCDE_MODE_TOP) && (p.grad_attention != nullptr);
    // timewise attention (one value per sample and knot): its cotangent is ONE number per sample and stage, so every lane
    // sums its share in a register while it walks the output-layer tiles -- no per-channel partial sums in shared memory
    const bool att_tw = need_att && lo.att_C == 1;
    const int64_t g_tile = (int64_t)TS * NCP;           // floats per (stage, tile) block of saved tanh outputs

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* bar_empty = bar_full + 8;
    uint64_t* bar_act = bar_full + 16;                   // [2] saved activations of a stage have landed (TMA bulk copy)
    float* smem = reinterpret_cast<float*>(smem_raw + 160);
    const uint4* sWhb = reinterpret_cast<const uint4*>(smem);   // !CHAIN: hidden layers, W_l^T as FP16 hi/lo A fragments
    const float* sWtb = smem;                            // CHAIN: transposed weights Wt_l[k][o] (FP32)
    float* sp = smem + (CHAIN ? (size_t)pad4(lo.hiddenb_floats) : (size_t)lo.hb_u4 * 4);
    uint4* sRing = reinterpret_cast<uint4*>(sp);         // [NSLOT][CT][MTK][64] W_out^T fragments
    sp += (size_t)NSLOT * CT * MTK * 64 * 4;
    float* sActBuf = sp; sp += 2 * (size_t)lo.act_fwd_floats;   // saved activations, double buffered
    float* sG = sp;     sp += HT;                        // cotangent of the stage derivative
    float* sLam = sp;   sp += HT;                        // cotangent of z at the end of the step
    float* sMu = sp;    sp += HT;                        // pending cotangent of z at the start
    float* sV4 = sp;    sp += HT;
    float* sV3 = sp;    sp += HT;
    float* sV2 = sp;    sp += HT;
    float* sZb = sp;    sp += HT;                        // cotangent of the stage input
    const int KP = 16 * lo.KSB + 8;                      // halfs per row of the cotangent operand matrices [sample][feature]
    __half* sBd = reinterpret_cast<__half*>(sp);         // !CHAIN: two buffers of {hi, lo} [8][KP]
    const int HP = pad4(lo.Hmax);                        // CHAIN: floats per sample row of the chain's buffers
    float* sD0 = sp;                                     // CHAIN: hidden-layer cotangents ping [sample][unit]
    float* sD1 = sp + TS * HP;                           //        pong
    sp += CHAIN ? (size_t)2 * TS * HP : (size_t)(4 * 8 * KP) / 2;
    float* sHbP = sp;   sp += (size_t)NW * MTK * 128;    // per-warp partial sums of the hidden cotangent (fragment order)
    float* sAttW = sp;  sp += need_att ? (size_t)NW * NJC * 64 : 0;   // per-warp partial sums of the dU cotangent [jc][b][8 j]
    float* sMax = sp;   sp += 32;
    int* sLp = reinterpret_cast<int*>(sp); sp += 8 * ANCDE_MAX_LAYERS;   // per-layer {mtiles, full, tail8, stride, off, M, act_off_h[l-1], act_off_delta[l-1]}
    float* sTmp = sp;   sp += (size_t)C4 * TS;           // [C4][TS]
    int4* sTile = reinterpret_cast<int4*>(sp);           // per output-layer tile: {G offset, stage-cotangent offset, dU offset, flags}

    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, CT);                // one arrive per warp of the group that owns the slot's chunk
        }
        mbar_init(bar_act + 0, 1);
        mbar_init(bar_act + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid >= NT) {
        // ---- producer warp: W_out^T tiles through the ring + L2 prefetch of the tanh outputs
        if (lane == 0) {
            const uint4* Wg = reinterpret_cast<const uint4*>(p.wpack + lo.off_wob);
            int slot = 0;
            uint32_t phase = 1;
            for (int n = lo.n_stages - 1; n >= 0; --n) {
                if (n >= 1)
                    bulk_prefetch_l2(p.save_G + ((int64_t)(n - 1) * lo.ntiles + tile) * g_tile, (uint32_t)(g_tile * 4));
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(bar_empty + slot, phase);
                    const int tiles = (ntl - c * CT < CT) ? ntl - c * CT : CT;
                    const uint32_t bytes = (uint32_t)(tiles * MTK * 64 * 16);
                    mbar_arrive_expect_tx(bar_full + slot, bytes);
                    bulk_g2s(sRing + (size_t)slot * CT * MTK * 64, Wg + (size_t)c * CT * MTK * 64, bytes, bar_full + slot);
                    if (++slot == NSLOT) { slot = 0; phase ^= 1; }
                }
            }
        }
        return;
    }
    auto cta_sync = [&]() { named_bar_sync(1, NT); };
    auto act_block = [&](int n) { return p.save_act + ((int64_t)n * lo.ntiles + tile) * lo.act_floats; };
    // The block of saved activations of a stage is one contiguous, 16-byte aligned run: ONE thread stages it with a TMA
    // bulk copy on the buffer's mbarrier (was: ~800 cp.async spread over the CTA plus a wait_group per stage).
    auto issue_act = [&](int n, int buf) {
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(lo.act_fwd_floats * 4);
            mbar_arrive_expect_tx(bar_act + buf, bytes);
            bulk_g2s(sActBuf + (size_t)buf * lo.act_fwd_floats, act_block(n), bytes, bar_act + buf);
        }
    };
    uint32_t act_parity = 0u;                            // bit b: parity of the next completion of bar_act[b]

    {
        if constexpr (CHAIN) {
            const float4* src = reinterpret_cast<const float4*>(p.wpack + lo.off_hiddenb);
            float4* dst = reinterpret_cast<float4*>(smem);
            for (int i = tid; i < lo.hiddenb_floats / 4; i += NT) dst[i] = __ldg(src + i);
            for (int i = tid; i < 2 * TS * HP; i += NT) sD0[i] = 0.f;    // sD0, sD1 (contiguous)
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(p.wpack + lo.off_whb);
            uint4* dst = reinterpret_cast<uint4*>(smem);
            for (int i = tid; i < lo.hb_u4; i += NT) dst[i] = __ldg(src + i);
            uint32_t* zb = reinterpret_cast<uint32_t*>(sBd);
            for (int i = tid; i < 2 * 8 * KP; i += NT) zb[i] = 0u;      // 4 matrices of 8*KP halfs
        }
        for (int item = tid; item < Hd * TS; item += NT) {
            sLam[item] = 0.f; sMu[item] = 0.f;
            sV4[item] = 0.f; sV3[item] = 0.f; sV2[item] = 0.f;       // read by stage 0 of the single-stage mode
        }
        if (tid < 32) sMax[tid] = 0.f;
        if (tid < lo.n_hidden) {
            // the layer loop reads its parameters from shared memory: register-indexed loads from the kernel-parameter
            // constant bank cost hundreds of cycles each on the critical path of every layer
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
        // Coordinates of the output-layer tiles, worked out once: tile tl = (ip, jc) covers field rows 2*ip, 2*ip + 1 and
        // control channels 8*jc .. 8*jc + 7; lane (g, t) owns sample g and channels 8*jc + 2t, + 1.
        //   x = offset of (row 2*ip, channel 8*jc) in a sample's saved tanh outputs,  y = offset of row 2*ip in [Hd][TS],
        //   z = dU offset of channel 8*jc,  w = bit t: channel 8*jc + 2t exists, bit 4: row 2*ip + 1 exists, bits 8..: jc
        for (int tl = tid; tl < ntl; tl += NT) {
            const int ip = tl / NJC, jc = tl - ip * NJC;
            int flags = jc << 8;
            for (int tt = 0; tt < 4; ++tt)
                if (8 * jc + 2 * tt < C4) flags |= 1 << tt;
            if (2 * ip + 1 < Hd) flags |= 16;
            sTile[tl] = make_int4(2 * ip * C4 + 8 * jc, 2 * ip * TS, 2 * jc * DUS, flags);
        }
        issue_act(lo.n_stages - 1, 0);
    }
    cta_sync();

    const int lane_g = g * NCP + 2 * t;                                   // this lane's offset inside a tile's tanh outputs
    const int lane_du = (t >> 1) * DUS + 2 * (t & 1) * TS + g;            // ... and inside a dU chunk pair (du_index of channel 2t)
    const bool b_ok = g < TS;
    const uint32_t bar_full_a = smem_u32(bar_full), bar_empty_a = smem_u32(bar_empty);
    int jb = lo.n_out - 1;   // last output time whose cotangent has not been consumed yet
    int abuf = 0;            // which half of sActBuf holds the current stage
    float gmax = 0.f;        // largest |stage cotangent| this thread wrote (scale of the FP16 weight-gradient operands)
    const int grp = warp / CT, wl = warp - grp * CT;     // this warp's group and tile slot inside a chunk
    int ring_base = 0;       // ring position of this stage's chunk 0, modulo 2*NSLOT (slot and phase parity in one number)
    PH2_DECL();

    for (int k = lo.n_steps - 1; k >= 0; --k) {
        const float t0 = grid_time(k, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float t1 = grid_time(k + 1, lo.n_steps, p.t_start, p.t_end, p.step, p.grid);
        const float dt = __fsub_rn(t1, t0);
        // ---- cotangents of the outputs emitted during this step: t_out[j] in (t0, t1]
        if (lo.n_stages > 1) {
            for (int item = tid; item < Hd * TS; item += NT) {
                const int i = item / TS, b = item - i * TS;
                const int bg = b0 + b;
                float lam = sLam[item], mu = sMu[item];
                for (int j = jb; j >= 1; --j) {
                    const float tj = __ldg(p.t_out + j);
                    if (!(tj > t0)) break;
                    const float gz = (bg < lo.B) ? __ldg(p.grad_z_out + ((int64_t)j * lo.B + bg) * Hd + i) : 0.f;
                    if (tj == t1) lam += gz;
                    else {
                        const float th = (tj - t0) / (t1 - t0);
                        lam += th * gz;
                        mu += (1.0f - th) * gz;
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
            if (n >= lo.n_stages) continue;          // single-stage mode: stage 0 only
            float* act = act_block(n);
            float* sAct = sActBuf + (size_t)abuf * lo.act_fwd_floats;
            PH2_T0();
            // ---- cotangent of the stage derivative k_{s+1}  (transpose of rk4_alt_step_func)
            float lmax = 0.f;
            for (int item = tid; item < Hd * TS; item += NT) {
                const float lam = sLam[item];
                float gv;
                if (s == 3) gv = dt * 0.125f * lam;
                else if (s == 2) gv = dt * (0.375f * lam + sV4[item]);
                else if (s == 1) gv = dt * (0.375f * lam - sV4[item] + sV3[item]);
                else gv = dt * (0.125f * lam + sV4[item] + (sV2[item] - sV3[item]) * (1.0f / 3.0f));
                if (lo.n_stages == 1) {              // the cotangent of the one stage derivative comes from the caller
                    const int i = item / TS, b = item - i * TS;
                    gv = (b0 + b < lo.B) ? __ldg(p.grad_k_out + (int64_t)(b0 + b) * Hd + i) : 0.f;
                }
                sG[item] = gv;
                act[lo.act_off_gbar + item] = gv;        // the stage cotangent goes to HBM for the weight-gradient kernels
                lmax = fmaxf(lmax, fabsf(gv));
            }
            gmax = fmaxf(gmax, lmax);
            for (int o = 16; o >= 1; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
            if (lane == 0) sMax[warp] = lmax;
            if (need_att && !att_tw) {
                float2* z = reinterpret_cast<float2*>(sAttW + (size_t)warp * NJC * 64) + lane;
                for (int jc = 0; jc < NJC; ++jc) z[jc * 32] = make_float2(0.f, 0.f);
            }
            mbar_wait(bar_act + abuf, (act_parity >> abuf) & 1u);      // saved activations of this stage have landed
            act_parity ^= 1u << abuf;
            cta_sync();
            // prefetch the next stage's activations (this kernel walks n downwards)
            if (n >= 1) issue_act(n - 1, abuf ^ 1);
            // power-of-two scale that brings the largest cotangent of this tile and stage to ~2^6
            float scale = 1.0f;
            {
                // 16 slots (NW <= 14, the rest stay zero): four independent 128-bit loads instead of a dependent loop
                const float4 m0 = *reinterpret_cast<const float4*>(sMax), m1 = *reinterpret_cast<const float4*>(sMax + 4);
                const float4 m2 = *reinterpret_cast<const float4*>(sMax + 8), m3 = *reinterpret_cast<const float4*>(sMax + 12);
                const float m = fmaxf(fmaxf(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m0.z, m0.w)), fmaxf(fmaxf(m1.x, m1.y), fmaxf(m1.z, m1.w))),
                                      fmaxf(fmaxf(fmaxf(m2.x, m2.y), fmaxf(m2.z, m2.w)), fmaxf(fmaxf(m3.x, m3.y), fmaxf(m3.z, m3.w))));
                if (m > 1e-30f && m < 1e30f) {
                    const int e = (int)((__float_as_uint(m) >> 23) & 0xffu) - 126;     // m = f * 2^e, f in [0.5, 1)
                    scale = __uint_as_float((uint32_t)(127 + 6 - e) << 23);
                }
            }
            const float* sDU = sAct + lo.act_off_du;
            PH2_MARK(0);      // stage cotangent + saved activations

            // ---- phases 1+2: G -> pre_bar (registers, B fragments) -> hbar[k][b] += W_out^T pre_bar on the tensor cores
            float acc[BW_MTK_MAX][4];
#pragma unroll
            for (int m = 0; m < BW_MTK_MAX; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][e] = 0.f;
            {
                const float* gbase = p.save_G + ((int64_t)n * lo.ntiles + tile) * g_tile + lane_g;
                auto load_g = [&](int tl, float2& ga, float2& gb2) {
                    ga = make_float2(0.f, 0.f);
                    gb2 = make_float2(0.f, 0.f);
                    if (tl < ntl && b_ok) {
                        const int4 e = sTile[tl];
                        if ((e.w >> t) & 1) {
                            const float* gp = gbase + e.x;
                            ga = __ldg(reinterpret_cast<const float2*>(gp));
                            if (e.w & 16) gb2 = __ldg(reinterpret_cast<const float2*>(gp + C4));
                        }
                    }
                };
                // tanh outputs of the next two tiles of this warp are in flight while the current one is processed
                float2 g0n, g1n, g0nn, g1nn;
                load_g(grp * CT + wl, g0n, g1n);
                load_g(grp * CT + wl + 2 * CT, g0nn, g1nn);
                int rpos = (ring_base + grp) % (2 * NSLOT);
                float2* const attw = reinterpret_cast<float2*>(sAttW + (size_t)warp * NJC * 64) + lane;
                const float* const sGl = sG + g;
                const float* const sDUl = sDU + lane_du;
                const float* const sDAl = sAct + lo.act_off_duda + lane_du;      // dU/da of this lane's channels
                float att_acc = 0.f;
                const uint4* const ringl = sRing + (size_t)wl * MTK * 64 + lane;
                for (int c = grp; c < nchunks; c += 2) {
                    const int tl = c * CT + wl;
                    const float2 g0 = g0n, g1 = g1n;
                    g0n = g0nn; g1n = g1nn;
                    load_g(tl + 4 * CT, g0nn, g1nn);
                    const int slot = (rpos >= NSLOT) ? rpos - NSLOT : rpos;
                    const uint32_t parity = (rpos >= NSLOT) ? 1u : 0u;
                    rpos += 2;
                    if (rpos >= 2 * NSLOT) rpos -= 2 * NSLOT;
                    mbar_wait_a(bar_full_a + 8u * (uint32_t)slot, parity);
                    if (tl < ntl) {
                        float p00 = 0.f, p01 = 0.f, p10 = 0.f, p11 = 0.f;
                        const int4 e = sTile[tl];
                        if (b_ok && ((e.w >> t) & 1)) {
                            const float gb0 = sGl[e.y] * scale;
                            const float gb1 = (e.w & 16) ? sGl[e.y + TS] * scale : 0.f;
                            const float du0 = sDUl[e.z], du1 = sDUl[e.z + TS];
                            p00 = gb0 * du0 * fmaf(-g0.x, g0.x, 1.0f);
                            p01 = gb0 * du1 * fmaf(-g0.y, g0.y, 1.0f);
                            p10 = gb1 * du0 * fmaf(-g1.x, g1.x, 1.0f);
                            p11 = gb1 * du1 * fmaf(-g1.y, g1.y, 1.0f);
                            if (att_tw) {
                                // sum_j dUbar[j][b] * dU/da[j][b] with dUbar[j][b] = sum_i gbar[i][b] * G[i][j][b]
                                att_acc = fmaf(gb0 * g0.x + gb1 * g1.x, sDAl[e.z], att_acc);
                                att_acc = fmaf(gb0 * g0.y + gb1 * g1.y, sDAl[e.z + TS], att_acc);
                            } else if (need_att) {
                                // cotangent of dU: dUbar[j][b] = sum_i gbar[i][b] * G[i][j][b] (scaled like gbar)
                                float2* q = attw + (e.w >> 8) * 32;
                                float2 v = *q;
                                v.x += gb0 * g0.x + gb1 * g1.x;
                                v.y += gb0 * g0.y + gb1 * g1.y;
                                *q = v;
                            }
                        }
                        uint32_t bh0, bl0, bh1, bl1;
                        split_pair(p00, p01, bh0, bl0);      // rows 2t, 2t+1   = (i0, j), (i0, j+1)
                        split_pair(p10, p11, bh1, bl1);      // rows 2t+8, 2t+9 = (i0+1, j), (i0+1, j+1)
                        const uint4* A = ringl + (size_t)slot * CT * MTK * 64;
#pragma unroll
                        for (int m = 0; m < BW_MTK_MAX; ++m) {
                            if (m < MTK) {
                                const uint4 ah = A[m * 64], al = A[m * 64 + 32];
                                mma16x3(acc[m], ah, al, bh0, bh1, bl0, bl1);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(bar_empty_a + 8u * (uint32_t)slot);
                }
                if (att_tw) {
                    att_acc += __shfl_xor_sync(0xffffffffu, att_acc, 1);
                    att_acc += __shfl_xor_sync(0xffffffffu, att_acc, 2);
                    if (t == 0) sAttW[warp * 8 + g] = att_acc;              // this warp's share for sample g
                }
                ring_base = (ring_base + nchunks) % (2 * NSLOT);
                float4* hp4 = reinterpret_cast<float4*>(sHbP) + (size_t)warp * MTK * 32 + lane;
#pragma unroll
                for (int m = 0; m < BW_MTK_MAX; ++m)
                    if (m < MTK) hp4[m * 32] = make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]);
            }
            cta_sync();
            PH2_MARK(1);      // phases 1+2
            // ---- cotangent of the attention: a[q] += dUbar * dU/da   
            if (att_tw) {
                if (warp == 0) {
                    // lane = (quarter of the warps' partial sums, sample); two shuffles finish the sum
                    const int smp = lane & 7;
                    float a = 0.f;
                    for (int w = lane >> 3; w < NW; w += 4) a += sAttW[w * 8 + smp];
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    a += __shfl_xor_sync(0xffffffffu, a, 16);
                    const float ts = stage_time(t0, dt, s);
                    int qa = (int)floorf(ts) - 1;
                    if (qa < 0) qa += lo.att_L;
                    qa = qa < 0 ? 0 : (qa >= lo.att_L ? lo.att_L - 1 : qa);
                    // scale is a power of two: its reciprocal is an exponent flip
                    const float inv_scale = __uint_as_float((254u << 23) - __float_as_uint(scale));
                    if (lane < TS && b0 + lane < lo.B)
                        atomicAdd(p.grad_attention + (int64_t)qa * lo.B + b0 + lane, a * inv_scale);
                }
            } else if (need_att) {
                const float inv_s = 1.0f / scale;
                for (int item = tid; item < lo.C * TS; item += NT) {
                    const int j = item / TS, b = item - j * TS;
                    float a = 0.f;
                    for (int w = 0; w < NW; ++w) a += sAttW[(((size_t)w * NJC + (j >> 3)) * 8 + b) * 8 + (j & 7)];
                    sTmp[j * TS + b] = a * inv_s * sAct[lo.act_off_duda + du_index(j, b, TS, DUS)];
                }
                cta_sync();
                const float ts = stage_time(t0, dt, s);
                int qa = (int)floorf(ts) - 1;
                if (qa < 0) qa += lo.att_L;
                qa = qa < 0 ? 0 : (qa >= lo.att_L ? lo.att_L - 1 : qa);
                if (lo.att_C == 1) {
                    // timewise attention: one cotangent per sample = sum over the channels; warp b sums sample b
                    for (int b = warp; b < TS; b += NW) {
                        float a = 0.f;
                        for (int j = lane; j < lo.C; j += 32) a += sTmp[j * TS + b];
                        for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                        if (lane == 0 && b0 + b < lo.B) atomicAdd(p.grad_attention + (int64_t)qa * lo.B + b0 + b, a);
                    }
                } else {
                    for (int item = tid; item < lo.C * TS; item += NT) {
                        const int b = item / lo.C, j = item - b * lo.C;
                        if (b0 + b < lo.B)
                            atomicAdd(p.grad_attention + ((int64_t)qa * lo.B + b0 + b) * lo.C + j, sTmp[j * TS + b]);
                    }
                }
            }
            PH2_MARK(2);      // attention
            if constexpr (CHAIN) {
                // ---- phase 3: hidden layers backwards (FP32, warp-synchronous chain):
                // delta_l = (W_{l+1}^T delta_{l+1}) * relu'(h_l), the stage-input cotangent is W_0^T delta_0
                {
                    // delta of the last hidden layer from the warps' partial accumulator fragments: element (o, b) lives in
                    // m-tile o/16, lane (o%8)*4 + b/2, register ((o%16)/8)*2 + b%2; the tensor-core operands carried the
                    // factors 64 (weights) and `scale` (cotangent)
                    const float unscale = (1.0f / 64.0f) / scale;
                    const int l = lo.n_hidden - 1;
                    const int N = lo.width[l];
                    const float* hl = sAct + lo.act_off_h[l];
                    for (int item = tid; item < N * TS; item += NT) {
                        const int o = item / TS, b = item - o * TS;
                        const int idx = ((o >> 4) * 32 + (o & 7) * 4 + (b >> 1)) * 4 + ((o >> 3) & 1) * 2 + (b & 1);
                        float v = 0.f;
                        for (int w = 0; w < NW; ++w) v += sHbP[(size_t)w * MTK * 128 + idx];
                        v = (hl[item] > 0.f) ? v * unscale : 0.f;
                        act[lo.act_off_delta[l] + item] = v;     // for the hidden weight-gradient kernel
                        sD0[b * HP + o] = v;
                    }
                }
                cta_sync();
                hidden_chain_bwd<TS>(sLp, lo.n_hidden, sWtb, sD0, sD1, sAct, act, sZb, HP, warp, NW, lane);
                cta_sync();
            } else {
                // ---- phase 3: hidden layers backwards on the tensor cores.  Cotangents stay multiplied by `scale` inside the
                // chain (FP16 range); delta_l = (W_{l+1}^T delta_{l+1}) * relu'(h_l), the stage-input cotangent is W_0^T delta_0
                const float inv_s = 1.0f / scale;
                const uint32_t bd_base = smem_u32(sBd);
                const uint32_t bd_mat = (uint32_t)(8 * KP * 2);          // bytes of one FP16 matrix
                const uint32_t bd_lane = (uint32_t)(((lane & 7) * KP + ((lane >> 3) & 1) * 8) * 2) + ((lane >> 4) ? bd_mat : 0u);
                {
                    // delta of the last hidden layer from the warps' partial accumulator fragments: element (o, b) lives in
                    // m-tile o/16, lane (o%8)*4 + b/2, register ((o%16)/8)*2 + b%2
                    const int l = lo.n_hidden - 1;
                    const int N = lo.width[l];
                    const float* hl = sAct + lo.act_off_h[l];
                    for (int item = tid; item < N * TS; item += NT) {
                        const int o = item / TS, b = item - o * TS;
                        const int idx = ((o >> 4) * 32 + (o & 7) * 4 + (b >> 1)) * 4 + ((o >> 3) & 1) * 2 + (b & 1);
                        float v = 0.f;
                        for (int w = 0; w < NW; ++w) v += sHbP[(size_t)w * MTK * 128 + idx];
                        v = (hl[item] > 0.f) ? v * (1.0f / 64.0f) : 0.f;
                        act[lo.act_off_delta[l] + item] = v * inv_s;     // for the hidden weight-gradient kernel
                        const __half hi = __float2half_rn(v);
                        sBd[b * KP + o] = hi;
                        sBd[8 * KP + b * KP + o] = __float2half_rn(v - __half2float(hi));
                    }
                }
                cta_sync();
                int cur = 0;
                for (int l = lo.n_hidden - 1; l >= 0; --l) {
                    const int4 lp0 = *reinterpret_cast<const int4*>(sLp + 8 * l), lp1 = *reinterpret_cast<const int4*>(sLp + 8 * l + 4);
                    const int l_mtiles = lp0.x, l_full = lp0.y, l_tail8 = lp0.z, l_stride = lp0.w, l_off = lp1.x, l_M = lp1.y;
                    if (warp < l_mtiles) {
                        const int mt = warp;
                        const uint4* A = sWhb + l_off + mt * l_stride + lane;
                        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
                        uint32_t baddr = bd_base + (uint32_t)cur * 2u * bd_mat + bd_lane;
    #pragma unroll 4
                        for (int ks = 0; ks < l_full; ++ks) {
                            const uint4 ah = A[ks * 64], al = A[ks * 64 + 32];
                            uint32_t bh0, bh1, bl0, bl1;
                            ldmatrix_x4(baddr, bh0, bh1, bl0, bl1);
                            baddr += 32;
                            mma16(a0, al.x, al.y, al.z, al.w, bh0, bh1);
                            mma16(a1, ah.x, ah.y, ah.z, ah.w, bl0, bl1);
                            mma16(a2, ah.x, ah.y, ah.z, ah.w, bh0, bh1);
                        }
                        if (l_tail8) {
                            const uint4 a = A[l_full * 64];
                            uint32_t bh0, bh1, bl0, bl1;
                            ldmatrix_x4(baddr, bh0, bh1, bl0, bl1);
                            mma8(a0, a.z, a.w, bh0);
                            mma8(a1, a.x, a.y, bl0);
                            mma8(a2, a.x, a.y, bh0);
                        }
                        // rows = inputs of layer l (16*mt + g, + 8), columns = samples 2t, 2t+1
                        float dv[4];
    #pragma unroll
                        for (int e = 0; e < 4; ++e) dv[e] = (a0[e] + a1[e] + a2[e]) * (1.0f / 64.0f);
                        const int bcol = 2 * t;
                        if (l > 0) {
                            const float* hl = sAct + lp1.z;
                            float* dsave = act + lp1.w;
    #pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int kin = mt * 16 + g + (e >> 1) * 8, b = bcol + (e & 1);
                                const bool ok = kin < l_M && b < TS;
                                const float hv = ok ? hl[kin * TS + b] : 0.f;
                                dv[e] = (hv > 0.f) ? dv[e] : 0.f;
                                if (ok) dsave[kin * TS + b] = dv[e] * inv_s;
                            }
                            store_act_tile(bd_base + (uint32_t)(cur ^ 1) * 2u * bd_mat + bd_lane + (uint32_t)(mt * 16 * 2), dv);
                        } else {
    #pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int kin = mt * 16 + g + (e >> 1) * 8, b = bcol + (e & 1);
                                if (kin < l_M && b < TS) sZb[kin * TS + b] = dv[e] * inv_s;
                            }
                        }
                    }
                    cta_sync();
                    cur ^= 1;
                }
            }
            const float* hb = sZb;
            PH2_MARK(3);      // phase 3: hidden layers
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
            PH2_MARK(4);      // RK4 adjoint combine + barrier
        }
    }
    PH2_FLUSH(lo.mode == ANCDE_MODE_TOP ? 40 : 32);
    // ---- results
    for (int o = 16; o >= 1; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0 && gmax > 0.f && gmax < 3.0e38f)
        atomicMax(reinterpret_cast<unsigned*>(const_cast<float*>(p.wpack) + lo.off_gscale), __float_as_uint(gmax));
    for (int item = tid; item < Hd * TS; item += NT) {
        const int i = item / TS, b = item - i * TS;
        const int bg = b0 + b;
        if (bg < lo.B)
            p.grad_z0[(int64_t)bg * Hd + i] = sLam[item] + (lo.n_stages > 1 ? __ldg(p.grad_z_out + (int64_t)bg * Hd + i) : 0.f);
    }
}

// consumer warps of the backward: two groups of CT warps (CT tiles of the output layer per ring slot)
static int bwd_ct(const NcdeLayout& lo)
{
    int ct = (lo.n_wtiles + 1) / 2;
    if (ct > 7) ct = 7;
    if (ct < 1) ct = 1;
    // phase 3 gives every m-tile (16 inputs) of a hidden layer its own warp
    int maxmt = 1;
    for (int l = 0; l < lo.n_hidden; ++l) maxmt = lo.hbp[l].mtiles > maxmt ? lo.hbp[l].mtiles : maxmt;
    if (2 * ct < maxmt) ct = (maxmt + 1) / 2;
    return ct;
}

static bool bwd_chain(const NcdeLayout& lo)
{
    for (int l = 0; l < lo.n_hidden; ++l)
        if (lo.width[l] > 32) return false;
    return true;
}

static size_t bwd_smem_bytes(const NcdeLayout& lo, int ct, int slots, bool att)
{
    const int nw = 2 * ct;
    const bool chain = bwd_chain(lo);
    size_t f = chain ? (size_t)pad4(lo.hiddenb_floats) : (size_t)lo.hb_u4 * 4;
    f += (size_t)slots * ct * lo.MTK * 64 * 4;
    f += 2 * (size_t)lo.act_fwd_floats + 7 * (size_t)pad4(lo.Hd * lo.TS) +
         (chain ? 2 * (size_t)lo.TS * pad4(lo.Hmax) : (size_t)(4 * 8 * (16 * lo.KSB + 8)) / 2);
    f += (size_t)nw * lo.MTK * 128 + (att ? (size_t)nw * lo.NJC * 64 : 0) + 32 + 8 * ANCDE_MAX_LAYERS + (size_t)lo.C4 * lo.TS + 4 * (size_t)lo.n_wtiles;
    return f * 4 + 160;
}

// Used by make_layout's tile-size choice: does the backward fit with this many samples per CTA?
bool bwd_fits(const NcdeLayout& lo)
{
    return lo.MTK <= BW_MTK_MAX && bwd_smem_bytes(lo, bwd_ct(lo), 2, lo.mode == ANCDE_MODE_TOP) <= 227 * 1024;
}

// Side stream of the asynchronous weight gradients (ancde_ncde_desc::async_wgrad), one per device.
// (Measured alternative: deferring the launches until the NEXT recurrent kernel has been enqueued, so that it gets its
// SMs first, stretches the persistent tcgen05 weight-gradient kernel over the whole recurrent kernel on the 20 free SMs
// and its last wave then runs alone: 11.49 ms per step against 11.40 ms with the immediate launch below.)
struct SideStream {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, done = nullptr;
    bool pending = false;
};
static SideStream g_side[64];
static std::mutex g_side_mutex;       // creation and the pending flag; the library is otherwise single-threaded per device (header)

static int side_stream(SideStream** out)
{
    int dev = 0;
    ANCDE_CUDA(cudaGetDevice(&dev));
    ANCDE_CHECK_ARG(dev >= 0 && dev < 64, "ncde_backward: device index %d", dev);
    SideStream* ss = &g_side[dev];
    std::lock_guard<std::mutex> lock(g_side_mutex);
    if (!ss->s) {
        ANCDE_CUDA(cudaStreamCreateWithFlags(&ss->s, cudaStreamNonBlocking));
        ANCDE_CUDA(cudaEventCreateWithFlags(&ss->fork, cudaEventDisableTiming));
        ANCDE_CUDA(cudaEventCreateWithFlags(&ss->done, cudaEventDisableTiming));
    }
    *out = ss;
    return ANCDE_OK;
}

// The weight-gradient kernels of a backward call (they read the stage / layer cotangents the recurrent kernel left in
// save_act and depend on nothing else).
static int launch_wgrads(const ancde_ncde_desc* d, const SolveArgs& a, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    // With async_wgrad they go to the side stream and run next to whatever the caller enqueues after
    // this call -- in a training step the recurrent backward of the OTHER network, which leaves 20 of the 148 SMs idle
    {
        cudaStream_t ws = stream;
        SideStream* ss = nullptr;
        if (d->async_wgrad) {
            int rc = side_stream(&ss);
            if (rc != ANCDE_OK) return rc;
            ANCDE_CUDA(cudaEventRecord(ss->fork, stream));
            ANCDE_CUDA(cudaStreamWaitEvent(ss->s, ss->fork, 0));
            ws = ss->s;
        }
        int rc = launch_wgrad(a, d->grad_W[lo.n_hidden], d->grad_bias[lo.n_hidden], ws);
        if (rc != ANCDE_OK) return rc;
        HiddenGradPtrs hp;
        for (int l = 0; l < lo.n_hidden; ++l) { hp.gW[l] = d->grad_W[l]; hp.gB[l] = d->grad_bias[l]; }
        rc = launch_hidden_wgrad(a, hp, ws);
        if (rc != ANCDE_OK) return rc;
        if (ss) {
            ANCDE_CUDA(cudaEventRecord(ss->done, ss->s));
            std::lock_guard<std::mutex> lock(g_side_mutex);
            ss->pending = true;
        }
    }
    return ANCDE_OK;
}

template <int TS>
static int launch_bwd(const ancde_ncde_desc* d, SolveArgs& a, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    const size_t max_smem = 227 * 1024;
    if (lo.MTK > BW_MTK_MAX) {
        set_error("ncde_backward: last hidden width %d exceeds %d", lo.K, 16 * BW_MTK_MAX);
        return ANCDE_EUNSUPPORTED;
    }
    const int ct = bwd_ct(lo);
    if (2 * ct > 16) {
        set_error("ncde_backward: %d consumer warps (hidden layers wider than 256 units are not supported)", 2 * ct);
        return ANCDE_EUNSUPPORTED;
    }
    const int nt = 2 * ct * 32;
    const bool att = (lo.mode == ANCDE_MODE_TOP) && (a.grad_attention != nullptr);
    int slots = 8;      // as deep a ring as fits: the fills are latency bound (~1.5 us each)
    while (slots > 2 && bwd_smem_bytes(lo, ct, slots, att) > max_smem) --slots;
    const size_t smem = bwd_smem_bytes(lo, ct, slots, att);
    if (smem > max_smem) {
        set_error("ncde_backward: needs %zu bytes of shared memory per CTA (max %zu)", smem, max_smem);
        return ANCDE_EUNSUPPORTED;
    }
    a.ntc = nt;
    a.ring_kc = ct;
    a.ring_slots = slots;
    a.ring_pw = 0;
    static const char* names[3] = {"ncde_bwd_plain", "ncde_bwd_bottom", "ncde_bwd_top"};
    void (*kern)(SolveArgs) = bwd_chain(lo) ? ncde_bwd_kernel<TS, true> : ncde_bwd_kernel<TS, false>;
    ANCDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(names[lo.mode], stream);
    kern<<<lo.ntiles, nt + 32, smem, stream>>>(a);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());

    return launch_wgrads(d, a, stream);
}

}  // namespace ancde

using namespace ancde;

extern "C" int ancde_ncde_join(void* stream_)
{
    int dev = 0;
    ANCDE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return ANCDE_OK;
    ancde::SideStream* ss = &ancde::g_side[dev];
    std::lock_guard<std::mutex> lock(ancde::g_side_mutex);
    if (ss->s && ss->pending) {
        ANCDE_CUDA(cudaStreamWaitEvent((cudaStream_t)stream_, ss->done, 0));
        ss->pending = false;
    }
    return ANCDE_OK;
}

namespace ancde {
static int g_force_bwd_engine = 0;       // 0 = automatic, 1 = streamed kernel, 4 = cluster / tcgen05 kernel, 5 = cluster where supported
static int g_bwd4_launches = 0;
static int g_force_bwd_cs = 0;           // cluster size of the tcgen05 backward (0 = automatic)

int64_t bwd4_img_offset(const ancde_ncde_desc* d, const NcdeLayout& st);      // ncde_host.cu

// Which recurrent backward kernel runs for this descriptor (the layout is filled when the answer is 4).
static int bwd_engine_of(const ancde_ncde_desc* d, const NcdeLayout& lo, V4Layout* v4)
{
    if (g_force_bwd_engine == 1) return 1;
    const bool need_att = d->mode == ANCDE_MODE_TOP && d->grad_attention != nullptr;
    if (choose_layout4(d, lo, need_att, g_force_bwd_cs, v4) != ANCDE_OK) return g_force_bwd_engine == 4 ? -1 : 1;
    if (g_force_bwd_engine >= 4) return 4;
    // Measured on B200 (round 2): the cluster kernel wins where the streamed one is bound by re-streaming W_out^T through
    // shared memory for 8 samples -- a large output layer and enough samples to fill clusters; small shapes stay put.
    if ((int64_t)lo.NCP * lo.K >= 20000 && lo.B >= v4->NS) return 4;
    return 1;
}
}  // namespace ancde

extern "C" int ancde_debug_force_bwd_engine(int engine, int cluster_size)
{
    const int old = ancde::g_force_bwd_engine;
    ancde::g_force_bwd_engine = engine;
    ancde::g_force_bwd_cs = cluster_size;
    return old;
}

extern "C" int ancde_debug_bwd4_launches(void) { return ancde::g_bwd4_launches; }

extern "C" int ancde_ncde_backward(const ancde_ncde_desc* d, void* stream_)
{
    reset_launch_count();
    cudaStream_t stream = (cudaStream_t)stream_;
    ANCDE_CHECK_ARG(d != nullptr, "ncde_backward: null descriptor");
    NcdeLayout lo;
    int rc = make_layout(d, &lo);
    if (rc != ANCDE_OK) return rc;
    ANCDE_CHECK_ARG(d->save_act && d->save_G, "ncde_backward: needs the save_act / save_G buffers of the forward");
    ANCDE_CHECK_ARG((d->single_stage ? d->grad_k_out : d->grad_z_out) && d->grad_z0,
                    "ncde_backward: null grad_z_out (grad_k_out in single-stage mode) / grad_z0");
    for (int l = 0; l <= d->n_hidden; ++l)
        ANCDE_CHECK_ARG(d->grad_W[l] && d->grad_bias[l], "ncde_backward: null gradient buffer for layer %d", l);
    SolveArgs a;
    rc = fill_args(d, lo, &a);
    if (rc != ANCDE_OK) return rc;
    rc = pack_weights(d, lo, stream);   // the packed weights of the forward may have been reused
    if (rc != ANCDE_OK) return rc;
    if (engine_of(d) == ANCDE_ENGINE_LANE && g_force_bwd_engine == 0) {
        // small networks: one small CTA per sample (ncde_lane.cu); same saved layout, same weight-gradient kernels (a backward
        // kernel forced through ancde_debug_force_bwd_engine runs on the lane forward's saves: the layouts are the same)
        rc = launch_lane_bwd(a, stream);
        if (rc != ANCDE_OK) return rc;
        return launch_wgrads(d, a, stream);
    }
    V4Layout v4;
    const int engine = bwd_engine_of(d, lo, &v4);
    if (engine < 0) {
        set_error("ncde_backward: this shape is not supported by the cluster (tcgen05) backward");
        return ANCDE_EUNSUPPORTED;
    }
    if (engine == 4) {
        unsigned char* img = reinterpret_cast<unsigned char*>(d->wpack + bwd4_img_offset(d, lo));
        rc = pack_weights4(d, lo, v4, img, stream);
        if (rc != ANCDE_OK) return rc;
        rc = launch_bwd4(a, v4, img, stream);
        if (rc != ANCDE_OK) return rc;
        ++g_bwd4_launches;
        return launch_wgrads(d, a, stream);
    }
    switch (lo.TS) {
        case 8: return launch_bwd<8>(d, a, stream);
        case 4: return launch_bwd<4>(d, a, stream);
        case 2: return launch_bwd<2>(d, a, stream);
        default: return launch_bwd<1>(d, a, stream);
    }
}
