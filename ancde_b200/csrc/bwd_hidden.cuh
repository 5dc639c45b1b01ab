// Hidden ReLU layers of the vector-field MLP, backwards, for the TS samples a CTA owns: shared by the streamed backward
// (ncde_bwd.cu) and the cluster / tcgen05 backward (ncde_bwd4.cu).
#pragma once
#include "mma_common.cuh"

namespace ancde {

// The hidden ReLU layers backwards as ONE warp-synchronous chain (the mirror image of hidden_chain in ncde_fwd.cu):
// a warp owns SPW samples for all layers, lane = input unit k of layer l, and per 4 outputs reads one 128-bit piece of
// the transposed weight row Wt_l[k][o..o+3] plus one 128-bit broadcast per sample of the layer's cotangent
// [sample][unit]; delta_{l-1} = (W_l^T delta_l) * relu'(h_{l-1}) goes to the next layer through shared memory behind a
// __syncwarp.  FP32 FMA: these layers are far too small for the tensor cores (measured: 1.0-1.2 k cycles per layer
// with mma.sync, one CTA barrier per layer and FP16 hi/lo operand conversions, 5.1 k cycles per stage for the Sepsis
// bottom network).
//   sLp[8*l ..] = {ceil(width / 4), in_dim, row stride, off_wt, act_off_h[l-1], act_off_delta[l-1], -, -}
template <int TS, int SPW_ = 0>
__device__ __forceinline__ void hidden_chain_bwd(const int* __restrict__ sLp, int n_hidden, const float* __restrict__ sW,
                                                 float* __restrict__ sD0, float* __restrict__ sD1,
                                                 const float* __restrict__ sAct, float* __restrict__ act,
                                                 float* __restrict__ sZb, int HP, int warp, int NW, int lane)
{
    constexpr int SPW = SPW_ ? SPW_ : ((TS >= 2) ? 2 : 1);      // samples per warp
    constexpr int NCW = TS / SPW;
    for (int cw = warp; cw < NCW; cw += NW) {
        const int s0 = cw * SPW;
        const float* in = sD0 + s0 * HP;
        int pp = 1;
        for (int l = n_hidden - 1; l >= 0; --l) {
            const int4 q0 = *reinterpret_cast<const int4*>(sLp + 8 * l);
            const int2 q1 = *reinterpret_cast<const int2*>(sLp + 8 * l + 4);
            const int nk4 = q0.x, N = q0.y;
            float* out = (pp ? sD1 : sD0) + s0 * HP;
            const int N4 = pad4(N);
            for (int k0 = 0; k0 < N4; k0 += 32) {
                const int k = k0 + lane;
                const bool valid = k < N;
                const int kc = valid ? k : N - 1;
                const float4* wr = reinterpret_cast<const float4*>(sW + q0.w + kc * q0.z);
                const float4* x0 = reinterpret_cast<const float4*>(in);
                const float4* x1 = reinterpret_cast<const float4*>(in + (SPW - 1) * HP);
                float h[SPW];
#pragma unroll
                for (int e = 0; e < SPW; ++e) h[e] = 1.0f;
                if (l > 0) Vec<SPW>::ld(sAct + q1.x + kc * TS + s0, h);
                float accA[SPW], accB[SPW];
#pragma unroll
                for (int e = 0; e < SPW; ++e) { accA[e] = 0.f; accB[e] = 0.f; }
                float4 w = wr[0], xa = x0[0], xb = x1[0];
#pragma unroll 2
                for (int q = 1; q < nk4; ++q) {
                    const float4 wn = wr[q], xan = x0[q], xbn = x1[q];
                    accA[0] = fmaf(w.x, xa.x, accA[0]);
                    accB[0] = fmaf(w.y, xa.y, accB[0]);
                    accA[0] = fmaf(w.z, xa.z, accA[0]);
                    accB[0] = fmaf(w.w, xa.w, accB[0]);
                    if (SPW == 2) {
                        accA[SPW - 1] = fmaf(w.x, xb.x, accA[SPW - 1]);
                        accB[SPW - 1] = fmaf(w.y, xb.y, accB[SPW - 1]);
                        accA[SPW - 1] = fmaf(w.z, xb.z, accA[SPW - 1]);
                        accB[SPW - 1] = fmaf(w.w, xb.w, accB[SPW - 1]);
                    }
                    w = wn; xa = xan; xb = xbn;
                }
                accA[0] = fmaf(w.x, xa.x, accA[0]);
                accB[0] = fmaf(w.y, xa.y, accB[0]);
                accA[0] = fmaf(w.z, xa.z, accA[0]);
                accB[0] = fmaf(w.w, xa.w, accB[0]);
                if (SPW == 2) {
                    accA[SPW - 1] = fmaf(w.x, xb.x, accA[SPW - 1]);
                    accB[SPW - 1] = fmaf(w.y, xb.y, accB[SPW - 1]);
                    accA[SPW - 1] = fmaf(w.z, xb.z, accA[SPW - 1]);
                    accB[SPW - 1] = fmaf(w.w, xb.w, accB[SPW - 1]);
                }
                float v[SPW];
#pragma unroll
                for (int e = 0; e < SPW; ++e) v[e] = (valid && h[e] > 0.f) ? accA[e] + accB[e] : 0.f;
                if (l > 0) {
                    if (k < N4) {
#pragma unroll
                        for (int e = 0; e < SPW; ++e) out[e * HP + k] = v[e];
                    }
                    if (valid && act) Vec<SPW>::st(act + q1.y + k * TS + s0, v);     // for the hidden weight-gradient kernel
                } else if (valid) {
                    Vec<SPW>::st(sZb + k * TS + s0, v);                       // cotangent of the stage input
                }
            }
            __syncwarp();
            in = out;
            pp ^= 1;
        }
    }
}


// The same layers in FP32 with ONE (input unit k, sample b) output per thread: thread tid = 8*k + b of the calling
// warps computes  delta_{l-1}[k][b] = relu'(h_{l-1}[k][b]) * sum_o Wt_l[k][o] * delta_l[o][b]  from 128-bit pieces of its
// weight row (the 8 threads of a k share the address: broadcast) and of its sample's cotangent row, four accumulators,
// one barrier of the calling warps per layer.  For layers of any width up to 64: a 49-unit layer is ~100 instructions
// per thread and one barrier, against 15 mma.sync of ~50 cycles each on a fifth of the warps (hidden_mma_bwd below:
// measured 1.25 k cycles per layer inside the cluster backward).  Buffers and parameters as hidden_chain_bwd.
template <int TS, typename Sync>
__device__ __forceinline__ void hidden_wide_bwd(const int* __restrict__ sLp, int n_hidden, const float* __restrict__ sW,
                                                float* __restrict__ sD0, float* __restrict__ sD1,
                                                const float* __restrict__ sAct, float* __restrict__ act,
                                                float* __restrict__ sZb, int HP, int tid, Sync layer_sync)
{
    static_assert(TS == 8, "one output per thread: tid = 8 * unit + sample");
    const int k = tid >> 3, b = tid & 7;
    const float* in = sD0;
    int pp = 1;
    for (int l = n_hidden - 1; l >= 0; --l) {
        const int4 q0 = *reinterpret_cast<const int4*>(sLp + 8 * l);
        const int2 q1 = *reinterpret_cast<const int2*>(sLp + 8 * l + 4);
        const int nk4 = q0.x, N = q0.y;               // ceil(width / 4), inputs of the layer
        float* out = pp ? sD1 : sD0;
        if (k < N) {
            const float4* wr = reinterpret_cast<const float4*>(sW + q0.w + k * q0.z);
            const float4* xr = reinterpret_cast<const float4*>(in + b * HP);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
            for (int q = 0; q < nk4; ++q) {
                const float4 w = wr[q], x = xr[q];
                a0 = fmaf(w.x, x.x, a0);
                a1 = fmaf(w.y, x.y, a1);
                a2 = fmaf(w.z, x.z, a2);
                a3 = fmaf(w.w, x.w, a3);
            }
            float v = (a0 + a1) + (a2 + a3);
            if (l > 0) {
                v = (sAct[q1.x + tid] > 0.f) ? v : 0.f;       // h_{l-1}[k][b] sits at k * 8 + b = tid
                out[b * HP + k] = v;
                if (act) act[q1.y + tid] = v;                   // for the hidden weight-gradient kernel
            } else {
                sZb[tid] = v;                                   // cotangent of the stage input
            }
        }
        layer_sync();
        in = out;
        pp ^= 1;
    }
}

// The same layers on the tensor cores (mma.sync m16n8k16, FP16 hi/lo x3 = FP32-grade) for layers wider than 32 units:
// every m-tile (16 inputs of layer l) has its own warp, one CTA barrier per layer.  The cotangents stay multiplied by the
// stage's power-of-two `scale` inside the chain (FP16 range); delta_l = (W_{l+1}^T delta_{l+1}) * relu'(h_l), the
// stage-input cotangent W_0^T delta_0 goes to sZb (unscaled).  sBd: two ping-pong buffers of {hi, lo} matrices
// [8 samples][KP halfs], buffer 0 holds delta of the last hidden layer on entry.
//   sLp[8*l ..] = {mtiles, full, tail8, stride, off, M, act_off_h[l-1], act_off_delta[l-1]}   (APack of W_l^T)
template <int TS, typename Sync>
__device__ __forceinline__ void hidden_mma_bwd(const int* __restrict__ sLp, int n_hidden, const uint4* __restrict__ sWhb,
                                               __half* __restrict__ sBd, int KP, const float* __restrict__ sAct,
                                               float* __restrict__ act, float* __restrict__ sZb, float inv_s, int warp,
                                               int lane, Sync cta_sync)
{
    const int g = lane >> 2, t = lane & 3;
    const uint32_t bd_base = smem_u32(sBd);
    const uint32_t bd_mat = (uint32_t)(8 * KP * 2);          // bytes of one FP16 matrix
    const uint32_t bd_lane = (uint32_t)(((lane & 7) * KP + ((lane >> 3) & 1) * 8) * 2) + ((lane >> 4) ? bd_mat : 0u);
    int cur = 0;
    for (int l = n_hidden - 1; l >= 0; --l) {
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
                    if (ok && act) dsave[kin * TS + b] = dv[e] * inv_s;
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

}  // namespace ancde
