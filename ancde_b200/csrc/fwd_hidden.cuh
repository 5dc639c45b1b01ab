// Hidden layers of the forward pass shared by the forward kernels (ncde_fwd.cu, ncde3_fwd1.cu).
#pragma once
#include "ncde_common.cuh"

namespace ancde {

// The hidden ReLU layers of the vector-field MLP as ONE warp-synchronous chain: a warp owns SPW samples for all layers
// (lane = output unit o).  Per 4 inputs it reads one 128-bit piece of its weight row W[o][k..k+3] (rows of 4*odd floats:
// conflict free) and one 128-bit broadcast per sample of the layer input [sample][unit]; the result goes to the next
// layer through shared memory behind a __syncwarp -- no CTA barrier between layers, and almost no index arithmetic on
// the dependent path.  (Measured on B200, Sepsis bottom network, 5 layers of 20 units: 4.4 k cycles per stage with one
// CTA barrier per layer and the layer spread over all warps -- the layers are far too small for that.)
// Layer parameters come from shared memory (sLp): register-indexed loads from the kernel-parameter constant bank
// cost hundreds of cycles on this critical path.
//   sLp[8*l ..] = {ceil(in_dim / 4), width, row stride, off_w, off_b, act_off_h, -, -}
// The last layer is also written as [unit][TS] (sLast): the operand layout of the output-layer loop.
template <int TS, int SPW_ = (TS >= 2 ? 2 : 1)>
__device__ __forceinline__ void hidden_chain(const int* __restrict__ sLp, int n_hidden, const float* __restrict__ sHW,
                                             const float* __restrict__ sInT, float* __restrict__ sT0,
                                             float* __restrict__ sT1, float* __restrict__ sLast,
                                             float* __restrict__ act, int HP, int warp, int NW, int lane)
{
    constexpr int SPW = SPW_;                       // samples per chain warp (each weight load feeds SPW FMAs)
    constexpr int NCW = TS / SPW;
    for (int cw = warp; cw < NCW; cw += NW) {
        const int s0 = cw * SPW;
        const float* in = sInT + s0 * HP;
        for (int l = 0; l < n_hidden; ++l) {
            const int4 q0 = *reinterpret_cast<const int4*>(sLp + 8 * l);
            const int2 q1 = *reinterpret_cast<const int2*>(sLp + 8 * l + 4);
            const int nk4 = q0.x, N = q0.y;
            const bool last = (l == n_hidden - 1);
            float* out = ((l & 1) ? sT1 : sT0) + s0 * HP;
            const int N4 = pad4(N);
            for (int o0 = 0; o0 < N4; o0 += 32) {
                const int o = o0 + lane;
                const bool valid = o < N;
                const int oc = valid ? o : N - 1;
                const float4* wr = reinterpret_cast<const float4*>(sHW + q0.w + oc * q0.z);
                const float4* x0 = reinterpret_cast<const float4*>(in);
                const float4* x1 = reinterpret_cast<const float4*>(in + (SPW - 1) * HP);
                float accA[SPW], accB[SPW];
                {
                    const float bv = sHW[q1.x + oc];
#pragma unroll
                    for (int e = 0; e < SPW; ++e) { accA[e] = bv; accB[e] = 0.f; }
                }
                // software pipelined: the loads of piece q + 1 are in flight during the FMAs of piece q
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
                for (int e = 0; e < SPW; ++e) v[e] = valid ? fmaxf(accA[e] + accB[e], 0.f) : 0.f;
                if (o < N4) {
#pragma unroll
                    for (int e = 0; e < SPW; ++e) out[e * HP + o] = v[e];
                }
                if (valid) {
                    if (last) Vec<SPW>::st(sLast + o * TS + s0, v);
                    if (act) Vec<SPW>::st(act + q1.y + o * TS + s0, v);
                }
            }
            __syncwarp();
            in = out;
        }
    }
}

}  // namespace ancde
