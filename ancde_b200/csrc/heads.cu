// The small layers either side of the two solves, fused:
//  * HEAD  (experiments/models/metamodel.py:327-348 / :650-671): attention = act(time_attention(raw)) or act(raw),
//    act = sigmoid | round(Hardsigmoid(slope * x)) | round(sigmoid(x)) with the straight-through round of :148-161, and
//    y0 = feature_extractor(X(t0) * attention[0]).  One kernel forward, two backward (the feature extractor's weight
//    gradient is a reduction over the batch and gets its own grid).
//  * TAIL  (metamodel.py:360-371 / :683-696 + experiments/common.py:677-680,766): gather z at final_index (or the last
//    `output_time` knots), the final linear layer, the task loss (BCE-with-logits with pos_weight, cross entropy, MSE),
//    and in the SAME launch the gradients of all of it: dLoss/dz scattered into a zero-filled (n_out, B, H) tensor for the
//    top solve's backward, dW and db of the linear layer.
//  * REG   (common.py:47-58): loss + 0.03 * sum_p ||p||_2 over the parameters of func_g, value and gradient, on the flat
//    parameter / gradient buffers (two launches).
// The reference runs these as ~100 eager torch kernels per step; they are HBM-trivial, the point is launch count.
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"

namespace ancde {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_MAXC = 128;        // channels a lane keeps in registers: C <= 128 (4 per lane)

__device__ __forceinline__ float warp_sum(float v)
{
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// attention value and its derivative w.r.t. the pre-activation x (straight-through for the rounded kinds)
__device__ __forceinline__ float head_act(int kind, float slope, float x, float* dact)
{
    if (kind == ANCDE_HEAD_SOFT) {
        const float a = 1.0f / (1.0f + expf(-x));
        *dact = a * (1.0f - a);
        return a;
    }
    if (kind == ANCDE_HEAD_HARD_STE) {
        // Hardsigmoid = (hardtanh(slope * x) + 1) / 2 (metamodel.py:139-146); torch.round rounds half to even: 0.5 -> 0
        const float sx = slope * x;
        const float hs = (fminf(fmaxf(sx, -1.0f), 1.0f) + 1.0f) / 2.0f;
        *dact = (sx > -1.0f && sx < 1.0f) ? 0.5f * slope : 0.0f;      // hardtanh passes the gradient strictly inside (-1, 1)
        return rintf(hs);
    }
    const float a = 1.0f / (1.0f + expf(-x));
    *dact = a * (1.0f - a);
    return rintf(a);
}

// One warp per (knot l, sample b) row of the raw bottom states.
__global__ void __launch_bounds__(HEAD_THREADS) head_fwd_kernel(const ancde_head_desc d)
{
    const int lane = threadIdx.x & 31;
    const int64_t nrows = (int64_t)d.L * d.B;
    const int64_t wstep = (int64_t)gridDim.x * (HEAD_THREADS / 32);
    const float bias = d.timewise ? __ldg(d.ta_b) : 0.f;
    for (int64_t row = (int64_t)blockIdx.x * (HEAD_THREADS / 32) + (threadIdx.x >> 5); row < nrows; row += wstep) {
        const int l = (int)(row / d.B), b = (int)(row - (int64_t)l * d.B);
        const float* r = d.raw + row * d.C;
        float a[HEAD_MAXC / 32], dummy;
        if (d.timewise) {
            float s = 0.f;
            for (int c = lane; c < d.C; c += 32) s = fmaf(__ldg(r + c), __ldg(d.ta_w + c), s);
            s = warp_sum(s) + bias;
            const float av = head_act(d.kind, d.slope, s, &dummy);
            if (lane == 0) d.attention[row] = av;
#pragma unroll
            for (int q = 0; q < HEAD_MAXC / 32; ++q) a[q] = av;
        } else {
#pragma unroll
            for (int q = 0; q < HEAD_MAXC / 32; ++q) {
                const int c = lane + 32 * q;
                a[q] = 0.f;
                if (c < d.C) {
                    a[q] = head_act(d.kind, d.slope, __ldg(r + c), &dummy);
                    d.attention[row * d.C + c] = a[q];
                }
            }
        }
        if (l == 0) {
            // y0 = feature_extractor(X(t0) * attention[0])   (metamodel.py:345-348)
            float xa[HEAD_MAXC / 32];
#pragma unroll
            for (int q = 0; q < HEAD_MAXC / 32; ++q) {
                const int c = lane + 32 * q;
                xa[q] = (c < d.C) ? __ldg(d.x0 + (int64_t)b * d.x0_stride + c) * a[q] : 0.f;
            }
            for (int h = 0; h < d.H; ++h) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < HEAD_MAXC / 32; ++q) {
                    const int c = lane + 32 * q;
                    if (c < d.C) s = fmaf(__ldg(d.fe_w + (int64_t)h * d.C + c), xa[q], s);
                }
                s = warp_sum(s);
                if (lane == 0) d.y0[(int64_t)b * d.H + h] = s + __ldg(d.fe_b + h);
            }
        }
    }
}

// Backward of the attention: grad_raw, and (timewise) the gradients of time_attention.  One warp per row again.
__global__ void __launch_bounds__(HEAD_THREADS) head_bwd_kernel(const ancde_head_desc d)
{
    __shared__ float s_gw[HEAD_THREADS / 32][HEAD_MAXC];
    __shared__ float s_gb[HEAD_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nrows = (int64_t)d.L * d.B;
    const int64_t wstep = (int64_t)gridDim.x * (HEAD_THREADS / 32);
    const float bias = d.timewise ? __ldg(d.ta_b) : 0.f;
    float gw[HEAD_MAXC / 32], gbias = 0.f;
#pragma unroll
    for (int q = 0; q < HEAD_MAXC / 32; ++q) gw[q] = 0.f;
    for (int64_t row = (int64_t)blockIdx.x * (HEAD_THREADS / 32) + warp; row < nrows; row += wstep) {
        const int l = (int)(row / d.B), b = (int)(row - (int64_t)l * d.B);
        const float* r = d.raw + row * d.C;
        float rv[HEAD_MAXC / 32];
#pragma unroll
        for (int q = 0; q < HEAD_MAXC / 32; ++q) {
            const int c = lane + 32 * q;
            rv[q] = (c < d.C) ? __ldg(r + c) : 0.f;
        }
        // cotangent of attention[0] through y0: x0[c] * sum_h grad_y0[h] * fe_W[h][c]
        float g0[HEAD_MAXC / 32];
#pragma unroll
        for (int q = 0; q < HEAD_MAXC / 32; ++q) g0[q] = 0.f;
        if (l == 0) {
            for (int h = 0; h < d.H; ++h) {
                const float gy = __ldg(d.grad_y0 + (int64_t)b * d.H + h);
#pragma unroll
                for (int q = 0; q < HEAD_MAXC / 32; ++q) {
                    const int c = lane + 32 * q;
                    if (c < d.C) g0[q] = fmaf(gy, __ldg(d.fe_w + (int64_t)h * d.C + c), g0[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < HEAD_MAXC / 32; ++q) {
                const int c = lane + 32 * q;
                g0[q] = (c < d.C) ? g0[q] * __ldg(d.x0 + (int64_t)b * d.x0_stride + c) : 0.f;
            }
        }
        if (d.timewise) {
            float s = 0.f, gA = d.grad_attention ? __ldg(d.grad_attention + row) : 0.f;
#pragma unroll
            for (int q = 0; q < HEAD_MAXC / 32; ++q) {
                const int c = lane + 32 * q;
                if (c < d.C) s = fmaf(rv[q], __ldg(d.ta_w + c), s);
            }
            s = warp_sum(s) + bias;
            if (l == 0) {
                float t = 0.f;
#pragma unroll
                for (int q = 0; q < HEAD_MAXC / 32; ++q) t += g0[q];
                gA += warp_sum(t);
            }
            float da;
            head_act(d.kind, d.slope, s, &da);
            const float gr = gA * da;
            gbias += gr;
#pragma unroll
            for (int q = 0; q < HEAD_MAXC / 32; ++q) {
                const int c = lane + 32 * q;
                if (c < d.C) {
                    d.grad_raw[row * d.C + c] = gr * __ldg(d.ta_w + c);
                    gw[q] = fmaf(gr, rv[q], gw[q]);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < HEAD_MAXC / 32; ++q) {
                const int c = lane + 32 * q;
                if (c < d.C) {
                    float da;
                    head_act(d.kind, d.slope, rv[q], &da);
                    const float gA = (d.grad_attention ? __ldg(d.grad_attention + row * d.C + c) : 0.f) + g0[q];
                    d.grad_raw[row * d.C + c] = gA * da;
                }
            }
        }
    }
    if (d.timewise) {
        // block-level sums, then one atomic per channel and block
#pragma unroll
        for (int q = 0; q < HEAD_MAXC / 32; ++q) s_gw[warp][lane + 32 * q] = gw[q];
        if (lane == 0) s_gb[warp] = gbias;
        __syncthreads();
        for (int c = threadIdx.x; c < d.C; c += HEAD_THREADS) {
            float t = 0.f;
            for (int w = 0; w < HEAD_THREADS / 32; ++w) t += s_gw[w][c];
            atomicAdd(d.grad_ta_w + c, t);
        }
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < HEAD_THREADS / 32; ++w) t += s_gb[w];
            atomicAdd(d.grad_ta_b, t);
        }
    }
}

// Feature extractor: grad_W[h][c] = sum_b grad_y0[b][h] * x0[b][c] * attention[0][b][c|0], grad_b[h] = sum_b grad_y0[b][h].
// Grid (H, slices of the batch): one output unit h per block row, lanes over the channels, the block's slice of the batch
// split over its warps (x0 rows are a spline-coefficient stride apart: many short independent loops, not one long one).
__global__ void __launch_bounds__(HEAD_THREADS) head_fe_grad_kernel(const ancde_head_desc d)
{
    __shared__ float s_acc[HEAD_THREADS / 32][HEAD_MAXC + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = blockIdx.x;
    float acc[HEAD_MAXC / 32], gb = 0.f;
#pragma unroll
    for (int q = 0; q < HEAD_MAXC / 32; ++q) acc[q] = 0.f;
    const int attC = d.timewise ? 1 : d.C;
    for (int b = blockIdx.y * (HEAD_THREADS / 32) + warp; b < d.B; b += gridDim.y * (HEAD_THREADS / 32)) {
        const float gy = __ldg(d.grad_y0 + (int64_t)b * d.H + h);
        gb += gy;
#pragma unroll
        for (int q = 0; q < HEAD_MAXC / 32; ++q) {
            const int c = lane + 32 * q;
            if (c < d.C) {
                const float a = __ldg(d.attention + (int64_t)b * attC + (d.timewise ? 0 : c));     // attention[0, b, :]
                acc[q] = fmaf(gy, __ldg(d.x0 + (int64_t)b * d.x0_stride + c) * a, acc[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < HEAD_MAXC / 32; ++q) s_acc[warp][lane + 32 * q] = acc[q];
    if (lane == 0) s_acc[warp][HEAD_MAXC] = gb;
    __syncthreads();
    for (int c = threadIdx.x; c < d.C; c += HEAD_THREADS) {
        float t = 0.f;
        for (int w = 0; w < HEAD_THREADS / 32; ++w) t += s_acc[w][c];
        atomicAdd(d.grad_fe_w + (int64_t)h * d.C + c, t);
    }
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < HEAD_THREADS / 32; ++w) t += s_acc[w][HEAD_MAXC];
        atomicAdd(d.grad_fe_b + h, t);
    }
}

// ---- TAIL ---------------------------------------------------------------------------------------------------------
constexpr int TAIL_THREADS = 256;
constexpr int TAIL_MAXH = 128;       // hidden state <= 128 (4 per lane)
constexpr int TAIL_MAXO = 64;        // outputs of the final linear layer

// One warp per read-out row (sample b, forecast step q): pred = W z + b, loss term, dpred, dz; the linear layer's gradients
// are summed per block in shared memory.  Every block also zero-fills its share of grad_z_t first (rows that are not read
// out carry no cotangent); the read-out rows are written after a grid-wide ordering that the zero fill of OTHER blocks
// cannot violate: a row is zero-filled and written by the SAME block (rows are dealt to blocks, not elements).
__global__ void __launch_bounds__(TAIL_THREADS) tail_kernel(const ancde_tail_desc d)
{
    extern __shared__ float s_tail[];      // [O][H] dW | [O] db | [warps] loss
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NWP = TAIL_THREADS / 32;
    const int H = d.H, O = d.O;
    float* s_dw = s_tail;
    float* s_db = s_tail + (size_t)O * H;
    float* s_loss = s_db + O;
    for (int i = threadIdx.x; i < O * H + O + NWP; i += TAIL_THREADS) s_tail[i] = 0.f;
    __syncthreads();
    const int T = d.kind == ANCDE_TAIL_MSE ? d.T : 1;
    const int64_t nrows = (int64_t)d.B * T;                      // read-out rows
    const float inv_n = d.loss_scale / (float)(d.kind == ANCDE_TAIL_MSE ? nrows * O : d.B);    // mean reductions
    float lsum = 0.f;
    // the (knot, sample) rows of grad_z_t are dealt to the warps of the grid; the warp that owns a row zero-fills it or,
    // when it is a read-out row, computes it
    const int64_t zrows = (int64_t)d.n_out * d.B;
    const int64_t wstep = (int64_t)gridDim.x * NWP;
    for (int64_t zr = (int64_t)blockIdx.x * NWP + warp; zr < zrows; zr += wstep) {
        const int k = (int)(zr / d.B), b = (int)(zr - (int64_t)k * d.B);
        bool read = false;
        int q = 0;
        if (d.kind == ANCDE_TAIL_MSE) {
            read = k >= d.n_out - T;          // the last T knots of every sample (metamodel.py:694-696)
            q = k - (d.n_out - T);
        } else {
            read = (int64_t)k == __ldg(d.final_index + b);
        }
        float* gz = d.grad_z_t ? d.grad_z_t + zr * H : nullptr;
        if (!read) {
            if (gz)
                for (int h = lane; h < H; h += 32) gz[h] = 0.f;
            continue;
        }
        const float* z = d.z_t + zr * H;
        float zv[TAIL_MAXH / 32];
#pragma unroll
        for (int u = 0; u < TAIL_MAXH / 32; ++u) {
            const int h = lane + 32 * u;
            zv[u] = (h < H) ? __ldg(z + h) : 0.f;
        }
        // pred[o] = W[o] . z + b[o]; every lane ends up with all O logits of the row (O <= 64: two per lane slot)
        float p0 = 0.f, p1 = 0.f;            // logits lane, lane + 32
        for (int o = 0; o < O; ++o) {
            float s = 0.f;
#pragma unroll
            for (int u = 0; u < TAIL_MAXH / 32; ++u) {
                const int h = lane + 32 * u;
                if (h < H) s = fmaf(__ldg(d.W + (int64_t)o * H + h), zv[u], s);
            }
            s = warp_sum(s) + __ldg(d.b + o);
            if ((o & 31) == lane) { if (o < 32) p0 = s; else p1 = s; }
        }
        const int64_t prow = (d.kind == ANCDE_TAIL_MSE) ? ((int64_t)b * T + q) : (int64_t)b;
        if (lane < O) d.pred[prow * O + lane] = p0;
        if (lane + 32 < O) d.pred[prow * O + lane + 32] = p1;
        // loss term and dLoss/dpred
        float g0 = 0.f, g1 = 0.f;
        if (d.kind == ANCDE_TAIL_BCE) {
            // BCEWithLogits(pos_weight): l = -[pw * y * log s(x) + (1 - y) * log(1 - s(x))]  (common.py:677-678)
            if (lane == 0) {
                const float x = p0, y = __ldg(reinterpret_cast<const float*>(d.target) + b), pw = d.pos_weight;
                const float sp = fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));           // softplus(-x) = -log s(x)
                lsum += (1.0f - y) * x + (1.0f + (pw - 1.0f) * y) * sp;
                const float sg = 1.0f / (1.0f + expf(-x));
                g0 = ((1.0f - y) - (1.0f + (pw - 1.0f) * y) * (1.0f - sg)) * inv_n;
            }
        } else if (d.kind == ANCDE_TAIL_CE) {
            float m = fmaxf(lane < O ? p0 : -INFINITY, lane + 32 < O ? p1 : -INFINITY);
            for (int o2 = 16; o2 >= 1; o2 >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o2));
            const float e0 = lane < O ? expf(p0 - m) : 0.f, e1 = lane + 32 < O ? expf(p1 - m) : 0.f;
            const float se = warp_sum(e0 + e1);
            const int y = (int)__ldg(reinterpret_cast<const long long*>(d.target) + b);
            const float py = __shfl_sync(0xffffffffu, y < 32 ? p0 : p1, y & 31);
            if (lane == 0) lsum += logf(se) + m - py;
            g0 = (lane < O) ? (e0 / se - (lane == y ? 1.f : 0.f)) * inv_n : 0.f;
            g1 = (lane + 32 < O) ? (e1 / se - (lane + 32 == y ? 1.f : 0.f)) * inv_n : 0.f;
        } else {
            const float* y = reinterpret_cast<const float*>(d.target) + prow * O;
            if (lane < O) { const float e = p0 - __ldg(y + lane); lsum += e * e; g0 = 2.0f * e * inv_n; }
            if (lane + 32 < O) { const float e = p1 - __ldg(y + lane + 32); lsum += e * e; g1 = 2.0f * e * inv_n; }
        }
        // dz = W^T dpred; dW += dpred z^T; db += dpred
        float dz[TAIL_MAXH / 32];
#pragma unroll
        for (int u = 0; u < TAIL_MAXH / 32; ++u) dz[u] = 0.f;
        for (int o = 0; o < O; ++o) {
            const float g = __shfl_sync(0xffffffffu, o < 32 ? g0 : g1, o & 31);
            if (g == 0.f) continue;
#pragma unroll
            for (int u = 0; u < TAIL_MAXH / 32; ++u) {
                const int h = lane + 32 * u;
                if (h < H) {
                    dz[u] = fmaf(g, __ldg(d.W + (int64_t)o * H + h), dz[u]);
                    atomicAdd(s_dw + o * H + h, g * zv[u]);
                }
            }
            if (lane == 0) atomicAdd(s_db + o, g);
        }
        if (gz) {
#pragma unroll
            for (int u = 0; u < TAIL_MAXH / 32; ++u) {
                const int h = lane + 32 * u;
                if (h < H) gz[h] = dz[u];
            }
        }
    }
    lsum = warp_sum(lsum);
    if (lane == 0) s_loss[warp] = lsum;
    __syncthreads();
    for (int i = threadIdx.x; i < O * H; i += TAIL_THREADS)
        if (s_dw[i] != 0.f) atomicAdd(d.grad_W + i, s_dw[i]);
    for (int i = threadIdx.x; i < O; i += TAIL_THREADS)
        if (s_db[i] != 0.f) atomicAdd(d.grad_b + i, s_db[i]);
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < NWP; ++w) t += s_loss[w];
        if (t != 0.f) atomicAdd(d.loss, t * (inv_n / d.loss_scale));
    }
}

// ---- REG ----------------------------------------------------------------------------------------------------------
// sq[s] += sum of squares of segment s (segments = parameters, contiguous in the flat buffer): one block per chunk of 4096
__global__ void __launch_bounds__(256) reg_sumsq_kernel(const float* __restrict__ w, const int* __restrict__ seg_begin, int nseg,
                                                       float* __restrict__ sq)
{
    __shared__ float s_part[8];
    const int seg = blockIdx.y;
    const int b = seg_begin[seg], e = seg_begin[seg + 1];
    float s = 0.f;
    for (int i = b + blockIdx.x * 256 + threadIdx.x; i < e; i += gridDim.x * 256) s = fmaf(w[i], w[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_part[i];
        if (t != 0.f) atomicAdd(sq + seg, t);
    }
}
// grad += scaling * w / ||w|| per segment (0 for a zero parameter, like torch.norm's backward); loss += scaling * sum ||w||
__global__ void __launch_bounds__(256) reg_apply_kernel(const float* __restrict__ w, float* __restrict__ g, const int* __restrict__ seg_begin,
                                                       int nseg, const float* __restrict__ sq, float scaling, float* loss)
{
    const int seg = blockIdx.y;
    const int b = seg_begin[seg], e = seg_begin[seg + 1];
    const float n = sqrtf(sq[seg]);
    const float coef = n > 0.f ? scaling / n : 0.f;
    for (int i = b + blockIdx.x * 256 + threadIdx.x; i < e; i += gridDim.x * 256) g[i] = fmaf(coef, w[i], g[i]);
    if (loss && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(loss, scaling * n);
}

}  // namespace ancde

using namespace ancde;

static int head_check(const ancde_head_desc* d)
{
    ANCDE_CHECK_ARG(d != nullptr, "head: null descriptor");
    ANCDE_CHECK_ARG(d->L >= 1 && d->B >= 1 && d->C >= 1 && d->H >= 1, "head: bad shape L=%d B=%d C=%d H=%d", d->L, d->B, d->C, d->H);
    ANCDE_CHECK_ARG(d->C <= HEAD_MAXC, "head: C=%d exceeds %d channels", d->C, HEAD_MAXC);
    ANCDE_CHECK_ARG(d->kind >= ANCDE_HEAD_SOFT && d->kind <= ANCDE_HEAD_HARD_SIGMOID, "head: bad kind %d", d->kind);
    ANCDE_CHECK_ARG(d->raw && d->x0 && d->fe_w && d->fe_b && d->attention, "head: null pointer");
    ANCDE_CHECK_ARG(!d->timewise || (d->ta_w && d->ta_b), "head: timewise attention needs time_attention's weight and bias");
    ANCDE_CHECK_ARG(d->x0_stride >= d->C, "head: x0_stride %lld < C", (long long)d->x0_stride);
    return ANCDE_OK;
}

static int rows_grid(int64_t rows, int warps_per_block)
{
    int64_t blocks = (rows + warps_per_block - 1) / warps_per_block;
    const int64_t cap = 148 * 8;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

extern "C" int ancde_head_forward(const ancde_head_desc* d, void* stream_)
{
    reset_launch_count();
    int rc = head_check(d);
    if (rc != ANCDE_OK) return rc;
    ANCDE_CHECK_ARG(d->y0 != nullptr, "head_forward: null y0");
    cudaStream_t stream = (cudaStream_t)stream_;
    prof_begin("head_fwd", stream);
    head_fwd_kernel<<<rows_grid((int64_t)d->L * d->B, HEAD_THREADS / 32), HEAD_THREADS, 0, stream>>>(*d);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

extern "C" int ancde_head_backward(const ancde_head_desc* d, void* stream_)
{
    reset_launch_count();
    int rc = head_check(d);
    if (rc != ANCDE_OK) return rc;
    ANCDE_CHECK_ARG(d->grad_y0 && d->grad_raw && d->grad_fe_w && d->grad_fe_b, "head_backward: null gradient pointer");
    ANCDE_CHECK_ARG(!d->timewise || (d->grad_ta_w && d->grad_ta_b), "head_backward: null time_attention gradient pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    prof_begin("head_bwd", stream);
    head_bwd_kernel<<<rows_grid((int64_t)d->L * d->B, HEAD_THREADS / 32), HEAD_THREADS, 0, stream>>>(*d);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    prof_begin("head_fe_grad", stream);
    {
        int slices = (d->B + 63) / 64;             // ~8 samples per warp
        if (slices > 64) slices = 64;
        head_fe_grad_kernel<<<dim3((unsigned)d->H, (unsigned)slices), HEAD_THREADS, 0, stream>>>(*d);
    }
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

extern "C" int ancde_tail_forward_backward(const ancde_tail_desc* d, void* stream_)
{
    reset_launch_count();
    ANCDE_CHECK_ARG(d != nullptr, "tail: null descriptor");
    ANCDE_CHECK_ARG(d->n_out >= 1 && d->B >= 1 && d->H >= 1 && d->O >= 1, "tail: bad shape n_out=%d B=%d H=%d O=%d", d->n_out, d->B, d->H, d->O);
    ANCDE_CHECK_ARG(d->H <= TAIL_MAXH && d->O <= TAIL_MAXO, "tail: H=%d (max %d) O=%d (max %d)", d->H, TAIL_MAXH, d->O, TAIL_MAXO);
    ANCDE_CHECK_ARG(d->kind >= ANCDE_TAIL_BCE && d->kind <= ANCDE_TAIL_MSE, "tail: bad kind %d", d->kind);
    ANCDE_CHECK_ARG(d->kind != ANCDE_TAIL_BCE || d->O == 1, "tail: BCE needs one output (O=%d)", d->O);
    ANCDE_CHECK_ARG(d->kind != ANCDE_TAIL_MSE || (d->T >= 1 && d->T <= d->n_out), "tail: output_time %d of %d knots", d->T, d->n_out);
    ANCDE_CHECK_ARG(d->kind == ANCDE_TAIL_MSE || d->final_index, "tail: null final_index");
    ANCDE_CHECK_ARG(d->z_t && d->W && d->b && d->target && d->pred && d->loss && d->grad_W && d->grad_b, "tail: null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t smem = ((size_t)d->O * d->H + d->O + TAIL_THREADS / 32) * sizeof(float);
    ANCDE_CHECK_ARG(smem <= 48 * 1024, "tail: O*H = %d does not fit the block's shared memory", d->O * d->H);
    ANCDE_CUDA(cudaMemsetAsync(d->loss, 0, sizeof(float), stream));
    prof_begin("tail_fwd_bwd", stream);
    tail_kernel<<<rows_grid((int64_t)d->n_out * d->B, TAIL_THREADS / 32), TAIL_THREADS, smem, stream>>>(*d);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

extern "C" int ancde_weight_regulariser(const float* flat_param, float* flat_grad, const int* seg_begin, int nseg, float* sumsq,
                                        float scaling, float* loss, void* stream_)
{
    reset_launch_count();
    ANCDE_CHECK_ARG(flat_param && flat_grad && seg_begin && sumsq && nseg >= 1, "weight_regulariser: null pointer / no segments");
    cudaStream_t stream = (cudaStream_t)stream_;
    ANCDE_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float) * (size_t)nseg, stream));
    prof_begin("reg_sumsq", stream);
    reg_sumsq_kernel<<<dim3(16, (unsigned)nseg), 256, 0, stream>>>(flat_param, seg_begin, nseg, sumsq);
    prof_end(stream);
    count_launch();
    prof_begin("reg_apply", stream);
    reg_apply_kernel<<<dim3(16, (unsigned)nseg), 256, 0, stream>>>(flat_param, flat_grad, seg_begin, nseg, sumsq, scaling, loss);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}
