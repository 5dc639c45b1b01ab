// Hidden-layer weight gradients of the fused NCDE solve (the output layer's is in ncde_wgrad_mma.cu).
#include <stdlib.h>

#include "ncde_common.cuh"
#include "pipeline.cuh"

namespace ancde {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;     // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------
// Hidden-layer weight gradients: dW_l[o][k] = sum over (stage, sample) of delta_l[o] * in_l[k], where the
// recurrent backward left the masked layer cotangents delta_l in save_act and in_l is the saved input of
// the layer (stage input for l = 0, previous activation otherwise).  One CTA column per layer
// (blockIdx.x = l), rows split over blockIdx.y; thread = output o x a quarter of the inputs, accumulators
// in registers, rows through a two-stage cp.async pipeline.  Results are added to the torch-layout
// gradients directly.
// ---------------------------------------------------------------------------------------------
struct HiddenGradPtrs { float* gW[ANCDE_MAX_LAYERS]; float* gB[ANCDE_MAX_LAYERS]; };
constexpr int HW_KQ_MAX = 32;  // inputs per thread (4 input quarters x 32 = 128 inputs max)
constexpr int HW_BPC = 4;      // (stage, tile) blocks per pipeline chunk

template <int TS, int HW_KQ>
__global__ void __launch_bounds__(256) ncde_hidden_wgrad_kernel(const __grid_constant__ SolveArgs p,
                                                                const __grid_constant__ HiddenGradPtrs hp, int chunks_per_cta)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    constexpr int NQ = TS / VW;
    const NcdeLayout& lo = p.lo;
    const int l = blockIdx.x;
    const int N = lo.width[l], Kin = lo.in_dim[l];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int o = tid & 63, kq = tid >> 6;
    const int KQ = (Kin + 3) >> 2;
    const int kbeg = kq * KQ;
    const int off_delta = lo.act_off_delta[l];
    const int off_in = (l == 0) ? 0 : lo.act_off_h[l - 1];

    extern __shared__ __align__(16) float smem[];
    const int szD = pad4(N * TS), szI = pad4(Kin * TS);
    const int blk_floats = szD + szI;
    const int stage_floats = HW_BPC * blk_floats;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);

    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + HW_BPC - 1) / HW_BPC;
    const int64_t c0 = (int64_t)blockIdx.y * chunks_per_cta;
    int64_t c1 = c0 + chunks_per_cta;
    if (c1 > nchunks) c1 = nchunks;

    auto issue = [&](int64_t ch, int s) {
        for (int bl = 0; bl < HW_BPC; ++bl) {
            const int64_t blk = ch * HW_BPC + bl;
            const bool ok = blk < nblocks;
            const float* act = p.save_act + (ok ? blk : 0) * lo.act_floats;
            const uint32_t dD = smem_base + (uint32_t)(s * stage_floats + bl * blk_floats) * 4u;
            const uint32_t dI = dD + (uint32_t)szD * 4u;
            const int sz16 = ok ? 16 : 0, sz4 = ok ? 4 : 0;
            if (VW == 4) {
                for (int i = tid; i < N * TS / 4; i += NT)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dD + i * 16u), "l"(act + off_delta + i * 4), "r"(sz16) : "memory");
                for (int i = tid; i < Kin * TS / 4; i += NT)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dI + i * 16u), "l"(act + off_in + i * 4), "r"(sz16) : "memory");
            } else {
                for (int i = tid; i < N * TS; i += NT)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dD + i * 4u), "l"(act + off_delta + i), "r"(sz4) : "memory");
                for (int i = tid; i < Kin * TS; i += NT)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dI + i * 4u), "l"(act + off_in + i), "r"(sz4) : "memory");
            }
        }
    };

    float acc[HW_KQ];
#pragma unroll
    for (int j = 0; j < HW_KQ; ++j) acc[j] = 0.f;
    float bacc = 0.f;

    if (c0 < c1) issue(c0, 0);
    cp_async_commit();
    if (c0 + 1 < c1) issue(c0 + 1, 1);
    cp_async_commit();
    for (int64_t ch = c0; ch < c1; ++ch) {
        const int s = (int)((ch - c0) & 1);
        cp_async_wait<1>();
        __syncthreads();
        if (o < N) {
#pragma unroll 1
            for (int bl = 0; bl < HW_BPC; ++bl) {
                const float* sD = smem + s * stage_floats + bl * blk_floats;
                const float* sI = sD + szD;
                float dv[TS];
#pragma unroll
                for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sD + o * TS + q * VW, dv + q * VW);
                if (kq == 0) {
#pragma unroll
                    for (int b = 0; b < TS; ++b) bacc += dv[b];
                }
#pragma unroll
                for (int j = 0; j < HW_KQ; ++j) {
                    const int kk = kbeg + j;
                    if (j < KQ && kk < Kin) {
                        float iv[TS];
#pragma unroll
                        for (int q = 0; q < NQ; ++q) Vec<VW>::ld(sI + kk * TS + q * VW, iv + q * VW);
#pragma unroll
                        for (int b = 0; b < TS; ++b) acc[j] = fmaf(dv[b], iv[b], acc[j]);
                    }
                }
            }
        }
        __syncthreads();
        if (ch + 2 < c1) issue(ch + 2, s);
        cp_async_commit();
    }
    cp_async_wait<0>();
    if (o < N && c0 < c1) {
#pragma unroll
        for (int j = 0; j < HW_KQ; ++j) {
            const int kk = kbeg + j;
            if (j < KQ && kk < Kin) atomicAdd(hp.gW[l] + (int64_t)o * Kin + kk, acc[j]);
        }
        if (kq == 0) atomicAdd(hp.gB[l] + o, bacc);
    }
}

// ---------------------------------------------------------------------------------------------
// The same sums on the tensor cores (mma.sync m16n8k8, 3xTF32 = FP32-grade: x = hi + lo with both halves in TF32, the
// products lo*hi, hi*lo and hi*hi accumulated in FP32).  Per (stage, tile) block the sum over the 8 samples of a tile
// is exactly the k = 8 of one instruction: A = delta_l[16 outputs][8 samples], B = in_l[8 samples][8 inputs], both read
// straight from the saved [unit][TS] blocks (conflict free).  A CTA works on one layer; its 8 warps form two groups
// that take alternate blocks, warp w of a group owns outputs 16w .. 16w+15 and every input tile, accumulators in
// registers for the whole kernel, one atomicAdd per element at the end.  Measured on B200 (Sepsis shape, B = 1024):
// 0.41 + 0.29 ms with the FP32-FMA kernel above (issue bound: 278 warp instructions per block and warp).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo)
{
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int TS, int NT8>
__global__ void __launch_bounds__(256, (NT8 <= 8) ? 4 : 2) ncde_hidden_wgrad_mma_kernel(const __grid_constant__ SolveArgs p,
                                                                    const __grid_constant__ HiddenGradPtrs hp, int chunks_per_cta)
{
    constexpr int VW = (TS >= 4) ? 4 : TS;
    const NcdeLayout& lo = p.lo;
    const int l = blockIdx.x;
    const int N = lo.width[l], Kin = lo.in_dim[l];
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int grp = warp >> 2, mt = warp & 3;            // block parity this warp works on, its m-tile of 16 outputs
    const int ntiles8 = (Kin + 7) >> 3;
    const int off_delta = lo.act_off_delta[l];
    const int off_in = (l == 0) ? 0 : lo.act_off_h[l - 1];

    extern __shared__ __align__(16) float smem[];
    const int szD = pad4(N * TS), szI = pad4(Kin * TS);
    const int blk_floats = szD + szI;
    const int stage_floats = HW_BPC * blk_floats;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    // TS = 8: a block's delta_l and in_l are two contiguous, 16-byte aligned runs of the saved activations -- one thread
    // stages a chunk with 2 TMA bulk copies per block on an mbarrier per pipeline stage instead of ~200 cp.async spread
    // over the CTA (the staging loop was a fifth of this kernel's instructions)
    constexpr bool BULK = (TS == 8);
    __shared__ __align__(8) uint64_t bar_full[2];
    if (BULK && tid == 0) {
        mbar_init(bar_full + 0, 1);
        mbar_init(bar_full + 1, 1);
        mbar_fence_init();
    }
    if (BULK) __syncthreads();

    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + HW_BPC - 1) / HW_BPC;
    const int64_t c0 = (int64_t)blockIdx.y * chunks_per_cta;
    int64_t c1 = c0 + chunks_per_cta;
    if (c1 > nchunks) c1 = nchunks;

    auto issue = [&](int64_t ch, int s) {
        if (BULK) {
            if (tid == 0) {
                int nb = 0;
                for (int bl = 0; bl < HW_BPC; ++bl) nb += (ch * HW_BPC + bl < nblocks) ? 1 : 0;
                mbar_arrive_expect_tx(bar_full + s, (uint32_t)(nb * (N + Kin) * TS * 4));
                for (int bl = 0; bl < nb; ++bl) {
                    const float* act = p.save_act + (ch * HW_BPC + bl) * lo.act_floats;
                    float* dD = smem + s * stage_floats + bl * blk_floats;
                    bulk_g2s(dD, act + off_delta, (uint32_t)(N * TS * 4), bar_full + s);
                    bulk_g2s(dD + szD, act + off_in, (uint32_t)(Kin * TS * 4), bar_full + s);
                }
            }
            return;
        }
        for (int bl = 0; bl < HW_BPC; ++bl) {
            const int64_t blk = ch * HW_BPC + bl;
            const bool ok = blk < nblocks;
            const float* act = p.save_act + (ok ? blk : 0) * lo.act_floats;
            const uint32_t dD = smem_base + (uint32_t)(s * stage_floats + bl * blk_floats) * 4u;
            const uint32_t dI = dD + (uint32_t)szD * 4u;
            const int sz16 = ok ? 16 : 0, sz4 = ok ? 4 : 0;
            if (VW == 4) {
                for (int i = tid; i < N * TS / 4; i += NT)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dD + i * 16u), "l"(act + off_delta + i * 4), "r"(sz16) : "memory");
                for (int i = tid; i < Kin * TS / 4; i += NT)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dI + i * 16u), "l"(act + off_in + i * 4), "r"(sz16) : "memory");
            } else {
                for (int i = tid; i < N * TS; i += NT)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dD + i * 4u), "l"(act + off_delta + i), "r"(sz4) : "memory");
                for (int i = tid; i < Kin * TS; i += NT)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dI + i * 4u), "l"(act + off_in + i), "r"(sz4) : "memory");
            }
        }
    };

    float acc[NT8][4];
#pragma unroll
    for (int j = 0; j < NT8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    float bs0 = 0.f, bs1 = 0.f;                           // bias gradient of outputs 16*mt + g and + 8 (this lane's samples)

    // fragment coordinates: A rows o0 = 16*mt + g and o0 + 8, columns (samples) t and t + 4; B rows (samples) t, t + 4,
    // column (input) 8*j + g.  Samples beyond TS and units beyond N / Kin read as zero.
    const int o0 = 16 * mt + g;
    const bool a_row0 = o0 < N, a_row1 = o0 + 8 < N;
    const bool s_lo = t < TS, s_hi = t + 4 < TS;
    const bool active = 16 * mt < N;

    if (c0 < c1) issue(c0, 0);
    cp_async_commit();
    if (c0 + 1 < c1) issue(c0 + 1, 1);
    cp_async_commit();
    for (int64_t ch = c0; ch < c1; ++ch) {
        const int s = (int)((ch - c0) & 1);
        if (BULK) {
            mbar_wait(bar_full + s, (uint32_t)(((ch - c0) >> 1) & 1));
        } else {
            cp_async_wait<1>();
            __syncthreads();
        }
        if (active) {
#pragma unroll 1
            for (int bl = grp; bl < HW_BPC; bl += 2) {
                if (BULK && ch * HW_BPC + bl >= nblocks) break;      // past the last block: nothing was copied
                const float* sD = smem + s * stage_floats + bl * blk_floats;
                const float* sI = sD + szD;
                const float d0 = (a_row0 && s_lo) ? sD[o0 * TS + t] : 0.f;
                const float d1 = (a_row1 && s_lo) ? sD[(o0 + 8) * TS + t] : 0.f;
                const float d2 = (a_row0 && s_hi) ? sD[o0 * TS + t + 4] : 0.f;
                const float d3 = (a_row1 && s_hi) ? sD[(o0 + 8) * TS + t + 4] : 0.f;
                bs0 += d0 + d2;
                bs1 += d1 + d3;
                uint32_t ah[4], al[4];
                split_tf32(d0, ah[0], al[0]);
                split_tf32(d1, ah[1], al[1]);
                split_tf32(d2, ah[2], al[2]);
                split_tf32(d3, ah[3], al[3]);
#pragma unroll
                for (int j = 0; j < NT8; ++j) {
                    if (j < ntiles8) {
                        const int k = 8 * j + g;
                        const bool k_ok = k < Kin;
                        const float x0 = (k_ok && s_lo) ? sI[k * TS + t] : 0.f;
                        const float x1 = (k_ok && s_hi) ? sI[k * TS + t + 4] : 0.f;
                        uint32_t bh0, bl0, bh1, bl1;
                        split_tf32(x0, bh0, bl0);
                        split_tf32(x1, bh1, bl1);
                        mma_tf32(acc[j], al[0], al[1], al[2], al[3], bh0, bh1);
                        mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                        mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
                    }
                }
            }
        }
        __syncthreads();
        if (ch + 2 < c1) issue(ch + 2, s);
        cp_async_commit();
    }
    cp_async_wait<0>();
    if (active && c0 < c1) {
        // C fragment: rows o0, o0 + 8; columns 8*j + 2t, + 1
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
            if (j < ntiles8) {
                const int k = 8 * j + 2 * t;
                if (a_row0 && k < Kin) atomicAdd(hp.gW[l] + (int64_t)o0 * Kin + k, acc[j][0]);
                if (a_row0 && k + 1 < Kin) atomicAdd(hp.gW[l] + (int64_t)o0 * Kin + k + 1, acc[j][1]);
                if (a_row1 && k < Kin) atomicAdd(hp.gW[l] + (int64_t)(o0 + 8) * Kin + k, acc[j][2]);
                if (a_row1 && k + 1 < Kin) atomicAdd(hp.gW[l] + (int64_t)(o0 + 8) * Kin + k + 1, acc[j][3]);
            }
        }
        bs0 += __shfl_xor_sync(0xffffffffu, bs0, 1);
        bs0 += __shfl_xor_sync(0xffffffffu, bs0, 2);
        bs1 += __shfl_xor_sync(0xffffffffu, bs1, 1);
        bs1 += __shfl_xor_sync(0xffffffffu, bs1, 2);
        if (t == 0) {
            if (a_row0) atomicAdd(hp.gB[l] + o0, bs0);
            if (a_row1) atomicAdd(hp.gB[l] + o0 + 8, bs1);
        }
    }
}

template <int TS>
static int launch_hidden_wgrad_ts(const SolveArgs& a, const HiddenGradPtrs& hp, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    size_t smem = 0;
    for (int l = 0; l < lo.n_hidden; ++l) {
        if (lo.width[l] > 64 || lo.in_dim[l] > 4 * HW_KQ_MAX) {
            set_error("ncde_backward: hidden layer %d is %dx%d; the hidden weight-gradient kernel handles up to 64 outputs x 128 inputs",
                      l, lo.width[l], lo.in_dim[l]);
            return ANCDE_EUNSUPPORTED;
        }
        const size_t f = 2 * (size_t)HW_BPC * (pad4(lo.width[l] * TS) + pad4(lo.in_dim[l] * TS)) * sizeof(float);
        if (f > smem) smem = f;
    }
    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + HW_BPC - 1) / HW_BPC;
    int64_t gy = (148 * 4 + lo.n_hidden - 1) / lo.n_hidden;
    if (gy > nchunks) gy = nchunks;
    if (gy < 1) gy = 1;
    const int per = (int)((nchunks + gy - 1) / gy);
    gy = (nchunks + per - 1) / per;
    int kmax = 0;
    for (int l = 0; l < lo.n_hidden; ++l) kmax = lo.in_dim[l] > kmax ? lo.in_dim[l] : kmax;
    static const char* names[3] = {"ncde_hidden_wgrad_plain", "ncde_hidden_wgrad_bottom", "ncde_hidden_wgrad_top"};
    // tensor-core kernel unless ANCDE_HIDDEN_WGRAD_FP32=1 (A/B timing and the cross-check test)
    static const bool force_fp32 = getenv("ANCDE_HIDDEN_WGRAD_FP32") && getenv("ANCDE_HIDDEN_WGRAD_FP32")[0] == '1';
    if (!force_fp32) {
        if (kmax <= 64) {
            ANCDE_CUDA(cudaFuncSetAttribute(ncde_hidden_wgrad_mma_kernel<TS, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            prof_begin(names[lo.mode], stream);
            ncde_hidden_wgrad_mma_kernel<TS, 8><<<dim3(lo.n_hidden, (unsigned)gy), 256, smem, stream>>>(a, hp, per);
        } else {
            ANCDE_CUDA(cudaFuncSetAttribute(ncde_hidden_wgrad_mma_kernel<TS, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            prof_begin(names[lo.mode], stream);
            ncde_hidden_wgrad_mma_kernel<TS, 16><<<dim3(lo.n_hidden, (unsigned)gy), 256, smem, stream>>>(a, hp, per);
        }
    } else if (kmax <= 64) {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_hidden_wgrad_kernel<TS, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_hidden_wgrad_kernel<TS, 16><<<dim3(lo.n_hidden, (unsigned)gy), 256, smem, stream>>>(a, hp, per);
    } else {
        ANCDE_CUDA(cudaFuncSetAttribute(ncde_hidden_wgrad_kernel<TS, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(names[lo.mode], stream);
        ncde_hidden_wgrad_kernel<TS, 32><<<dim3(lo.n_hidden, (unsigned)gy), 256, smem, stream>>>(a, hp, per);
    }
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

bool hidden_wgrad_umma_ok(const NcdeLayout& lo);
int launch_hidden_wgrad_umma(const SolveArgs& a, const HiddenGradPtrs& hp, cudaStream_t stream);

int launch_hidden_wgrad(const SolveArgs& a, const HiddenGradPtrs& hp, cudaStream_t stream)
{
    // round 2: tiles of 8 samples and <= 4 layers of <= 64 units go to the tcgen05 kernel (ncde_hidden_wgrad_umma.cu) unless
    // ANCDE_HIDDEN_WGRAD_MMA=1 / ANCDE_HIDDEN_WGRAD_FP32=1 ask for the mma.sync / FP32 kernels (A/B timing, cross-check test)
    static const bool legacy = (getenv("ANCDE_HIDDEN_WGRAD_MMA") && getenv("ANCDE_HIDDEN_WGRAD_MMA")[0] == '1') ||
                               (getenv("ANCDE_HIDDEN_WGRAD_FP32") && getenv("ANCDE_HIDDEN_WGRAD_FP32")[0] == '1');
    if (!legacy && hidden_wgrad_umma_ok(a.lo)) {
        const int rc = launch_hidden_wgrad_umma(a, hp, stream);
        if (rc != ANCDE_EUNSUPPORTED) return rc;
    }
    switch (a.lo.TS) {
        case 8: return launch_hidden_wgrad_ts<8>(a, hp, stream);
        case 4: return launch_hidden_wgrad_ts<4>(a, hp, stream);
        case 2: return launch_hidden_wgrad_ts<2>(a, hp, stream);
        default: return launch_hidden_wgrad_ts<1>(a, hp, stream);
    }
}

int launch_wgrad_mma(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream);   // ncde_wgrad_mma.cu

int launch_wgrad_umma(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream);  // ncde_wgrad_umma.cu

// tcgen05 kernel for tiles of 8 samples; the mma.sync kernel otherwise, or when forced (ANCDE_WGRAD_MMA_SYNC=1 in the
// environment or ancde_debug_force_wgrad_mma_sync(): A/B timing and tests)
static int g_force_sync = -1;
int launch_wgrad(const SolveArgs& a, float* gW, float* gB, cudaStream_t stream)
{
    if (g_force_sync < 0) {
        const char* e = getenv("ANCDE_WGRAD_MMA_SYNC");
        g_force_sync = (e && e[0] == '1') ? 1 : 0;
    }
    if (!g_force_sync) {
        const int rc = launch_wgrad_umma(a, gW, gB, stream);
        if (rc != ANCDE_EUNSUPPORTED) return rc;
    }
    return launch_wgrad_mma(a, gW, gB, stream);
}

}  // namespace ancde

extern "C" int ancde_debug_force_wgrad_mma_sync(int on)
{
    const int old = ancde::g_force_sync > 0 ? 1 : 0;
    ancde::g_force_sync = on ? 1 : 0;
    return old;
}
