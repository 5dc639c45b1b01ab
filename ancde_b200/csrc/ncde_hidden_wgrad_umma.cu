// Hidden-layer weight gradients of the fused NCDE solve on the tcgen05 tensor cores (tiles of 8 samples; the layers'
// accumulators must fit the 512 columns of tensor memory: 128 per layer, 64 where in_dim + 1 <= 32).
//
//   dW_l[o][k] = sum over every (stage, sample) row r of  delta_l[r][o] * in_l[r][k],      db_l[o] = sum_r delta_l[r][o]
//
// delta_l (masked layer cotangents) and in_l (stage input for l = 0, previous activation otherwise) were left in the saved
// activation blocks by the recurrent kernels as [unit][8 samples]: for a fixed unit the 8 samples of a tile are 8
// consecutive elements of the REDUCTION index -- exactly one row of a K-major core matrix.  So a (unit, tile) item is two
// 128-bit loads, an FP16 hi/lo split and two 128-bit stores, and per chunk of 16 rows (two tiles) and layer ONE
// tcgen05.mma (M = 128, N = 128, K = 16) accumulates all four hi/lo quadrants
//     [delta_hi ; delta_lo] (128 rows)  x  [in_hi | 1 ; in_lo] (128 columns)
// in tensor memory (N = 128 columns per layer, or 64 = [32 | 32] when in_dim + 1 <= 32); the lo x lo quadrant is simply not read.
// The bias gradient is the extra column in = 1.  delta is scaled by a power of two taken from max |gbar| (tracked by the
// recurrent backward) so that the cotangents sit in FP16's range.  A CTA owns a contiguous range of chunks of 32 rows; the
// items of a chunk are loaded straight into registers one chunk ahead (consecutive threads read consecutive 32-byte
// pieces: the HBM latency is covered by registers -- a TMA ring of raw blocks in shared memory measured 0.21 ms for the
// top network, per-chunk synchronisation bound); two operand buffers so that the conversion of chunk c+1 overlaps the
// products of chunk c; at the end the accumulators are read back (tcgen05.ld) and added to the torch-layout gradients
// with atomics.
// The mma.sync kernel it replaces (ncde_wgrad.cu: 3xTF32, 0.29 + 0.19 ms for the two Sepsis networks) was issue bound.
#include <cstdlib>

#include "mma_common.cuh"
#include "umma.cuh"

namespace ancde {

struct HiddenGradPtrs { float* gW[ANCDE_MAX_LAYERS]; float* gB[ANCDE_MAX_LAYERS]; };

constexpr int HU_THREADS = 512;
constexpr int HU_TPC = 4;            // tiles of 8 samples per chunk = two K = 16 steps
constexpr int HU_LBO = 128;          // bytes between the core matrices of the two tiles
constexpr int HU_SBO = HU_TPC * HU_LBO;               // bytes between groups of 8 rows
constexpr int HU_MAT = 16 * HU_SBO;  // one operand matrix: 128 rows
constexpr int HU_MAXL = ANCDE_MAX_LAYERS;

// HU_IPT: (unit, tile) items per thread and chunk; DEEP: loaded two chunks ahead instead of one (few items per thread: the
// bytes in flight decide the HBM rate)
template <int HU_IPT, bool DEEP>
__global__ void __launch_bounds__(HU_THREADS, 1) ncde_hidden_wgrad_umma_kernel(const __grid_constant__ SolveArgs p,
                                                                               const __grid_constant__ HiddenGradPtrs hp, int chunks_per_cta,
                                                                               const unsigned* __restrict__ gmax_bits)
{
    constexpr int NT = HU_THREADS;
    const NcdeLayout& lo = p.lo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int nl = lo.n_hidden;
    // accumulator columns of layer l: [cb[l], cb[l] + nb[l]), in_hi | 1 in the first half, in_lo in the second
    __shared__ int nb[HU_MAXL], cb[HU_MAXL];
    int ncols = 0;
    for (int l = 0; l < nl; ++l) {
        const int w = (lo.in_dim[l] + 1 <= 32) ? 64 : 128;
        if (tid == 0) { nb[l] = w; cb[l] = ncols; }
        ncols += w;
    }
    const uint32_t tmem_cols = ncols <= 32 ? 32u : (ncols <= 64 ? 64u : (ncols <= 128 ? 128u : (ncols <= 256 ? 256u : 512u)));

    // power-of-two scale that brings the largest stage cotangent to [1, 2): the layer cotangents are sums of such terms and
    // stay far below FP16's 65504
    float scale = 1.0f;
    {
        const float gm = __uint_as_float(__ldg(gmax_bits));
        if (gm > 0.f) {
            int e;
            frexpf(gm, &e);                    // gm = f * 2^e, f in [0.5, 1)
            scale = ldexpf(1.0f, 1 - e);
        }
    }

    __shared__ __align__(8) uint64_t mma_bar[2];           // products that read operand buffer b have completed
    __shared__ uint32_t tslot;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sOp = smem_raw;                                          // [2 buffers][layer][A, B][HU_MAT]
    const int op_bytes = nl * 2 * HU_MAT;
    int n_items = 0;
    for (int l = 0; l < nl; ++l) n_items += lo.width[l] + lo.in_dim[l];

    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + HU_TPC - 1) / HU_TPC;
    const int64_t c0 = (int64_t)blockIdx.x * chunks_per_cta;
    int64_t c1 = c0 + chunks_per_cta;
    if (c1 > nchunks) c1 = nchunks;
    const int n_it = (c1 > c0) ? (int)(c1 - c0) : 0;

    // this thread's items of a chunk: item = tid + u * NT  ->  (tile t, entry e); source offset inside the saved block,
    // destination of the hi row inside an operand buffer (bit 31: scaled, i.e. a delta row); the same for every chunk
    int i_src[HU_IPT], i_dst[HU_IPT], i_tile[HU_IPT];
#pragma unroll
    for (int u = 0; u < HU_IPT; ++u) {
        const int idx = tid + u * NT;
        i_tile[u] = -1; i_src[u] = 0; i_dst[u] = 0;
        if (idx < HU_TPC * n_items) {
            const int t = idx / n_items, e = idx - t * n_items;
            int l = 0, base = 0;
            while (e >= base + lo.width[l] + lo.in_dim[l]) { base += lo.width[l] + lo.in_dim[l]; ++l; }
            const int r = e - base;
            i_tile[u] = t;
            if (r < lo.width[l]) {                             // A rows: delta_l, scaled
                i_src[u] = lo.act_off_delta[l] + r * 8;
                i_dst[u] = (l * 2 * HU_MAT + (r >> 3) * HU_SBO + (r & 7) * 16 + t * HU_LBO) | (int)0x80000000;
            } else {                                           // B rows: in_l
                const int k = r - lo.width[l];
                i_src[u] = ((l == 0) ? 0 : lo.act_off_h[l - 1]) + k * 8;
                // (bit 30: narrow layer -- its lo rows sit 32 rows down, not 64)
                i_dst[u] = (l * 2 * HU_MAT + HU_MAT + (k >> 3) * HU_SBO + (k & 7) * 16 + t * HU_LBO) | ((lo.in_dim[l] + 1 <= 32) ? 0x40000000 : 0);
            }
        }
    }
    auto load_items = [&](int64_t ch, float4 (&g)[2 * HU_IPT]) {
#pragma unroll
        for (int u = 0; u < HU_IPT; ++u) {
            g[2 * u] = g[2 * u + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int64_t blk = ch * HU_TPC + i_tile[u];
            if (i_tile[u] >= 0 && ch < c1 && blk < nblocks) {
                const float4* src = reinterpret_cast<const float4*>(p.save_act + blk * lo.act_floats + i_src[u]);
                g[2 * u] = __ldcs(src);
                g[2 * u + 1] = __ldcs(src + 1);
            }
        }
    };

    if (tid == 0) {
        mbar_init(mma_bar + 0, 1);
        mbar_init(mma_bar + 1, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tslot, tmem_cols);
    // operand buffers: zero (rows beyond a layer's units stay zero), the table
    for (int i = tid; i < 2 * op_bytes / 16; i += NT) reinterpret_cast<uint4*>(sOp)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // the bias column of every B matrix (row k = in_dim): ones in the hi half, both tiles, both buffers
    if (tid < 2 * nl * HU_TPC) {
        const int ob = tid / (nl * HU_TPC), r = tid - ob * nl * HU_TPC, l = r / HU_TPC, t = r - l * HU_TPC;
        const int k = lo.in_dim[l];
        uint4* dst = reinterpret_cast<uint4*>(sOp + ob * op_bytes + l * 2 * HU_MAT + HU_MAT + (k >> 3) * HU_SBO + t * HU_LBO + (k & 7) * 16);
        const uint32_t one2 = 0x3C003C00u;                   // two FP16 ones
        *dst = make_uint4(one2, one2, one2, one2);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;

    float4 g0[2 * HU_IPT], g1[2 * HU_IPT], g2[DEEP ? 2 * HU_IPT : 1];
    load_items(c0, g0);
    if (DEEP) load_items(c0 + 1, g1);
    for (int it = 0; it < n_it; ++it) {
        const int64_t ch = c0 + it;
        const int ob = it & 1;
        unsigned char* sO = sOp + ob * op_bytes;
        if constexpr (DEEP) load_items(ch + 2, reinterpret_cast<float4 (&)[2 * HU_IPT]>(g2));      // two chunks ahead
        else load_items(ch + 1, g1);                              // one chunk ahead
        if (it >= 2) mbar_wait(mma_bar + ob, (uint32_t)(((it - 2) >> 1) & 1));       // the products of chunk it-2 have read this buffer
        // ---- (unit, tile) items -> FP16 hi/lo rows of the operands (tiles beyond the last block were loaded as zeros)
#pragma unroll
        for (int u = 0; u < HU_IPT; ++u) {
            if (i_tile[u] < 0) continue;
            const float f = i_dst[u] < 0 ? scale : 1.0f;
            const float4 a = g0[2 * u], c = g0[2 * u + 1];
            uint32_t hh[4], hl[4];
            split_pair(a.x * f, a.y * f, hh[0], hl[0]);
            split_pair(a.z * f, a.w * f, hh[1], hl[1]);
            split_pair(c.x * f, c.y * f, hh[2], hl[2]);
            split_pair(c.z * f, c.w * f, hh[3], hl[3]);
            unsigned char* dst = sO + (i_dst[u] & 0x3fffffff);
            *reinterpret_cast<uint4*>(dst) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(dst + ((i_dst[u] & 0x40000000) ? 4 : 8) * HU_SBO) = make_uint4(hl[0], hl[1], hl[2], hl[3]);      // the lo half
        }
        fence_async_smem();
        __syncthreads();
        // ---- D_l += [delta_hi ; delta_lo] x [in_hi | 1 ; in_lo]^T over the 16 rows
        if (warp_u == 1) {
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int l = 0; l < HU_MAXL; ++l) {
                    if (l < nl) {
                        const uint64_t dA = umma_smem_desc(smem_u32(sO) + l * 2 * HU_MAT, HU_LBO, HU_SBO);
                        const uint64_t dB = umma_smem_desc(smem_u32(sO) + l * 2 * HU_MAT + HU_MAT, HU_LBO, HU_SBO);
#pragma unroll
                        for (int ks = 0; ks < HU_TPC / 2; ++ks)      // a K = 16 step = two tiles = 2 * HU_LBO bytes = 16 descriptor units
                            umma_f16(tbase + cb[l], dA + (uint64_t)(ks * 16), dB + (uint64_t)(ks * 16), umma_idesc_f16(128, nb[l], false), (it | ks) != 0);
                    }
                }
                umma_commit(mma_bar + ob);
            }
        }
#pragma unroll
        for (int e = 0; e < 2 * HU_IPT; ++e) {
            g0[e] = g1[e];
            if (DEEP) g1[e] = g2[e];
        }
    }

    // ---- results: TMEM lane = accumulator row (this warp's quarter), 32 columns per warp group and layer
    if (n_it > 0) {
        for (int ob = 0; ob < 2; ++ob)
            if (n_it > ob) mbar_wait(mma_bar + ob, (uint32_t)(((n_it - 1 - ob) >> 1) & 1));
        tc_fence_after();
        const float inv = 1.0f / scale;
        const int wq = warp & 3, cg = warp >> 2;
        const int m = 32 * wq + lane;
        const bool lo_row = m >= 64;                 // rows 64.. hold delta_lo: only their product with in_hi (first half) counts
        const int o = lo_row ? m - 64 : m;
        for (int l = 0; l < nl; ++l) {
            const int N = lo.width[l], Kin = lo.in_dim[l], half = nb[l] >> 1;
            for (int c16 = cg; c16 < (nb[l] >> 4); c16 += 4) {           // 16 columns at a time, the four warp groups in turn
                float v[16];
                tmem_ld16(tbase + ((uint32_t)(32 * wq) << 16) + cb[l] + 16 * c16, v);
                const bool second = 16 * c16 >= half;                       // in_lo columns
                if (o < N && !(lo_row && second)) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int k = 16 * c16 + u - (second ? half : 0);
                        const float val = v[u] * inv;
                        if (k < Kin) atomicAdd(hp.gW[l] + (int64_t)o * Kin + k, val);
                        else if (k == Kin && !second) atomicAdd(hp.gB[l] + o, val);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, tmem_cols);
}

// true when this kernel takes the shape
bool hidden_wgrad_umma_ok(const NcdeLayout& lo)
{
    if (lo.TS != 8 || lo.n_hidden < 1 || lo.n_hidden > HU_MAXL) return false;
    int ncols = 0;
    for (int l = 0; l < lo.n_hidden; ++l) {
        if (lo.width[l] > 64 || lo.in_dim[l] + 1 > 64) return false;
        ncols += (lo.in_dim[l] + 1 <= 32) ? 64 : 128;
    }
    if (ncols > 512) return false;
    if (lo.act_off_delta[0] < lo.act_off_du) return false;
    return true;
}

int launch_hidden_wgrad_umma(const SolveArgs& a, const HiddenGradPtrs& hp, cudaStream_t stream)
{
    const NcdeLayout& lo = a.lo;
    int n_items = 0;
    for (int l = 0; l < lo.n_hidden; ++l) n_items += lo.width[l] + lo.in_dim[l];
    const size_t smem = 2 * (size_t)lo.n_hidden * 2 * HU_MAT + 128;
    if (smem > 227 * 1024 || HU_TPC * n_items > 4 * HU_THREADS) return ANCDE_EUNSUPPORTED;
    const bool few = HU_TPC * n_items <= 2 * HU_THREADS;
    const int64_t nblocks = (int64_t)lo.n_stages * lo.ntiles;
    const int64_t nchunks = (nblocks + HU_TPC - 1) / HU_TPC;
    // Every CTA ends with one atomic per weight (its accumulators are partial sums over its rows): with few rows (small batches of
    // small networks) 148 CTAs spend longer on 148 x (weights) contended atomics than on their rows, so a CTA gets at least
    // `min_per` chunks of 32 rows.  Measured on B200 (min_per 1 / 8 / 32 / 128): UEA B = 32 4.41 / 4.15 / 4.14 / 4.21 ms per step,
    // Stock B = 200 0.57 / 0.54 / 0.55 / 0.61, Mujoco B = 1024 1.05 / 1.05 / 1.03 / 1.08; the Sepsis batch has 62 chunks per CTA anyway.
    static const int min_per = getenv("ANCDE_HWGRAD_MIN_CHUNKS") ? atoi(getenv("ANCDE_HWGRAD_MIN_CHUNKS")) : 16;
    int64_t grid = 148;
    if (grid > nchunks) grid = nchunks;
    int per = (int)((nchunks + grid - 1) / grid);
    if (per < min_per) per = min_per;
    grid = (nchunks + per - 1) / per;
    const unsigned* gmax = reinterpret_cast<const unsigned*>(a.wpack + lo.off_gscale);   // written by the recurrent backward
    auto kern = few ? ncde_hidden_wgrad_umma_kernel<2, true> : ncde_hidden_wgrad_umma_kernel<4, false>;
    ANCDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const char* names[3] = {"ncde_hidden_wgrad_plain", "ncde_hidden_wgrad_bottom", "ncde_hidden_wgrad_top"};
    prof_begin(names[lo.mode], stream);
    kern<<<(unsigned)grid, HU_THREADS, smem, stream>>>(a, hp, per, gmax);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

}  // namespace ancde
