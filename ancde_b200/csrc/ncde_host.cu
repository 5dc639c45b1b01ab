// Host side of the fused NCDE solve: layout of packed weights / saved activations, weight packing,
// argument marshalling and the C-ABI entry points for the forward.
#include "mma_common.cuh"
#include "ncde3_common.cuh"
#include "ncde_bwd4.cuh"
#include "ncde_lane.cuh"

namespace ancde {

int launch_fwd(const SolveArgs& a, cudaStream_t stream);   // ncde_fwd.cu
int engine_of(const ancde_ncde_desc* d);

// ---- weight packing: torch (out, in) row-major -> transposed, padded layout of NcdeLayout
struct PackArgs {
    NcdeLayout lo;
    const float* W[ANCDE_MAX_LAYERS + 1];
    const float* bias[ANCDE_MAX_LAYERS + 1];
    float* wpack;
};

__global__ void pack_weights_kernel(const __grid_constant__ PackArgs a)
{
    const NcdeLayout& lo = a.lo;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int l = 0; l < lo.n_hidden; ++l) {
        const int in = lo.in_dim[l], out = lo.width[l], S = lo.stride[l], T = lo.tstride[l];
        for (int idx = gtid; idx < out * S; idx += gsz) {
            const int o = idx / S, k = idx - o * S;
            a.wpack[lo.off_w[l] + idx] = (k < in) ? a.W[l][(int64_t)o * in + k] : 0.f;
        }
        for (int idx = gtid; idx < in * T; idx += gsz) {
            const int k = idx / T, o = idx - k * T;
            a.wpack[lo.off_wt[l] + idx] = (o < out) ? a.W[l][(int64_t)o * in + k] : 0.f;
        }
        for (int o = gtid; o < pad4(out); o += gsz) a.wpack[lo.off_b[l] + o] = (o < out) ? a.bias[l][o] : 0.f;
    }
    if (gtid < 4) a.wpack[lo.off_gscale + gtid] = 0.f;
    for (int which = 0; which < 2 * lo.n_hidden; ++which) {
        // hidden layers as FP16 hi/lo A fragments: forward (which < L) A[m][k] = 64*W_l[m][k], backward A[m][k] = 64*W_l[k][m]
        const bool fwd = which < lo.n_hidden;
        const int l = fwd ? which : which - lo.n_hidden;
        const APack ap = fwd ? lo.hfp[l] : lo.hbp[l];
        uint4* base = reinterpret_cast<uint4*>(a.wpack + (fwd ? lo.off_whf : lo.off_whb)) + ap.off;
        const float* Wl = a.W[l];
        const int in = lo.in_dim[l], out = lo.width[l];
        auto el = [&](int m, int k) -> float {
            if (fwd) return (m < out && k < in) ? 64.0f * Wl[(int64_t)m * in + k] : 0.f;
            return (m < in && k < out) ? 64.0f * Wl[(int64_t)k * in + m] : 0.f;
        };
        const int steps = ap.full + ap.tail8;
        const int total = ap.mtiles * steps * 32;
        for (int idx = gtid; idx < total; idx += gsz) {
            const int lane = idx & 31, ks = (idx >> 5) % steps, mt = (idx >> 5) / steps;
            const int g = lane >> 2, t = lane & 3;
            const int m0 = mt * 16 + g, m1 = m0 + 8, k0 = ks * 16 + 2 * t;
            uint4* dst = base + mt * ap.stride + ks * 64;
            uint32_t h[4], lw[4];
            split_pair(el(m0, k0), el(m0, k0 + 1), h[0], lw[0]);
            split_pair(el(m1, k0), el(m1, k0 + 1), h[1], lw[1]);
            if (ks < ap.full) {
                split_pair(el(m0, k0 + 8), el(m0, k0 + 9), h[2], lw[2]);
                split_pair(el(m1, k0 + 8), el(m1, k0 + 9), h[3], lw[3]);
                dst[lane] = make_uint4(h[0], h[1], h[2], h[3]);
                dst[32 + lane] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            } else {
                dst[lane] = make_uint4(h[0], h[1], lw[0], lw[1]);
            }
        }
    }
    {
        int o = lo.off_hbias64;
        for (int l = 0; l < lo.n_hidden; ++l) {
            for (int i = gtid; i < lo.hfp[l].mtiles * 16; i += gsz) a.wpack[o + i] = (i < lo.width[l]) ? 64.0f * a.bias[l][i] : 0.f;
            o += lo.hfp[l].mtiles * 16;
        }
    }
    {
        // W_out^T tiles as FP16 hi/lo A fragments (m16n8k16 row-major A: a0 = (m g, k 2t..2t+1), a1 = (m g+8, ..),
        // a2 = (m g, k 2t+8..), a3 = (m g+8, k 2t+8..)); m = hidden index, k = tile row
        uint4* dst = reinterpret_cast<uint4*>(a.wpack + lo.off_wob);
        const float* Wo = a.W[lo.n_hidden];
        const int total = lo.n_wtiles * lo.MTK * 32;
        for (int idx = gtid; idx < total; idx += gsz) {
            const int lane = idx & 31, mt = (idx >> 5) % lo.MTK, tile = (idx >> 5) / lo.MTK;
            const int g = lane >> 2, t = lane & 3;
            const int ip = tile / lo.NJC, jc = tile - ip * lo.NJC;
            auto el = [&](int m, int r) -> float {
                const int i = 2 * ip + (r >> 3), j = 8 * jc + (r & 7);
                return (m < lo.K && i < lo.Hd && j < lo.C) ? 64.0f * Wo[((int64_t)i * lo.C + j) * lo.K + m] : 0.f;
            };
            const int m0 = mt * 16 + g, m1 = m0 + 8, r0 = 2 * t;
            uint32_t h[4], l[4];
            split_pair(el(m0, r0), el(m0, r0 + 1), h[0], l[0]);
            split_pair(el(m1, r0), el(m1, r0 + 1), h[1], l[1]);
            split_pair(el(m0, r0 + 8), el(m0, r0 + 9), h[2], l[2]);
            split_pair(el(m1, r0 + 8), el(m1, r0 + 9), h[3], l[3]);
            uint4* d = dst + ((int64_t)tile * lo.MTK + mt) * 64;
            d[lane] = make_uint4(h[0], h[1], h[2], h[3]);
            d[32 + lane] = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
    const int K = lo.K, NCP = lo.NCP, C4 = lo.C4, C = lo.C;
    const float* Wout = a.W[lo.n_hidden];
    const float* bout = a.bias[lo.n_hidden];
    for (int64_t idx = gtid; idx < (int64_t)K * NCP; idx += gsz) {
        const int k = (int)(idx / NCP), col = (int)(idx - (int64_t)k * NCP);
        const int i = col / C4, j = col - i * C4;
        a.wpack[lo.off_wo + idx] = (j < C) ? Wout[((int64_t)i * C + j) * K + k] : 0.f;
    }
    for (int col = gtid; col < NCP; col += gsz) {
        const int i = col / C4, j = col - i * C4;
        a.wpack[lo.off_bo + col] = (j < C) ? bout[i * C + j] : 0.f;
    }
    for (int64_t idx = gtid; idx < (int64_t)NCP * lo.KPS; idx += gsz) {
        const int col = (int)(idx / lo.KPS), k = (int)(idx - (int64_t)col * lo.KPS);
        const int i = col / C4, j = col - i * C4;
        a.wpack[lo.off_wot + idx] = (j < C && k < K) ? Wout[((int64_t)i * C + j) * K + k] : 0.f;
    }
}

// ---- host side -------------------------------------------------------------------------------
bool bwd_fits(const NcdeLayout& lo);

static void fill_layout(const ancde_ncde_desc* d, int TS, NcdeLayout* lo)
{
    *lo = NcdeLayout();
    lo->B = d->B; lo->L = d->L; lo->C = d->C; lo->Hd = d->Hd; lo->n_hidden = d->n_hidden; lo->mode = d->mode;
    lo->att_C = d->att_C; lo->att_L = d->att_L; lo->n_hp = d->n_hp; lo->n_steps = d->n_steps; lo->n_out = d->n_out;
    lo->n_stages = d->single_stage ? 1 : 4 * d->n_steps;      // single stage: the vector field at (t_start, z0) only
    lo->Hmax = d->Hd;
    for (int l = 0; l < d->n_hidden; ++l) {
        lo->width[l] = d->width[l];
        lo->in_dim[l] = (l == 0) ? d->Hd : d->width[l - 1];
        lo->stride[l] = pad4odd(lo->in_dim[l]);
        lo->tstride[l] = pad4odd(d->width[l]);
        if (d->width[l] > lo->Hmax) lo->Hmax = d->width[l];
    }
    lo->TS = TS;
    lo->ntiles = (d->B + TS - 1) / TS;
    lo->C4 = pad4(d->C);
    lo->NCH = lo->C4 / 4;
    lo->NCG = d->Hd * lo->NCH;
    lo->NCP = d->Hd * lo->C4;
    lo->K = d->width[d->n_hidden - 1];
    int off = 0;
    for (int l = 0; l < d->n_hidden; ++l) {
        lo->off_w[l] = off; off += lo->width[l] * lo->stride[l];
        lo->off_b[l] = off; off += pad4(lo->width[l]);
    }
    lo->hidden_floats = off;
    lo->off_hiddenb = off;
    for (int l = 0; l < d->n_hidden; ++l) { lo->off_wt[l] = off; off += lo->in_dim[l] * lo->tstride[l]; }
    lo->hiddenb_floats = off - lo->off_hiddenb;
    lo->off_wo = off; off += lo->K * lo->NCP;
    lo->off_bo = off; off += lo->NCP;
    lo->KPS = 8 * ((lo->K + 7) / 8) + 4;   // +4: 16-byte rows whose stride is bank-conflict free for LDS.128
    lo->off_wot = off; off += lo->NCP * lo->KPS;
    lo->NIP = (d->Hd + 1) / 2;
    lo->NJC = (d->C + 7) / 8;
    lo->MTK = (lo->K + 15) / 16;
    lo->n_wtiles = lo->NIP * lo->NJC;
    lo->magicNJC = (unsigned)((((uint64_t)1 << 32) + lo->NJC - 1) / lo->NJC);
    lo->off_wob = off; off += lo->n_wtiles * lo->MTK * 64 * 4;
    {
        int u = 0, kmax = d->Hd;
        for (int l = 0; l < d->n_hidden; ++l) { lo->hfp[l] = make_apack(lo->width[l], lo->in_dim[l], u); u += lo->hfp[l].mtiles * lo->hfp[l].stride; }
        lo->hf_u4 = u;
        lo->off_whf = off; off += u * 4;
        u = 0;
        for (int l = 0; l < d->n_hidden; ++l) { lo->hbp[l] = make_apack(lo->in_dim[l], lo->width[l], u); u += lo->hbp[l].mtiles * lo->hbp[l].stride; }
        lo->hb_u4 = u;
        lo->off_whb = off; off += u * 4;
        lo->off_hbias64 = off;
        for (int l = 0; l < d->n_hidden; ++l) { off += lo->hfp[l].mtiles * 16; if (lo->width[l] > kmax) kmax = lo->width[l]; }
        lo->KSB = (kmax + 15) / 16;
    }
    lo->off_gscale = off; off += 4;
    lo->wpack_floats = off;
    int a = d->Hd * TS;
    for (int l = 0; l < d->n_hidden; ++l) { lo->act_off_h[l] = a; a += d->width[l] * TS; }
    lo->DUS = 4 * TS + 4;
    lo->du_floats = pad4(lo->NCH * lo->DUS);
    a = pad4(a);
    lo->act_off_du = a; a += lo->du_floats;
    lo->act_off_duda = a;
    if (d->mode == ANCDE_MODE_TOP) a += lo->du_floats;
    a = pad4(a);
    lo->act_fwd_floats = a;
    lo->act_off_gbar = a; a += pad4(d->Hd * TS);
    for (int l = 0; l < d->n_hidden; ++l) { lo->act_off_delta[l] = a; a += pad4(d->width[l] * TS); }
    lo->act_floats = pad4(a);
}

int make_layout(const ancde_ncde_desc* d, NcdeLayout* lo)
{
    ANCDE_CHECK_ARG(d != nullptr, "ncde: null descriptor");
    ANCDE_CHECK_ARG(d->B >= 1 && d->L >= 2 && d->C >= 1 && d->Hd >= 1, "ncde: bad shape B=%d L=%d C=%d Hd=%d",
                    d->B, d->L, d->C, d->Hd);
    ANCDE_CHECK_ARG(d->n_hidden >= 1 && d->n_hidden <= ANCDE_MAX_LAYERS, "ncde: n_hidden=%d out of range",
                    d->n_hidden);
    ANCDE_CHECK_ARG(d->mode >= ANCDE_MODE_PLAIN && d->mode <= ANCDE_MODE_TOP, "ncde: bad mode %d", d->mode);
    ANCDE_CHECK_ARG(d->n_steps >= 1 && d->n_out >= 1, "ncde: n_steps=%d n_out=%d", d->n_steps, d->n_out);
    ANCDE_CHECK_ARG(d->grid != nullptr || d->step > 0.f, "ncde: step must be positive");
    if (d->mode == ANCDE_MODE_TOP) {
        ANCDE_CHECK_ARG(d->att_C == 1 || d->att_C == d->C, "ncde: attention channels %d (need 1 or C=%d)", d->att_C, d->C);
        ANCDE_CHECK_ARG(d->att_L >= 1 && d->n_hp >= 2, "ncde: att_L=%d n_hp=%d", d->att_L, d->n_hp);
    }
    for (int l = 0; l < d->n_hidden; ++l) ANCDE_CHECK_ARG(d->width[l] >= 1, "ncde: width[%d]=%d", l, d->width[l]);
    if (d->single_stage)
        ANCDE_CHECK_ARG(d->n_steps == 1 && d->n_out == 1 && d->mode != ANCDE_MODE_TOP,
                        "ncde: single_stage needs n_steps = n_out = 1 and a plain / bottom field (n_steps=%d n_out=%d mode=%d)",
                        d->n_steps, d->n_out, d->mode);
    int TS = d->samples_per_cta;
    const int engine = engine_of(d);
    if (engine == ANCDE_ENGINE_LANE) TS = lane_choice(d);         // the lane engine: tiles of 1 sample, or 8 for large batches
    if (TS == 0 && engine == ANCDE_ENGINE_UMMA) TS = 8;           // the UMMA forward saves in tiles of 8 samples
    if (TS == 0) {
        // the largest tile that still gives every SM work (148 SMs); weight reuse grows with TS.
        // The backward's shared-memory footprint grows with TS too, so shrink until it fits.
        TS = 8;
        while (TS > 1 && (d->B + TS - 1) / TS < 120) TS /= 2;
        for (;;) {
            fill_layout(d, TS, lo);
            if (TS == 1 || bwd_fits(*lo)) break;
            TS /= 2;
        }
    }
    ANCDE_CHECK_ARG(TS == 8 || TS == 4 || TS == 2 || TS == 1, "ncde: samples_per_cta=%d (need 8, 4, 2 or 1)", TS);
    fill_layout(d, TS, lo);
    return ANCDE_OK;
}

int pack_weights(const ancde_ncde_desc* d, const NcdeLayout& lo, cudaStream_t stream)
{
    PackArgs pa;
    pa.lo = lo;
    for (int l = 0; l <= d->n_hidden; ++l) {
        ANCDE_CHECK_ARG(d->W[l] && d->bias[l], "ncde: null weight/bias pointer for layer %d", l);
        pa.W[l] = d->W[l];
        pa.bias[l] = d->bias[l];
    }
    pa.wpack = d->wpack;
    prof_begin("pack_weights", stream);
    pack_weights_kernel<<<64, 256, 0, stream>>>(pa);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

int fill_args(const ancde_ncde_desc* d, const NcdeLayout& lo, SolveArgs* a)
{
    ANCDE_CHECK_ARG(d->times && d->t_out && d->coeff_b && d->coeff_c2 && d->coeff_d3 && d->z0 && d->z_out && d->wpack,
                    "ncde: null input/output pointer");
    if (d->mode == ANCDE_MODE_TOP) ANCDE_CHECK_ARG(d->attention && d->h_prime, "ncde: TOP mode needs attention and h_prime");
    *a = SolveArgs();
    a->lo = lo;
    a->t_start = d->t_start; a->t_end = d->t_end; a->step = d->step; a->grid = d->grid;
    a->times = d->times; a->t_out = d->t_out; a->cb = d->coeff_b; a->cc2 = d->coeff_c2; a->cd3 = d->coeff_d3;
    a->z0 = d->z0; a->attention = d->attention; a->h_prime = d->h_prime; a->wpack = d->wpack;
    a->z_out = d->z_out; a->k_out = d->k_out; a->save_act = d->save_act; a->save_G = d->save_G;
    a->grad_z_out = d->grad_z_out; a->grad_z0 = d->grad_z0; a->grad_attention = d->grad_attention;
    a->grad_k_out = d->grad_k_out;
    a->dbg = debug_buffer();
    return ANCDE_OK;
}

}  // namespace ancde

using namespace ancde;

// Engine choice: the descriptor's `engine`; 0 = the UMMA (tcgen05) forward where the shape supports it and the output
// layer is large enough to fill 128-row tensor-core tiles, else the streamed engine.
namespace ancde {
static bool umma_layouts(const ancde_ncde_desc* d, NcdeLayout* st, V3Layout* l3)
{
    if (d->B < 1 || d->L < 2 || d->C < 1 || d->Hd < 1 || d->n_hidden < 1 || d->n_hidden > ANCDE_MAX_LAYERS || d->n_steps < 1) return false;
    for (int l = 0; l < d->n_hidden; ++l)
        if (d->width[l] < 1) return false;
    if (d->samples_per_cta != 0 && d->samples_per_cta != 8) return false;
    fill_layout(d, 8, st);
    if (!bwd_fits(*st)) return false;
    return make_layout3(d, st, l3) == ANCDE_OK;
}
int engine_of(const ancde_ncde_desc* d)
{
    if (!d) return ANCDE_ENGINE_STREAMED;
    if (d->single_stage) return (d->engine == 0 || d->engine == ANCDE_ENGINE_STREAMED) ? ANCDE_ENGINE_STREAMED : -1;
    if (d->engine == ANCDE_ENGINE_STREAMED) return d->engine;
    // round 2, the latency regime and small networks at any batch size: one small CTA per sample (ncde_lane.cu; the choice
    // and its measurements are in lane_choice)
    if (d->engine == ANCDE_ENGINE_LANE) return lane_choice(d) > 0 ? ANCDE_ENGINE_LANE : -1;
    if (d->engine == 0 && lane_choice(d) > 0) return ANCDE_ENGINE_LANE;
    if (d->engine != 0 && d->engine != ANCDE_ENGINE_UMMA) return -1;
    NcdeLayout st;
    V3Layout l3;
    if (d->engine == ANCDE_ENGINE_UMMA) return umma_layouts(d, &st, &l3) ? ANCDE_ENGINE_UMMA : -1;
    // measured on B200 (Sepsis shapes, B = 1024): the UMMA forward wins once the output layer has >= ~50 k weights
    // (top network 49*35*49: 2.1 ms against 3.3 ms); below that the per-layer latency of the tensor-core round trip
    // (commit -> mbarrier -> tcgen05.ld, ~1.2 k cycles) outweighs it (bottom network 35*35*20: 2.2 ms against 1.8 ms)
    if ((int64_t)d->Hd * d->C * d->width[d->n_hidden - 1] >= 50000 && umma_layouts(d, &st, &l3)) return ANCDE_ENGINE_UMMA;
    // round 2: networks with narrow hidden layers whose whole output layer fits one SM's tensor memory run on the single-CTA
    // UMMA variant (ncde3_fwd1.cu; Sepsis bottom network 1.62 -> see DESIGN.md) once there are enough tiles of 8 samples to
    // fill the GPU
    if ((int64_t)d->Hd * d->C * d->width[d->n_hidden - 1] >= 15000 && d->B >= 8 * 100 && umma_layouts(d, &st, &l3) && l3.chain1)
        return ANCDE_ENGINE_UMMA;
    return ANCDE_ENGINE_STREAMED;
}
// float offset of the UMMA operand image inside wpack (behind the streamed engine's packed weights, 16-byte aligned)
static int64_t umma_img_offset(const NcdeLayout& st) { return (st.wpack_floats + 3) & ~(int64_t)3; }
}

extern "C" int ancde_ncde_engine(const ancde_ncde_desc* d)
{
    ANCDE_CHECK_ARG(d != nullptr, "ncde_engine: null descriptor");
    NcdeLayout lo;
    if (make_layout(d, &lo) != ANCDE_OK) return -1;
    return engine_of(d);
}

namespace ancde {
// float offset of the cluster backward's operand image inside wpack: behind everything the forward engines use
int64_t bwd4_img_offset(const ancde_ncde_desc* d, const NcdeLayout& st)
{
    int64_t end = st.wpack_floats;
    if (engine_of(d) == ANCDE_ENGINE_UMMA) {
        V3Layout l3;
        if (make_layout3(d, &st, &l3) == ANCDE_OK) end = umma_img_offset(st) + (l3.img_bytes + 3) / 4;
    }
    return (end + 31) & ~(int64_t)31;          // 128-byte aligned
}
// floats the cluster backward's image needs (0 when the shape is not supported by it, whatever the attention mode)
static int64_t bwd4_img_floats(const ancde_ncde_desc* d, const NcdeLayout& st)
{
    V4Layout v4;
    int64_t best = 0;
    for (int att = 0; att < 2; ++att) {
        if (att && (d->mode != ANCDE_MODE_TOP || d->att_C != 1)) continue;
        if (make_layout4(d, st, att != 0, 0, &v4) == ANCDE_OK && v4.img_bytes > best) best = v4.img_bytes;
        for (int cs = 2; cs <= 8; cs *= 2)
            if (make_layout4(d, st, att != 0, cs, &v4) == ANCDE_OK && v4.img_bytes > best) best = v4.img_bytes;
    }
    return (best + 3) / 4;
}
}  // namespace ancde

extern "C" int64_t ancde_ncde_wpack_floats(const ancde_ncde_desc* d)
{
    if (engine_of(d) < 0) {
        set_error("ncde: this shape is not supported by the requested engine (%s)", d->engine == ANCDE_ENGINE_LANE ? "lane" : "UMMA");
        return -1;
    }
    NcdeLayout lo;
    if (make_layout(d, &lo) != ANCDE_OK) return -1;
    if (engine_of(d) == ANCDE_ENGINE_UMMA) {
        V3Layout l3;
        if (make_layout3(d, &lo, &l3) != ANCDE_OK) return -1;
    }
    return bwd4_img_offset(d, lo) + bwd4_img_floats(d, lo);
}

extern "C" int64_t ancde_ncde_save_act_floats(const ancde_ncde_desc* d)
{
    NcdeLayout lo;
    if (make_layout(d, &lo) != ANCDE_OK) return -1;
    return (int64_t)lo.n_stages * lo.ntiles * lo.act_floats;
}

extern "C" int64_t ancde_ncde_save_G_floats(const ancde_ncde_desc* d)
{
    NcdeLayout lo;
    if (make_layout(d, &lo) != ANCDE_OK) return -1;
    return (int64_t)lo.n_stages * lo.ntiles * lo.TS * lo.NCP;
}

extern "C" int ancde_ncde_forward(const ancde_ncde_desc* d, void* stream_)
{
    reset_launch_count();
    cudaStream_t stream = (cudaStream_t)stream_;
    ANCDE_CHECK_ARG(d != nullptr, "ncde_forward: null descriptor");
    ANCDE_CHECK_ARG((d->save_act == nullptr) == (d->save_G == nullptr), "ncde_forward: save_act and save_G go together");
    const int engine = engine_of(d);
    if (engine < 0) {
        set_error("ncde_forward: this shape is not supported by the requested engine (UMMA: state / hidden widths <= 63, C <= 128, backward "
                  "tile of 8; lane: C <= 32 with C4 a power of two, widths <= 64, Hd * C4 <= 256, one sample per tile)");
        return ANCDE_EUNSUPPORTED;
    }
    NcdeLayout lo;
    int rc = make_layout(d, &lo);
    if (rc != ANCDE_OK) return rc;
    if (engine == ANCDE_ENGINE_UMMA) {
        ANCDE_CHECK_ARG(d->wpack != nullptr, "ncde: null wpack");
        V3Layout l3;
        rc = make_layout3(d, &lo, &l3);
        if (rc != ANCDE_OK) { set_error("ncde_forward: UMMA layout failed"); return rc; }
        unsigned char* img = reinterpret_cast<unsigned char*>(d->wpack + umma_img_offset(lo));
        rc = pack_weights3(d, l3, img, stream);     // (the backward packs the streamed engine's weights itself)
        if (rc != ANCDE_OK) return rc;
        return launch_fwd3(d, l3, img, stream);
    }
    SolveArgs a;
    rc = fill_args(d, lo, &a);
    if (rc != ANCDE_OK) return rc;
    rc = pack_weights(d, lo, stream);
    if (rc != ANCDE_OK) return rc;
    if (engine == ANCDE_ENGINE_LANE) return launch_lane_fwd(a, stream);
    return launch_fwd(a, stream);
}
