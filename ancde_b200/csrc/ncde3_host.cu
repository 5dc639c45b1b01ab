// Host side of the UMMA (tcgen05) NCDE forward: cluster / row tiling, the packed FP16 hi/lo operand images and the
// packing kernel.
#include <stdlib.h>

#include "ncde3_common.cuh"

namespace ancde {

static int pow2ceil(int x)
{
    int p = 1;
    while (p < x) p *= 2;
    return p;
}
static int ilog2(int x)
{
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

int make_layout3(const ancde_ncde_desc* d, const NcdeLayout* st, V3Layout* lo)
{
    *lo = V3Layout();
    lo->B = d->B; lo->L = d->L; lo->C = d->C; lo->Hd = d->Hd; lo->n_hidden = d->n_hidden; lo->mode = d->mode;
    lo->att_C = d->att_C; lo->att_L = d->att_L; lo->n_hp = d->n_hp; lo->n_steps = d->n_steps; lo->n_out = d->n_out;
    lo->n_stages = 4 * d->n_steps;
    if (d->Hd > 63 || d->C > 128 || d->C < 1) return ANCDE_EUNSUPPORTED;
    if ((int64_t)d->B * (d->L - 1) * d->C >= ((int64_t)1 << 31)) return ANCDE_EUNSUPPORTED;
    for (int l = 0; l < d->n_hidden; ++l) {
        if (d->width[l] > 63) return ANCDE_EUNSUPPORTED;
        lo->width[l] = d->width[l];
        lo->in_dim[l] = (l == 0) ? d->Hd : d->width[l - 1];
    }
    lo->K = d->width[d->n_hidden - 1];
    lo->nfull = d->C / 32;
    lo->rem = d->C % 32;
    lo->seg = lo->rem ? pow2ceil(lo->rem) : 1;
    lo->seg_log2 = ilog2(lo->seg);
    lo->ipb = lo->rem ? 32 / lo->seg : 0;
    lo->nchunks = lo->nfull + (lo->rem ? 1 : 0);
    lo->KPo = 16 * ((lo->K + 1 + 15) / 16);
    lo->SBOo = lo->KPo / 8 * 128;
    lo->KPa = lo->KPo;
    int off = 0;
    for (int l = 0; l < d->n_hidden; ++l) {
        lo->KPh[l] = 16 * ((lo->in_dim[l] + 1 + 15) / 16);
        lo->SBOh[l] = lo->KPh[l] / 8 * 128;
        lo->rows_h[l] = 8 * ((lo->width[l] + 7) / 8);
        if (lo->KPh[l] > lo->KPa) lo->KPa = lo->KPh[l];
        lo->off_hhi[l] = off; off += lo->rows_h[l] / 8 * lo->SBOh[l];
        lo->off_hlo[l] = off; off += lo->rows_h[l] / 8 * lo->SBOh[l];
    }
    lo->hidden_bytes = off;
    lo->SBOa = lo->KPa / 8 * 128;
    if (st) {
        lo->sv_ntiles = st->ntiles; lo->sv_act_floats = st->act_floats; lo->sv_off_du = st->act_off_du;
        lo->sv_off_duda = st->act_off_duda; lo->sv_DUS = st->DUS; lo->sv_NCP = st->NCP; lo->sv_C4 = st->C4;
        for (int l = 0; l < d->n_hidden; ++l) lo->sv_off_h[l] = st->act_off_h[l];
    }
    // ---- single-CTA variant: narrow hidden layers run as the FP32 warp chain, the whole W_out lives in tensor memory
    if (st && st->TS == 8 && !getenv("ANCDE_FWD3_NOSINGLE")) {
        bool narrow = true;
        for (int l = 0; l < d->n_hidden; ++l) narrow = narrow && d->width[l] <= 32;
        if (narrow) {
            V3Layout s1 = *lo;
            s1.CS = 1; s1.NS = 8; s1.ns_log2 = 3;
            s1.nclusters = (d->B + 7) / 8;
            s1.i_begin[0] = 0;
            for (int r = 0; r < V3_MAX_CS; ++r) s1.i_begin[r + 1] = d->Hd;
            s1.ni_max = d->Hd;
            const int nblk = d->Hd * s1.nfull + (s1.rem ? (d->Hd + s1.ipb - 1) / s1.ipb : 0);
            s1.mtiles = (nblk + 3) / 4;
            s1.ts = 1; s1.chain1 = 1;
            s1.hidden_bytes = 0;
            s1.a_col = s1.mtiles * 16;
            s1.tcol_h = 0;
            s1.C4 = pad4(d->C);
            s1.sg_str = d->Hd * s1.C4;
            if ((s1.sg_str / 4) % 2 == 0) s1.sg_str += 4;      // odd number of 16-byte units per row: conflict-free 128-bit reads over the samples
            s1.nip_log2 = ilog2(pow2ceil(d->Hd));
            s1.wo_mat_bytes = 16 * s1.mtiles * s1.SBOo;
            s1.img_bytes = 2 * (int64_t)s1.wo_mat_bytes;
            s1.DUSTR = s1.C4; s1.du_floats = 8 * s1.C4;
            s1.KPa = s1.KPo; s1.SBOa = s1.KPa / 8 * 128;
            s1.c_hidden_floats = st->hidden_floats; s1.c_HP = pad4(st->Hmax);
            for (int l = 0; l < d->n_hidden; ++l) { s1.c_off_w[l] = st->off_w[l]; s1.c_off_b[l] = st->off_b[l]; s1.c_stride[l] = st->stride[l]; }
            if (s1.mtiles <= V3_MAX_TILES && s1.mtiles * 16 + s1.mtiles * s1.KPo <= 512 && (8 << s1.nip_log2) <= V3_THREADS &&
                v3_s1_smem(s1).total <= 227 * 1024) {
                *lo = s1;
                return ANCDE_OK;
            }
        }
    }
    static const int cand[3] = {2, 4, 8};
    for (int c = 0; c < 3; ++c) {
        const int CS = cand[c];
        if (d->Hd < CS) continue;
        lo->CS = CS;
        lo->NS = 8 * CS;
        lo->ns_log2 = ilog2(lo->NS);
        lo->nclusters = (d->B + lo->NS - 1) / lo->NS;
        const int base = d->Hd / CS, extra = d->Hd % CS;
        lo->i_begin[0] = 0;
        for (int r = 0; r < CS; ++r) lo->i_begin[r + 1] = lo->i_begin[r] + base + (r < extra ? 1 : 0);
        for (int r = CS; r < V3_MAX_CS; ++r) lo->i_begin[r + 1] = lo->i_begin[CS];
        lo->ni_max = base + (extra ? 1 : 0);
        const int nblk = lo->ni_max * lo->nfull + (lo->rem ? (lo->ni_max + lo->ipb - 1) / lo->ipb : 0);
        lo->mtiles = (nblk + 3) / 4;
        if (lo->mtiles > V3_MAX_TILES) continue;
        // TMEM columns: accumulators (hi and lo halves per tile), then either the hidden layers' accumulator or -- when
        // they fit -- the W_out slice itself (the hidden layers' accumulator then shares tile 0's columns)
        lo->ts = lo->mtiles * 2 * lo->NS + lo->mtiles * lo->KPo <= 512 && !getenv("ANCDE_FWD3_SS");
        if (!lo->ts && (lo->mtiles + 1) * 2 * lo->NS > 512) continue;
        lo->a_col = lo->mtiles * 2 * lo->NS;
        lo->C4 = pad4(d->C);
        lo->sg_str = ((lo->ni_max * lo->C4 - 8 + 31) / 32) * 32 + 8;     // = 8 (mod 32): conflict-free 128-bit reads of the contraction
        lo->nip_log2 = ilog2(pow2ceil(lo->ni_max));
        lo->wo_mat_bytes = 16 * lo->mtiles * lo->SBOo;
        lo->img_bytes = (int64_t)lo->hidden_bytes + (int64_t)CS * 2 * lo->wo_mat_bytes;
        lo->DUSTR = lo->ts ? lo->C4 : lo->NS + 4;
        lo->du_floats = lo->ts ? lo->NS * lo->C4 : pad4(d->C * lo->DUSTR);
        lo->tcol_h = lo->ts ? 0 : lo->mtiles * 2 * lo->NS;
        if (lo->ts && v3_smem(*lo).total > 227 * 1024) {      // no room for the staged rows: weights from shared memory instead
            if ((lo->mtiles + 1) * 2 * lo->NS > 512) continue;
            lo->ts = 0;
            lo->DUSTR = lo->NS + 4;
            lo->du_floats = pad4(d->C * lo->DUSTR);
            lo->tcol_h = lo->mtiles * 2 * lo->NS;
        }
        if (v3_smem(*lo).total > 227 * 1024) continue;
        bool ok = true;
        // the RK4 state lives in registers: one thread per (field row, chunk of 8 samples)
        if (d->Hd * CS > V3_THREADS) ok = false;
        if (ok) return ANCDE_OK;
    }
    return ANCDE_EUNSUPPORTED;
}

struct Pack3Args {
    V3Layout lo;
    const float* W[ANCDE_MAX_LAYERS + 1];
    const float* bias[ANCDE_MAX_LAYERS + 1];
    unsigned char* img;
    float* wpack;
};

__device__ __forceinline__ void put_hilo(unsigned char* hi_mat, unsigned char* lo_mat, int sbo, int r, int k, float v)
{
    const size_t o = (size_t)(r >> 3) * sbo + (size_t)(k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    *reinterpret_cast<__half*>(hi_mat + o) = h;
    *reinterpret_cast<__half*>(lo_mat + o) = l;
}

__global__ void pack_weights3_kernel(const __grid_constant__ Pack3Args a)
{
    const V3Layout& lo = a.lo;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (lo.chain1) {
        // FP32 rows of the hidden layers where the streamed engine keeps them (pack_weights_kernel writes the same values)
        for (int l = 0; l < lo.n_hidden; ++l) {
            const int in = lo.in_dim[l], out = lo.width[l], S = lo.c_stride[l];
            for (int idx = gtid; idx < out * S; idx += gsz) {
                const int o = idx / S, k = idx - o * S;
                a.wpack[lo.c_off_w[l] + idx] = (k < in) ? a.W[l][(int64_t)o * in + k] : 0.f;
            }
            for (int o = gtid; o < pad4(out); o += gsz) a.wpack[lo.c_off_b[l] + o] = (o < out) ? a.bias[l][o] : 0.f;
        }
    }
    for (int l = 0; l < (lo.chain1 ? 0 : lo.n_hidden); ++l) {
        const int KP = lo.KPh[l], in = lo.in_dim[l], out = lo.width[l];
        for (int idx = gtid; idx < lo.rows_h[l] * KP; idx += gsz) {
            const int m = idx / KP, k = idx - m * KP;
            float v = 0.f;
            if (m < out) v = (k < in) ? a.W[l][(int64_t)m * in + k] : (k == in ? a.bias[l][m] : 0.f);
            put_hilo(a.img + lo.off_hhi[l], a.img + lo.off_hlo[l], lo.SBOh[l], m, k, V3_WSCALE * v);
        }
    }
    const int rows = 128 * lo.mtiles, KP = lo.KPo, K = lo.K;
    const float* Wo = a.W[lo.n_hidden];
    const float* bo = a.bias[lo.n_hidden];
    for (int64_t idx = gtid; idx < (int64_t)lo.CS * rows * KP; idx += gsz) {
        const int k = (int)(idx % KP);
        const int r = (int)((idx / KP) % rows);
        const int rank = (int)(idx / ((int64_t)KP * rows));
        int i, j;
        float v = 0.f;
        if (v3_row(lo, rank, r, &i, &j)) v = (k < K) ? Wo[((int64_t)i * lo.C + j) * K + k] : (k == K ? bo[i * lo.C + j] : 0.f);
        unsigned char* hi = a.img + lo.hidden_bytes + (size_t)rank * 2 * lo.wo_mat_bytes;
        put_hilo(hi, hi + lo.wo_mat_bytes, lo.SBOo, r, k, V3_WSCALE * v);
    }
}

int pack_weights3(const ancde_ncde_desc* d, const V3Layout& lo, unsigned char* img, cudaStream_t stream)
{
    Pack3Args pa;
    pa.lo = lo;
    for (int l = 0; l <= d->n_hidden; ++l) {
        ANCDE_CHECK_ARG(d->W[l] && d->bias[l], "ncde: null weight/bias pointer for layer %d", l);
        pa.W[l] = d->W[l];
        pa.bias[l] = d->bias[l];
    }
    pa.img = img;
    pa.wpack = d->wpack;
    prof_begin("pack_weights", stream);
    pack_weights3_kernel<<<148, 256, 0, stream>>>(pa);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

}  // namespace ancde
