// Host side of the tensor-core (v2) NCDE solve: layouts of the packed FP16 hi/lo weight fragments and of the
// buffers saved for the backward, the packing kernel and argument marshalling.
#include "ncde2_common.cuh"

namespace ancde {

int make_layout2(const ancde_ncde_desc* d, V2Layout* lo)
{
    ANCDE_CHECK_ARG(d != nullptr, "ncde: null descriptor");
    ANCDE_CHECK_ARG(d->B >= 1 && d->L >= 2 && d->C >= 1 && d->Hd >= 1, "ncde: bad shape B=%d L=%d C=%d Hd=%d", d->B, d->L, d->C,
                    d->Hd);
    ANCDE_CHECK_ARG(d->n_hidden >= 1 && d->n_hidden <= ANCDE_MAX_LAYERS, "ncde: n_hidden=%d out of range", d->n_hidden);
    ANCDE_CHECK_ARG(d->mode >= ANCDE_MODE_PLAIN && d->mode <= ANCDE_MODE_TOP, "ncde: bad mode %d", d->mode);
    ANCDE_CHECK_ARG(d->n_steps >= 1 && d->n_out >= 1, "ncde: n_steps=%d n_out=%d", d->n_steps, d->n_out);
    ANCDE_CHECK_ARG(d->step > 0.f, "ncde: step must be positive");
    if (d->mode == ANCDE_MODE_TOP) {
        ANCDE_CHECK_ARG(d->att_C == 1 || d->att_C == d->C, "ncde: attention channels %d (need 1 or C=%d)", d->att_C, d->C);
        ANCDE_CHECK_ARG(d->att_L >= 1 && d->n_hp >= 2, "ncde: att_L=%d n_hp=%d", d->att_L, d->n_hp);
    }
    for (int l = 0; l < d->n_hidden; ++l) ANCDE_CHECK_ARG(d->width[l] >= 1, "ncde: width[%d]=%d", l, d->width[l]);
    *lo = V2Layout();
    lo->B = d->B; lo->L = d->L; lo->C = d->C; lo->Hd = d->Hd; lo->n_hidden = d->n_hidden; lo->mode = d->mode;
    lo->att_C = d->att_C; lo->att_L = d->att_L; lo->n_hp = d->n_hp; lo->n_steps = d->n_steps; lo->n_out = d->n_out;
    lo->n_stages = 4 * d->n_steps;
    int kmax = d->Hd;
    for (int l = 0; l < d->n_hidden; ++l) {
        lo->width[l] = d->width[l];
        lo->in_dim[l] = (l == 0) ? d->Hd : d->width[l - 1];
        if (d->width[l] > kmax) kmax = d->width[l];
    }
    lo->K = d->width[d->n_hidden - 1];
    if (lo->K > 16 * V2_MAX_KSTEPS) {
        set_error("ncde: last hidden width %d exceeds %d (tensor-core solver limit)", lo->K, 16 * V2_MAX_KSTEPS);
        return ANCDE_EUNSUPPORTED;
    }
    lo->KSB = (kmax + 15) / 16;
    lo->NIP = (d->Hd + 1) / 2;
    lo->NJC = (d->C + 7) / 8;
    lo->C8 = 8 * lo->NJC;
    lo->ntiles = lo->NIP * lo->NJC;
    int off = 0;   // uint4 units
    for (int l = 0; l < d->n_hidden; ++l) {
        lo->hf[l] = make_apack(lo->width[l], lo->in_dim[l], off);
        off += lo->hf[l].mtiles * lo->hf[l].stride;
    }
    lo->hidden_u4 = off;
    for (int l = 0; l < d->n_hidden; ++l) {
        lo->hb[l] = make_apack(lo->in_dim[l], lo->width[l], off);
        off += lo->hb[l].mtiles * lo->hb[l].stride;
    }
    lo->hidden_b_u4 = off - lo->hidden_u4;
    lo->of = make_apack(lo->ntiles * 16, lo->K, off);
    off += lo->of.mtiles * lo->of.stride;
    lo->ob = make_apack(lo->K, 16, off);
    lo->ob_tile = lo->ob.mtiles * 64;
    off += lo->ntiles * lo->ob_tile;
    int foff = off * 4;   // float units from here on
    for (int l = 0; l < d->n_hidden; ++l) { lo->off_hbias[l] = foff; foff += lo->hf[l].mtiles * 16; }
    lo->off_obias = foff; foff += lo->ntiles * 16;
    lo->wpack_floats = pad4(foff);
    lo->MTz = (d->Hd + 15) / 16;
    // batch tiling: decided by the forward's shared-memory fit (ncde2_fwd.cu)
    V2Launch la;
    lo->NT = 0;
    lo->magicC = (unsigned)((((uint64_t)1 << 32) + d->C - 1) / d->C);
    const int rc = choose_launch2_fwd(lo, &la);
    if (rc != ANCDE_OK) return rc;
    int o = 0;
    const int frag = lo->NT * 128;
    lo->fo_z = o; o += lo->MTz * frag;
    for (int l = 0; l < d->n_hidden; ++l) { lo->fo_h[l] = o; o += lo->hf[l].mtiles * frag; }
    lo->fo_du = o; o += pad4(d->C * lo->NS);
    lo->fo_duda = o;
    if (d->mode == ANCDE_MODE_TOP) o += pad4(d->C * lo->NS);
    lo->fo_gbar = o; o += lo->MTz * frag;
    for (int l = 0; l < d->n_hidden; ++l) { lo->fo_delta[l] = o; o += lo->hf[l].mtiles * frag; }
    lo->act_block = o;
    return ANCDE_OK;
}

// ---- weight packing -------------------------------------------------------------------------------------------
struct Pack2Args {
    V2Layout lo;
    const float* W[ANCDE_MAX_LAYERS + 1];
    const float* bias[ANCDE_MAX_LAYERS + 1];
    uint4* wpack;
};

// element (m, k) of the logical A matrix of pack `which`: 0..L-1 hidden forward, L..2L-1 hidden backward, 2L output
// forward, 2L+1 output backward (m = tile*mtiles(K)*16 + k index, k = tile row)
__device__ __forceinline__ float a_elem(const Pack2Args& a, int which, int m, int k)
{
    const V2Layout& lo = a.lo;
    const int L = lo.n_hidden;
    if (which < L) {
        const int l = which;
        return (m < lo.width[l] && k < lo.in_dim[l]) ? a.W[l][(int64_t)m * lo.in_dim[l] + k] : 0.f;
    }
    if (which < 2 * L) {
        const int l = which - L;
        return (m < lo.in_dim[l] && k < lo.width[l]) ? a.W[l][(int64_t)k * lo.in_dim[l] + m] : 0.f;
    }
    int tile, r, kk;
    if (which == 2 * L) { tile = m >> 4; r = m & 15; kk = k; }
    else { const int per = lo.ob.mtiles * 16; tile = m / per; kk = m - tile * per; r = k; }
    const int ip = tile / lo.NJC, jc = tile - ip * lo.NJC;
    const int i = 2 * ip + (r >> 3), j = 8 * jc + (r & 7);
    return (i < lo.Hd && j < lo.C && kk < lo.K) ? a.W[L][((int64_t)i * lo.C + j) * lo.K + kk] : 0.f;
}

__device__ __forceinline__ uint32_t pack_h2(float x0, float x1, bool want_lo)
{
    uint32_t hi, lo;
    split_pair(x0 * V2_WSCALE, x1 * V2_WSCALE, hi, lo);
    return want_lo ? lo : hi;
}

__global__ void pack_weights2_kernel(const __grid_constant__ Pack2Args a)
{
    const V2Layout& lo = a.lo;
    const int L = lo.n_hidden;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int which = 0; which < 2 * L + 2; ++which) {
        APack ap;
        int mt_total;
        if (which < L) { ap = lo.hf[which]; mt_total = ap.mtiles; }
        else if (which < 2 * L) { ap = lo.hb[which - L]; mt_total = ap.mtiles; }
        else if (which == 2 * L) { ap = lo.of; mt_total = ap.mtiles; }
        else { ap = lo.ob; mt_total = lo.ntiles * ap.mtiles; }
        const int steps = ap.full + ap.tail8;
        const int64_t total = (int64_t)mt_total * steps * 32;
        for (int64_t idx = gtid; idx < total; idx += gsz) {
            const int lane = (int)(idx & 31);
            const int ks = (int)((idx >> 5) % steps);
            const int mt = (int)((idx >> 5) / steps);
            const int g = lane >> 2, t = lane & 3;
            const int m0 = mt * 16 + g, m1 = m0 + 8, k0 = ks * 16 + 2 * t;
            uint4* dst = a.wpack + ap.off + (int64_t)mt * ap.stride + ks * 64;
            const float e00 = a_elem(a, which, m0, k0), e01 = a_elem(a, which, m0, k0 + 1);
            const float e10 = a_elem(a, which, m1, k0), e11 = a_elem(a, which, m1, k0 + 1);
            if (ks < ap.full) {
                const float f00 = a_elem(a, which, m0, k0 + 8), f01 = a_elem(a, which, m0, k0 + 9);
                const float f10 = a_elem(a, which, m1, k0 + 8), f11 = a_elem(a, which, m1, k0 + 9);
                dst[lane] = make_uint4(pack_h2(e00, e01, false), pack_h2(e10, e11, false), pack_h2(f00, f01, false), pack_h2(f10, f11, false));
                dst[32 + lane] = make_uint4(pack_h2(e00, e01, true), pack_h2(e10, e11, true), pack_h2(f00, f01, true), pack_h2(f10, f11, true));
            } else {
                dst[lane] = make_uint4(pack_h2(e00, e01, false), pack_h2(e10, e11, false), pack_h2(e00, e01, true), pack_h2(e10, e11, true));
            }
        }
    }
    float* wf = reinterpret_cast<float*>(a.wpack);
    for (int l = 0; l < L; ++l)
        for (int o = gtid; o < lo.hf[l].mtiles * 16; o += gsz) wf[lo.off_hbias[l] + o] = (o < lo.width[l]) ? V2_WSCALE * a.bias[l][o] : 0.f;
    for (int idx = gtid; idx < lo.ntiles * 16; idx += gsz) {
        const int tile = idx >> 4, r = idx & 15;
        const int ip = tile / lo.NJC, jc = tile - ip * lo.NJC;
        const int i = 2 * ip + (r >> 3), j = 8 * jc + (r & 7);
        wf[lo.off_obias + idx] = (i < lo.Hd && j < lo.C) ? V2_WSCALE * a.bias[L][i * lo.C + j] : 0.f;
    }
}

int pack_weights2(const ancde_ncde_desc* d, const V2Layout& lo, cudaStream_t stream)
{
    Pack2Args pa;
    pa.lo = lo;
    for (int l = 0; l <= d->n_hidden; ++l) {
        ANCDE_CHECK_ARG(d->W[l] && d->bias[l], "ncde: null weight/bias pointer for layer %d", l);
        pa.W[l] = d->W[l];
        pa.bias[l] = d->bias[l];
    }
    pa.wpack = reinterpret_cast<uint4*>(d->wpack);
    prof_begin("pack_weights", stream);
    pack_weights2_kernel<<<148, 256, 0, stream>>>(pa);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

int fill_args2(const ancde_ncde_desc* d, const V2Layout& lo, V2Args* a)
{
    ANCDE_CHECK_ARG(d->times && d->t_out && d->coeff_b && d->coeff_c2 && d->coeff_d3 && d->z0 && d->z_out && d->wpack,
                    "ncde: null input/output pointer");
    ANCDE_CHECK_ARG((reinterpret_cast<uintptr_t>(d->wpack) & 15) == 0, "ncde: wpack must be 16-byte aligned");
    if (d->mode == ANCDE_MODE_TOP) ANCDE_CHECK_ARG(d->attention && d->h_prime, "ncde: TOP mode needs attention and h_prime");
    *a = V2Args();
    a->lo = lo;
    a->t_start = d->t_start; a->t_end = d->t_end; a->step = d->step;
    a->times = d->times; a->t_out = d->t_out; a->cb = d->coeff_b; a->cc2 = d->coeff_c2; a->cd3 = d->coeff_d3;
    a->z0 = d->z0; a->attention = d->attention; a->h_prime = d->h_prime;
    a->wpack = reinterpret_cast<const uint4*>(d->wpack);
    a->z_out = d->z_out; a->k_out = d->k_out; a->save_act = d->save_act; a->save_G = d->save_G;
    a->grad_z_out = d->grad_z_out; a->grad_z0 = d->grad_z0; a->grad_attention = d->grad_attention;
    a->dbg = debug_buffer();
    return ANCDE_OK;
}

}  // namespace ancde
