// Natural cubic spline coefficients with missing values, one scalar path (sample, channel) per
// thread, batched Thomas solve over the observed knots.
//
// Replaces controldiffeq/interpolate.py:7-53 (coefficients on a fully observed grid),
// :78-153 (per-scalar-path NaN handling: end-point imputation, compression to the observed
// knots, expansion of each observed piece to every interval with the offset rule) and
// controldiffeq/misc.py:12-66 (tridiagonal_solve).  The arithmetic is done op by op without
// FMA contraction, in the reference's order, so the result is bit-comparable to the CPU path.
//
// Memory: X is (N, L, C) and every output (N, L-1, C), channel innermost.  Consecutive threads
// own consecutive channels of one sample, so at a fixed knot a warp touches one contiguous run
// of 32 values: loads and stores are sector-coalesced without staging.  The three per-knot
// scratch arrays (imputed x, modified diagonal, modified rhs -> knot derivative) live in shared
// memory as [knot][thread] (bank-conflict free).  HBM-bound: algorithmic traffic is
// (L + 4(L-1)) * sizeof(T) bytes per path.
#include "common.cuh"

namespace ancde {

template <typename T>
__global__ void __launch_bounds__(128)
spline_coeffs_kernel(const T* __restrict__ times, const T* __restrict__ X, T* __restrict__ A,
                     T* __restrict__ Bc, T* __restrict__ C2, T* __restrict__ D3, int64_t P, int L, int C,
                     int tf32)
{
    typedef Ops<T> O;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    T* sT = reinterpret_cast<T*>(smem_raw);
    const int Lp = (L + 3) & ~3;
    T* sx = sT + Lp;                 // [L][NT] imputed values (NaN = unobserved)
    T* sd = sx + (size_t)L * NT;     // [L][NT] modified diagonal at observed knots
    T* sb = sd + (size_t)L * NT;     // [L][NT] modified rhs, then the knot derivative

    for (int i = tid; i < L; i += NT) sT[i] = times[i];
    __syncthreads();

    const int64_t p = (int64_t)blockIdx.x * NT + tid;
    if (p >= P) return;
    const int64_t n = p / C;
    const int c = (int)(p - n * C);
    const T* xp = X + n * (int64_t)L * C + c;
    const int64_t obase = n * (int64_t)(L - 1) * C + c;

    // ---- pass 0: load the path, find the first / last observation
    int first = -1, last = -1;
#pragma unroll 8
    for (int i = 0; i < L; ++i) {
        T v = xp[(int64_t)i * C];
        sx[i * NT + tid] = v;
        if (v == v) {
            if (first < 0) first = i;
            last = i;
        }
    }
    if (first < 0) {   // all NaN: zero coefficients (interpolate.py:93-101)
        for (int i = 0; i < L - 1; ++i) {
            int64_t o = obase + (int64_t)i * C;
            A[o] = T(0); Bc[o] = T(0); C2[o] = T(0); D3[o] = T(0);
        }
        return;
    }
    // end-point imputation (interpolate.py:104-114)
    if (first != 0) sx[tid] = sx[first * NT + tid];
    if (last != L - 1) sx[(L - 1) * NT + tid] = sx[last * NT + tid];

    // reciprocal of a time difference in the dtype of `times` (float32 for the Stock fixtures)
    auto recip = [&](int i1, int i0, T& r, T& r2) {
        if (tf32) {
            float dt = __fsub_rn((float)sT[i1], (float)sT[i0]);
            float rf = __fdiv_rn(1.0f, dt);
            r = (T)rf;
            r2 = (T)__fmul_rn(rf, rf);
        } else {
            T dt = O::sub(sT[i1], sT[i0]);
            r = O::div(T(1), dt);
            r2 = O::mul(r, r);
        }
    };

    // ---- pass 1: forward sweep over the observed knots (misc.py:57-61).  Row j of the system
    // (interpolate.py:31-43) is D_j = 2 (r_j + r_{j-1}), rhs_j = pds_j + pds_{j-1}; it is
    // completed when the NEXT observed knot is reached.
    int prev = -1, pprev = -1, m = 0;
    T r_prev = T(0), pds_prev = T(0);
    for (int i = 0; i < L; ++i) {
        T v = sx[i * NT + tid];
        if (!(v == v)) continue;
        ++m;
        if (prev >= 0) {
            T r, r2;
            recip(i, prev, r, r2);
            T pds = O::mul(O::mul(T(3), O::sub(v, sx[prev * NT + tid])), r2);
            T D = O::mul(O::add(r, r_prev), T(2));
            T rhs = O::add(pds, pds_prev);
            if (pprev >= 0) {
                T w = O::div(r_prev, sd[pprev * NT + tid]);
                D = O::sub(D, O::mul(w, r_prev));
                rhs = O::sub(rhs, O::mul(w, sb[pprev * NT + tid]));
            }
            sd[prev * NT + tid] = D;
            sb[prev * NT + tid] = rhs;
            pprev = prev;
            r_prev = r;
            pds_prev = pds;
        }
        prev = i;
    }
    {   // last row: D = 2 r_{m-2}, rhs = pds_{m-2}
        T D = O::mul(O::add(T(0), r_prev), T(2));
        T rhs = pds_prev;
        if (pprev >= 0) {
            T w = O::div(r_prev, sd[pprev * NT + tid]);
            D = O::sub(D, O::mul(w, r_prev));
            rhs = O::sub(rhs, O::mul(w, sb[pprev * NT + tid]));
        }
        sd[prev * NT + tid] = D;
        sb[prev * NT + tid] = rhs;
    }
    // ---- pass 2: back substitution (misc.py:63-65); sb becomes the knot derivative
    {
        int nxt = -1;
        for (int i = L - 1; i >= 0; --i) {
            T v = sx[i * NT + tid];
            if (!(v == v)) continue;
            if (nxt < 0) {
                sb[i * NT + tid] = O::div(sb[i * NT + tid], sd[i * NT + tid]);
            } else {
                T r, r2;
                recip(nxt, i, r, r2);
                sb[i * NT + tid] = O::div(O::sub(sb[i * NT + tid], O::mul(r, sb[nxt * NT + tid])),
                                          sd[i * NT + tid]);
            }
            nxt = i;
        }
    }
    // ---- pass 3: coefficients of each observed piece (interpolate.py:46-53), expanded to every
    // interval with offset = t_observed - t_i (interpolate.py:136-148)
    const bool two_knots = (m == 2);   // a single linear piece (interpolate.py:17-21)
    T pa = T(0), pb = T(0), pc = T(0), pd = T(0);
    int cur = 0;
    for (int i = 0; i < L - 1; ++i) {
        T v = sx[i * NT + tid];
        if (v == v) {
            cur = i;
            int ni = i + 1;
            while (!(sx[ni * NT + tid] == sx[ni * NT + tid])) ++ni;
            T x1 = sx[ni * NT + tid];
            pa = v;
            if (two_knots) {
                T dt = tf32 ? (T)__fsub_rn((float)sT[ni], (float)sT[i]) : O::sub(sT[ni], sT[i]);
                pb = O::div(O::sub(x1, v), dt);
                pc = T(0);
                pd = T(0);
            } else {
                T r, r2;
                recip(ni, i, r, r2);
                T k0 = sb[i * NT + tid], k1 = sb[ni * NT + tid];
                T six_pd = O::mul(T(2), O::mul(T(3), O::sub(x1, v)));
                pb = k0;
                // two_c = (six_pd * r - 4 k0 - 2 k1) * r
                pc = O::mul(O::sub(O::sub(O::mul(six_pd, r), O::mul(T(4), k0)), O::mul(T(2), k1)), r);
                // three_d = (-six_pd * r + 3 (k0 + k1)) * r^2
                pd = O::mul(O::add(O::mul(-six_pd, r), O::mul(T(3), O::add(k0, k1))), r2);
            }
        }
        T off = tf32 ? (T)__fsub_rn((float)sT[cur], (float)sT[i]) : O::sub(sT[cur], sT[i]);
        // a_inner = (0.5 two_c - three_d * off / 3) * off
        T a_inner = O::mul(O::sub(O::mul(T(0.5), pc), O::div(O::mul(pd, off), T(3))), off);
        int64_t o = obase + (int64_t)i * C;
        A[o] = O::add(pa, O::mul(O::sub(a_inner, pb), off));
        Bc[o] = O::add(pb, O::mul(O::sub(O::mul(pd, off), pc), off));
        C2[o] = O::sub(pc, O::mul(O::mul(T(2), pd), off));
        D3[o] = pd;
    }
}

template <typename T>
static int launch_spline(const T* times, const T* X, T* a, T* b, T* c2, T* d3, int64_t N, int L, int C,
                         int tf32, cudaStream_t stream)
{
    ANCDE_CHECK_ARG(times && X && a && b && c2 && d3, "spline_coeffs: null pointer");
    ANCDE_CHECK_ARG(L >= 2, "spline_coeffs: Must have a time dimension of size at least 2.");
    ANCDE_CHECK_ARG(C >= 1 && N >= 0, "spline_coeffs: bad shape N=%lld C=%d", (long long)N, C);
    if (N == 0) return ANCDE_OK;
    const int64_t P = N * (int64_t)C;
    const int Lp = (L + 3) & ~3;
    int nt = 128;
    size_t smem = 0;
    for (;; nt /= 2) {
        smem = ((size_t)3 * L * nt + Lp) * sizeof(T);
        if (smem <= 200 * 1024 || nt == 32) break;
    }
    if (smem > 227 * 1024) {
        set_error("spline_coeffs: L=%d needs %zu bytes of shared memory per CTA (max 227 KB)", L, smem);
        return ANCDE_EUNSUPPORTED;
    }
    ANCDE_CUDA(cudaFuncSetAttribute(spline_coeffs_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    const int64_t grid = (P + nt - 1) / nt;
    ANCDE_CHECK_ARG(grid <= 0x7fffffffLL, "spline_coeffs: too many paths");
    prof_begin(sizeof(T) == 4 ? "spline_coeffs_f32" : "spline_coeffs_f64", stream);
    spline_coeffs_kernel<T><<<(unsigned)grid, nt, smem, stream>>>(times, X, a, b, c2, d3, P, L, C, tf32);
    prof_end(stream);
    count_launch();
    ANCDE_CUDA(cudaGetLastError());
    return ANCDE_OK;
}

}  // namespace ancde

extern "C" int ancde_spline_coeffs_f32(const float* times, const float* X, float* a, float* b, float* two_c,
                                       float* three_d, int64_t N, int L, int C, void* stream)
{
    ancde::reset_launch_count();
    return ancde::launch_spline<float>(times, X, a, b, two_c, three_d, N, L, C, 0, (cudaStream_t)stream);
}

extern "C" int ancde_spline_coeffs_f64(const double* times, const double* X, double* a, double* b,
                                       double* two_c, double* three_d, int64_t N, int L, int C,
                                       int times_were_f32, void* stream)
{
    ancde::reset_launch_count();
    return ancde::launch_spline<double>(times, X, a, b, two_c, three_d, N, L, C, times_were_f32 ? 1 : 0,
                                        (cudaStream_t)stream);
}
