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
// of 32 values: loads and stores are sector-coalesced without staging.  HBM-bound: algorithmic
// traffic is (L + 4(L-1)) * sizeof(T) bytes per path.
//
// SIMT efficiency: with 90 % of the entries missing (Sepsis) nearly every knot is observed by SOME
// lane of a warp, so loops over knots with a per-lane "observed?" branch run at ~10 % lane
// utilisation.  Instead pass 0 compacts the observed knot indices of each path into shared memory
// ([ordinal][thread], 16 bit; the values are re-read from X), the two Thomas sweeps iterate over ORDINALS (all lanes busy until
// the shortest path ends) and the expansion pass walks the L-1 intervals in lock step.
//
// Measured alternatives (B200, 29 959 x 72 x 35 floats, 90 % missing; this kernel: 0.825 ms = 28 % of the HBM roofline,
// 16.5 of 32 threads active per instruction, issue slots 47 % busy):
//  * piece coefficients computed in the ordinal loop of pass 2 and parked in the output slots (saves ~45 divergent
//    instructions per interval, costs three stores and three L2 loads per ordinal): 0.916 ms;
//  * a tile kernel -- one CTA per sample, X block and every scratch value in shared memory, thread = channel for the
//    compaction and the Thomas sweeps, thread = (interval, channel) element for the expansion, fully coalesced 128-bit
//    loads and 128-byte stores, bit-identical results: 0.951 ms (the per-channel phases keep 35 of 128 threads busy and
//    serialise ~15 k cycles per sample);
//  * (round 2) the same tile idea with one WARP per path -- ballot/popc compaction, per-ordinal terms and piece
//    coefficients one per lane, lane 0 alone in the two Thomas sweeps, intervals one per lane, outputs staged in shared
//    memory and stored as contiguous runs, bit-identical: 1.56 ms.  The serial sweeps (one division per ordinal on ONE
//    lane) and five paths per warp in sequence leave ~44 k cycles per sample and CTA; 32 independent paths per warp, as
//    here, hide that latency far better.
// What this kernel does instead about the expensive piece set-up in the interval loop: paths with at most SPL_MAXP pieces
// (sparse data: Sepsis has ~8 of 71) compute every piece's coefficients in the back-substitution loop, where all lanes of
// a warp are busy, and park them in shared memory; the interval loop then enters a piece with four loads.
#include <stdlib.h>

#include "common.cuh"

namespace ancde {

constexpr int SPL_MAXP = 16;     // pieces per path whose coefficients are parked in shared memory

// IDX: type of the ordinal -> knot index list in shared memory (8 bit for L <= 256 knots: shared memory per CTA bounds the
// resident warps of this latency-bound kernel)
template <typename T, typename IDX>
__global__ void __launch_bounds__(256)
spline_coeffs_kernel(const T* __restrict__ times, const T* __restrict__ X, T* __restrict__ A,
                     T* Bc, T* __restrict__ C2, T* D3, int64_t P, int L, int C, int tf32)
{
    typedef Ops<T> O;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    T* sT = reinterpret_cast<T*>(smem_raw);
    const int Lp = (L + 3) & ~3;
    T* sPB = sT + Lp;                                                    // [SPL_MAXP][NT] parked piece coefficients b, two_c, three_d
    T* sPC = sPB + SPL_MAXP * NT;
    T* sPD = sPC + SPL_MAXP * NT;
    IDX* sidx = reinterpret_cast<IDX*>(sPD + SPL_MAXP * NT);   // [L+2][NT] knot index per ordinal

    for (int i = tid; i < L; i += NT) sT[i] = times[i];
    __syncthreads();

    const int64_t p = (int64_t)blockIdx.x * NT + tid;
    if (p >= P) return;
    const int64_t n = p / C;
    const int c = (int)(p - n * C);
    const T* xp = X + n * (int64_t)L * C + c;
    const int64_t obase = n * (int64_t)(L - 1) * C + c;

    // ---- pass 0: compact the observed knots (slot 0 is kept free for an imputed first knot)
    int pos = 1, first = -1, last = -1;
    T xfirst = T(0), xlast = T(0);
#pragma unroll 8
    for (int i = 0; i < L; ++i) {
        const T v = xp[(int64_t)i * C];
        if (v == v) {
            if (first < 0) { first = i; xfirst = v; }
            last = i;
            xlast = v;
            sidx[pos * NT + tid] = (IDX)i;
            ++pos;
        }
    }
    if (first < 0) {   // all NaN: zero coefficients (interpolate.py:93-101)
        for (int i = 0; i < L - 1; ++i) {
            const int64_t o = obase + (int64_t)i * C;
            A[o] = T(0); Bc[o] = T(0); C2[o] = T(0); D3[o] = T(0);
        }
        return;
    }
    // end-point imputation (interpolate.py:104-114): knot 0 takes the first observed value, knot L-1 the last
    const bool imp_first = first != 0, imp_last = last != L - 1;
    int base = 1;
    if (imp_first) { sidx[tid] = 0; base = 0; }
    if (imp_last) { sidx[pos * NT + tid] = (IDX)(L - 1); ++pos; }
    const int m = pos - base;                        // observed knots after imputation, >= 2
    const IDX* jx = sidx + base * NT + tid;     // jx[j*NT] = knot index of ordinal j
    // value of ordinal j (knot i): read again from X (L1/L2 resident after pass 0) instead of a shared-memory copy --
    // shared memory per thread drops from 6 to 2 bytes per knot, which doubles the resident warps of this
    // latency-bound kernel (measured: 0.883 -> 0.825 ms on the Sepsis-size split); the two imputed ends come from
    // registers.  (Also measured, and rejected: computing the piece coefficients in the ordinal loop of pass 2 and
    // parking them in the output slots saves ~45 instructions per interval but adds three stores and three L2 loads
    // per ordinal: 0.916 ms.)
    auto xval = [&](int j, int i) -> T {
        if (j == 0 && imp_first) return xfirst;
        if (j == m - 1 && imp_last) return xlast;
        return __ldg(xp + (int64_t)i * C);
    };

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

    // Scratch of the Thomas solve: the modified diagonal / rhs of ordinal j live in interval slot j of this
    // path's OWN output arrays (D3 / Bc, coalesced, L1/L2 resident) until pass 3 overwrites them; only m of the
    // L-1 slots are touched.  The last ordinal (m-1 can be L-1, one past the slots) stays in registers.
    // Paths with at most SPL_MAXP pieces keep that scratch in SHARED memory instead (the slots that later hold the parked
    // piece coefficients): no global round trip inside the two dependent sweeps.
    const bool park = (m - 1 <= SPL_MAXP) && (m > 2);     // (a single linear piece takes the general path)
    T* sd = park ? sPC + tid : D3 + obase;
    T* sb = park ? sPB + tid : Bc + obase;
    const int64_t SC = park ? NT : C;       // stride between ordinal slots
    T k_last;
    // ---- pass 1: forward sweep over the ordinals (misc.py:57-61).  Row j of the system
    // (interpolate.py:31-43) is D_j = 2 (r_j + r_{j-1}), rhs_j = pds_j + pds_{j-1}.
    {
        int iprev = jx[0];
        T xprev = xval(0, iprev);
        T r_prev = T(0), pds_prev = T(0), Dpp = T(0), bpp = T(0);
        for (int j = 1; j < m; ++j) {
            const int ij = jx[j * NT];
            const T xj = xval(j, ij);
            T r, r2;
            recip(ij, iprev, r, r2);
            const T pds = O::mul(O::mul(T(3), O::sub(xj, xprev)), r2);
            T D = O::mul(O::add(r, r_prev), T(2));
            T rhs = O::add(pds, pds_prev);
            if (j >= 2) {
                const T w = O::div(r_prev, Dpp);
                D = O::sub(D, O::mul(w, r_prev));
                rhs = O::sub(rhs, O::mul(w, bpp));
            }
            sd[(j - 1) * SC] = D;
            sb[(j - 1) * SC] = rhs;
            Dpp = D; bpp = rhs;
            r_prev = r; pds_prev = pds;
            iprev = ij; xprev = xj;
        }
        // last row: D = 2 r_{m-2}, rhs = pds_{m-2}
        T D = O::mul(O::add(T(0), r_prev), T(2));
        T rhs = pds_prev;
        {
            const T w = O::div(r_prev, Dpp);
            D = O::sub(D, O::mul(w, r_prev));
            rhs = O::sub(rhs, O::mul(w, bpp));
        }
        // ---- pass 2: back substitution (misc.py:63-65); slot j of Bc becomes the knot derivative k_j
        T knext = O::div(rhs, D);
        k_last = knext;
        int inext = iprev;
        T xnext = park ? xval(m - 1, inext) : T(0);
#pragma unroll 4
        for (int j = m - 2; j >= 0; --j) {
            const int ij = jx[j * NT];
            T r, r2;
            recip(inext, ij, r, r2);
            const T k1 = knext;
            knext = O::div(O::sub(sb[j * SC], O::mul(r, knext)), sd[j * SC]);
            if (!park) sb[j * SC] = knext;
            if (park) {
                // coefficients of piece j (interpolate.py:46-53), the same operations as in the interval loop below
                const T v = xval(j, ij);
                const T six_pd = O::mul(T(2), O::mul(T(3), O::sub(xnext, v)));
                sPB[j * NT + tid] = knext;
                sPC[j * NT + tid] = O::mul(O::sub(O::sub(O::mul(six_pd, r), O::mul(T(4), knext)), O::mul(T(2), k1)), r);
                sPD[j * NT + tid] = O::mul(O::add(O::mul(-six_pd, r), O::mul(T(3), O::add(knext, k1))), r2);
                xnext = v;
            }
            inext = ij;
        }
    }
    const bool parked = park;
    // ---- pass 3: coefficients of each observed piece (interpolate.py:46-53), expanded to every
    // interval with offset = t_observed - t_i (interpolate.py:136-148).  All lanes walk i in lock step,
    // BACKWARDS: piece jj is entered at its last interval idx[jj+1]-1 >= jj, when slots jj+1.. are the only
    // ones overwritten so far; k_{jj+1} is carried in a register from the piece above, k_jj is read from slot jj.
    const bool two_knots = (m == 2);   // a single linear piece (interpolate.py:17-21)
    T pa = T(0), pb = T(0), pc = T(0), pd = T(0);
    int jj = m - 1, icur = L - 1;      // current piece ordinal (none yet) and its first knot
    T k_hi = k_last;                   // derivative at the upper knot of the piece about to be entered
    for (int i = L - 2; i >= 0; --i) {
        if (i < icur) {
            // enter piece jj-1: intervals [idx[jj-1], idx[jj])
            --jj;
            const int ni = icur;
            icur = jx[jj * NT];
            if (parked) {
                pa = xval(jj, icur);
                pb = sPB[jj * NT + tid];
                pc = sPC[jj * NT + tid];
                pd = sPD[jj * NT + tid];
            } else {
            const T v = xval(jj, icur), x1 = xval(jj + 1, ni);
            pa = v;
            const T k0 = sb[jj * SC], k1 = k_hi;
            k_hi = k0;
            if (two_knots) {
                T dt = tf32 ? (T)__fsub_rn((float)sT[ni], (float)sT[icur]) : O::sub(sT[ni], sT[icur]);
                pb = O::div(O::sub(x1, v), dt);
                pc = T(0);
                pd = T(0);
            } else {
                T r, r2;
                recip(ni, icur, r, r2);
                const T six_pd = O::mul(T(2), O::mul(T(3), O::sub(x1, v)));
                pb = k0;
                // two_c = (six_pd * r - 4 k0 - 2 k1) * r
                pc = O::mul(O::sub(O::sub(O::mul(six_pd, r), O::mul(T(4), k0)), O::mul(T(2), k1)), r);
                // three_d = (-six_pd * r + 3 (k0 + k1)) * r^2
                pd = O::mul(O::add(O::mul(-six_pd, r), O::mul(T(3), O::add(k0, k1))), r2);
            }
            }
        }
        const T off = tf32 ? (T)__fsub_rn((float)sT[icur], (float)sT[i]) : O::sub(sT[icur], sT[i]);
        // a_inner = (0.5 two_c - three_d * off / 3) * off
        const T a_inner = O::mul(O::sub(O::mul(T(0.5), pc), O::div(O::mul(pd, off), T(3))), off);
        const int64_t o = obase + (int64_t)i * C;
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
    ANCDE_CHECK_ARG(L <= 65535, "spline_coeffs: L=%d too long (max 65535 knots)", L);
    const int64_t P = N * (int64_t)C;
    const int Lp = (L + 3) & ~3;
    // the only shared scratch is the 16-bit ordinal -> knot index list; the Thomas scratch lives in the outputs
    const size_t isz = (L <= 256) ? 1 : 2;        // bytes per ordinal -> knot index
    int nt = 128;
    size_t smem = 0;
    for (;; nt /= 2) {
        smem = (size_t)Lp * sizeof(T) + 3 * (size_t)SPL_MAXP * nt * sizeof(T) + (size_t)(L + 2) * nt * isz;
        if (smem <= 56 * 1024 || nt == 32) break;
    }
    if (smem > 227 * 1024) {
        set_error("spline_coeffs: L=%d needs %zu bytes of shared memory per CTA (max 227 KB)", L, smem);
        return ANCDE_EUNSUPPORTED;
    }
    auto kern = (isz == 1) ? spline_coeffs_kernel<T, unsigned char> : spline_coeffs_kernel<T, unsigned short>;
    ANCDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (P + nt - 1) / nt;
    ANCDE_CHECK_ARG(grid <= 0x7fffffffLL, "spline_coeffs: too many paths");
    prof_begin(sizeof(T) == 4 ? "spline_coeffs_f32" : "spline_coeffs_f64", stream);
    kern<<<(unsigned)grid, nt, smem, stream>>>(times, X, a, b, c2, d3, P, L, C, tf32);
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
