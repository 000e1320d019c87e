// kernels_apply.cuh -- the subtract column pass (Construct_FDIFF + the axis-0 half of ifft2), done WITHOUT a
// column FFT.
//
// Reference: Kab_Wla / Kab_Wmb twiddle planes + Construct_FDIFF + ifft2 (sfft/sfftcore/SFFTSubtract.py:433-461,
// kernel sfft/sfftcore/SFFTConfigure.py:737-809): Fab x Fij complex MACs per pixel in the 2-D Fourier domain.
// The matching kernel is only 2 w0 + 1 rows tall, so after the row transforms the axis-0 part of the convolution
// is a (2 w0 + 1)-tap circular FIR along the (contiguous, transposed) columns of the row spectra, with taps that
// depend on the column only through h_A[a; k1] = sum_b a_Aab e^{-2 pi i b k1 / N1}:
//     d[r; k1] = g_J[r; k1] - (1/N) ( sum_A sum_a h_A[a; k1] g_A[(r - a) % N0; k1]  -  sum_A c_A g_A[r; k1] ),
//     g_A[r] = cx(r)^i g_j[r],  c_A = sum_ab a_Aab - a_A00   (the "-1" of the modified delta basis).
// d is the row spectrum of the difference image; the background term sum_pq b_pq T_pq is subtracted in real space
// by row_inv_kernel.  Per output: (2 w0 + 1) * sum_j (2 (DK - j) + 4) fp64 FMAs, no shared-memory exchange.
#pragma once
#include "common.cuh"

#define FIR_NT 256
#define FIR_CHUNK 2048        // rows of one column handled by one CTA

struct FirArgs {
    int N0, N1, NH;
    int DK, Fij, nj, w0, w1;
    unsigned char plane_of[4][4];
    const cd* tw1;
};

// smem: h[Fij][L0] (cd) | cA[Fij] (double)
template <typename TSt>
__global__ void __launch_bounds__(FIR_NT) apply_fir_kernel(FirArgs a, const TSt* __restrict__ gI, const TSt* gJ,
                                                           const double* __restrict__ sol, TSt* outD)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L0 = 2 * a.w0 + 1, L1 = 2 * a.w1 + 1, Fab = L0 * L1;
    cd* h = reinterpret_cast<cd*>(smem_raw);
    double* cA = reinterpret_cast<double*>(h + a.Fij * L0);
    const int tid = threadIdx.x;
    const int k1 = blockIdx.x;
    const int rbeg = blockIdx.y * FIR_CHUNK;
    const int rend = min(a.N0, rbeg + FIR_CHUNK);
    const double invN = 1.0 / ((double)a.N0 * (double)a.N1);
    const double inv0 = 1.0 / (double)a.N0;

    for (int idx = tid; idx < a.Fij * L0; idx += FIR_NT) {
        const int A = idx / L0, ia = idx - A * L0;
        const double* s = sol + (size_t)A * Fab + (size_t)ia * L1;
        cd acc = cmake(0, 0);
        for (int ib = 0; ib < L1; ++ib) {
            const int b = ib - a.w1;
            const cd w = a.tw1[imod((int)(((long long)b * k1) % a.N1), a.N1)];       // e^{-2 pi i b k1 / N1}
            acc.x = fma(s[ib], w.x, acc.x);
            acc.y = fma(s[ib], w.y, acc.y);
        }
        h[idx] = cscale(acc, invN);
    }
    for (int A = tid; A < a.Fij; A += FIR_NT) {
        const double* s = sol + (size_t)A * Fab;
        double t = 0.0;
        for (int ab = 0; ab < Fab; ++ab) t += s[ab];
        cA[A] = (t - s[a.w0 * L1 + a.w1]) * invN;
    }
    __syncthreads();

    const TSt* colJ = gJ + (size_t)k1 * a.N0;
    for (int r = rbeg + tid; r < rend; r += FIR_NT) {
        cd acc = load_c(colJ + r);
        for (int j = 0; j < a.nj; ++j) {
            const TSt* col = gI + ((size_t)j * a.NH + k1) * a.N0;
            const int ni = a.DK - j + 1;
            const int A0 = a.plane_of[0][j];
            const int A1 = ni > 1 ? a.plane_of[1][j] : 0, A2 = ni > 2 ? a.plane_of[2][j] : 0, A3 = ni > 3 ? a.plane_of[3][j] : 0;
            // + c_A g_A[r]
            {
                const double cx = (r + 1) * inv0;
                double t = 0.0;
                if (ni > 3) t = cA[A3];
                if (ni > 2) t = fma(t, cx, cA[A2]);
                if (ni > 1) t = fma(t, cx, cA[A1]);
                t = fma(t, cx, cA[A0]);
                const cd g = load_c(col + r);
                acc.x = fma(t, g.x, acc.x);
                acc.y = fma(t, g.y, acc.y);
            }
            int rs = r + a.w0;                       // source row r - sh for sh = -w0
            if (rs >= a.N0) rs -= a.N0;
            for (int ia = 0; ia < L0; ++ia) {
                const double cx = (rs + 1) * inv0;
                cd t = cmake(0, 0);
                if (ni > 3) t = h[A3 * L0 + ia];
                if (ni > 2) { const cd hh = h[A2 * L0 + ia]; t = cmake(fma(t.x, cx, hh.x), fma(t.y, cx, hh.y)); }
                if (ni > 1) { const cd hh = h[A1 * L0 + ia]; t = cmake(fma(t.x, cx, hh.x), fma(t.y, cx, hh.y)); }
                { const cd hh = h[A0 * L0 + ia]; t = cmake(fma(t.x, cx, hh.x), fma(t.y, cx, hh.y)); }
                const cd g = load_c(col + rs);
                // acc -= t * g
                acc.x = fma(-t.x, g.x, acc.x); acc.x = fma(t.y, g.y, acc.x);
                acc.y = fma(-t.x, g.y, acc.y); acc.y = fma(-t.y, g.x, acc.y);
                rs = (rs == 0) ? a.N0 - 1 : rs - 1;
            }
        }
        store_c(outD + (size_t)k1 * a.N0 + r, acc);
    }
}

// index of plane (i, j) in REF_ij order (i-major, j = 0 .. DK - i; sfft/sfftcore/SFFTSubtract.py:61-70)
__host__ __device__ constexpr int fir_plane(int DK, int i, int j) { return i * (DK + 1) - (i * (i - 1)) / 2 + j; }

// ---- version 3: register sliding window -----------------------------------------------------------------------------
// h_A[a; k1] and c_A are tabulated once per apply by fir_taps_kernel (they depend on the solution and k1 only).
// A thread owns FIR3_R CONSECUTIVE output rows; walking the taps a = -w0 .. w0 moves the source window down by one
// row per tap, so only one new source row (DK + 1 shared loads) enters the register window per tap for FIR3_R
// outputs.  The staged chunk is stored de-interleaved (row n at (n % R) * WP + n / R) so that the per-thread
// consecutive rows are unit stride across a warp.
#define FIR3_NT 256
#define FIR3_R 4
#define FIR3_CH (FIR3_NT * FIR3_R)

// taps[k1][A][ia] = (1/N) sum_b a_Aab e^{-2 pi i b k1 / N1};  cAout[A] = (sum_ab a_Aab - a_A00) / N
#ifdef SFFTB_TU_APPLY
__global__ void __launch_bounds__(128) fir_taps_kernel(FirArgs a, const double* __restrict__ sol, cd* __restrict__ taps,
                                                       double* __restrict__ cAout)
{
    const int L0 = 2 * a.w0 + 1, L1 = 2 * a.w1 + 1, Fab = L0 * L1;
    const int k1 = blockIdx.x;
    const double invN = 1.0 / ((double)a.N0 * (double)a.N1);
    for (int idx = threadIdx.x; idx < a.Fij * L0; idx += blockDim.x) {
        const int A = idx / L0, ia = idx - A * L0;
        const double* s = sol + (size_t)A * Fab + (size_t)ia * L1;
        cd acc = cmake(0, 0);
        for (int ib = 0; ib < L1; ++ib) {
            const int b = ib - a.w1;
            const cd w = a.tw1[imod((int)(((long long)b * k1) % a.N1), a.N1)];
            acc.x = fma(s[ib], w.x, acc.x);
            acc.y = fma(s[ib], w.y, acc.y);
        }
        taps[(size_t)k1 * a.Fij * L0 + idx] = cscale(acc, invN);
    }
    if (k1 == 0)
        for (int A = threadIdx.x; A < a.Fij; A += blockDim.x) {
            const double* s = sol + (size_t)A * Fab;
            double t = 0.0;
            for (int ab = 0; ab < Fab; ++ab) t += s[ab];
            cAout[A] = (t - s[a.w0 * L1 + a.w1]) * invN;
        }
}
#endif  // SFFTB_TU_APPLY

// smem: h[Fij][L0] cd | cA[16] double | st[nj][R * WP] cd | cxs[R * WP] double,  WP = ceil((CH + 2 w0) / R)
template <typename TSt, int DK>
__global__ void __launch_bounds__(FIR3_NT, 2) apply_fir3_kernel(FirArgs a, const TSt* __restrict__ gI, const TSt* gJ,
                                                                const cd* __restrict__ taps, const double* __restrict__ cAin, TSt* outD)
{
    constexpr int NJ = DK + 1, Fij = (DK + 1) * (DK + 2) / 2, R = FIR3_R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L0 = 2 * a.w0 + 1;
    const int W = FIR3_CH + 2 * a.w0;
    const int WP = (W + R - 1) / R;
    cd* h = reinterpret_cast<cd*>(smem_raw);
    double* cA = reinterpret_cast<double*>(h + Fij * L0);
    cd* st = reinterpret_cast<cd*>(cA + 16);
    double* cxs = reinterpret_cast<double*>(st + (size_t)NJ * R * WP);
    const int tid = threadIdx.x;
    const int k1 = blockIdx.x;
    const int rbeg = blockIdx.y * FIR3_CH;
    const double inv0 = 1.0 / (double)a.N0;

    for (int idx = tid; idx < W; idx += FIR3_NT) {
        int r = rbeg - a.w0 + idx;
        r %= a.N0; if (r < 0) r += a.N0;
        const int pos = (idx % R) * WP + idx / R;
        cxs[pos] = (r + 1) * inv0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) st[(size_t)j * R * WP + pos] = load_c(gI + ((size_t)j * a.NH + k1) * a.N0 + r);
    }
    for (int idx = tid; idx < Fij * L0; idx += FIR3_NT) h[idx] = taps[(size_t)k1 * Fij * L0 + idx];
    if (tid < Fij) cA[tid] = cAin[tid];
    __syncthreads();

    // staged index of (output o, tap ia): n = R tid + 2 w0 + (o - ia);  window slot of u = o - ia is (u mod R)
    const int nb = R * tid + 2 * a.w0;
    cd acc[R];
    cd win[NJ][R];
    double cxw[R];
#pragma unroll
    for (int o = 0; o < R; ++o) {
        const int n = nb + o;                              // u = o (tap ia = 0)
        const int pos = (n % R) * WP + n / R;
        cxw[o] = cxs[pos];
#pragma unroll
        for (int j = 0; j < NJ; ++j) win[j][o] = st[(size_t)j * R * WP + pos];
    }
    // start value: J[r] + sum_A c_A G_A[r]   (output row r_o = staged index R tid + o + w0)
#pragma unroll
    for (int o = 0; o < R; ++o) {
        const int r = rbeg + R * tid + o;
        cd v = cmake(0.0, 0.0);
        if (r < a.N0) v = load_c(gJ + (size_t)k1 * a.N0 + r);
        const int n = R * tid + o + a.w0;
        const int pos = (n % R) * WP + n / R;
        const double cx = cxs[pos];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            double t = 0.0;
#pragma unroll
            for (int i = DK - j; i >= 0; --i) t = fma(t, cx, cA[fir_plane(DK, i, j)]);
            const cd g = st[(size_t)j * R * WP + pos];
            v.x = fma(t, g.x, v.x);
            v.y = fma(t, g.y, v.y);
        }
        acc[o] = v;
    }
    for (int a0 = 0; a0 < L0; a0 += R) {
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int ia = a0 + q;
            if (ia < L0) {
                if (ia > 0) {
                    // new window element u = -ia enters slot (-q mod R)
                    const int n = nb - ia;
                    const int pos = (n % R) * WP + n / R;
                    cxw[(R - q) % R] = cxs[pos];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) win[j][(R - q) % R] = st[(size_t)j * R * WP + pos];
                }
                cd hh[Fij];
#pragma unroll
                for (int A = 0; A < Fij; ++A) hh[A] = h[A * L0 + ia];
#pragma unroll
                for (int o = 0; o < R; ++o) {
                    const int sl = (o - q + R) % R;
                    const double cx = cxw[sl];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        cd t = cmake(0.0, 0.0);
#pragma unroll
                        for (int i = DK - j; i >= 0; --i) {
                            const cd c = hh[fir_plane(DK, i, j)];
                            t = cmake(fma(t.x, cx, c.x), fma(t.y, cx, c.y));
                        }
                        const cd g = win[j][sl];
                        acc[o].x = fma(-t.x, g.x, acc[o].x); acc[o].x = fma(t.y, g.y, acc[o].x);
                        acc[o].y = fma(-t.x, g.y, acc[o].y); acc[o].y = fma(-t.y, g.x, acc[o].y);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < R; ++o) {
        const int r = rbeg + R * tid + o;
        if (r < a.N0) store_c(outD + (size_t)k1 * a.N0 + r, acc[o]);
    }
}
