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
