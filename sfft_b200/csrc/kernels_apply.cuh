// kernels_apply.cuh -- the subtract column pass (Construct_FDIFF + the axis-0 half of ifft2).
//
// Reference: Kab_Wla / Kab_Wmb twiddle planes + Construct_FDIFF + ifft2 (sfft/sfftcore/SFFTSubtract.py:433-461,
// kernel sfft/sfftcore/SFFTConfigure.py:737-809).  The reference evaluates sum_ab a_ijab (W^a W^b - 1) with
// Fab x Fij complex MACs per pixel from 2L full twiddle planes.  Here one CTA owns one column k1: the kernel
// spectrum is separable per column, h_A[a] = sum_b a_Aab W1^{b k1}, and its axis-0 transform is an FFT of a
// (2 w0 + 1)-sparse vector, done per DIF slice next to the slice spectra of the images.
//     FDIFF[k0,k1] = FJ - (1/N) sum_A F_A[k0,k1] (K_A[k0,k1] - c_A),   c_A = sum_ab a_Aab - a_A00
// The background term sum_pq b_pq T_pq is subtracted in real space by row_inv_kernel.
#pragma once
#include "kernels_fit.cuh"

// smem (cd): S[(2 Fij + 1) * pitch] | E[N0] | h[Fij * L0] | cA[Fij]
// planes of S: 0..Fij-1 image slices, Fij = J slice (becomes FDIFF), Fij+1 .. 2Fij = kernel-spectrum slices
template <typename TSt>
__global__ void __launch_bounds__(NT_COL) apply_col_kernel(ColArgs a, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
                                                           const double* __restrict__ sol, const cd* __restrict__ tw1,
                                                           TSt* __restrict__ outD)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* S = reinterpret_cast<cd*>(smem_raw);
    cd* E = S + (size_t)(2 * a.Fij + 1) * a.pitch;
    cd* h = E + a.N0;
    cd* cA = h + a.Fij * (2 * a.w0 + 1);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int L0 = 2 * a.w0 + 1, L1 = 2 * a.w1 + 1, Fab = L0 * L1;
    const double invN = 1.0 / ((double)a.N0 * (double)a.N1);
    cd* KS = S + (size_t)(a.Fij + 1) * a.pitch;
    cd* SJ = S + (size_t)a.Fij * a.pitch;

    for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x) {
        // per-column separable kernel factors
        for (int idx = tid; idx < a.Fij * L0; idx += nthr) {
            const int A = idx / L0, ia = idx - A * L0;
            const double* s = sol + (size_t)A * Fab + (size_t)ia * L1;
            cd acc = cmake(0, 0);
            for (int ib = 0; ib < L1; ++ib) {
                const int b = ib - a.w1;
                const cd w = tw1[imod((int)(((long long)b * k1) % a.N1), a.N1)];    // e^{-2 pi i b k1 / N1}
                acc.x = fma(s[ib], w.x, acc.x);
                acc.y = fma(s[ib], w.y, acc.y);
            }
            h[idx] = acc;
        }
        for (int A = tid; A < a.Fij; A += nthr) {
            const double* s = sol + (size_t)A * Fab;
            double t = 0.0;
            for (int ab = 0; ab < Fab; ++ab) t += s[ab];
            cA[A] = cmake(t - s[a.w0 * L1 + a.w1], 0.0);
        }
        __syncthreads();

        for (int t = 0; t < a.V; ++t) {
            fold_slice(a, gI, gJ, k1, t, S, true);
            // sparse kernel-spectrum inputs: KS_A[a mod M] += h_A[a] W_N0^{a t}
            for (int idx = tid; idx < a.Fij * a.M; idx += nthr) {
                const int A = idx / a.M, n = idx - A * a.M;
                KS[(size_t)A * a.pitch + n] = cmake(0, 0);
            }
            __syncthreads();
            if (tid < a.Fij) {
                const int A = tid;
                for (int ia = 0; ia < L0; ++ia) {
                    const int sh = ia - a.w0;
                    const cd w = a.tw0[imod(sh * t, a.N0)];
                    cd* dst = KS + (size_t)A * a.pitch + imod(sh, a.M);
                    *dst = cadd(*dst, cmul(h[A * L0 + ia], w));
                }
            }
            __syncthreads();
            fft_planes(S, a.pitch, 2 * a.Fij + 1, a.fd, a.twM, -1.0);
            for (int u = tid; u < a.M; u += nthr) {
                cd acc = cmake(0, 0);
                for (int A = 0; A < a.Fij; ++A) {
                    const cd k = csub(KS[(size_t)A * a.pitch + u], cA[A]);
                    cfma(acc, S[(size_t)A * a.pitch + u], k);
                }
                const cd fj = SJ[u];
                SJ[u] = cmake(fj.x - invN * acc.x, fj.y - invN * acc.y);
            }
            __syncthreads();
            fft_planes(SJ, a.pitch, 1, a.fd, a.twM, +1.0);
            for (int n = tid; n < a.M; n += nthr) E[t * a.M + n] = SJ[n];
            __syncthreads();
        }
        // DIT unfold: d[r] = sum_t e^{+2 pi i t r / N0} e_t[r mod M]
        for (int r = tid; r < a.N0; r += nthr) {
            const int n = r % a.M;
            cd acc = E[n];
            for (int t = 1; t < a.V; ++t) {
                const cd w = a.tw0[(int)(((long long)t * r) % a.N0)];
                cfma(acc, E[t * a.M + n], cconj(w));
            }
            store_c(outD + (size_t)k1 * a.N0 + r, acc);
        }
        __syncthreads();
    }
}
