// kernels_row.cuh -- row passes.
//
// row_fwd_kernel : real rows -> half spectra along axis 1, with the spatial polynomial factor cy(c)^j fused
//                  into the load (reference: SpatialPoly + fft2, sfft/sfftcore/SFFTConfigure.py:112-145,
//                  sfft/sfftcore/SFFTSubtract.py:127-161).  Output is stored TRANSPOSED, g[j][k1][r], so that
//                  the column passes read contiguous columns.
// row_inv_kernel : transposed half spectra -> real rows (C2R), scaling, background polynomial subtraction
//                  (reference: ifft2 + the b_pq T_pq term of Construct_FDIFF, SFFTSubtract.py:452-461).
#pragma once
#include "fft_smem.cuh"

struct RowArgs {
    int N0, N1, NH;
    int H;            // FFT length: N1/2 when packed (even N1), N1 otherwise
    int packed;
    int RB;           // rows per CTA
    int pitch;        // smem plane pitch in elements
    FftDesc fd;       // plan for length H
    const cd* twH;    // exp(-2 pi i e / H)
    const cd* tw1;    // exp(-2 pi i e / N1)
    // Bluestein (chirp-z) path for lengths H with a prime factor > 13: blu_M = power of two >= 2 H - 1 (0 = unused)
    int blu_M;
    FftDesc blu_fd;       // plan for length blu_M
    const cd* blu_tw;     // exp(-2 pi i e / blu_M)
    const cd* blu_c;      // chirp c[n] = exp(-pi i n^2 / H), n < H
    const cd* blu_B;      // FFT_M of conj(c) wrapped to length M (the chirp filter), already divided by M
    // general-basis plans: plane j is the row times the table vtab[j][c] (B-spline or any other 1-D function of the
    // column, kernels_gen.cuh) instead of cy(c)^j; NULL = powers
    const double* vtab;
};

// Length-H transform of `nplanes` planes through Bluestein's identity n k = (n^2 + k^2 - (k - n)^2) / 2:
//   X[k] = c[k] * sum_n (x[n] c[n]) conj(c)[k - n]      (forward);  inverse = conj(forward(conj(x))).
// Planes must have room for blu_M elements.  Same barrier contract as fft_planes.
__device__ __forceinline__ void row_transform(cd* buf, int nplanes, const RowArgs& a, double sgn) {
    if (a.blu_M == 0) { fft_planes(buf, a.pitch, nplanes, a.fd, a.twH, sgn); return; }
    const int tid = threadIdx.x, nthr = blockDim.x, M = a.blu_M;
    const bool inv = sgn > 0.0;
    for (int idx = tid; idx < nplanes * M; idx += nthr) {
        const int pl = idx / M, n = idx - pl * M;
        cd* q = buf + (size_t)pl * a.pitch + n;
        cd v = cmake(0.0, 0.0);
        if (n < a.H) { v = *q; if (inv) v.y = -v.y; v = cmul(v, a.blu_c[n]); }
        *q = v;
    }
    __syncthreads();
    fft_planes(buf, a.pitch, nplanes, a.blu_fd, a.blu_tw, -1.0);
    for (int idx = tid; idx < nplanes * M; idx += nthr) {
        const int pl = idx / M, n = idx - pl * M;
        cd* q = buf + (size_t)pl * a.pitch + n;
        *q = cmul(*q, a.blu_B[n]);
    }
    __syncthreads();
    fft_planes(buf, a.pitch, nplanes, a.blu_fd, a.blu_tw, +1.0);
    for (int idx = tid; idx < nplanes * a.H; idx += nthr) {
        const int pl = idx / a.H, n = idx - pl * a.H;
        cd* q = buf + (size_t)pl * a.pitch + n;
        cd v = cmul(*q, a.blu_c[n]);
        if (inv) v.y = -v.y;
        *q = v;
    }
    __syncthreads();
}

template <typename TIn, typename TSt>
__global__ void __launch_bounds__(512) row_fwd_kernel(RowArgs a, const TIn* __restrict__ img, TSt* __restrict__ out, int nj)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* buf = reinterpret_cast<cd*>(smem_raw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int r0 = blockIdx.x * a.RB;
    const double inv1 = 1.0 / (double)a.N1;

    for (int j = 0; j < nj; ++j) {
        if (a.packed) {
            for (int idx = tid; idx < a.RB * a.H; idx += nthr) {
                const int row = idx / a.H, n = idx - row * a.H;
                const int r = r0 + row;
                cd z = cmake(0.0, 0.0);
                if (r < a.N0) {
                    const TIn* p = img + (size_t)r * a.N1 + 2 * n;
                    double x0 = (double)p[0], x1 = (double)p[1];
                    if (a.vtab) {
                        const double* vt = a.vtab + (size_t)j * a.N1 + 2 * n;
                        x0 *= vt[0]; x1 *= vt[1];
                    } else if (j > 0) {
                        x0 *= ipow((2 * n + 1) * inv1, j);
                        x1 *= ipow((2 * n + 2) * inv1, j);
                    }
                    z = cmake(x0, x1);
                }
                buf[(size_t)row * a.pitch + n] = z;
            }
        } else {
            for (int idx = tid; idx < a.RB * a.H; idx += nthr) {
                const int row = idx / a.H, n = idx - row * a.H;
                const int r = r0 + row;
                double x0 = 0.0;
                if (r < a.N0) {
                    x0 = (double)img[(size_t)r * a.N1 + n];
                    if (a.vtab) x0 *= a.vtab[(size_t)j * a.N1 + n];
                    else if (j > 0) x0 *= ipow((n + 1) * inv1, j);
                }
                buf[(size_t)row * a.pitch + n] = cmake(x0, 0.0);
            }
        }
        __syncthreads();
        row_transform(buf, a.RB, a, -1.0);
        // untangle + transposed store: consecutive threads -> consecutive rows of the same k1
        for (int idx = tid; idx < a.RB * a.NH; idx += nthr) {
            const int k = idx / a.RB, row = idx - k * a.RB;
            const int r = r0 + row;
            if (r >= a.N0) continue;
            const cd* pl = buf + (size_t)row * a.pitch;
            cd g;
            if (a.packed) {
                const cd zk = pl[k == a.H ? 0 : k];
                const cd zm = cconj(pl[k == 0 ? 0 : a.H - k]);
                const cd s = cadd(zk, zm), d = csub(zk, zm);
                const cd w = a.tw1[k];
                // g = 0.5 s - 0.5 i w d
                const cd wd = cmul(w, d);
                g = cmake(0.5 * (s.x + wd.y), 0.5 * (s.y - wd.x));
            } else {
                g = pl[k];
            }
            store_c(out + ((size_t)j * a.NH + k) * a.N0 + r, g);
        }
        __syncthreads();
    }
}

struct RowInvArgs {
    RowArgs r;
    double scale;          // applied to the real output
    int DB, Fpq;
    unsigned char p_of[16], q_of[16];
};

// bpq: Fpq background coefficients (device), may be NULL (no background term)
template <typename TSt, typename TOut>
__global__ void __launch_bounds__(512) row_inv_kernel(RowInvArgs ia, const TSt* __restrict__ spec, const double* __restrict__ bpq,
                                                      TOut* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* buf = reinterpret_cast<cd*>(smem_raw);
    const RowArgs& a = ia.r;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int r0 = blockIdx.x * a.RB;

    if (a.packed) {
        // Z[k] = Ze[k] + i Zo[k],  Ze = (g[k] + conj g[H-k]) / 2,  Zo = (g[k] - conj g[H-k]) conj(W^k) / 2
        for (int idx = tid; idx < a.RB * a.H; idx += nthr) {
            const int k = idx / a.RB, row = idx - k * a.RB;
            const int r = r0 + row;
            cd z = cmake(0.0, 0.0);
            if (r < a.N0) {
                const cd gk = load_c(spec + (size_t)k * a.N0 + r);
                const cd gm = cconj(load_c(spec + (size_t)(a.H - k) * a.N0 + r));
                const cd ze = cscale(cadd(gk, gm), 0.5);
                const cd zo = cscale(cmul(csub(gk, gm), cconj(a.tw1[k])), 0.5);
                z = cmake(ze.x - zo.y, ze.y + zo.x);
            }
            buf[(size_t)row * a.pitch + k] = z;
        }
    } else {
        for (int idx = tid; idx < a.RB * a.H; idx += nthr) {
            const int k = idx / a.RB, row = idx - k * a.RB;
            const int r = r0 + row;
            cd z = cmake(0.0, 0.0);
            if (r < a.N0) {
                z = (k < a.NH) ? load_c(spec + (size_t)k * a.N0 + r)
                               : cconj(load_c(spec + (size_t)(a.N1 - k) * a.N0 + r));
            }
            buf[(size_t)row * a.pitch + k] = z;
        }
    }
    __syncthreads();
    row_transform(buf, a.RB, a, +1.0);

    double b[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) b[k] = (bpq != nullptr && k < ia.Fpq) ? bpq[k] : 0.0;
    const double inv0 = 1.0 / (double)a.N0, inv1 = 1.0 / (double)a.N1;
    for (int idx = tid; idx < a.RB * a.H; idx += nthr) {
        const int row = idx / a.H, n = idx - row * a.H;
        const int r = r0 + row;
        if (r >= a.N0) continue;
        const cd z = buf[(size_t)row * a.pitch + n];
        const double cx = (r + 1) * inv0;
        if (a.packed) {
            double x0 = z.x * ia.scale, x1 = z.y * ia.scale;
            if (bpq != nullptr) {
                const double cy0 = (2 * n + 1) * inv1, cy1 = (2 * n + 2) * inv1;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (k < ia.Fpq) {
                        const double px = ipow(cx, ia.p_of[k]);
                        x0 -= b[k] * px * ipow(cy0, ia.q_of[k]);
                        x1 -= b[k] * px * ipow(cy1, ia.q_of[k]);
                    }
                }
            }
            TOut* p = out + (size_t)r * a.N1 + 2 * n;
            p[0] = (TOut)x0;
            p[1] = (TOut)x1;
        } else {
            double x0 = z.x * ia.scale;
            if (bpq != nullptr) {
                const double cy0 = (n + 1) * inv1;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k < ia.Fpq) x0 -= b[k] * ipow(cx, ia.p_of[k]) * ipow(cy0, ia.q_of[k]);
            }
            out[(size_t)r * a.N1 + n] = (TOut)x0;
        }
    }
}
