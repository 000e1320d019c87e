// standalone timing of cc_potrf_inv (one CTA), cycles per call
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../sfft_b200/csrc/kernels_chol.cuh"
__device__ void pv0(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC_NB; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = cc_fast_rcp(piv);
        if (own > c) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncthreads();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}


__device__ void pv1(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC_NB; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = cc_fast_rcp(piv);
        if (own > c) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncwarp();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}


__device__ void pv2(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC_NB; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = piv * 1e-3;
        if (own > c) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncthreads();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}


__device__ void pv3(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC_NB; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = cc_fast_rcp(piv);
        if (own > c) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1 && part == 7) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1 && part == 7) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncthreads();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}


__device__ void pv4(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC_NB; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = cc_fast_rcp(piv);
        if (own > c && part == 7) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncthreads();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}


__device__ void pv5(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC_NB; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = cc_fast_rcp(piv);
        if (own > c) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c && part == 7) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncthreads();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}


__device__ void pv6(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC_NB; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = cc_fast_rcp(piv);
        if (own > c && part == 7) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c && part == 7) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncthreads();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}


__device__ void pv7(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < 0; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = cc_fast_rcp(piv);
        if (own > c) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncthreads();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}


template <int V> __global__ void __launch_bounds__(CC_NT, 1) kbench(CholArgs a, long long* cyc, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + CC_NB * CC_PITCH;
    for (int rep = 0; rep < reps; ++rep) {
        // restore the matrix
        for (int i = threadIdx.x; i < 64 * 64; i += CC_NT) a.A[(i >> 6) * a.ld + (i & 63)] = a.yv[i];
        __syncthreads();
        long long t0 = clock64();
        if (V == 0) pv0(a, 0, As, Bs); if (V == 1) pv1(a, 0, As, Bs); if (V == 2) pv2(a, 0, As, Bs); if (V == 3) pv3(a, 0, As, Bs); if (V == 4) pv4(a, 0, As, Bs); if (V == 5) pv5(a, 0, As, Bs); if (V == 6) pv6(a, 0, As, Bs); if (V == 7) pv7(a, 0, As, Bs);
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[rep] = t1 - t0;
    }
}
int main() {
    const int n = 64, ld = 65;
    std::vector<double> M(64 * 64), A(65 * 65, 0.0);
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k < 64; ++k) s += sin(0.1 * (i + 1) * (k + 1)) * sin(0.1 * (j + 1) * (k + 1)); M[i * 64 + j] = s + (i == j ? 10.0 : 0.0); }
    CholArgs a; memset(&a, 0, sizeof a);
    double *dA, *dW, *dM; int* info; long long* cyc;
    cudaMalloc(&dA, sizeof(double) * 65 * 65); cudaMalloc(&dW, sizeof(double) * 4096); cudaMalloc(&dM, sizeof(double) * 4096);
    cudaMalloc(&info, 16); cudaMemset(info, 0, 16); cudaMalloc(&cyc, 8 * 16);
    cudaMemcpy(dM, M.data(), sizeof(double) * 4096, cudaMemcpyHostToDevice);
    a.A = dA; a.ld = ld; a.n = n; a.ntot = n + 1; a.W = dW; a.yv = dM; a.info = info;
    size_t sm = sizeof(double) * 2 * CC_NB * CC_PITCH;
    long long h[8];
#define RUNV(V) cudaFuncSetAttribute(kbench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); kbench<V><<<1, CC_NT, sm>>>(a, cyc, 4); cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost); printf("variant %d: %s cycles %lld %lld\n", V, cudaGetErrorString(cudaDeviceSynchronize()), h[2], h[3]);
    RUNV(0) RUNV(1) RUNV(2) RUNV(3) RUNV(4) RUNV(5) RUNV(6) RUNV(7)
    // check L L^T = M
    std::vector<double> L(65 * 65), W(4096); cudaMemcpy(L.data(), dA, sizeof(double) * 65 * 65, cudaMemcpyDeviceToHost); cudaMemcpy(W.data(), dW, sizeof(double) * 4096, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < 64; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += L[i * 65 + k] * L[j * 65 + k]; e1 = fmax(e1, fabs(s - M[i * 64 + j])); }
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k < 64; ++k) s += (k <= i ? W[i * 64 + k] : 0.0) * (j <= k ? L[k * 65 + j] : 0.0); e2 = fmax(e2, fabs(s - (i == j))); }
    printf("max |LL^T - M| = %.3e, max |W L - I| = %.3e\n", e1, e2);
}
