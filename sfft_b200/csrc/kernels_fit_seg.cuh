// kernels_fit_seg.cuh -- the fit column pass by segmented (overlap-save) cross-correlation, KerPolyOrder <= 2.
//
// Same mathematics as fit_col_kernel (kernels_fit.cuh): per column k1 of the transposed row spectra,
//     kappa_AB[m0; k1] = sum_r conj(G_A[r]) G_B[(r + m0) % N0],   G_A[r] = cx(r)^i g_j[r; k1],
// for the lags FillLS_* reads (|m0| <= 2 w0 for I x I pairs, |m0| <= w0 for I x J;
// sfft/sfftcore/SFFTConfigure.py:198-275, 590-634).  The reference gets them from full-size FFTs of every
// cross-spectrum plane (sfft/sfftcore/SFFTSubtract.py:224-383).  Here the column is cut into segments of S rows;
// for every segment the plane restricted to the segment ("A role", zero padded to M = 256) and the plane on the
// segment extended by the halo h = 2 w0 on both sides ("B role", S + 2h <= M) are transformed, the cross spectra
// conj(FA) FB are ACCUMULATED OVER SEGMENTS in registers, and a single inverse transform per pair at the end of the
// column yields the lags.  No aliasing: an A-role sample at window position [h, h+S) shifted by |m| <= h stays in
// [0, M).  Per column: nseg (2 Fij + 1) forward FFTs + (npairs + Fij) inverse FFTs of length 256, all on the
// register FFT engine (16 threads x 16 values, one shared-memory exchange).
//
// Output layout: kap[k1][row], rows = Omega pairs x (4 w0 + 1) | Theta planes x (2 w0 + 1) |
//                conj(lam_(A,p,ia)) Q_q rows (cross terms with the background basis, see column_poly_rows) | nuJ rows.
// lag_reduce2_kernel contracts k1 with the axis-1 twiddles into the lag tables R, RJ, RT, RJT.
#pragma once
#include "kernels_fit.cuh"


struct SegFitArgs {
    ColArgs c;
    int S, nseg, h;                  // core rows per segment, number of segments, halo = 2 w0
    int nrows;                       // rows per column of kap
    int nOm, nK, nLT;                // Omega rows | + Theta rows | lam*Q rows (then Fpq nuJ*Q rows)
    const cd* tabA;                  // engine table of the second radix-16 pass (240 entries)
    const cd* Q;                     // Q[q][k1] = DFT_c(cy(c)^q), q = 0..DB
    signed char pq_of[4][4];
    const cd* momg;                  // column moments [NH][planes * SFFTB_MAXE] (col_moments_kernel -> fit_seg4_kernel)
    int mom_external;                // 1: the background cross-term rows are written by col_poly_rows_kernel, not inside fit_seg4_kernel
};

// asynchronous element copy global -> shared (LDGSTS); 8 bytes for fp32 spectra, 16 for fp64
__device__ __forceinline__ void cp_async_elem(float2* dst, const float2* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(double2* dst, const double2* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ int wrap_row(int r, int N0) {
    if (r < 0) { r += N0; if (r < 0) { r %= N0; if (r < 0) r += N0; } }
    else if (r >= N0) { r -= N0; if (r >= N0) r %= N0; }
    return r;
}

// ---- axis-1 contraction of the lag rows -----------------------------------------------------------------------------
// part[ks][row][l] = (1/N1) sum_{k1 in chunk ks} wt(k1) Re(kap[k1][row] e^{+2 pi i k1 (l - 2 w1) / N1}),  l = 0 .. 4 w1
// (wt = 1 for k1 = 0 and the Nyquist column, 2 otherwise: Hermitian half spectrum).  Every row gets all 4 w1 + 1 lags;
// lag_finish_kernel sums the chunks and scatters the lags each table needs.
#define LR2_LB 9          // lags per pass: m = mb .. mb + LB - 1, both signs
#define LR2_KC 512        // k1 columns per CTA (chunk); their twiddles for the lags of a pass are staged in shared memory
struct LagReduce2Args {
    int N1, NH, nrows, w1, ksplit;
    int rb0;                 // first 16-row block of this launch (shared-template tiles reduce only the rows that changed)
    const cd* tw1;
};

#ifdef SFFTB_TU_FIT
__global__ void __launch_bounds__(256) lag_reduce2_kernel(LagReduce2Args a, const cd* __restrict__ kap, double* __restrict__ part)
{
    __shared__ double red[16][16][2 * LR2_LB + 1];
    extern __shared__ __align__(16) unsigned char lr2_smem[];
    cd (*twc)[LR2_LB] = reinterpret_cast<cd (*)[LR2_LB]>(lr2_smem);   // twc[k - kbeg][t] = wt(k)/N1 exp(-2 pi i k (mb + t) / N1)
    const int tid = threadIdx.x;
    const int rl = tid & 15, kl = tid >> 4;
    const int rb = blockIdx.x + a.rb0;
    const int row = rb * 16 + rl;
    const int ks = blockIdx.y;
    const int kbeg = ks * LR2_KC, kend = min(a.NH, kbeg + LR2_KC);
    const double inv1 = 1.0 / (double)a.N1;
    const int nl = 4 * a.w1 + 1;
    for (int mb = 0; mb <= 2 * a.w1; mb += LR2_LB) {
        for (int idx = tid; idx < (kend - kbeg) * LR2_LB; idx += 256) {
            const int kk = idx / LR2_LB, t = idx - kk * LR2_LB;
            const int k = kbeg + kk;
            const double wt = (k == 0 || 2 * k == a.N1) ? inv1 : 2.0 * inv1;
            const int e = (int)(((long long)k * (mb + t)) % a.N1);
            twc[kk][t] = cscale(a.tw1[e], wt);
        }
        __syncthreads();
        double accP[LR2_LB], accN[LR2_LB];
#pragma unroll
        for (int t = 0; t < LR2_LB; ++t) { accP[t] = 0.0; accN[t] = 0.0; }
        if (row < a.nrows) {
            for (int k = kbeg + kl; k < kend; k += 16) {
                const cd v = kap[(size_t)k * a.nrows + row];
#pragma unroll
                for (int t = 0; t < LR2_LB; ++t) {
                    const cd w = twc[k - kbeg][t];            // (cos, -sin) of 2 pi k m / N1, weighted
                    const double t1 = v.x * w.x, t2 = v.y * w.y;
                    accP[t] += t1 + t2;                       // Re(v e^{+i th m}) = vr cos - vi sin = vr w.x + vi w.y
                    accN[t] += t1 - t2;                       // Re(v e^{-i th m})
                }
            }
        }
#pragma unroll
        for (int t = 0; t < LR2_LB; ++t) { red[kl][rl][t] = accP[t]; red[kl][rl][LR2_LB + t] = accN[t]; }
        __syncthreads();
        for (int idx = tid; idx < 16 * 2 * LR2_LB; idx += 256) {
            const int r2 = idx / (2 * LR2_LB), t2 = idx - r2 * (2 * LR2_LB);
            const int rr = rb * 16 + r2;
            const bool neg = t2 >= LR2_LB;
            const int m = mb + (neg ? t2 - LR2_LB : t2);
            if (rr < a.nrows && m <= 2 * a.w1 && !(neg && m == 0)) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 16; ++q) s += red[q][r2][t2];
                part[((size_t)ks * a.nrows + rr) * nl + (neg ? 2 * a.w1 - m : 2 * a.w1 + m)] = s;
            }
        }
        __syncthreads();
    }
}
#endif  // SFFTB_TU_FIT

struct LagFinishArgs {
    int nrows, nOm, nK, nLT, w1, ksplit;
    double* R; double* RJ; double* RT; double* RJT;
};

#ifdef SFFTB_TU_FIT
__global__ void lag_finish_kernel(LagFinishArgs a, const double* __restrict__ part)
{
    const int nl = 4 * a.w1 + 1, nlj1 = 2 * a.w1 + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.nrows * nl) return;
    const int row = idx / nl, l = idx - row * nl;
    double s = 0.0;
    for (int ks = 0; ks < a.ksplit; ++ks) s += part[((size_t)ks * a.nrows + row) * nl + l];
    const int m1 = l - 2 * a.w1;
    if (row < a.nOm) { a.R[(size_t)row * nl + l] = s; return; }
    if (m1 < -a.w1 || m1 > a.w1) return;
    if (row < a.nK) a.RJ[(size_t)(row - a.nOm) * nlj1 + (m1 + a.w1)] = s;
    else if (row < a.nK + a.nLT) a.RT[(size_t)(row - a.nK) * nlj1 + (m1 + a.w1)] = s;
    else if (m1 == 0) a.RJT[row - a.nK - a.nLT] = s;
}
#endif  // SFFTB_TU_FIT
