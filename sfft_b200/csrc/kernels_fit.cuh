// kernels_fit.cuh -- the fit column pass and the lag-table reductions.
//
// The reference materialises every cross-spectrum plane F(A) conj(F(B)), runs a full 2-D FFT on each and reads
// (4w+1)^2 numbers out of it (HadProd_* + fft2 + FillLS_*, sfft/sfftcore/SFFTSubtract.py:224-383).  Here one CTA
// owns one column k1 of the transposed row spectra and, per DIF-folded slice t of the axis-0 frequencies
// (k0 = V u + t), forms the slice spectra in shared memory, multiplies pairs, inverse-transforms the product and
// keeps only the lags FillLS_* would read:
//     kappa_AB[m0; k1] = sum_r conj(g_A[r;k1]) g_B[(r+m0)%N0; k1]
// A second tiny kernel reduces over k1 with the axis-1 twiddles and Hermitian weights to the lag tables
//     R_AB[m0, m1] = sum_x A[x] B[x+m]   (circular).
// Cross terms with the background basis T_pq = cx^p cy^q never touch an FFT: they come from polynomial moments of
// the columns with explicit wrap-row corrections (T_pq is separable; the reference FFTs it as a full plane,
// SFFTSubtract.py:157-161).
#pragma once
#include "fft_smem.cuh"

#define NT_COL 512

// ---- fold: z_t[n] = W_N0^{t n} sum_v W_V^{t v} cx(n + M v)^i g_j[n + M v]  for every I-plane (i, j) and for J ----
template <typename TSt>
__device__ __forceinline__ void fold_slice(const ColArgs& a, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
                                           int k1, int t, cd* S, bool withJ)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    const double inv0 = 1.0 / (double)a.N0;
    const int nsrc = a.nj + (withJ ? 1 : 0);
    for (int idx = tid; idx < nsrc * a.M; idx += nthr) {
        const int jj = idx / a.M, n = idx - jj * a.M;
        const bool isJ = (jj == a.nj);
        const TSt* col = isJ ? (gJ + (size_t)k1 * a.N0) : (gI + ((size_t)jj * a.NH + k1) * a.N0);
        const int ni = isJ ? 1 : (a.DK - jj + 1);       // number of powers i = 0 .. DK - j
        cd s0 = cmake(0, 0), s1 = s0, s2 = s0, s3 = s0;
        for (int v = 0; v < a.V; ++v) {
            const int r = n + a.M * v;
            cd g = load_c(col + r);
            if (t != 0 && v != 0) g = cmul(g, a.tw0[((t * v) % a.V) * a.M]);
            const double cx = (r + 1) * inv0;
            s0 = cadd(s0, g);
            if (ni > 1) { g = cscale(g, cx); s1 = cadd(s1, g); }
            if (ni > 2) { g = cscale(g, cx); s2 = cadd(s2, g); }
            if (ni > 3) { g = cscale(g, cx); s3 = cadd(s3, g); }
        }
        const cd wn = (t == 0) ? cmake(1.0, 0.0) : a.tw0[t * n];
        if (isJ) {
            S[(size_t)a.Fij * a.pitch + n] = cmul(s0, wn);
        } else {
            S[(size_t)a.plane_of[0][jj] * a.pitch + n] = cmul(s0, wn);
            if (ni > 1) S[(size_t)a.plane_of[1][jj] * a.pitch + n] = cmul(s1, wn);
            if (ni > 2) S[(size_t)a.plane_of[2][jj] * a.pitch + n] = cmul(s2, wn);
            if (ni > 3) S[(size_t)a.plane_of[3][jj] * a.pitch + n] = cmul(s3, wn);
        }
    }
}

// ---- polynomial moments of one column: nu[jj][e] = sum_r cx(r)^e g_jj[r]  (jj = nj is J) ----
#define SFFTB_MAXE 7      // e = 0 .. DK + DB <= 6
template <typename TSt>
__device__ void column_moments(const ColArgs& a, const TSt* __restrict__ gI, const TSt* __restrict__ gJ, int k1,
                               cd* mom /* [(nj+1)][SFFTB_MAXE] */, cd* red /* [nwarps][SFFTB_MAXE] */)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const double inv0 = 1.0 / (double)a.N0;
    for (int jj = 0; jj <= a.nj; ++jj) {
        const bool isJ = (jj == a.nj);
        const TSt* col = isJ ? (gJ + (size_t)k1 * a.N0) : (gI + ((size_t)jj * a.NH + k1) * a.N0);
        const int ne = isJ ? (a.DB + 1) : (a.DK - jj + a.DB + 1);
        cd acc[SFFTB_MAXE];
#pragma unroll
        for (int e = 0; e < SFFTB_MAXE; ++e) acc[e] = cmake(0, 0);
        for (int r = tid; r < a.N0; r += nthr) {
            cd g = load_c(col + r);
            const double cx = (r + 1) * inv0;
#pragma unroll
            for (int e = 0; e < SFFTB_MAXE; ++e) {
                if (e < ne) { acc[e] = cadd(acc[e], g); g = cscale(g, cx); }
            }
        }
#pragma unroll
        for (int e = 0; e < SFFTB_MAXE; ++e) {
            const double sx = warp_sum(acc[e].x), sy = warp_sum(acc[e].y);
            if (lane == 0) red[warp * SFFTB_MAXE + e] = cmake(sx, sy);
        }
        __syncthreads();
        if (tid < SFFTB_MAXE) {
            cd s = cmake(0, 0);
            for (int w = 0; w < nwarps; ++w) s = cadd(s, red[w * SFFTB_MAXE + tid]);
            mom[jj * SFFTB_MAXE + tid] = s;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ double binom_small(int n, int k) {
    // n <= 3
    const double tbl[4][4] = {{1, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 1, 0}, {1, 3, 3, 1}};
    return tbl[n][k];
}

// lam[(A, p, ia)][k1] = sum_r cx(r)^i cx((r+a)%N0)^p g_j[r]   with (i, j) = plane A, a = ia - w0
template <typename TSt>
__device__ void column_poly_terms(const ColArgs& a, const TSt* __restrict__ gI, int k1, const cd* mom,
                                  cd* __restrict__ lam, cd* __restrict__ nuJ)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    const double inv0 = 1.0 / (double)a.N0;
    const int np = a.DB + 1;
    for (int idx = tid; idx < a.Fij * np * a.nlj0; idx += nthr) {
        const int ia = idx % a.nlj0;
        const int p = (idx / a.nlj0) % np;
        const int A = idx / (a.nlj0 * np);
        const int i = a.pl_i[A], j = a.pl_j[A];
        const int sh = ia - a.w0;
        const double beta = sh * inv0;
        cd v = cmake(0, 0);
        for (int e = 0; e <= p; ++e) {
            const double c = binom_small(p, e) * ipow(beta, p - e);
            const cd m = mom[j * SFFTB_MAXE + i + e];
            v.x += c * m.x; v.y += c * m.y;
        }
        if (p > 0 && sh != 0) {
            const TSt* col = gI + ((size_t)j * a.NH + k1) * a.N0;
            const int rbeg = sh > 0 ? a.N0 - sh : 0;
            const int rend = sh > 0 ? a.N0 : -sh;
            const double wrap = sh > 0 ? -1.0 : 1.0;
            for (int r = rbeg; r < rend; ++r) {
                const double cx = (r + 1) * inv0;
                const double corr = ipow(cx + beta + wrap, p) - ipow(cx + beta, p);
                const double c = ipow(cx, i) * corr;
                const cd g = load_c(col + r);
                v.x += c * g.x; v.y += c * g.y;
            }
        }
        lam[(size_t)idx * a.NH + k1] = v;
    }
    for (int p = tid; p < np; p += nthr) nuJ[(size_t)p * a.NH + k1] = mom[a.nj * SFFTB_MAXE + p];
}

// smem layout (cd units): S[(Fij+1)*pitch] | Wk[PB*pitch] | acc[nacc] | mom[(nj+1)*MAXE] | red[16*MAXE]
template <typename TSt>
__global__ void __launch_bounds__(NT_COL) fit_col_kernel(ColArgs a, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
                                                         cd* __restrict__ kap, cd* __restrict__ lam, cd* __restrict__ nuJ)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* S = reinterpret_cast<cd*>(smem_raw);
    cd* Wk = S + (size_t)(a.Fij + 1) * a.pitch;
    const int nOm = a.npairs * a.nl0;
    const int nacc = nOm + a.Fij * a.nlj0;
    cd* acc = Wk + (size_t)a.PB * a.pitch;
    cd* mom = acc + nacc;
    cd* red = mom + (a.nj + 1) * SFFTB_MAXE;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int ntot = a.npairs + a.Fij;
    const double invN0 = 1.0 / (double)a.N0;

    for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x) {
        for (int idx = tid; idx < nacc; idx += nthr) acc[idx] = cmake(0, 0);
        column_moments(a, gI, gJ, k1, mom, red);
        column_poly_terms(a, gI, k1, mom, lam, nuJ);

        for (int t = 0; t < a.V; ++t) {
            fold_slice(a, gI, gJ, k1, t, S, true);
            __syncthreads();
            fft_planes(S, a.pitch, a.Fij + 1, a.fd, a.twM, -1.0);
            for (int g0 = 0; g0 < ntot; g0 += a.PB) {
                const int np = min(a.PB, ntot - g0);
                for (int idx = tid; idx < np * a.M; idx += nthr) {
                    const int p = idx / a.M, u = idx - p * a.M;
                    const int q = g0 + p;
                    const int A = q < a.npairs ? a.pairA[q] : q - a.npairs;
                    const int B = q < a.npairs ? a.pairB[q] : a.Fij;
                    Wk[(size_t)p * a.pitch + u] = cmulcj(S[(size_t)A * a.pitch + u], S[(size_t)B * a.pitch + u]);
                }
                __syncthreads();
                fft_planes(Wk, a.pitch, np, a.fd, a.twM, +1.0);
                for (int idx = tid; idx < np * a.nl0; idx += nthr) {
                    const int p = idx / a.nl0, l = idx - p * a.nl0;
                    const int q = g0 + p;
                    int m0, ai;
                    if (q < a.npairs) { m0 = l - 2 * a.w0; ai = q * a.nl0 + l; }
                    else if (l < a.nlj0) { m0 = l - a.w0; ai = nOm + (q - a.npairs) * a.nlj0 + l; }
                    else continue;
                    const cd y = Wk[(size_t)p * a.pitch + imod(m0, a.M)];
                    const cd w = a.tw0[imod(t * m0, a.N0)];            // e^{+2 pi i t m0 / N0} = conj(w)
                    acc[ai] = cadd(acc[ai], cmul(y, cconj(w)));
                }
                __syncthreads();
            }
        }
        for (int idx = tid; idx < nacc; idx += nthr) kap[(size_t)idx * a.NH + k1] = cscale(acc[idx], invN0);
        __syncthreads();
    }
}

// ---- axis-1 reductions -------------------------------------------------------------------------------------------
struct ReduceArgs {
    int N1, NH;
    int w1;
    int nOm;          // rows of kap that belong to Omega pairs (4 w1 + 1 output lags); the rest are Theta rows (2 w1 + 1)
    int nrows;
    const cd* tw1;
};

// R[row][l1] = (1/N1) sum_{k1} wt(k1) Re(kap[row][k1] e^{+2 pi i k1 m1 / N1})
#ifdef SFFTB_TU_FIT
__global__ void __launch_bounds__(256) lag_reduce_kernel(ReduceArgs a, const cd* __restrict__ kap, double* __restrict__ R,
                                                         double* __restrict__ RJ)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* rowbuf = reinterpret_cast<cd*>(smem_raw);
    const int row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const double inv1 = 1.0 / (double)a.N1;
    for (int k = tid; k < a.NH; k += blockDim.x) {
        double wt = (k == 0 || (2 * k == a.N1)) ? inv1 : 2.0 * inv1;
        rowbuf[k] = cscale(kap[(size_t)row * a.NH + k], wt);
    }
    __syncthreads();
    const bool om = row < a.nOm;
    const int nl1 = om ? 4 * a.w1 + 1 : 2 * a.w1 + 1;
    const int base = om ? 2 * a.w1 : a.w1;
    double* out = om ? (R + (size_t)row * nl1) : (RJ + (size_t)(row - a.nOm) * nl1);
    for (int l1 = warp; l1 < nl1; l1 += nwarps) {
        const int m1 = l1 - base;
        double s = 0.0;
        for (int k = lane; k < a.NH; k += 32) {
            const int e = (int)(((long long)k * m1) % a.N1);
            const cd w = a.tw1[e < 0 ? e + a.N1 : e];        // conj(w) = e^{+...}
            const cd v = rowbuf[k];
            s += v.x * w.x + v.y * w.y;                      // Re(v conj(w))
        }
        s = warp_sum(s);
        if (lane == 0) out[l1] = s;
    }
}
#endif  // SFFTB_TU_FIT

struct PolyReduceArgs {
    int N1, NH, w1, DB, Fpq, Fij;
    int nlj0, nlj1;
    int nrowsL;                 // Fij * (DB+1) * nlj0 rows of lam, followed by DB+1 rows of nuJ
    const cd* tw1;
    const cd* Q;                // Q[q][k1] = DFT(cy^q), q = 0..DB
    signed char pq_of[4][4];    // (p, q) -> pq index or -1
};

// RT[A][pq][ia][ib] = (1/N1) sum_k1 wt Re(lam[(A,p,ia)][k1] conj(Q_q[k1]) e^{-2 pi i k1 b / N1});  RJT[pq] likewise
#ifdef SFFTB_TU_FIT
__global__ void __launch_bounds__(256) poly_reduce_kernel(PolyReduceArgs a, const cd* __restrict__ lam, const cd* __restrict__ nuJ,
                                                          double* __restrict__ RT, double* __restrict__ RJT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* rowbuf = reinterpret_cast<cd*>(smem_raw);
    const int row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const double inv1 = 1.0 / (double)a.N1;
    const bool isJ = row >= a.nrowsL;
    const cd* src = isJ ? (nuJ + (size_t)(row - a.nrowsL) * a.NH) : (lam + (size_t)row * a.NH);
    for (int k = tid; k < a.NH; k += blockDim.x) {
        double wt = (k == 0 || (2 * k == a.N1)) ? inv1 : 2.0 * inv1;
        rowbuf[k] = cscale(src[k], wt);
    }
    __syncthreads();
    const int np = a.DB + 1;
    int A = 0, p, ia = 0;
    if (isJ) { p = row - a.nrowsL; }
    else { ia = row % a.nlj0; p = (row / a.nlj0) % np; A = row / (a.nlj0 * np); }
    const int nb = isJ ? 1 : a.nlj1;
    for (int c = warp; c < np * nb; c += nwarps) {
        const int q = c / nb, ib = c - q * nb;
        if (p + q > a.DB) continue;
        const int b = isJ ? 0 : ib - a.w1;
        const cd* Qq = a.Q + (size_t)q * a.NH;
        double s = 0.0;
        for (int k = lane; k < a.NH; k += 32) {
            const cd v = cmul(rowbuf[k], cconj(Qq[k]));
            const int e = (int)(((long long)k * b) % a.N1);
            const cd w = a.tw1[e < 0 ? e + a.N1 : e];        // e^{-2 pi i k b / N1}
            s += v.x * w.x - v.y * w.y;                      // Re(v w)
        }
        s = warp_sum(s);
        if (lane == 0) {
            const int pq = a.pq_of[p][q];
            if (isJ) RJT[pq] = s;
            else RT[(((size_t)A * a.Fpq + pq) * a.nlj0 + ia) * a.nlj1 + ib] = s;
        }
    }
}
#endif  // SFFTB_TU_FIT

// ---- normal-equation fill (FillLS_* + Remove_LSFStripes restated through LHMAT = D^T D / N) -------------------------
struct FillArgs {
    int Fij, Fpq, Fab, Fijab, L0, L1, w0, w1;
    int nl0, nl1, nlj0, nlj1;
    double invN, invN2, invN3;
    const double* R; const double* RJ; const double* RT; const double* RJT; const double* PHI;
    // kernel regulariser (sfft/BSplineSFFT.py:3570-3700): LHMAT[(k,c),(k',c')] += regw * SST[k,k'] * iREG[c,c'],
    // regw = LAMBDA_REGULARIZE * SCALE^2; null pointers = off
    const double* SST; const double* iREG; double regw;
    // SEPARATE-VARYING scaling: the centre taps are penalised through the scaling basis (fill_regmat, :2122-2166):
    // CSST[k][k'] = kernel basis k x scaling basis k' (one tap is a centre tap), DSST = scaling x scaling (both are)
    const double* CSST; const double* DSST;
    // SEPARATE-VARYING scaling with polynomial bases (sfft/BSplineSFFT.py:2487-2495, 3733-3747): the centre-tap unknown of
    // plane k scales the image times the k-th SCALING basis function, i.e. the unshifted plane sca[k] of the kernel's own
    // plane set (-1: no such unknown; its row / column of the exported LHMAT is zero like the reference's placeholder)
    int sca_on; signed char sca[16];
};

#ifdef SFFTB_TU_MAIN
__device__ __forceinline__ double fill_R(const FillArgs& f, int A, int B, int m0, int m1) {
    if (A > B) { int t = A; A = B; B = t; m0 = -m0; m1 = -m1; }
    const int pidx = A * f.Fij - (A * (A - 1)) / 2 + (B - A);
    return f.R[((size_t)pidx * f.nl0 + (m0 + 2 * f.w0)) * f.nl1 + (m1 + 2 * f.w1)];
}

__device__ __forceinline__ double fill_lh_kernel_block(const FillArgs& f, int A, int B, int Ak, int Bk, int ab8, int ab,
                                                        int a8, int b8, int a0, int b0, bool nz8, bool nz) {
    double v = fill_R(f, A, B, a8 - a0, b8 - b0);
    if (nz) v -= fill_R(f, A, B, a8, b8);
    if (nz8) v -= fill_R(f, A, B, -a0, -b0);
    if (nz && nz8) v += fill_R(f, A, B, 0, 0);
    v *= f.invN3;
    if (f.SST) {
        double g = f.SST[Ak * f.Fij + Bk];
        if (f.CSST) {
            if (!nz8 && !nz) g = f.DSST[Ak * f.Fij + Bk];
            else if (!nz) g = f.CSST[Ak * f.Fij + Bk];          // row tap off-centre, column tap = centre
            else if (!nz8) g = f.CSST[Bk * f.Fij + Ak];
        }
        v = fma(f.regw * g, f.iREG[(size_t)ab8 * f.Fab + ab], v);
    }
    return v;
}

__device__ double fill_lh_entry(const FillArgs& f, int fr, int fc) {
    if (fr < f.Fijab && fc < f.Fijab) {
        const int A = fr / f.Fab, ab8 = fr - A * f.Fab;
        const int B = fc / f.Fab, ab = fc - B * f.Fab;
        const int a8 = ab8 / f.L1 - f.w0, b8 = ab8 % f.L1 - f.w1;
        const int a0 = ab / f.L1 - f.w0, b0 = ab % f.L1 - f.w1;
        const bool nz8 = (a8 != 0) || (b8 != 0), nz = (a0 != 0) || (b0 != 0);
        const int Ak = A, Bk = B;                      // kernel-plane indices (the regulariser acts on those)
        int Ap = A, Bp = B;
        if (f.sca_on) {
            if (!nz8) Ap = f.sca[A];
            if (!nz) Bp = f.sca[B];
            if (Ap < 0 || Bp < 0) return 0.0;
        }
        return fill_lh_kernel_block(f, Ap, Bp, Ak, Bk, ab8, ab, a8, b8, a0, b0, nz8, nz);
    }

    if (fr >= f.Fijab && fc >= f.Fijab) return f.PHI[(fr - f.Fijab) * f.Fpq + (fc - f.Fijab)] * f.invN;
    if (fr >= f.Fijab) { int t = fr; fr = fc; fc = t; }
    int A = fr / f.Fab;
    const int ab8 = fr - A * f.Fab;
    const int a8 = ab8 / f.L1, b8 = ab8 % f.L1;
    const int pq = fc - f.Fijab;
    if (f.sca_on && a8 == f.w0 && b8 == f.w1) { A = f.sca[A]; if (A < 0) return 0.0; }
    const double* T = f.RT + ((size_t)A * f.Fpq + pq) * f.nlj0 * f.nlj1;
    double v = T[a8 * f.nlj1 + b8];
    if (a8 != f.w0 || b8 != f.w1) v -= T[f.w0 * f.nlj1 + f.w1];
    return v * f.invN2;
}

__device__ double fill_rhs_entry(const FillArgs& f, int fr) {
    if (fr >= f.Fijab) return f.RJT[fr - f.Fijab] * f.invN;
    int A = fr / f.Fab;
    const int ab = fr - A * f.Fab;
    const int a0 = ab / f.L1, b0 = ab % f.L1;
    if (f.sca_on && a0 == f.w0 && b0 == f.w1) { A = f.sca[A]; if (A < 0) return 0.0; }
    const double* T = f.RJ + (size_t)A * f.nlj0 * f.nlj1;
    double v = T[a0 * f.nlj1 + b0];
    if (a0 != f.w0 || b0 != f.w1) v -= T[f.w0 * f.nlj1 + f.w1];
    return v * f.invN2;
}

// sc[rr] = 1 / sqrt(diag) for the symmetric diagonal scaling; flags non-finite / non-positive diagonals
__global__ void fill_diag_kernel(FillArgs f, const int* __restrict__ idx, int n, double* __restrict__ sc, int* __restrict__ info)
{
    const int rr = blockIdx.x * blockDim.x + threadIdx.x;
    if (rr >= n) return;
    const double d = fill_lh_entry(f, idx[rr], idx[rr]);
    if (!isfinite(d)) { atomicExch(&info[1], 1); sc[rr] = 1.0; }
    else if (!(d > 0.0)) { atomicExch(&info[0], rr + 1); sc[rr] = 1.0; }
    else sc[rr] = rsqrt(d);
}

// scaled right-hand side only (cached-factor path): b[rr] = RHb[idx[rr]] * sc[rr]
__global__ void fill_rhs_kernel(FillArgs f, const int* __restrict__ idx, int n, const double* __restrict__ sc,
                                double* __restrict__ b, int* __restrict__ info)
{
    const int rr = blockIdx.x * blockDim.x + threadIdx.x;
    if (rr >= n) return;
    const double v = fill_rhs_entry(f, idx[rr]) * sc[rr];
    if (!isfinite(v)) atomicExch(&info[1], 1);
    b[rr] = v;
}

// Aug is (n+1) x ld row-major: rows/cols 0..n-1 = scaled compact LHMAT, row n = scaled RHb (and column n mirrors it).
// sc == nullptr -> unscaled (used for the export hook with idx = identity).
__global__ void fill_matrix_kernel(FillArgs f, const int* __restrict__ idx, int n, const double* __restrict__ sc,
                                   double* __restrict__ Aug, int ld, int* __restrict__ info)
{
    const int cc = blockIdx.x * blockDim.x + threadIdx.x;
    const int rr = blockIdx.y * blockDim.y + threadIdx.y;
    if (rr > n || cc > n) return;
    double v;
    if (rr < n && cc < n) {
        v = fill_lh_entry(f, idx[rr], idx[cc]);
        if (sc) v *= sc[rr] * sc[cc];
    } else if (rr == n && cc == n) {
        v = 0.0;
    } else {
        const int k = rr == n ? cc : rr;
        v = fill_rhs_entry(f, idx[k]);
        if (sc) v *= sc[k];
    }
    if (!isfinite(v)) atomicExch(&info[1], 1);
    Aug[(size_t)rr * ld + cc] = v;
}
#endif  // SFFTB_TU_MAIN
