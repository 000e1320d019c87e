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
#include "fft_regs.cuh"
#include "kernels_fit.cuh"

#define FSG_NT 256
#define FSG_M 256
#define FSG_PITCH 272
#define FSG_NBUF 16

struct SegFitArgs {
    ColArgs c;
    int S, nseg, h;                  // core rows per segment, number of segments, halo = 2 w0
    int nrows;                       // rows per column of kap
    int nOm, nK, nLT;                // Omega rows | + Theta rows | lam*Q rows (then Fpq nuJ*Q rows)
    const cd* tabA;                  // engine table of the second radix-16 pass (240 entries)
    const cd* Q;                     // Q[q][k1] = DFT_c(cy(c)^q), q = 0..DB
    signed char pq_of[4][4];
};

// Rows of the background cross terms for one column (T_pq = cx^p cy^q is separable, so no T plane is ever transformed;
// the reference FFTs them as full planes, sfft/sfftcore/SFFTSubtract.py:157-161):
//   lam_(A,p,ia) = sum_r cx(r)^i cx((r+a)%N0)^p g_j[r]     (from the column moments + explicit wrap rows)
//   row (A,pq,ia) = conj(lam_(A,p,ia)) Q_q[k1];   row pq of the last block = conj(nuJ_p) Q_q[k1]
template <typename TSt>
__device__ void column_poly_rows(const SegFitArgs& fa, const TSt* __restrict__ gI, int k1, const cd* mom, cd* __restrict__ kaprow)
{
    const ColArgs& a = fa.c;
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
        for (int q = 0; q + p <= a.DB; ++q) {
            const int pq = fa.pq_of[p][q];
            kaprow[fa.nK + (A * a.Fpq + pq) * a.nlj0 + ia] = cmulcj(v, fa.Q[(size_t)q * a.NH + k1]);
        }
    }
    for (int p = tid; p < np; p += nthr)
        for (int q = 0; q + p <= a.DB; ++q)
            kaprow[fa.nK + fa.nLT + fa.pq_of[p][q]] = cmulcj(mom[a.nj * SFFTB_MAXE + p], fa.Q[(size_t)q * a.NH + k1]);
}

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

#define FSG_NMOM 48                  // threads of the idle transform groups that accumulate the column moments
#define FSG_MSLOTS (4 * SFFTB_MAXE)  // (nj + 1 <= 4 source planes) x (e = 0 .. MAXE-1)

// smem: spec[FSG_NBUF * FSG_PITCH] cd | mom[4 * MAXE] cd | macc[FSG_MSLOTS * FSG_NMOM] cd | tabA[240] cd |
//       stage[2][nj + 1][FSG_M] TSt
template <typename TSt, int DK>
__global__ void __launch_bounds__(FSG_NT, 1) fit_seg_kernel(SegFitArgs fa, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
                                                            cd* __restrict__ kap)
{
    constexpr int Fij = (DK + 1) * (DK + 2) / 2;
    constexpr int NPAIR = Fij * (Fij + 1) / 2;
    constexpr int NACC = NPAIR + Fij;
    constexpr int NP = 2 * Fij + 1;
    constexpr int NSRC = DK + 2;             // stored row-spectrum planes: g_0 .. g_DK, J
    static_assert(NP <= FSG_NBUF - 3, "KerPolyOrder too large for the segmented fit kernel");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ColArgs& a = fa.c;
    cd* spec = reinterpret_cast<cd*>(smem_raw);
    cd* mom = spec + FSG_NBUF * FSG_PITCH;
    cd* macc = mom + 4 * SFFTB_MAXE;
    cd* tabA = macc + FSG_MSLOTS * FSG_NMOM;
    TSt* stage = reinterpret_cast<TSt*>(tabA + 240);
    const int tid = threadIdx.x;
    const int grp = tid >> 4, lane = tid & 15;
    GroupSync gs;
    gs.mask = 0xffffu << (16 * (grp & 1));
    gs.bar_id = 0; gs.count = 0;
    cd* plane = spec + grp * FSG_PITCH;
    const double inv0 = 1.0 / (double)a.N0;
    const int h = fa.h, S = fa.S;

    for (int i = tid; i < 240; i += FSG_NT) tabA[i] = fa.tabA[i];

    // role of this transform group: A-role plane grp | B-role plane grp - Fij | B-role J | moment accumulation
    const bool roleA = grp < Fij, isJ = grp == 2 * Fij, active = grp < NP;
    const int pl = roleA ? grp : (grp < 2 * Fij ? grp - Fij : 0);
    const int my_i = isJ ? 0 : a.pl_i[pl];
    const int my_src = isJ ? DK + 1 : a.pl_j[pl];
    const int mt = tid - (FSG_NT - FSG_NMOM);        // >= 0: this thread accumulates moments

    for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x) {
        cd acc[NACC];
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[q] = cmake(0.0, 0.0);
        cd* kaprow = kap + (size_t)k1 * fa.nrows;
        if (mt >= 0)
            for (int s = 0; s < FSG_MSLOTS; ++s) macc[s * FSG_NMOM + mt] = cmake(0.0, 0.0);

        // prefetch the window of segment 0 (element tid of every source plane)
        {
            const int r = wrap_row(-h + tid, a.N0);
#pragma unroll
            for (int jj = 0; jj < NSRC; ++jj) {
                const TSt* col = (jj == DK + 1) ? gJ + (size_t)k1 * a.N0 : gI + ((size_t)jj * a.NH + k1) * a.N0;
                cp_async_elem(stage + jj * FSG_M + tid, col + r);
            }
            cp_async_commit();
        }
        for (int seg = 0; seg < fa.nseg; ++seg) {
            const int c0 = seg * S;
            const int Sc = min(S, a.N0 - c0);
            const TSt* st = stage + (seg & 1) * NSRC * FSG_M;
            if (seg + 1 < fa.nseg) {
                TSt* nx = stage + ((seg + 1) & 1) * NSRC * FSG_M;
                const int r = wrap_row(c0 + S - h + tid, a.N0);
#pragma unroll
                for (int jj = 0; jj < NSRC; ++jj) {
                    const TSt* col = (jj == DK + 1) ? gJ + (size_t)k1 * a.N0 : gI + ((size_t)jj * a.NH + k1) * a.N0;
                    cp_async_elem(nx + jj * FSG_M + tid, col + r);
                }
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();          // window of this segment visible; products of the previous segment done
            if (active) {
                cd v[16];
                const TSt* src = st + my_src * FSG_M;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int n = lane + 16 * q;
                    cd g = cmake(0.0, 0.0);
                    if (!roleA || (n >= h && n < h + Sc)) {
                        g = load_c(src + n);
                        if (my_i > 0) {
                            const double cx = (wrap_row(c0 - h + n, a.N0) + 1) * inv0;
                            g = cscale(g, my_i == 1 ? cx : (my_i == 2 ? cx * cx : cx * cx * cx));
                        }
                    }
                    v[q] = g;
                }
                reg_fft<FSG_M>(v, plane, lane, tabA, nullptr, nullptr, -1.0, gs);
#pragma unroll
                for (int q = 0; q < 16; ++q) plane[RPAD(lane + 16 * q)] = v[q];
            } else if (mt >= 0) {
                // column moments nu[jj][e] += sum_{core rows} cx(r)^e g_jj[r]   (e <= DK - jj + DB for I planes, <= DB for J)
                for (int n = h + mt; n < h + Sc; n += FSG_NMOM) {
                    const double cx = (c0 + (n - h) + 1) * inv0;
#pragma unroll
                    for (int jj = 0; jj < NSRC; ++jj) {
                        const int ne = (jj == DK + 1) ? a.DB + 1 : DK - jj + a.DB + 1;
                        cd g = load_c(st + jj * FSG_M + n);
                        for (int e = 0; e < ne; ++e) {
                            cd* slot = macc + (jj * SFFTB_MAXE + e) * FSG_NMOM + mt;
                            *slot = cadd(*slot, g);
                            g = cscale(g, cx);
                        }
                    }
                }
            }
            __syncthreads();
            {
                cd fA[Fij], fB[Fij];
#pragma unroll
                for (int A = 0; A < Fij; ++A) {
                    fA[A] = spec[A * FSG_PITCH + RPAD(tid)];
                    fB[A] = spec[(Fij + A) * FSG_PITCH + RPAD(tid)];
                }
                const cd fJ = spec[2 * Fij * FSG_PITCH + RPAD(tid)];
                int q = 0;
#pragma unroll
                for (int A = 0; A < Fij; ++A)
#pragma unroll
                    for (int B = A; B < Fij; ++B) {
                        // acc += conj(fA) fB
                        acc[q].x = fma(fA[A].x, fB[B].x, acc[q].x); acc[q].x = fma(fA[A].y, fB[B].y, acc[q].x);
                        acc[q].y = fma(fA[A].x, fB[B].y, acc[q].y); acc[q].y = fma(-fA[A].y, fB[B].x, acc[q].y);
                        ++q;
                    }
#pragma unroll
                for (int A = 0; A < Fij; ++A) {
                    acc[NPAIR + A].x = fma(fA[A].x, fJ.x, acc[NPAIR + A].x); acc[NPAIR + A].x = fma(fA[A].y, fJ.y, acc[NPAIR + A].x);
                    acc[NPAIR + A].y = fma(fA[A].x, fJ.y, acc[NPAIR + A].y); acc[NPAIR + A].y = fma(-fA[A].y, fJ.x, acc[NPAIR + A].y);
                }
            }
        }
        __syncthreads();

        // ---- column moments -> background cross-term rows ----
        if (tid < FSG_MSLOTS) {
            cd s = cmake(0.0, 0.0);
            for (int t = 0; t < FSG_NMOM; ++t) s = cadd(s, macc[tid * FSG_NMOM + t]);
            mom[tid] = s;
        }
        __syncthreads();
        column_poly_rows(fa, gI, k1, mom, kaprow);

        // ---- one inverse transform per pair; keep the lags FillLS_* reads ----
        const double invM = 1.0 / (double)FSG_M;
#pragma unroll
        for (int b0 = 0; b0 < NACC; b0 += FSG_NBUF) {
#pragma unroll
            for (int q = 0; q < FSG_NBUF; ++q)
                if (b0 + q < NACC) spec[q * FSG_PITCH + RPAD(tid)] = acc[b0 + q];
            __syncthreads();
            const int job = b0 + grp;
            if (job < NACC) {
                cd v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = plane[RPAD(lane + 16 * q)];
                gs.sync<16>();
                reg_fft<FSG_M>(v, plane, lane, tabA, nullptr, nullptr, +1.0, gs);
#pragma unroll
                for (int q = 0; q < 16; ++q) plane[RPAD(lane + 16 * q)] = v[q];
                gs.sync<16>();
                const bool om = job < NPAIR;
                const int lim = om ? 2 * a.w0 : a.w0;
                const int rowbase = om ? job * a.nl0 : fa.nOm + (job - NPAIR) * a.nlj0;
                for (int l = lane; l <= 2 * lim; l += 16) {
                    const int m0 = l - lim;
                    kaprow[rowbase + l] = cscale(plane[RPAD(m0 & (FSG_M - 1))], invM);
                }
            }
            __syncthreads();
        }
    }
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

struct LagFinishArgs {
    int nrows, nOm, nK, nLT, w1, ksplit;
    double* R; double* RJ; double* RT; double* RJT;
};

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
