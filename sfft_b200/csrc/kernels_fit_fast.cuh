// kernels_fit_fast.cuh -- the fit column pass on the register FFT engine (N0 a multiple of 256).
//
// Same mathematics and same outputs (kap, lam, nuJ) as fit_col_kernel in kernels_fit.cuh; the axis-0 frequencies are
// split k0 = V u + t with V = Vo * VI and slice length 256:
//   * outer fold (factor Vo): direct sums over the column read from HBM/L2 (Vo passes over the column per CTA);
//   * inner fold (factor VI <= 4): one register radix-VI DIF butterfly that yields VI sub-slices at once;
//   * 256-point slice transforms: 16 threads x 16 registers, radix 16 x 16, one shared-memory exchange each;
//   * pair products are formed in registers while loading the inverse transform's inputs; the two-real-in-one
//     trick packs the (real) auto-spectra |F_A|^2, |F_A'|^2 of two planes into one complex inverse transform;
//   * only the lags FillLS_* reads leave the transform; they are accumulated in shared memory per column.
// CTA = 256 threads = 16 transform groups; one CTA per SM (about 210 KB of shared memory).
#pragma once
#include "fft_regs.cuh"
#include "kernels_fit.cuh"

#define FCF_NT 256
#define FCF_M 256
#define FCF_PITCH 272
#define FCF_GROUPS 16

struct FastFitArgs {
    ColArgs c;                       // V = Vo * VI, M = 256
    int Vo, VI;
    int noff, ndg, njob;             // off-diagonal pairs, packed diagonal jobs, total jobs per sub-slice
    int pack_rounds;                 // 1: jobs of different sub-slices may share a round (njob >= 16)
    const cd* tabA;                  // engine table for the second radix-16 pass (240 entries)
    unsigned char offA[48], offB[48];
};

// ---- two-level DIF fold of one column into VI sub-slices of 256 for every plane ------------------------------------
template <typename TSt, int VI>
__device__ __forceinline__ void fold_two_level(const FastFitArgs& fa, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
                                               int k1, int to, cd* F)
{
    const ColArgs& a = fa.c;
    const int np = threadIdx.x;                  // n' in [0, 256)
    const int Mo = FCF_M * VI;
    const double inv0 = 1.0 / (double)a.N0;
    const int nplanes = a.Fij + 1;
    for (int jj = 0; jj <= a.nj; ++jj) {
        const bool isJ = (jj == a.nj);
        const TSt* col = isJ ? (gJ + (size_t)k1 * a.N0) : (gI + ((size_t)jj * a.NH + k1) * a.N0);
        const int ni = isJ ? 1 : (a.DK - jj + 1);
        cd z[VI][4];
#pragma unroll
        for (int q = 0; q < VI; ++q) {
            cd s0 = cmake(0, 0), s1 = s0, s2 = s0, s3 = s0;
            const int nq = np + FCF_M * q;
            for (int v = 0; v < fa.Vo; ++v) {
                const int r = nq + Mo * v;
                cd g = load_c(col + r);
                if (to != 0 && v != 0) g = cmul(g, a.tw0[((to * v) % fa.Vo) * Mo]);
                const double cx = (r + 1) * inv0;
                s0 = cadd(s0, g);
                if (ni > 1) { g = cscale(g, cx); s1 = cadd(s1, g); }
                if (ni > 2) { g = cscale(g, cx); s2 = cadd(s2, g); }
                if (ni > 3) { g = cscale(g, cx); s3 = cadd(s3, g); }
            }
            if (to != 0) {
                const cd wn = a.tw0[to * nq];
                s0 = cmul(s0, wn);
                if (ni > 1) s1 = cmul(s1, wn);
                if (ni > 2) s2 = cmul(s2, wn);
                if (ni > 3) s3 = cmul(s3, wn);
            }
            z[q][0] = s0; z[q][1] = s1; z[q][2] = s2; z[q][3] = s3;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < ni) {
                cd t[VI];
#pragma unroll
                for (int q = 0; q < VI; ++q) t[q] = z[q][i];
                if (VI > 1) bfly_r<VI>(t, -1.0);
                const int plane = isJ ? a.Fij : a.plane_of[i][jj];
#pragma unroll
                for (int s = 0; s < VI; ++s) {
                    cd val = t[s];
                    if (s > 0) val = cmul(val, a.tw0[s * np * fa.Vo]);
                    F[(size_t)(s * nplanes + plane) * FCF_PITCH + RPAD(np)] = val;
                }
            }
        }
    }
}

// smem (cd): F[VI*(Fij+1)*PITCH] | exch[16*PITCH] | acc[nacc] | mom[(nj+1)*MAXE] | red[16*MAXE] | tabA[240]
template <typename TSt, int VI>
__global__ void __launch_bounds__(FCF_NT) fit_col_fast_kernel(FastFitArgs fa, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
                                                              cd* __restrict__ kap, cd* __restrict__ lam, cd* __restrict__ nuJ)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ColArgs& a = fa.c;
    const int nplanes = a.Fij + 1;
    cd* F = reinterpret_cast<cd*>(smem_raw);
    cd* exch = F + (size_t)VI * nplanes * FCF_PITCH;
    const int nOm = a.npairs * a.nl0;
    const int nacc = nOm + a.Fij * a.nlj0;
    cd* acc = exch + (size_t)FCF_GROUPS * FCF_PITCH;
    cd* mom = acc + nacc;
    cd* red = mom + (a.nj + 1) * SFFTB_MAXE;
    cd* tabA = red + 16 * SFFTB_MAXE;
    const int tid = threadIdx.x;
    const int grp = tid >> 4, lane = tid & 15;
    GroupSync gs;
    gs.mask = 0xffffu << (16 * (grp & 1));
    gs.bar_id = 0; gs.count = 0;
    cd* myx = exch + (size_t)grp * FCF_PITCH;
    const double invN0 = 1.0 / (double)a.N0;

    for (int i = tid; i < 240; i += FCF_NT) tabA[i] = fa.tabA[i];

    for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x) {
        for (int idx = tid; idx < nacc; idx += FCF_NT) acc[idx] = cmake(0, 0);
        column_moments(a, gI, gJ, k1, mom, red);
        column_poly_terms(a, gI, k1, mom, lam, nuJ);

        for (int to = 0; to < fa.Vo; ++to) {
            fold_two_level<TSt, VI>(fa, gI, gJ, k1, to, F);
            __syncthreads();
            // ---- forward transforms, in place ----
            for (int p0 = 0; p0 < VI * nplanes; p0 += FCF_GROUPS) {
                const int P = p0 + grp;
                if (P < VI * nplanes) {
                    cd v[16];
                    cd* plane = F + (size_t)P * FCF_PITCH;
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = plane[RPAD(lane + 16 * q)];
                    reg_fft<FCF_M>(v, myx, lane, tabA, nullptr, nullptr, -1.0, gs);
#pragma unroll
                    for (int q = 0; q < 16; ++q) plane[RPAD(lane + 16 * q)] = v[q];
                }
            }
            __syncthreads();
            // ---- pair products -> inverse transforms -> lag accumulation ----
            const int total = VI * fa.njob;
            const int rps = (fa.njob + FCF_GROUPS - 1) / FCF_GROUPS;          // rounds per sub-slice when not packed
            const int nrounds = fa.pack_rounds ? (total + FCF_GROUPS - 1) / FCF_GROUPS : VI * rps;
            for (int rd = 0; rd < nrounds; ++rd) {
                int s, jb;
                if (fa.pack_rounds) { const int J = rd * FCF_GROUPS + grp; s = J / fa.njob; jb = J - s * fa.njob; if (J >= total) jb = -1; }
                else { s = rd / rps; jb = (rd - s * rps) * FCF_GROUPS + grp; if (jb >= fa.njob) jb = -1; }
                if (jb >= 0) {
                    const cd* Fs = F + (size_t)s * nplanes * FCF_PITCH;
                    int A, B, kind;                      // kind 0: off-diagonal, 1: packed diagonal, 2: theta
                    if (jb < fa.noff) { A = fa.offA[jb]; B = fa.offB[jb]; kind = 0; }
                    else if (jb < fa.noff + fa.ndg) { A = 2 * (jb - fa.noff); B = A + 1; kind = 1; }
                    else { A = jb - fa.noff - fa.ndg; B = a.Fij; kind = 2; }
                    cd v[16];
                    const cd* pa = Fs + (size_t)A * FCF_PITCH;
                    if (kind == 1) {
                        const bool hasB = B < a.Fij;
                        const cd* pb = Fs + (size_t)(hasB ? B : A) * FCF_PITCH;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const cd fa_ = pa[RPAD(lane + 16 * q)];
                            const cd fb_ = pb[RPAD(lane + 16 * q)];
                            v[q] = cmake(fa_.x * fa_.x + fa_.y * fa_.y, hasB ? (fb_.x * fb_.x + fb_.y * fb_.y) : 0.0);
                        }
                    } else {
                        const cd* pb = Fs + (size_t)B * FCF_PITCH;
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] = cmulcj(pa[RPAD(lane + 16 * q)], pb[RPAD(lane + 16 * q)]);
                    }
                    reg_fft<FCF_M>(v, myx, lane, tabA, nullptr, nullptr, +1.0, gs);
                    gs.sync<16>();
#pragma unroll
                    for (int q = 0; q < 16; ++q) myx[RPAD(lane + 16 * q)] = v[q];
                    gs.sync<16>();
                    const int t = fa.Vo * s + to;                    // global slice index: k0 = V u + t
                    const int lim = (kind == 2) ? a.w0 : 2 * a.w0;
                    const int nl = 2 * lim + 1;
                    for (int l = lane; l < nl; l += 16) {
                        const int m0 = l - lim;
                        const cd y = myx[RPAD(m0 & (FCF_M - 1))];
                        const cd w = cconj(a.tw0[imod(t * m0, a.N0)]);       // e^{+2 pi i t m0 / N0}
                        if (kind == 1) {
                            const cd y2 = cconj(myx[RPAD((-m0) & (FCF_M - 1))]);
                            const cd ya = cmake(0.5 * (y.x + y2.x), 0.5 * (y.y + y2.y));
                            const cd yb = cmake(0.5 * (y.y - y2.y), -0.5 * (y.x - y2.x));     // (y - y2) / (2 i)
                            const int ia = (A * a.Fij - (A * (A - 1)) / 2) * a.nl0 + l;
                            acc[ia] = cadd(acc[ia], cmul(ya, w));
                            if (B < a.Fij) {
                                const int ib = (B * a.Fij - (B * (B - 1)) / 2) * a.nl0 + l;
                                acc[ib] = cadd(acc[ib], cmul(yb, w));
                            }
                        } else if (kind == 0) {
                            const int ia = (A * a.Fij - (A * (A - 1)) / 2 + (B - A)) * a.nl0 + l;
                            acc[ia] = cadd(acc[ia], cmul(y, w));
                        } else {
                            const int ia = nOm + A * a.nlj0 + l;
                            acc[ia] = cadd(acc[ia], cmul(y, w));
                        }
                    }
                }
                __syncthreads();
            }
        }
        for (int idx = tid; idx < nacc; idx += FCF_NT) kap[(size_t)idx * a.NH + k1] = cscale(acc[idx], invN0);
        __syncthreads();
    }
}
