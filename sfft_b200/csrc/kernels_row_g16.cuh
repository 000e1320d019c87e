// kernels_row_g16.cuh -- row passes for N1 = 512 R with R NOT a power of two (R in {3, 5, 6, 10, 12}: N1 = 1536, 2560, 3072,
// 5120, 6144 -- e.g. a 3072-pixel ZTF axis, the 6144 x 6144 pair of BASELINE config 3) on the half-warp engine.
//
// Same contract and the same R x 256 decomposition as kernels_row_h16.cuh (reference: SpatialPoly + fft2 / ifft2,
// sfft/sfftcore/SFFTConfigure.py:112-145, SFFTSubtract.py:127-161, 452-461): H = N1 / 2 = 256 R complex points per row,
//   forward : n = 256 a + b, k = c + R d:  X[c + R d] = sum_b W_256^{b d} [ W_H^{b c} sum_a W_R^{a c} x[256 a + b] ]
//             pass A = radix-R butterflies over a (thread-local, mixed radix 3 / 5 / 2x3 / 2x5 / 4x3) and the twiddles W_H^{b c},
//             pass B = one 256-point transform per plane c by a half warp (hfft256), then the real-pair untangle step;
//   inverse : the mirror image (untangle -> 256-point inverse transforms per plane -> conj twiddles -> radix-R over c).
// T = 16 R threads per row (R half warps), RBI rows per CTA; 256 butterflies of a row are spread over the T threads
// (ceil(16 / R) rounds, the last one partly idle).  Before this kernel these widths ran through the shared-memory Stockham
// kernels (kernels_row.cuh) at a quarter of the speed of the power-of-two widths.
#pragma once
#include "kernels_row_h16.cuh"

// ---- composite butterflies R = R1 R2 (Cooley-Tukey inside a thread, compile-time twiddles) ---------------------------------
#define CT_S3 0.86602540378443864676
#define CT_C36 0.80901699437494742410
#define CT_S36 0.58778525229247312917
#define CT_C72 0.30901699437494742410
#define CT_S72 0.95105651629515357212
__device__ constexpr double CT6_C[6] = {1.0, 0.5, -0.5, -1.0, -0.5, 0.5};
__device__ constexpr double CT6_S[6] = {0.0, CT_S3, CT_S3, 0.0, -CT_S3, -CT_S3};
__device__ constexpr double CT10_C[10] = {1.0, CT_C36, CT_C72, -CT_C72, -CT_C36, -1.0, -CT_C36, -CT_C72, CT_C72, CT_C36};
__device__ constexpr double CT10_S[10] = {0.0, CT_S36, CT_S72, CT_S72, CT_S36, 0.0, -CT_S36, -CT_S72, -CT_S72, -CT_S36};
__device__ constexpr double CT12_C[12] = {1.0, CT_S3, 0.5, 0.0, -0.5, -CT_S3, -1.0, -CT_S3, -0.5, 0.0, 0.5, CT_S3};
__device__ constexpr double CT12_S[12] = {0.0, 0.5, CT_S3, 1.0, CT_S3, 0.5, 0.0, -0.5, -CT_S3, -1.0, -CT_S3, -0.5};

// y * exp(sgn 2 pi i m / N); m is a compile-time constant after unrolling, so the special cases fold away
template <int N>
__device__ __forceinline__ cd ct_mul(cd y, int m, double sgn) {
    const double c = N == 6 ? CT6_C[m % 6] : (N == 10 ? CT10_C[m % 10] : CT12_C[m % 12]);
    const double s = N == 6 ? CT6_S[m % 6] : (N == 10 ? CT10_S[m % 10] : CT12_S[m % 12]);
    if (s == 0.0) return c > 0.0 ? y : cmake(-y.x, -y.y);
    if (c == 0.0) return cmuli(y, s > 0.0 ? sgn : -sgn);
    return cmul(y, cmake(c, sgn * s));
}

// input index n = R2 n1 + n2, output index k = k1 + R1 k2
template <int R1, int R2>
__device__ __forceinline__ void bfly_ct(cd* v, double sgn) {
    constexpr int N = R1 * R2;
    cd y[R2][R1];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
        cd t[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) t[n1] = v[R2 * n1 + n2];
        butterfly<R1>(t, sgn);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) y[n2][k1] = (n2 > 0 && k1 > 0) ? ct_mul<N>(t[k1], (n2 * k1) % N, sgn) : t[k1];
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
        cd t[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) t[n2] = y[n2][k1];
        butterfly<R2>(t, sgn);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
    }
}
template <> __device__ __forceinline__ void bfly_r<6>(cd* t, double sgn) { bfly_ct<3, 2>(t, sgn); }
template <> __device__ __forceinline__ void bfly_r<10>(cd* t, double sgn) { bfly_ct<5, 2>(t, sgn); }
template <> __device__ __forceinline__ void bfly_r<12>(cd* t, double sgn) { bfly_ct<4, 3>(t, sgn); }

template <int R> struct RowgCfg {
    static const int rbi = R <= 3 ? 8 : (R <= 6 ? 4 : 2);           // rows per CTA
    static const int T = 16 * R, nt = rbi * T;                       // threads per row, per CTA (384 / 320 / 384 / 320 / 384)
    static const int lr = R > 8 ? 4 : (R > 4 ? 3 : (R > 2 ? 2 : 1));  // stored twiddle powers W^(b 2^l), 2^l < R
    static const int rowp = R * ROWH_PP + 4;
};
static inline bool rowg_supported(int R) { return R == 3 || R == 5 || R == 6 || R == 10 || R == 12; }
static inline int rowg_rbi(int R) { return R <= 3 ? 8 : (R <= 6 ? 4 : 2); }
// vt: the plan has column tables (general-basis plans): one table row (N1 = 512 R doubles) is staged in shared memory
static inline size_t rowg_smem_bytes(int R, bool vt = false) {
    const int lr = R > 8 ? 4 : (R > 4 ? 3 : (R > 2 ? 2 : 1));
    return sizeof(cd) * ((size_t)rowg_rbi(R) * (R * ROWH_PP + 4) + (size_t)lr * 256 + 128 * R + 2) + (vt ? sizeof(double) * 512 * (size_t)R : 0);
}
// loads that stay where they are written (plain loads are sunk to their first use, i.e. behind the untangle step they should overlap)
__device__ __forceinline__ float2 rowg_ld_early(const float2* p) { float2 v; asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ double2 rowg_ld_early(const double2* p) { double2 v; asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ void rowg_cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// W_H^{b c}, c = 1 .. R-1, from the stored powers c = 1, 2, 4, 8 (at most two more products each)
template <int R>
__device__ __forceinline__ void rowg_twiddles(cd (&w)[R], const cd* twp, int b) {
    constexpr int LR = RowgCfg<R>::lr;
#pragma unroll
    for (int l = 0; l < LR; ++l) w[1 << l] = twp[l * 256 + b];
#pragma unroll
    for (int c = 3; c < R; ++c) {
        const int hi = c >= 8 ? 8 : (c >= 4 ? 4 : 2);
        if (c != hi) w[c] = cmul(w[hi], w[c - hi]);
    }
}

template <typename TIn, typename TSt, int R>
__global__ void __launch_bounds__(RowgCfg<R>::nt, 1) row_fwd_g16_kernel(RowH16Args a, const TIn* __restrict__ img, TSt* __restrict__ out, int nj)
{
    constexpr int H = 256 * R, T = RowgCfg<R>::T, RBI = RowgCfg<R>::rbi, NT = RowgCfg<R>::nt, LR = RowgCfg<R>::lr, ROWP = RowgCfg<R>::rowp;
    constexpr int NBF = (256 + T - 1) / T;
    typedef typename In2<TIn>::type TIn2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* zbuf = reinterpret_cast<cd*>(smem_raw);              // [RBI][R planes][ROWH_PP]
    cd* twp = zbuf + (size_t)RBI * ROWP;                     // [LR][256]: W_H^{b 2^l}
    cd* tw1s = twp + LR * 256;                               // [H/2 + 1] untangle factors
    double* vts = reinterpret_cast<double*>(tw1s + H / 2 + 2);     // [N1] column table of the current plane (general-basis plans only)
    const int tid = threadIdx.x;
    for (int i = tid; i < LR * 256; i += NT) {
        const int l = i >> 8, b = i & 255;
        twp[i] = a.twP[(size_t)((1 << l) - 1) * 256 + b];
    }
    for (int i = tid; i <= H / 2; i += NT) tw1s[i] = a.tw1[i];
    // the column table of plane j sits in shared memory; the row of the next plane is copied (cp.async) under pass B and the
    // untangle step of this one -- read from global memory inside pass A its L2 latency was exposed once per plane
    auto stage_vtab = [&](int j) {
        const double* src = a.vtab + (size_t)j * a.N1;
        for (int i = tid; i < H; i += NT) rowg_cp16(vts + 2 * i, src + 2 * i);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (a.vtab) { stage_vtab(0); asm volatile("cp.async.wait_group 0;" ::: "memory"); }
    const int grp = tid / T, t = tid - grp * T;              // row slot of the CTA, thread of the row
    const int hw = t >> 4, hl = tid & 15;                    // plane of this half warp (T is a multiple of 16)
    cd* zrow = zbuf + (size_t)grp * ROWP;
    H16Tw htw;
    h16_load(htw, a.tabA, hl);
    __syncthreads();
    const double inv1 = 1.0 / (double)a.N1;
    const int ngroups = (a.N0 + RBI - 1) / RBI;
    const bool aligned = (a.N0 % 2 == 0);
    // fp32 images: the packed samples of a thread stay in registers for all planes j, and the next row group is requested under the
    // last untangle step; fp64 images are re-read per plane (L2 hits)
    constexpr bool KEEP = sizeof(TIn) == 4;
    TIn2 xk[KEEP ? NBF * R : 1];
    auto load_group = [&](int gb) {
        const int r = gb * RBI + grp;
#pragma unroll
        for (int i = 0; i < NBF; ++i)
#pragma unroll
            for (int aa = 0; aa < R; ++aa) {
                const int b = t + T * i;
                if (b < 256 && r < a.N0) xk[(KEEP ? i * R + aa : 0)] = rowg_ld_early(reinterpret_cast<const TIn2*>(img + (size_t)r * a.N1 + 2 * (256 * aa + b)));
                else { xk[(KEEP ? i * R + aa : 0)].x = 0; xk[(KEEP ? i * R + aa : 0)].y = 0; }
            }
    };
    if (KEEP && (int)blockIdx.x < ngroups) load_group(blockIdx.x);
    for (int gb = blockIdx.x; gb < ngroups; gb += gridDim.x) {
        const int r0 = gb * RBI, r = r0 + grp;
        for (int j = 0; j < nj; ++j) {
            // ---- pass A: radix-R butterflies over a for b = t, t + T, ... ----
#pragma unroll
            for (int i = 0; i < NBF; ++i) {
                const int b = t + T * i;
                if (b < 256) {
                    cd y[R];
#pragma unroll
                    for (int aa = 0; aa < R; ++aa) {
                        const int n = 256 * aa + b;
                        TIn2 x;
                        if (KEEP) x = xk[KEEP ? i * R + aa : 0];
                        else if (r < a.N0) x = *reinterpret_cast<const TIn2*>(img + (size_t)r * a.N1 + 2 * n);
                        else { x.x = 0; x.y = 0; }
                        double x0 = (double)x.x, x1 = (double)x.y;
                        if (a.vtab) {
                            const double2 vv = *reinterpret_cast<const double2*>(vts + 2 * n);
                            x0 *= vv.x; x1 *= vv.y;
                        } else if (j > 0) {
                            const double c0 = (2 * n + 1) * inv1, c1 = (2 * n + 2) * inv1;
                            x0 *= (j == 1) ? c0 : (j == 2 ? c0 * c0 : c0 * c0 * c0);
                            x1 *= (j == 1) ? c1 : (j == 2 ? c1 * c1 : c1 * c1 * c1);
                        }
                        y[aa] = cmake(x0, x1);
                    }
                    bfly_r<R>(y, -1.0);
                    cd w[R];
                    rowg_twiddles<R>(w, twp, b);
                    zrow[HPAD(b)] = y[0];
#pragma unroll
                    for (int c = 1; c < R; ++c) zrow[c * ROWH_PP + HPAD(b)] = cmul(y[c], w[c]);
                }
            }
            __syncthreads();
            if (a.vtab && nj > 1) stage_vtab(j + 1 < nj ? j + 1 : 0);          // pass A of this plane has consumed the table row
            // ---- pass B: 256-point transform of plane hw by this half warp ----
            {
                cd* plane = zrow + hw * ROWH_PP;
                cd v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = plane[HPAD(hl + 16 * q)];
                __syncwarp();
                hfft256(v, plane, hl, htw, -1.0);
#pragma unroll
                for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];        // X[hw + R (hl + 16 q)]
            }
            if (KEEP && j == nj - 1 && gb + (int)gridDim.x < ngroups) load_group(gb + gridDim.x);
            __syncthreads();
            // ---- untangle k and H - k together; Z[k] sits at plane k % R, position k / R ----
            constexpr int EPL = 16 / (int)sizeof(TSt) < RBI ? 16 / (int)sizeof(TSt) : RBI, LPC = RBI / EPL;
            const int nvalid = min(RBI, a.N0 - r0);
            for (int idx = tid; idx < (H / 2 + 1) * LPC; idx += NT) {
                const int k = idx / LPC, p0 = (idx - k * LPC) * EPL;
                const cd w = tw1s[k];
                const int km = k == 0 ? 0 : H - k;
                const int ia = (k % R) * ROWH_PP + HPAD(k / R), ib = (km % R) * ROWH_PP + HPAD(km / R);
                cd gk[EPL], gm[EPL];
#pragma unroll
                for (int p = 0; p < EPL; ++p) {
                    const cd A = zbuf[(size_t)(p0 + p) * ROWP + ia];
                    const cd B = zbuf[(size_t)(p0 + p) * ROWP + ib];
                    const cd s = cmake(A.x + B.x, A.y - B.y), d = cmake(A.x - B.x, A.y + B.y);
                    const cd wd = cmul(w, d);
                    gk[p] = cmake(0.5 * (s.x + wd.y), 0.5 * (s.y - wd.x));
                    gm[p] = cmake(0.5 * (s.x - wd.y), 0.5 * (-s.y - wd.x));
                }
                const int nv = max(0, min(EPL, nvalid - p0));
                store_rows<TSt, EPL>(out + ((size_t)j * a.NH + k) * a.N0 + r0 + p0, gk, nv, aligned);
                if (k != H - k) store_rows<TSt, EPL>(out + ((size_t)j * a.NH + (H - k)) * a.N0 + r0 + p0, gm, nv, aligned);
            }
            if (a.vtab) asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
    }
}

// Inverse: transposed half spectra -> real rows, scaling, background polynomial (the contract of row_inv_fast_kernel).
template <typename TSt, typename TOut, int R>
__global__ void __launch_bounds__(RowgCfg<R>::nt, 1) row_inv_g16_kernel(RowInvFastArgs ia, const cd* __restrict__ twP, const TSt* __restrict__ spec,
                                                                        const double* __restrict__ bpq, TOut* __restrict__ out)
{
    constexpr int H = 256 * R, T = RowgCfg<R>::T, RBI = RowgCfg<R>::rbi, NT = RowgCfg<R>::nt, LR = RowgCfg<R>::lr, ROWP = RowgCfg<R>::rowp;
    constexpr int NBF = (256 + T - 1) / T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* zbuf = reinterpret_cast<cd*>(smem_raw);
    cd* twp = zbuf + (size_t)RBI * ROWP;
    const RowFastArgs& a = ia.r;
    const int tid = threadIdx.x;
    for (int i = tid; i < LR * 256; i += NT) {
        const int l = i >> 8, b = i & 255;
        twp[i] = twP[(size_t)((1 << l) - 1) * 256 + b];
    }
    const int grp = tid / T, t = tid - grp * T;
    const int hw = t >> 4, hl = tid & 15;
    cd* zrow = zbuf + (size_t)grp * ROWP;
    H16Tw htw;
    h16_load(htw, a.tabA, hl);
    const double inv0 = 1.0 / (double)a.N0, inv1 = 1.0 / (double)a.N1;
    const int ngroups = (a.N0 + RBI - 1) / RBI;
    for (int gb = blockIdx.x; gb < ngroups; gb += gridDim.x) {
        const int r0 = gb * RBI, r = r0 + grp;
        // gather RBI rows of the transposed spectrum and build the packed half-length spectrum Z = Ze + i Zo
        for (int idx = tid; idx < RBI * H; idx += NT) {
            const int k = idx / RBI, row = idx - k * RBI;
            const int rr = r0 + row;
            cd z = cmake(0.0, 0.0);
            if (rr < a.N0) {
                const cd gk = load_c(spec + (size_t)k * a.N0 + rr);
                const cd gm = cconj(load_c(spec + (size_t)(H - k) * a.N0 + rr));
                const cd ze = cscale(cadd(gk, gm), 0.5);
                const cd zo = cscale(cmul(csub(gk, gm), cconj(a.tw1[k])), 0.5);
                z = cmake(ze.x - zo.y, ze.y + zo.x);
            }
            zbuf[(size_t)row * ROWP + (k % R) * ROWH_PP + HPAD(k / R)] = z;
        }
        __syncthreads();
        {   // y_c[b] = sum_d Z[c + R d] W_256^{-b d} on plane c = hw
            cd* plane = zrow + hw * ROWH_PP;
            cd v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = plane[HPAD(hl + 16 * q)];
            __syncwarp();
            hfft256(v, plane, hl, htw, +1.0);
#pragma unroll
            for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];
        }
        __syncthreads();
        if (r < a.N0) {
            const double cx = (r + 1) * inv0;
            double cq[4] = {0.0, 0.0, 0.0, 0.0};
            if (bpq != nullptr) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k < ia.Fpq) {
                        const double tt = bpq[k] * ipow(cx, ia.p_of[k]);
                        const int qq = ia.q_of[k];
                        cq[0] += (qq == 0) ? tt : 0.0; cq[1] += (qq == 1) ? tt : 0.0;
                        cq[2] += (qq == 2) ? tt : 0.0; cq[3] += (qq == 3) ? tt : 0.0;
                    }
            }
#pragma unroll
            for (int i = 0; i < NBF; ++i) {
                const int b = t + T * i;
                if (b < 256) {
                    cd w[R], y[R];
                    rowg_twiddles<R>(w, twp, b);
                    y[0] = zrow[HPAD(b)];
#pragma unroll
                    for (int c = 1; c < R; ++c) y[c] = cmulcj(w[c], zrow[c * ROWH_PP + HPAD(b)]);
                    bfly_r<R>(y, +1.0);
#pragma unroll
                    for (int aa = 0; aa < R; ++aa) {
                        const int n = 256 * aa + b;
                        const double cy0 = (2 * n + 1) * inv1, cy1 = (2 * n + 2) * inv1;
                        const double x0 = fma(y[aa].x, ia.scale, -fma(fma(fma(cq[3], cy0, cq[2]), cy0, cq[1]), cy0, cq[0]));
                        const double x1 = fma(y[aa].y, ia.scale, -fma(fma(fma(cq[3], cy1, cq[2]), cy1, cq[1]), cy1, cq[0]));
                        store2(out + (size_t)r * a.N1 + 2 * n, x0, x1);
                    }
                }
            }
        }
        __syncthreads();
    }
}
