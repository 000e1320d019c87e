// kernels_row_blu.cuh -- row passes for ANY even image width N1 <= 8192 that is neither a power of two nor 512 R (R in {3, 5, 6, 10,
// 12}): a 4088-pixel Roman row, a trimmed 4094-pixel DECam row, the 4072 / 4000-pixel LSST axes, a 3080-pixel ZTF axis ...
//
// Same contract as the other row kernels (reference: SpatialPoly + fft2 / ifft2 with pyFFTW / cuFFT, which take any size:
// sfft/sfftcore/SFFTConfigure.py:112-145, SFFTSubtract.py:127-161, 452-461).  The length-H transform (H = N1 / 2 packed complex
// points, any H <= 4096) runs as a chirp-z (Bluestein) convolution of length M = 256 R >= 2 H (R = 4, 8, 16; 32 at the end of this
// file) on the half-warp engine:
//   Z[k] = c[k] sum_n (z[n] c[n]) conj(c)[k - n],  c[n] = exp(-pi i n^2 / H)
//        = c[k] IFFT_M( FFT_M(z c, zero padded) . B )[k],   B = FFT_M(conj(c) wrapped) / M   (table, once per plan)
// Both M-point transforms use the R x 256 decomposition of kernels_row_h16.cuh, and they share the middle: after the forward
// 256-point transform of plane c a half warp multiplies by its slice of B (stored in the plane layout) and runs the inverse 256-point
// transform straight away -- the spectrum of length M never leaves the registers.  Inputs n >= H are zero and outputs k >= H are
// not needed: a thread owns the eight positions n = 256 a + b, a < R / 2, on both sides (with the same chirp factors).
// Before this kernel these widths ran through the shared-memory Stockham kernels (kernels_row.cuh: 4.2 ms per four-plane pass at 4088^2
// against 0.30 ms at 4096^2).
#pragma once
#include "kernels_row_g16.cuh"

#define BLU_NT 512
struct RowBluArgs {
    int N0, N1, NH, H;
    const cd* tabA;          // half-warp engine powers exp(-2 pi i r k / 256)
    const cd* twP;           // exp(-2 pi i r k / M), [(r-1) 256 + k], r = 1 .. R-1 (upload_engine_table(256, R))
    const cd* tw1;           // exp(-2 pi i e / N1)
    const cd* chirp;         // c[n], n < H
    const cd* Bp;            // B in the plane layout: Bp[c * 256 + d] = B[c + R d]
    const double* vtab;
};
static inline int blu_radix(int H) { return H <= 512 ? 4 : (H <= 1024 ? 8 : (H <= 2048 ? 16 : (H <= 4096 ? 32 : 0))); }
static inline size_t blu32_smem_bytes();
static inline size_t blu_smem_bytes(int R) {
    if (R == 32) return blu32_smem_bytes();
    const int LR = R == 4 ? 2 : (R == 8 ? 3 : 4), RBI = BLU_NT / (16 * R);
    return sizeof(cd) * ((size_t)RBI * (R * ROWH_PP + 4) + (size_t)LR * 256 + 64 * R + 2);
}

// M-point circular convolution with the chirp filter for one row (T = 16 R threads, all of the CTA in lockstep: __syncthreads).
// in / out: z[i * (R / 2) + a] belongs to position n = 256 a + (t + T i), a < R / 2.
template <int R>
__device__ __forceinline__ void blu_core(cd (&z)[8], cd* zrow, const cd* twp, const cd* __restrict__ Bp, const H16Tw& htw, int t, int hl)
{
    constexpr int T = 16 * R, NB = 16 / R, HI = R / 2, LR = R == 4 ? 2 : (R == 8 ? 3 : 4);
    // ---- forward pass A: radix-R butterflies over a (the upper half of the inputs is the zero padding) ----
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + T * i;
        cd y[R];
#pragma unroll
        for (int a = 0; a < R; ++a) y[a] = a < HI ? z[i * HI + a] : cmake(0.0, 0.0);
        bfly_r<R>(y, -1.0);
        cd w[R];
#pragma unroll
        for (int l = 0; l < LR; ++l) w[1 << l] = twp[l * 256 + b];
#pragma unroll
        for (int c = 3; c < R; ++c) {
            const int hi = c >= 8 ? 8 : (c >= 4 ? 4 : 2);
            if (c != hi) w[c] = cmul(w[hi], w[c - hi]);
        }
        zrow[HPAD(b)] = y[0];
#pragma unroll
        for (int c = 1; c < R; ++c) zrow[c * ROWH_PP + HPAD(b)] = cmul(y[c], w[c]);
    }
    __syncthreads();
    // ---- plane c = half warp: forward 256-point transform, filter, inverse 256-point transform ----
    {
        const int c = t >> 4;
        cd* plane = zrow + c * ROWH_PP;
        cd v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = plane[HPAD(hl + 16 * q)];
        __syncwarp();
        hfft256(v, plane, hl, htw, -1.0);                    // Y[c + R d], d = hl + 16 q
        const cd* bp = Bp + c * 256 + hl;
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], bp[16 * q]);
        __syncwarp();
        hfft256(v, plane, hl, htw, +1.0);                    // sum_d Y'[c + R d] W_256^{-b d}, b = hl + 16 q
#pragma unroll
        for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];
    }
    __syncthreads();
    // ---- inverse pass A: conjugate twiddles, radix-R over c; only the outputs n < M / 2 are kept ----
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + T * i;
        cd w[R], y[R];
#pragma unroll
        for (int l = 0; l < LR; ++l) w[1 << l] = twp[l * 256 + b];
#pragma unroll
        for (int c = 3; c < R; ++c) {
            const int hi = c >= 8 ? 8 : (c >= 4 ? 4 : 2);
            if (c != hi) w[c] = cmul(w[hi], w[c - hi]);
        }
        y[0] = zrow[HPAD(b)];
#pragma unroll
        for (int c = 1; c < R; ++c) y[c] = cmulcj(w[c], zrow[c * ROWH_PP + HPAD(b)]);
        bfly_r<R>(y, +1.0);
#pragma unroll
        for (int a = 0; a < HI; ++a) z[i * HI + a] = y[a];
    }
}

template <typename TIn, typename TSt, int R>
__global__ void __launch_bounds__(BLU_NT, 1) row_fwd_blu_kernel(RowBluArgs a, const TIn* __restrict__ img, TSt* __restrict__ out, int nj)
{
    constexpr int T = 16 * R, RBI = BLU_NT / T, NB = 16 / R, HI = R / 2, LR = R == 4 ? 2 : (R == 8 ? 3 : 4), ROWP = R * ROWH_PP + 4;
    typedef typename In2<TIn>::type TIn2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* zbuf = reinterpret_cast<cd*>(smem_raw);              // [RBI][R planes][ROWH_PP]; afterwards the row's spectrum Z[0 .. H)
    cd* twp = zbuf + (size_t)RBI * ROWP;                     // [LR][256]: W_M^{b 2^l}
    cd* tw1s = twp + LR * 256;                               // [H/2 + 1] untangle factors
    const int tid = threadIdx.x, H = a.H;
    for (int i = tid; i < LR * 256; i += BLU_NT) {
        const int l = i >> 8, b = i & 255;
        twp[i] = a.twP[(size_t)((1 << l) - 1) * 256 + b];
    }
    for (int i = tid; i <= H / 2; i += BLU_NT) tw1s[i] = a.tw1[i];
    const int grp = tid / T, t = tid - grp * T, hl = tid & 15;
    cd* zrow = zbuf + (size_t)grp * ROWP;
    H16Tw htw;
    h16_load(htw, a.tabA, hl);
    __syncthreads();
    const double inv1 = 1.0 / (double)a.N1;
    const int ngroups = (a.N0 + RBI - 1) / RBI;
    const bool aligned = (a.N0 % 2 == 0);
    for (int gb = blockIdx.x; gb < ngroups; gb += gridDim.x) {
        const int r0 = gb * RBI, r = r0 + grp;
        for (int j = 0; j < nj; ++j) {
            cd z[8];
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int aa = 0; aa < HI; ++aa) {
                    const int n = 256 * aa + t + T * i;
                    cd v = cmake(0.0, 0.0);
                    if (n < H && r < a.N0) {
                        const TIn2 x = *reinterpret_cast<const TIn2*>(img + (size_t)r * a.N1 + 2 * n);
                        double x0 = (double)x.x, x1 = (double)x.y;
                        if (a.vtab) {
                            const double2 vv = *reinterpret_cast<const double2*>(a.vtab + (size_t)j * a.N1 + 2 * n);
                            x0 *= vv.x; x1 *= vv.y;
                        } else if (j > 0) {
                            const double c0 = (2 * n + 1) * inv1, c1 = (2 * n + 2) * inv1;
                            x0 *= (j == 1) ? c0 : (j == 2 ? c0 * c0 : c0 * c0 * c0);
                            x1 *= (j == 1) ? c1 : (j == 2 ? c1 * c1 : c1 * c1 * c1);
                        }
                        v = cmul(cmake(x0, x1), a.chirp[n]);
                    }
                    z[i * HI + aa] = v;
                }
            blu_core<R>(z, zrow, twp, a.Bp, htw, t, hl);
            __syncthreads();                                 // every plane has been read: the row buffer now takes Z[n], n < H
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int aa = 0; aa < HI; ++aa) {
                    const int n = 256 * aa + t + T * i;
                    if (n < H) zrow[n] = cmul(z[i * HI + aa], a.chirp[n]);
                }
            __syncthreads();
            // ---- untangle k and H - k together (as in row_fwd_h16_kernel: LPC lanes share the RBI rows of a column) ----
            constexpr int EPL = 16 / (int)sizeof(TSt) < RBI ? 16 / (int)sizeof(TSt) : RBI, LPC = RBI / EPL;
            const int nvalid = min(RBI, a.N0 - r0);
            for (int idx = tid; idx < (H / 2 + 1) * LPC; idx += BLU_NT) {
                const int k = idx / LPC, p0 = (idx - k * LPC) * EPL;
                const cd w = tw1s[k];
                const int km = k == 0 ? 0 : H - k;
                cd gk[EPL], gm[EPL];
#pragma unroll
                for (int p = 0; p < EPL; ++p) {
                    const cd A = zbuf[(size_t)(p0 + p) * ROWP + k];
                    const cd B = zbuf[(size_t)(p0 + p) * ROWP + km];
                    const cd s = cmake(A.x + B.x, A.y - B.y), d = cmake(A.x - B.x, A.y + B.y);
                    const cd wd = cmul(w, d);
                    gk[p] = cmake(0.5 * (s.x + wd.y), 0.5 * (s.y - wd.x));
                    gm[p] = cmake(0.5 * (s.x - wd.y), 0.5 * (-s.y - wd.x));
                }
                const int nv = max(0, min(EPL, nvalid - p0));
                store_rows<TSt, EPL>(out + ((size_t)j * a.NH + k) * a.N0 + r0 + p0, gk, nv, aligned);
                if (k != H - k) store_rows<TSt, EPL>(out + ((size_t)j * a.NH + (H - k)) * a.N0 + r0 + p0, gm, nv, aligned);
            }
            __syncthreads();
        }
    }
}

// Inverse: transposed half spectra -> real rows, scaling, background polynomial.  z[n] = conj(DFT_H(conj Z))[n].
template <typename TSt, typename TOut, int R>
__global__ void __launch_bounds__(BLU_NT, 1) row_inv_blu_kernel(RowBluArgs a, RowInvFastArgs ia, const TSt* __restrict__ spec,
                                                                const double* __restrict__ bpq, TOut* __restrict__ out)
{
    constexpr int T = 16 * R, RBI = BLU_NT / T, NB = 16 / R, HI = R / 2, LR = R == 4 ? 2 : (R == 8 ? 3 : 4), ROWP = R * ROWH_PP + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* zbuf = reinterpret_cast<cd*>(smem_raw);
    cd* twp = zbuf + (size_t)RBI * ROWP;
    const int tid = threadIdx.x, H = a.H;
    for (int i = tid; i < LR * 256; i += BLU_NT) {
        const int l = i >> 8, b = i & 255;
        twp[i] = a.twP[(size_t)((1 << l) - 1) * 256 + b];
    }
    const int grp = tid / T, t = tid - grp * T, hl = tid & 15;
    cd* zrow = zbuf + (size_t)grp * ROWP;
    H16Tw htw;
    h16_load(htw, a.tabA, hl);
    const double inv0 = 1.0 / (double)a.N0, inv1 = 1.0 / (double)a.N1;
    const int ngroups = (a.N0 + RBI - 1) / RBI;
    for (int gb = blockIdx.x; gb < ngroups; gb += gridDim.x) {
        const int r0 = gb * RBI, r = r0 + grp;
        // gather RBI rows of the transposed spectrum: Z = Ze + i Zo; the row buffer takes conj(Z[k]) c[k]
        for (int idx = tid; idx < RBI * H; idx += BLU_NT) {
            const int k = idx / RBI, row = idx - k * RBI;
            const int rr = r0 + row;
            cd z = cmake(0.0, 0.0);
            if (rr < a.N0) {
                const cd gk = load_c(spec + (size_t)k * a.N0 + rr);
                const cd gm = cconj(load_c(spec + (size_t)(H - k) * a.N0 + rr));
                const cd ze = cscale(cadd(gk, gm), 0.5);
                const cd zo = cscale(cmul(csub(gk, gm), cconj(a.tw1[k])), 0.5);
                z = cmul(cmake(ze.x - zo.y, -(ze.y + zo.x)), a.chirp[k]);
            }
            zbuf[(size_t)row * ROWP + k] = z;
        }
        __syncthreads();
        cd z[8];
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int aa = 0; aa < HI; ++aa) {
                const int n = 256 * aa + t + T * i;
                z[i * HI + aa] = n < H ? zrow[n] : cmake(0.0, 0.0);
            }
        __syncthreads();
        blu_core<R>(z, zrow, twp, a.Bp, htw, t, hl);
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
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int aa = 0; aa < HI; ++aa) {
                    const int n = 256 * aa + t + T * i;
                    if (n < H) {
                        const cd d = cmul(z[i * HI + aa], a.chirp[n]);             // DFT_H(conj Z)[n]; z[n] = conj(d)
                        const double cy0 = (2 * n + 1) * inv1, cy1 = (2 * n + 2) * inv1;
                        const double x0 = fma(d.x, ia.scale, -fma(fma(fma(cq[3], cy0, cq[2]), cy0, cq[1]), cy0, cq[0]));
                        const double x1 = fma(-d.y, ia.scale, -fma(fma(fma(cq[3], cy1, cq[2]), cy1, cq[1]), cy1, cq[0]));
                        store2(out + (size_t)r * a.N1 + 2 * n, x0, x1);
                    }
                }
        }
        __syncthreads();
    }
}

// ---- M = 8192 (even widths in (4096, 8192] outside the two direct families, e.g. the 4176-pixel HSC axis): 32 x 256 ----------------
// One row per CTA of 256 threads; thread b owns the positions n = 256 a + b, a < 16 (the upper half of the 32 inputs of its radix-32
// butterfly is the zero padding).  With half of the inputs zero the radix-32 butterfly is two radix-16 butterflies on the same data:
//   X[2 c'] = DFT16(x)[c'],  X[2 c' + 1] = DFT16(x_a W32^a)[c'];  the inverse keeps the outputs a < 16 only:
//   y[a] = IDFT16(u_even)[a] + W32^{-a} IDFT16(u_odd)[a].
// The 32 planes are transformed / filtered / transformed back by the 16 half warps in two rounds.
#define BLU32_NT 256
static inline size_t blu32_smem_bytes() { return sizeof(cd) * ((size_t)32 * ROWH_PP + 4 + 5 * 256 + 2050); }

// twiddles W_M^{b (2 c')} (par = 0) or W_M^{b (2 c' + 1)} (par = 1), c' < 16, applied to v; sgn = -1: as stored, +1: conjugated
__device__ __forceinline__ void blu32_twiddle(cd (&v)[16], const cd* twp, int b, int par, double sgn) {
    const cd w2 = twp[256 + b], w4 = twp[512 + b], w8 = twp[768 + b], w16 = twp[1024 + b];
    h16_twiddle(v, cmake(w2.x, -sgn * w2.y), cmake(w4.x, -sgn * w4.y), cmake(w8.x, -sgn * w8.y), cmake(w16.x, -sgn * w16.y));
    if (par) {
        const cd w1 = twp[b], w = cmake(w1.x, -sgn * w1.y);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = cmul(v[c], w);
    }
}

// in / out: z[a] belongs to position n = 256 a + b, a < 16.  All 256 threads of the CTA.
__device__ __forceinline__ void blu_core32(cd (&z)[16], cd* zrow, const cd* twp, const cd* __restrict__ Bp, const H16Tw& htw, int b, int hl)
{
    {   // even planes
        cd v[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) v[a] = z[a];
        butterfly16(v, -1.0);
        blu32_twiddle(v, twp, b, 0, -1.0);
#pragma unroll
        for (int c = 0; c < 16; ++c) zrow[(2 * c) * ROWH_PP + HPAD(b)] = v[c];
    }
    {   // odd planes: inputs times W32^a
#pragma unroll
        for (int a = 1; a < 16; ++a) z[a] = cmake(z[a].x * ROWH_W32C[a] + z[a].y * ROWH_W32S[a], z[a].y * ROWH_W32C[a] - z[a].x * ROWH_W32S[a]);
        butterfly16(z, -1.0);
        blu32_twiddle(z, twp, b, 1, -1.0);
#pragma unroll
        for (int c = 0; c < 16; ++c) zrow[(2 * c + 1) * ROWH_PP + HPAD(b)] = z[c];
    }
    __syncthreads();
#pragma unroll 1
    for (int rnd = 0; rnd < 2; ++rnd) {
        const int c = (b >> 4) + 16 * rnd;
        cd* plane = zrow + c * ROWH_PP;
        cd v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = plane[HPAD(hl + 16 * q)];
        __syncwarp();
        hfft256(v, plane, hl, htw, -1.0);
        const cd* bp = Bp + c * 256 + hl;
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], bp[16 * q]);
        __syncwarp();
        hfft256(v, plane, hl, htw, +1.0);
#pragma unroll
        for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];
    }
    __syncthreads();
    {
        cd ye[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) ye[c] = zrow[(2 * c) * ROWH_PP + HPAD(b)];
        blu32_twiddle(ye, twp, b, 0, +1.0);
        butterfly16(ye, +1.0);
#pragma unroll
        for (int c = 0; c < 16; ++c) z[c] = zrow[(2 * c + 1) * ROWH_PP + HPAD(b)];
        blu32_twiddle(z, twp, b, 1, +1.0);
        butterfly16(z, +1.0);
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            const cd o = a == 0 ? z[0] : cmake(z[a].x * ROWH_W32C[a] - z[a].y * ROWH_W32S[a], z[a].y * ROWH_W32C[a] + z[a].x * ROWH_W32S[a]);   // z[a] * exp(+2 pi i a / 32)
            z[a] = cadd(ye[a], o);
        }
    }
}

template <typename TIn, typename TSt>
__global__ void __launch_bounds__(BLU32_NT, 1) row_fwd_blu32_kernel(RowBluArgs a, const TIn* __restrict__ img, TSt* __restrict__ out, int nj)
{
    typedef typename In2<TIn>::type TIn2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* zrow = reinterpret_cast<cd*>(smem_raw);              // [32 planes][ROWH_PP]; afterwards the row's spectrum Z[0 .. H)
    cd* twp = zrow + (size_t)32 * ROWH_PP + 4;               // [5][256]: W_M^{b 2^l}
    cd* tw1s = twp + 5 * 256;                                // [H/2 + 1]
    const int tid = threadIdx.x, H = a.H, hl = tid & 15;
    for (int i = tid; i < 5 * 256; i += BLU32_NT) {
        const int l = i >> 8, b = i & 255;
        twp[i] = a.twP[(size_t)((1 << l) - 1) * 256 + b];
    }
    for (int i = tid; i <= H / 2; i += BLU32_NT) tw1s[i] = a.tw1[i];
    H16Tw htw;
    h16_load(htw, a.tabA, hl);
    __syncthreads();
    const double inv1 = 1.0 / (double)a.N1;
    for (int r = blockIdx.x; r < a.N0; r += gridDim.x) {
        for (int j = 0; j < nj; ++j) {
            cd z[16];
#pragma unroll
            for (int aa = 0; aa < 16; ++aa) {
                const int n = 256 * aa + tid;
                cd v = cmake(0.0, 0.0);
                if (n < H) {
                    const TIn2 x = *reinterpret_cast<const TIn2*>(img + (size_t)r * a.N1 + 2 * n);
                    double x0 = (double)x.x, x1 = (double)x.y;
                    if (a.vtab) {
                        const double2 vv = *reinterpret_cast<const double2*>(a.vtab + (size_t)j * a.N1 + 2 * n);
                        x0 *= vv.x; x1 *= vv.y;
                    } else if (j > 0) {
                        const double c0 = (2 * n + 1) * inv1, c1 = (2 * n + 2) * inv1;
                        x0 *= (j == 1) ? c0 : (j == 2 ? c0 * c0 : c0 * c0 * c0);
                        x1 *= (j == 1) ? c1 : (j == 2 ? c1 * c1 : c1 * c1 * c1);
                    }
                    v = cmul(cmake(x0, x1), a.chirp[n]);
                }
                z[aa] = v;
            }
            blu_core32(z, zrow, twp, a.Bp, htw, tid, hl);
            __syncthreads();
#pragma unroll
            for (int aa = 0; aa < 16; ++aa) {
                const int n = 256 * aa + tid;
                if (n < H) zrow[n] = cmul(z[aa], a.chirp[n]);
            }
            __syncthreads();
            for (int k = tid; k <= H / 2; k += BLU32_NT) {
                const cd w = tw1s[k];
                const int km = k == 0 ? 0 : H - k;
                const cd A = zrow[k], B = zrow[km];
                const cd s = cmake(A.x + B.x, A.y - B.y), d = cmake(A.x - B.x, A.y + B.y);
                const cd wd = cmul(w, d);
                store_c(out + ((size_t)j * a.NH + k) * a.N0 + r, cmake(0.5 * (s.x + wd.y), 0.5 * (s.y - wd.x)));
                if (k != H - k) store_c(out + ((size_t)j * a.NH + (H - k)) * a.N0 + r, cmake(0.5 * (s.x - wd.y), 0.5 * (-s.y - wd.x)));
            }
            __syncthreads();
        }
    }
}

template <typename TSt, typename TOut>
__global__ void __launch_bounds__(BLU32_NT, 1) row_inv_blu32_kernel(RowBluArgs a, RowInvFastArgs ia, const TSt* __restrict__ spec,
                                                                    const double* __restrict__ bpq, TOut* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* zrow = reinterpret_cast<cd*>(smem_raw);
    cd* twp = zrow + (size_t)32 * ROWH_PP + 4;
    const int tid = threadIdx.x, H = a.H, hl = tid & 15;
    for (int i = tid; i < 5 * 256; i += BLU32_NT) {
        const int l = i >> 8, b = i & 255;
        twp[i] = a.twP[(size_t)((1 << l) - 1) * 256 + b];
    }
    H16Tw htw;
    h16_load(htw, a.tabA, hl);
    const double inv0 = 1.0 / (double)a.N0, inv1 = 1.0 / (double)a.N1;
    for (int r = blockIdx.x; r < a.N0; r += gridDim.x) {
        for (int k = tid; k < H; k += BLU32_NT) {
            const cd gk = load_c(spec + (size_t)k * a.N0 + r);
            const cd gm = cconj(load_c(spec + (size_t)(H - k) * a.N0 + r));
            const cd ze = cscale(cadd(gk, gm), 0.5);
            const cd zo = cscale(cmul(csub(gk, gm), cconj(a.tw1[k])), 0.5);
            zrow[k] = cmul(cmake(ze.x - zo.y, -(ze.y + zo.x)), a.chirp[k]);
        }
        __syncthreads();
        cd z[16];
#pragma unroll
        for (int aa = 0; aa < 16; ++aa) {
            const int n = 256 * aa + tid;
            z[aa] = n < H ? zrow[n] : cmake(0.0, 0.0);
        }
        __syncthreads();
        blu_core32(z, zrow, twp, a.Bp, htw, tid, hl);
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
        for (int aa = 0; aa < 16; ++aa) {
            const int n = 256 * aa + tid;
            if (n < H) {
                const cd d = cmul(z[aa], a.chirp[n]);
                const double cy0 = (2 * n + 1) * inv1, cy1 = (2 * n + 2) * inv1;
                const double x0 = fma(d.x, ia.scale, -fma(fma(fma(cq[3], cy0, cq[2]), cy0, cq[1]), cy0, cq[0]));
                const double x1 = fma(-d.y, ia.scale, -fma(fma(fma(cq[3], cy1, cq[2]), cy1, cq[1]), cy1, cq[0]));
                store2(out + (size_t)r * a.N1 + 2 * n, x0, x1);
            }
        }
        __syncthreads();
    }
}
