// kernels_row_h16.cuh -- forward row pass as R x 256: one thread-local radix-R pass, ONE group-wide exchange, then every half
// warp runs a 256-point transform on the half-warp engine (fft_h16.cuh: one more exchange, __syncwarp only).
//
// Same contract as row_fwd_v8_kernel (kernels_row_v8.cuh): real rows -> half spectra along axis 1 with cy(c)^j (or a column
// table of a general-basis plan) fused into the load (SpatialPoly + fft2, sfft/sfftcore/SFFTConfigure.py:112-145,
// SFFTSubtract.py:127-161), two real samples packed per complex point, output stored TRANSPOSED g[j][k1][r].
// H = N1 / 2 = 256 R complex points per row, R in {4, 8, 16} (N1 = 2048, 4096, 8192), T = H / 16 threads per row:
//   n = 256 a + b,  k = c + R d:   X[c + R d] = sum_b W_256^{b d} [ W_H^{b c} sum_a W_R^{a c} x[256 a + b] ]
//   pass A : thread t owns b = t + T i (i < 16 / R): radix-R butterflies over a, twiddles W_H^{b c} (the powers c = 1, 2, 4, 8
//            from a shared-memory table, the others by one or two multiplications), written to plane c of the row buffer;
//   pass B : half warp c transforms plane c (hfft256) -> X[c + R d] at position d of plane c.
// Against the 8-values engine (radix 8 / 8 / 8 / 4 on named barriers) a row crosses its 256-thread group barrier once instead
// of six times and moves through shared memory three times instead of four; RBI = 512 / T rows are in flight per CTA, so the
// transposed stores write RBI consecutive rows of a column: full 32-byte sectors for fp32 spectra at N1 = 4096.
// The plane pitch is 273 elements (= 1 mod 8): the untangle step reads k and H - k of all planes without bank conflicts.
#pragma once
#include "fft_h16.cuh"
#include "kernels_row_v8.cuh"

#define ROWH_NT 512
#define ROWH_PP 273
#ifndef ROWH_SETS
#define ROWH_SETS(R) 1
#endif

struct RowH16Args {
    int N0, N1, NH, H;
    const cd* tabA;          // half-warp engine powers: exp(-2 pi i r k / 256), [(r-1) 16 + k]
    const cd* twP;           // exp(-2 pi i r k / H), [(r-1) 256 + k], r = 1 .. R-1 (upload_engine_table(256, R))
    const cd* tw1;           // exp(-2 pi i e / N1)
    const double* vtab;      // general-basis plans: column tables instead of cy^j (or NULL)
};

// threads per CTA: 512 (four rows of 2048 points in flight).  fp64 spectra fill a 32-byte sector with TWO rows, so the apply-step
// pass could run 256-thread CTAs, two per SM, whose transform and store phases interleave (ROWH_NT64 256); measured slightly
// slower (2.22 against 2.20 ms per pair), so both storage types use 512
#ifndef ROWH_NT64
#define ROWH_NT64 512
#endif
template <typename TSt, int R> struct RowhCfg { static const int nt = (sizeof(TSt) == 16 && R <= 8) ? ROWH_NT64 : ROWH_NT; };
static inline int rowh_threads(int H, bool st64) { return (st64 && H <= 2048) ? ROWH_NT64 : ROWH_NT; }

static inline size_t rowh_smem_bytes(int H, int nt = ROWH_NT) {
    const int R = H / 256, T = H / 16, RBI = nt / T;          // rows in flight per CTA (all sets)
    int LR = 0;
    while ((1 << LR) < R) ++LR;
    return sizeof(cd) * ((size_t)RBI * (R * ROWH_PP + 4) + (size_t)LR * 256 + H / 2 + 1);
}

// SETS: the CTA works as SETS independent sets of 512 / SETS threads (own named barrier, own row groups), so that the untangle /
// store phase of one set runs under the transform phase of the other
template <typename TIn, typename TSt, int R, int SETS = ROWH_SETS(R)>
__global__ void __launch_bounds__(RowhCfg<TSt, R>::nt, ROWH_NT / RowhCfg<TSt, R>::nt) row_fwd_h16_kernel(RowH16Args a, const TIn* __restrict__ img, TSt* __restrict__ out, int nj)
{
    constexpr int H = 256 * R, T = H / 16, NB = 16 / R, NT = RowhCfg<TSt, R>::nt;
    constexpr int TS = NT / SETS, RBI = TS / T;          // threads per set, rows per set
    constexpr int LR = R == 4 ? 2 : (R == 8 ? 3 : 4);
    constexpr int ROWP = R * ROWH_PP + (sizeof(TSt) == 8 ? 4 : 2);     // row pitch: the LPC lanes of a column read different bank groups
    typedef typename In2<TIn>::type TIn2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* zbuf = reinterpret_cast<cd*>(smem_raw);              // [RBI][R planes][ROWH_PP]
    cd* twp = zbuf + (size_t)(NT / T) * ROWP;           // [LR][256]: W_H^{b 2^l}
    cd* tw1s = twp + LR * 256;                               // [H/2 + 1] untangle factors
    const int tid = threadIdx.x;
    for (int i = tid; i < LR * 256; i += NT) {
        const int l = i >> 8, b = i & 255;
        twp[i] = a.twP[(size_t)((1 << l) - 1) * 256 + b];
    }
    for (int i = tid; i <= H / 2; i += NT) tw1s[i] = a.tw1[i];
    const int grp = tid / T, t = tid - grp * T;               // row slot of the CTA, thread of the row
    const int set = tid / TS, ts = tid - set * TS, gs = grp - set * RBI;     // set, thread of the set, row of the set
    const int lane = tid & 31, half = lane >> 4, hl = lane & 15;
    const int csub = 2 * (t >> 5) + half;                    // plane / sub-transform of this half warp
    cd* zrow = zbuf + (size_t)grp * ROWP;
    H16Tw htw;
    h16_load(htw, a.tabA, hl);
    __syncthreads();
    const double inv1 = 1.0 / (double)a.N1;
    const int ngroups = (a.N0 + RBI - 1) / RBI;              // row groups of RBI rows; set s of CTA c takes groups c SETS + s + m gridDim SETS
    cd* zset = zbuf + (size_t)set * RBI * ROWP;
    auto set_sync = [&]() {
        if (SETS == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + set), "n"(TS) : "memory");
    };
    const int gb0 = blockIdx.x * SETS + set, gstep = gridDim.x * SETS;
    const bool aligned = (a.N0 % 2 == 0);

    // fp32 images: the 16 packed samples of a thread stay in registers for all planes j and the next row group is requested
    // under the last untangle step; fp64 images (64 registers) are re-read per plane instead (L2 hits)
    constexpr bool KEEP = sizeof(TIn) == 4;
    TIn2 x[16];
    auto load_row = [&](int gb) {
        const int r = gb * RBI + gs;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (r < a.N0) x[q] = *reinterpret_cast<const TIn2*>(img + (size_t)r * a.N1 + 2 * (t + q * T));
            else { x[q].x = 0; x[q].y = 0; }
        }
    };
    if (KEEP && gb0 < ngroups) load_row(gb0);
    for (int gb = gb0; gb < ngroups; gb += gstep) {
        const int r0 = gb * RBI;
        for (int j = 0; j < nj; ++j) {
            if (!KEEP) load_row(gb);
            cd v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int n = t + q * T;
                double x0 = (double)x[q].x, x1 = (double)x[q].y;
                if (a.vtab) {
                    const double2 vv = *reinterpret_cast<const double2*>(a.vtab + (size_t)j * a.N1 + 2 * n);
                    x0 *= vv.x; x1 *= vv.y;
                } else if (j > 0) {
                    const double c0 = (2 * n + 1) * inv1, c1 = (2 * n + 2) * inv1;
                    x0 *= (j == 1) ? c0 : (j == 2 ? c0 * c0 : c0 * c0 * c0);
                    x1 *= (j == 1) ? c1 : (j == 2 ? c1 * c1 : c1 * c1 * c1);
                }
                v[q] = cmake(x0, x1);
            }
            // ---- pass A: radix-R butterflies over a (element q = NB a + i belongs to butterfly i, b = t + T i) ----
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int b = t + T * i;
                cd y[R];
#pragma unroll
                for (int aa = 0; aa < R; ++aa) y[aa] = v[NB * aa + i];
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
            asm volatile("bar.sync %0, %1;" ::"r"(1 + SETS + grp), "n"(T) : "memory");
            // ---- pass B: 256-point transform of plane csub by this half warp ----
            cd* plane = zrow + csub * ROWH_PP;
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = plane[HPAD(hl + 16 * q)];
            __syncwarp();
            hfft256(v, plane, hl, htw, -1.0);
#pragma unroll
            for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];        // X[csub + R (hl + 16 q)]
            // the next row group's samples are requested now and arrive under the untangle step of the last plane
            if (KEEP && j == nj - 1 && gb + gstep < ngroups) load_row(gb + gstep);
            set_sync();
            // ---- untangle k and H - k together; Z[k] sits at plane k % R, position k / R.  A lane owns EPL consecutive rows of a
            // column (16 bytes of the transposed output), LPC adjacent lanes own the RBI rows of the same column, so every store
            // instruction writes whole 32-byte sectors (a lane per column would leave every sector half written per instruction)
            constexpr int EPL = 16 / (int)sizeof(TSt) < RBI ? 16 / (int)sizeof(TSt) : RBI, LPC = RBI / EPL;
            const int nvalid = min(RBI, a.N0 - r0);
            for (int idx = ts; idx < (H / 2 + 1) * LPC; idx += TS) {
                const int k = idx / LPC, p0 = (idx - k * LPC) * EPL;
                const cd w = tw1s[k];
                const int km = (H - k) & (H - 1);
                const int ia = (k & (R - 1)) * ROWH_PP + HPAD(k / R), ib = (km & (R - 1)) * ROWH_PP + HPAD(km / R);
                cd gk[EPL], gm[EPL];
#pragma unroll
                for (int p = 0; p < EPL; ++p) {
                    const cd A = zset[(size_t)(p0 + p) * ROWP + ia];
                    const cd B = zset[(size_t)(p0 + p) * ROWP + ib];
                    // G[k]   = 0.5 (A + conj B) - 0.5 i W^k (A - conj B)
                    // G[H-k] = 0.5 (B + conj A) + 0.5 i conj(W^k) (B - conj A)
                    const cd s = cmake(A.x + B.x, A.y - B.y), d = cmake(A.x - B.x, A.y + B.y);
                    const cd wd = cmul(w, d);
                    gk[p] = cmake(0.5 * (s.x + wd.y), 0.5 * (s.y - wd.x));
                    gm[p] = cmake(0.5 * (s.x - wd.y), 0.5 * (-s.y - wd.x));
                }
                const int nv = max(0, min(EPL, nvalid - p0));
                store_rows<TSt, EPL>(out + ((size_t)j * a.NH + k) * a.N0 + r0 + p0, gk, nv, aligned);
                if (k != H - k) store_rows<TSt, EPL>(out + ((size_t)j * a.NH + (H - k)) * a.N0 + r0 + p0, gm, nv, aligned);
            }
            set_sync();
        }
    }
}

// ---- H = 8192 (N1 = 16384): 32 x 256 ---------------------------------------------------------------------------------------------
// One row per CTA (512 threads).  Thread t owns b = t >> 1 and the sixteen a = 2 a' + p of parity p = t & 1: a thread-local
// radix-16 over a', then the radix-2 step X[c'] = E_0[c'] + W32^{c'} E_1[c'], X[c' + 16] = E_0[c'] - W32^{c'} E_1[c'] with the
// neighbouring lane (shuffles), twiddles W_H^{b c}, planes c = c' + 16 p; pass B and the untangle step as above (32 half warps,
// 32 planes).  The untangle factors are read from global memory (the planes take the shared memory).
// cos / sin of 2 pi c / 32, c = 0 .. 15 (compile-time constants of the unrolled radix-2 step)
__device__ constexpr double ROWH_W32C[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                                  0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0, -0.19509032201612826785,
                                  -0.38268343236508977173, -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                                  -0.92387953251128675613, -0.98078528040323044913};
__device__ constexpr double ROWH_W32S[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474, 0.70710678118654752440,
                                  0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913, 1.0, 0.98078528040323044913,
                                  0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                                  0.38268343236508977173, 0.19509032201612826785};
static inline size_t rowh32_smem_bytes() { return sizeof(cd) * ((size_t)32 * ROWH_PP + 4 + 5 * 256); }

template <typename TIn, typename TSt>
__global__ void __launch_bounds__(ROWH_NT, 1) row_fwd_h16x32_kernel(RowH16Args a, const TIn* __restrict__ img, TSt* __restrict__ out, int nj)
{
    constexpr int R = 32, H = 8192;
    typedef typename In2<TIn>::type TIn2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* zrow = reinterpret_cast<cd*>(smem_raw);              // [32 planes][ROWH_PP]
    cd* twp = zrow + (size_t)R * ROWH_PP + 4;                // [5][256]: W_H^{b 2^l}
    const int tid = threadIdx.x;
    for (int i = tid; i < 5 * 256; i += ROWH_NT) {
        const int l = i >> 8, b = i & 255;
        twp[i] = a.twP[(size_t)((1 << l) - 1) * 256 + b];
    }
    const int b = tid >> 1, par = tid & 1;
    const int lane = tid & 31, half = lane >> 4, hl = lane & 15;
    const int cpl = 2 * (tid >> 5) + half;                   // plane / sub-transform of this half warp
    H16Tw htw;
    h16_load(htw, a.tabA, hl);
    __syncthreads();
    const double inv1 = 1.0 / (double)a.N1;
    for (int r = blockIdx.x; r < a.N0; r += gridDim.x) {
        for (int j = 0; j < nj; ++j) {
            cd v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int n = 256 * (2 * q + par) + b;
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
                v[q] = cmake(x0, x1);
            }
            butterfly16(v, -1.0);                                  // E_par[c'], c' = 0 .. 15
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const double ox = __shfl_xor_sync(0xffffffffu, v[c].x, 1), oy = __shfl_xor_sync(0xffffffffu, v[c].y, 1);
                const cd e0 = par ? cmake(ox, oy) : v[c], e1 = par ? v[c] : cmake(ox, oy);
                const cd we = cmake(e1.x * ROWH_W32C[c] + e1.y * ROWH_W32S[c], e1.y * ROWH_W32C[c] - e1.x * ROWH_W32S[c]);     // e1 * exp(-2 pi i c / 32)
                v[c] = par ? csub(e0, we) : cadd(e0, we);          // X[c] (par = 0) or X[c + 16] (par = 1)
            }
            if (par) {
                const cd w16 = twp[4 * 256 + b];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = cmul(v[c], w16);
            }
            h16_twiddle(v, twp[b], twp[256 + b], twp[512 + b], twp[768 + b]);
#pragma unroll
            for (int c = 0; c < 16; ++c) zrow[(c + 16 * par) * ROWH_PP + HPAD(b)] = v[c];
            __syncthreads();
            cd* plane = zrow + cpl * ROWH_PP;
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = plane[HPAD(hl + 16 * q)];
            __syncwarp();
            hfft256(v, plane, hl, htw, -1.0);
#pragma unroll
            for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];        // X[cpl + 32 (hl + 16 q)]
            __syncthreads();
            for (int k = tid; k <= H / 2; k += ROWH_NT) {
                const cd w = a.tw1[k];
                const int km = (H - k) & (H - 1);
                const cd A = zrow[(k & (R - 1)) * ROWH_PP + HPAD(k / R)];
                const cd B = zrow[(km & (R - 1)) * ROWH_PP + HPAD(km / R)];
                const cd s = cmake(A.x + B.x, A.y - B.y), d = cmake(A.x - B.x, A.y + B.y);
                const cd wd = cmul(w, d);
                store_c(out + ((size_t)j * a.NH + k) * a.N0 + r, cmake(0.5 * (s.x + wd.y), 0.5 * (s.y - wd.x)));
                if (k != H - k) store_c(out + ((size_t)j * a.NH + (H - k)) * a.N0 + r, cmake(0.5 * (s.x - wd.y), 0.5 * (-s.y - wd.x)));
            }
            __syncthreads();
        }
    }
}
