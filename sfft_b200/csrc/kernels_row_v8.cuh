// kernels_row_v8.cuh -- forward row pass on the 8-values-per-thread register FFT engine (even N1, N1/2 in {256..2048}).
//
// Same contract as row_fwd_kernel (kernels_row.cuh): real rows -> half spectra along axis 1 with cy(c)^j fused
// into the load (SpatialPoly + fft2, sfft/sfftcore/SFFTConfigure.py:112-145, SFFTSubtract.py:127-161), two real
// samples packed per complex point, output stored TRANSPOSED g[j][k1][r].
// A CTA of 512 threads transforms RBI = 512 / (H/8) rows at a time and walks `nit` consecutive row groups, so that
// one CTA writes RBI * nit consecutive rows of every column (full 32-byte sectors within a few microseconds).  The
// image row is read once (vector loads, prefetched one group ahead) and reused for all powers j; the untangle step
// handles k and H - k together (one twiddle, two shared-memory reads per pair of outputs) and writes all RBI rows
// of a column with 16-byte stores.
#pragma once
#include "fft_vpt.cuh"
#include "kernels_row_fast.cuh"

#define ROWV_NT 512

struct RowV8Args {
    int N0, N1, NH, H;
    int nit;                 // row groups per CTA
    VTabs tabs;
    const cd* tw1;           // exp(-2 pi i e / N1)
    const double* vtab;      // general-basis plans: column tables instead of cy^j (or NULL)
};

template <typename TIn> struct In2;
template <> struct In2<float> { typedef float2 type; };
template <> struct In2<double> { typedef double2 type; };

template <typename TSt, int RBI>
__device__ __forceinline__ void store_rows(TSt* dst, const cd* g, int nvalid, bool aligned) {
    if (sizeof(TSt) == 8 && aligned && nvalid == RBI && RBI >= 2) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int p = 0; p < RBI / 2; ++p)
            d4[p] = make_float4((float)g[2 * p].x, (float)g[2 * p].y, (float)g[2 * p + 1].x, (float)g[2 * p + 1].y);
    } else {
#pragma unroll
        for (int p = 0; p < RBI; ++p)
            if (p < nvalid) store_c(dst + p, g[p]);
    }
}

template <typename TIn, typename TSt, int H>
__global__ void __launch_bounds__(ROWV_NT, 1) row_fwd_v8_kernel(RowV8Args a, const TIn* __restrict__ img, TSt* __restrict__ out, int nj)
{
    constexpr int T = H / 8, RBI = ROWV_NT / T, PITCH = H + H / 8 + 8;
    typedef typename In2<TIn>::type TIn2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* buf = reinterpret_cast<cd*>(smem_raw);
    const int tid = threadIdx.x;
    // twiddle tables of the engine and the untangle factors live in shared memory (an L1 hit still costs a long
    // scoreboard round trip in the middle of a pass)
    cd* tabs = buf + (size_t)RBI * PITCH;
    VTabs vt = a.tabs;
    {
        cd* p8 = tabs; cd* p64_8 = p8 + 56; cd* p64_4 = p64_8 + 448; cd* p256 = p64_4 + 192; cd* p512 = p256 + 768;
        for (int i = tid; i < 56; i += ROWV_NT) p8[i] = a.tabs.t8_8[i];
        if (H == 512 || H == 2048) for (int i = tid; i < 448; i += ROWV_NT) p64_8[i] = a.tabs.t64_8[i];
        if (H == 256 || H == 1024) for (int i = tid; i < 192; i += ROWV_NT) p64_4[i] = a.tabs.t64_4[i];
        if (H == 1024) for (int i = tid; i < 768; i += ROWV_NT) p256[i] = a.tabs.t256_4[i];
        if (H == 2048) for (int i = tid; i < 1536; i += ROWV_NT) p512[i] = a.tabs.t512_4[i];
        vt.t8_8 = p8; vt.t64_8 = p64_8; vt.t64_4 = p64_4; vt.t256_4 = p256; vt.t512_4 = p512;
    }
    cd* tw1s = tabs + 3000;
    for (int i = tid; i <= H / 2; i += ROWV_NT) tw1s[i] = a.tw1[i];
    __syncthreads();
    const int grp = tid / T, lane = tid - grp * T;
    cd* scratch = buf + (size_t)grp * PITCH;
    const int bar_id = 1 + grp;
    const double inv1 = 1.0 / (double)a.N1;
    const int ngroups = (a.N0 + RBI - 1) / RBI;
    const bool aligned = (a.N0 % 2 == 0);

    for (int gb = blockIdx.x * a.nit; gb < ngroups; gb += gridDim.x * a.nit) {
        TIn2 x[8], xn[8];
        {
            const int r = gb * RBI + grp;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int n = lane + q * T;
                if (r < a.N0) xn[q] = *reinterpret_cast<const TIn2*>(img + (size_t)r * a.N1 + 2 * n);
                else { xn[q].x = 0; xn[q].y = 0; }
            }
        }
        for (int it = 0; it < a.nit && gb + it < ngroups; ++it) {
            const int r0 = (gb + it) * RBI;
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = xn[q];
            if (it + 1 < a.nit && gb + it + 1 < ngroups) {
                const int r = r0 + RBI + grp;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = lane + q * T;
                    if (r < a.N0) xn[q] = *reinterpret_cast<const TIn2*>(img + (size_t)r * a.N1 + 2 * n);
                    else { xn[q].x = 0; xn[q].y = 0; }
                }
            }
            for (int j = 0; j < nj; ++j) {
                cd v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = lane + q * T;
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
                vfft<H>(v, scratch, lane, vt, -1.0, bar_id);
#pragma unroll
                for (int q = 0; q < 8; ++q) scratch[VPAD(lane + q * T)] = v[q];
                __syncthreads();
                // untangle k and H - k together for all RBI rows of the group
                const int nvalid = min(RBI, a.N0 - r0);
                for (int k = tid; k <= H / 2; k += ROWV_NT) {
                    const cd w = tw1s[k];
                    cd gk[RBI], gm[RBI];
#pragma unroll
                    for (int p = 0; p < RBI; ++p) {
                        const cd* pl = buf + (size_t)p * PITCH;
                        const cd A = pl[VPAD(k)];
                        const cd B = pl[VPAD((H - k) & (H - 1))];
                        // G[k]   = 0.5 (A + conj B) - 0.5 i W^k (A - conj B)
                        // G[H-k] = 0.5 (B + conj A) + 0.5 i conj(W^k) (B - conj A)
                        const cd s = cmake(A.x + B.x, A.y - B.y), d = cmake(A.x - B.x, A.y + B.y);
                        const cd wd = cmul(w, d);
                        gk[p] = cmake(0.5 * (s.x + wd.y), 0.5 * (s.y - wd.x));
                        // B - conj A = -conj(d);  conj(W) (-conj d) = -conj(W d) = -conj(wd)
                        gm[p] = cmake(0.5 * (s.x - wd.y), 0.5 * (-s.y - wd.x));
                    }
                    store_rows<TSt, RBI>(out + ((size_t)j * a.NH + k) * a.N0 + r0, gk, nvalid, aligned);
                    if (k != H - k) store_rows<TSt, RBI>(out + ((size_t)j * a.NH + (H - k)) * a.N0 + r0, gm, nvalid, aligned);
                }
                __syncthreads();
            }
        }
    }
}
