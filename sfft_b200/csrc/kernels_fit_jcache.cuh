// kernels_fit_jcache.cuh -- shared-template tiles after the first, with the template's segment spectra CACHED in HBM.
//
// A tile of a shared-template batch (BASELINE config 4; the masked template I is fixed, only J changes) needs from the fit column
// pass only the cross spectra conj(F_A) F_J, A = the Fij planes cx^i cy^j I in the "A role" (zero-padded core rows of every
// segment), against the B-role spectrum of J.  fit_seg4_kernel<JONLY> recomputes the Fij A-role transforms of the TEMPLATE for
// every tile: Fij + 1 transforms per segment, of which Fij give the same result tile after tile.  Here
//   * aspec_cache_kernel      (once per template): cache[k1][s][A][256] = the A-role spectrum, computed with the very expressions of
//                              the transform warps of fit_seg4_kernel (same window weights, same half-warp engine): bit-identical;
//   * fit_jonly_cached_kernel (per tile): ONE transform per segment (J), the products read the cached spectra (coalesced 4 KB rows),
//                              Fij accumulators per frequency bin, Fij inverse transforms per column, the J x T rows from the column
//                              moments like the JONLY instantiation.  Accumulation order over the segments as in fit_seg4_kernel, so
//                              the lag rows -- and with them Solution and DIFF -- equal those of the uncached path bit for bit.
// Cache size: NH * nseg * Fij * 4 KB (252 MB at 2048^2, KerPolyOrder 2); it is rebuilt when the template changes.
// Reference: the cross-spectrum planes of ElementalSFFTSubtract that involve J only (HadProd_OMG / GAM with the science image,
// sfft/sfftcore/SFFTSubtract.py:164-331) -- recomputed in full per pair there.
#pragma once
#include "kernels_fit_seg4.cuh"

#define JC_NT 256
#define JC_NHW (JC_NT / 16)            // half warps per CTA = segments transformed per batch
static inline size_t jcache_smem_bytes() { return sizeof(cd) * ((size_t)JC_NHW * FS4_PITCH + 5 * SFFTB_MAXE) + 16; }

// window element n = hl + 16 q of segment s of column `col`, times keep(n) * cx(row)^pw -- the load of fit_seg4_kernel's transform warps
template <typename TSt>
__device__ __forceinline__ void jc_load_window(cd (&v)[16], const TSt* __restrict__ col, int N0, int c0, int h, int hl, int klo, int khi,
                                               int pw, double inv0, bool wrap1)
{
    const int row_l = c0 - h + hl;
    const double cx_l = (double)(row_l + 1) * inv0;
    const double e0 = pw == 0 ? 1.0 : 0.0, e1 = pw == 1 ? 1.0 : 0.0, e2 = pw == 2 ? 1.0 : 0.0, e3 = pw == 3 ? 1.0 : 0.0;
    int r = wrap_row(row_l, N0);
    const int step = 16 % N0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const cd gg = load_c(col + r);
        r += step;
        if (r >= N0) r -= N0;
        const int row = row_l + 16 * q;
        double cx = fma((double)(16 * q), inv0, cx_l);
        if (wrap1) cx += (row < 0) ? 1.0 : ((row >= N0) ? -1.0 : 0.0);
        else cx -= floor(fma(-0.5, inv0, cx));
        double sc = fma(cx, fma(cx, fma(cx, e3, e2), e1), e0);
        sc = (16 * q >= klo && 16 * q < khi) ? sc : 0.0;
        v[q] = cmake(gg.x * sc, gg.y * sc);
    }
}

template <typename TSt, int DK>
__global__ void __launch_bounds__(JC_NT) aspec_cache_kernel(SegFitArgs fa, const TSt* __restrict__ gI, cd* __restrict__ cache)
{
    constexpr int Fij = (DK + 1) * (DK + 2) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* planes = reinterpret_cast<cd*>(smem_raw);
    const ColArgs& a = fa.c;
    const int tid = threadIdx.x, hw = tid >> 4, hl = tid & 15;
    const int h = fa.h, S = fa.S, nseg = fa.nseg;
    const double inv0 = 1.0 / (double)a.N0;
    const bool wrap1 = (a.N0 >= FS3_M);
    H16Tw htw;
    h16_load(htw, fa.tabA, hl);
    cd* scratch = planes + (size_t)hw * FS4_PITCH;
    const long long njobs = (long long)a.NH * nseg * Fij;
    for (long long j0 = (long long)blockIdx.x * JC_NHW; j0 < njobs; j0 += (long long)gridDim.x * JC_NHW) {
        const bool active = j0 + hw < njobs;
        const long long job = active ? j0 + hw : j0;
        const int A = (int)(job % Fij);
        const long long ks = job / Fij;
        const int s = (int)(ks % nseg), k1 = (int)(ks / nseg);
        const int c0 = s * S, Sc = min(S, a.N0 - c0);
        const TSt* col = gI + ((size_t)a.pl_j[A] * a.NH + k1) * a.N0;
        cd v[16];
        jc_load_window(v, col, a.N0, c0, h, hl, h - hl, (active ? h + Sc : 0) - hl, a.pl_i[A], inv0, wrap1);
        hfft256(v, scratch, hl, htw, -1.0, active);
        if (active) {
            cd* dst = cache + (size_t)job * FS3_M;
#pragma unroll
            for (int q = 0; q < 16; ++q) dst[hl + 16 * q] = v[q];
        }
        __syncwarp();
    }
}

template <typename TSt, int DK>
__global__ void __launch_bounds__(JC_NT, 2) fit_jonly_cached_kernel(SegFitArgs fa, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
                                                                 const cd* __restrict__ cache, cd* __restrict__ kap)
{
    constexpr int Fij = (DK + 1) * (DK + 2) / 2, NPAIR = Fij * (Fij + 1) / 2;
    constexpr int NMS = Fs3Mom<DK>::npl * SFFTB_MAXE;
    static_assert(Fij <= JC_NHW, "one half warp per inverse transform");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* planes = reinterpret_cast<cd*>(smem_raw);                  // [JC_NHW][FS4_PITCH]
    cd* mom = planes + (size_t)JC_NHW * FS4_PITCH;
    const ColArgs& a = fa.c;
    const int tid = threadIdx.x, hw = tid >> 4, hl = tid & 15;
    const int h = fa.h, S = fa.S, nseg = fa.nseg;
    const double inv0 = 1.0 / (double)a.N0;
    const bool wrap1 = (a.N0 >= FS3_M);
    H16Tw htw;
    h16_load(htw, fa.tabA, hl);
    cd* myplane = planes + (size_t)hw * FS4_PITCH;
    for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x) {
        cd* kaprow = kap + (size_t)k1 * fa.nrows;
        const TSt* colJ = gJ + (size_t)k1 * a.N0;
        const cd* ck = cache + (size_t)k1 * nseg * Fij * FS3_M + tid;
        cd acc[Fij];
#pragma unroll
        for (int A = 0; A < Fij; ++A) acc[A] = cmake(0.0, 0.0);
        for (int s0 = 0; s0 < nseg; s0 += JC_NHW) {
            {   // B-role spectrum of J for segment s0 + hw (segment with its halo, no weight)
                const int s = s0 + hw;
                const bool active = s < nseg;
                const int c0 = (active ? s : 0) * S;
                cd v[16];
                jc_load_window(v, colJ, a.N0, c0, h, hl, -hl, (active ? FS3_M : 0) - hl, 0, inv0, wrap1);
                hfft256(v, myplane, hl, htw, -1.0, active);
                if (active) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) myplane[HPAD(hl + 16 * q)] = v[q];
                }
            }
            __syncthreads();
            const int nb = min(JC_NHW, nseg - s0);
#pragma unroll 2
            for (int w = 0; w < nb; ++w) {
                const cd fJ = planes[(size_t)w * FS4_PITCH + HPAD(tid)];
                const cd* cs = ck + (size_t)(s0 + w) * Fij * FS3_M;
                cd fA[Fij];
#pragma unroll
                for (int A = 0; A < Fij; ++A) fA[A] = cs[(size_t)A * FS3_M];
#pragma unroll
                for (int A = 0; A < Fij; ++A) {
                    acc[A].x = fma(fA[A].x, fJ.x, acc[A].x); acc[A].x = fma(fA[A].y, fJ.y, acc[A].x);
                    acc[A].y = fma(fA[A].x, fJ.y, acc[A].y); acc[A].y = fma(-fA[A].y, fJ.x, acc[A].y);
                }
            }
            __syncthreads();
        }
        // one inverse transform per (A, J) pair, lags |m0| <= w0 into the Theta rows of the column
#pragma unroll
        for (int A = 0; A < Fij; ++A) planes[(size_t)A * FS4_PITCH + HPAD(tid)] = acc[A];
        if (tid < NMS) mom[tid] = fa.momg[(size_t)k1 * NMS + tid];
        __syncthreads();
        {
            const bool act = hw < Fij;
            if (__any_sync(0xffffffffu, act)) fs4_inverse_job<NPAIR>(fa, htw, planes + (size_t)(act ? hw : 0) * FS4_PITCH, NPAIR + (act ? hw : 0), hl, act, kaprow);
        }
        // J x T rows from the column moments of J (col_moments_kernel, jonly)
        column_poly_rows_sub(fa, gI, k1, mom, kaprow, tid, JC_NT, true);
        __syncthreads();
    }
}
