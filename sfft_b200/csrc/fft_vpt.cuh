// fft_vpt.cuh -- register FFT engine with 8 complex doubles per thread (T = N / 8 threads per transform).
//
// Same Stockham auto-sort pass structure as fft_regs.cuh (layout v[q] = x[lane + q T] on entry and on exit, data
// crosses threads through a padded shared-memory scratch once per non-final pass), but with radix-8/4 passes so that
// a thread needs ~100 registers instead of ~170: a 256-point transform is one warp (no CTA-level barrier at all),
// a 2048-point transform 256 threads on a named barrier.  Padding i -> i + i/8 makes the stride-8 stores of the
// first pass, the stride-(8, 64) stores of the later passes and the unit-stride loads conflict-free for 16-byte
// elements.  Twiddle tables are laid out [r-1][k], k = 0..Ns-1:  tab[(r-1) Ns + k] = exp(-2 pi i r k / (Ns R)).
#pragma once
#include "fft_smem.cuh"

#define VPAD(i) ((i) + ((i) >> 3))

struct VTabs {
    const cd* t8_8;      // Ns = 8,   R = 8
    const cd* t64_8;     // Ns = 64,  R = 8
    const cd* t64_4;     // Ns = 64,  R = 4
    const cd* t256_4;    // Ns = 256, R = 4
    const cd* t512_4;    // Ns = 512, R = 4
};

template <int T>
__device__ __forceinline__ void vsync(int bar_id) {
    if (T <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(T) : "memory");
}

template <int R> __device__ __forceinline__ void vbfly(cd* t, double sgn) { butterfly<R>(t, sgn); }

template <int T, int R, int Ns, bool LAST>
__device__ __forceinline__ void vpass(cd (&v)[8], cd* scratch, int lane, const cd* __restrict__ tab, double sgn, int bar_id) {
    constexpr int NB = 8 / R;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = lane + b * T;
        const int k = j & (Ns - 1);
        cd t[R];
#pragma unroll
        for (int r = 0; r < R; ++r) t[r] = v[b + NB * r];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const cd w = tab[(r - 1) * Ns + k];
                t[r] = cmul(t[r], cmake(w.x, -sgn * w.y));
            }
        }
        vbfly<R>(t, sgn);
        if (LAST) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[b + NB * r] = t[r];
        } else {
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) scratch[VPAD(j0 + r * Ns)] = t[r];
        }
    }
    if (!LAST) {
        vsync<T>(bar_id);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = scratch[VPAD(lane + q * T)];
        vsync<T>(bar_id);
    }
}

// N-point transform by T = N / 8 threads; sgn = -1 forward, +1 unnormalised inverse.  scratch: VPAD(N) elements.
template <int N>
__device__ __forceinline__ void vfft(cd (&v)[8], cd* scratch, int lane, const VTabs& tb, double sgn, int bar_id) {
    constexpr int T = N / 8;
    static_assert(N == 256 || N == 512 || N == 1024 || N == 2048, "unsupported length");
    vpass<T, 8, 1, false>(v, scratch, lane, nullptr, sgn, bar_id);
    vpass<T, 8, 8, false>(v, scratch, lane, tb.t8_8, sgn, bar_id);
    if (N == 256) vpass<T, 4, 64, true>(v, scratch, lane, tb.t64_4, sgn, bar_id);
    if (N == 512) vpass<T, 8, 64, true>(v, scratch, lane, tb.t64_8, sgn, bar_id);
    if (N == 1024) {
        vpass<T, 4, 64, false>(v, scratch, lane, tb.t64_4, sgn, bar_id);
        vpass<T, 4, 256, true>(v, scratch, lane, tb.t256_4, sgn, bar_id);
    }
    if (N == 2048) {
        vpass<T, 8, 64, false>(v, scratch, lane, tb.t64_8, sgn, bar_id);
        vpass<T, 4, 512, true>(v, scratch, lane, tb.t512_4, sgn, bar_id);
    }
}
