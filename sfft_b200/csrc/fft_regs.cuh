// fft_regs.cuh -- register-resident power-of-two FFT engine.
//
// A group of T threads transforms N = 16 T points; every thread keeps 16 complex doubles in registers, in the
// layout v[q] = x[lane + q T] on entry AND on exit.  Passes are radix 16 (then one radix 2/4/8/16 tail pass); data
// crosses threads once per non-final pass through a padded shared-memory scratch (index i -> i + i/16, which makes
// both the stride-16 stores of the first pass and the unit-stride loads conflict-free for 16-byte elements).  The
// final pass leaves its outputs in registers.  Twiddles come from per-pass tables laid out [r-1][k] so that
// consecutive lanes read consecutive entries:
//     tabA[(r-1)*16  + k] = exp(-2 pi i r k / 256)          (second pass, Ns = 16, R = 16)
//     tabB[(r-1)*256 + k] = exp(-2 pi i r k / (256 R3))     (third pass,  Ns = 256)
//     tabC[(r-1)*4096 + k] = exp(-2 pi i r k / 8192)        (fourth pass, N = 8192 only)
#pragma once
#include "fft_smem.cuh"

#define RPAD(i) ((i) + ((i) >> 4))

__device__ __forceinline__ void butterfly16(cd* v, double sgn) {
    const double c = 0.92387953251128675613, s = 0.38268343236508977173, h = 0.70710678118654752440;
    cd y[4][4];
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) {
        cd t[4] = {v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]};
        butterfly<4>(t, sgn);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) y[n1][k1] = t[k1];
    }
    // y[n1][k1] *= exp(sgn 2 pi i n1 k1 / 16)
    y[1][1] = cmul(y[1][1], cmake(c, sgn * s));
    y[1][2] = cmul(y[1][2], cmake(h, sgn * h));
    y[1][3] = cmul(y[1][3], cmake(s, sgn * c));
    y[2][1] = cmul(y[2][1], cmake(h, sgn * h));
    y[2][2] = cmuli(y[2][2], sgn);
    y[2][3] = cmul(y[2][3], cmake(-h, sgn * h));
    y[3][1] = cmul(y[3][1], cmake(s, sgn * c));
    y[3][2] = cmul(y[3][2], cmake(-h, sgn * h));
    y[3][3] = cmul(y[3][3], cmake(-c, -sgn * s));
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        cd t[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
        butterfly<4>(t, sgn);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
    }
}

template <int R> __device__ __forceinline__ void bfly_r(cd* t, double sgn) { butterfly<R>(t, sgn); }
template <> __device__ __forceinline__ void bfly_r<16>(cd* t, double sgn) { butterfly16(t, sgn); }
template <> __device__ __forceinline__ void bfly_r<1>(cd*, double) {}

// group synchronisation: a (sub-)warp mask, or a named barrier shared by T >= 64 threads
struct GroupSync {
    unsigned mask;      // for T <= 32
    int bar_id, count;  // for T > 32
    template <int T> __device__ __forceinline__ void sync() const {
        if (T <= 32) __syncwarp(mask);
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(count) : "memory");
    }
};

template <int T, int R, int Ns, bool LAST>
__device__ __forceinline__ void reg_pass(cd (&v)[16], cd* scratch, int lane, const cd* __restrict__ tab, double sgn,
                                         const GroupSync& gs) {
    constexpr int NB = 16 / R;
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
        bfly_r<R>(t, sgn);
        if (LAST) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[b + NB * r] = t[r];
        } else {
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) scratch[RPAD(j0 + r * Ns)] = t[r];
        }
    }
    if (!LAST) {
        gs.sync<T>();
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = scratch[RPAD(lane + q * T)];
        gs.sync<T>();
    }
}

// N-point transform by T = N / 16 threads.  sgn = -1 forward, +1 unnormalised inverse.
template <int N>
__device__ __forceinline__ void reg_fft(cd (&v)[16], cd* scratch, int lane, const cd* __restrict__ tabA,
                                        const cd* __restrict__ tabB, const cd* __restrict__ tabC, double sgn, const GroupSync& gs) {
    constexpr int T = N / 16;
    static_assert(N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096 || N == 8192, "unsupported length");
    reg_pass<T, 16, 1, false>(v, scratch, lane, nullptr, sgn, gs);
    if (N == 256) {
        reg_pass<T, 16, 16, true>(v, scratch, lane, tabA, sgn, gs);
    } else {
        reg_pass<T, 16, 16, false>(v, scratch, lane, tabA, sgn, gs);
        if (N == 512) reg_pass<T, 2, 256, true>(v, scratch, lane, tabB, sgn, gs);
        if (N == 1024) reg_pass<T, 4, 256, true>(v, scratch, lane, tabB, sgn, gs);
        if (N == 2048) reg_pass<T, 8, 256, true>(v, scratch, lane, tabB, sgn, gs);
        if (N == 4096) reg_pass<T, 16, 256, true>(v, scratch, lane, tabB, sgn, gs);
        if (N == 8192) {
            reg_pass<T, 16, 256, false>(v, scratch, lane, tabB, sgn, gs);
            reg_pass<T, 2, 4096, true>(v, scratch, lane, tabC, sgn, gs);
        }
    }
}

static inline int reg_fft_tail_radix(int N) { return N == 512 ? 2 : N == 1024 ? 4 : N == 2048 ? 8 : (N == 4096 || N == 8192) ? 16 : 0; }
