// fft_h16.cuh -- 256-point register FFT by a HALF warp: 16 lanes x 16 complex doubles, two radix-16 passes, ONE exchange.
//
// The 8-values-per-thread engine (fft_vpt.cuh) moves a 256-point transform through shared memory twice (radix 8 / 8 / 4);
// the segmented fit kernel is bound by exactly that traffic.  Here a transform crosses lanes once: layout
// v[q] = x[hl + 16 q] on entry and on exit (hl = lane & 15), pass 1 = radix-16 butterflies without twiddles, exchange
// through a padded scratch (index i -> i + i/16: the stride-16 stores and the unit-stride loads are conflict-free for
// 16-byte elements), pass 2 = twiddles exp(-+2 pi i r hl / 256) and radix-16 butterflies, outputs stay in registers.
// The twiddles of a lane never change (k = hl for every transform): the four powers w^1, w^2, w^4, w^8 live in
// registers and the other eleven are formed by at most two multiplications each -- no twiddle traffic at all.
// Both halves of a warp run in lockstep on two independent transforms (separate scratch regions), so the only
// synchronisation is __syncwarp.
#pragma once
#include "fft_regs.cuh"

#define HPAD(i) ((i) + ((i) >> 4))
#define H16_SCRATCH 272                 // HPAD(255) + 1 elements per transform

struct H16Tw { cd w1, w2, w4, w8; };    // exp(-2 pi i r hl / 256), r = 1, 2, 4, 8 (forward sign)

__device__ __forceinline__ void h16_init(H16Tw& t, int hl) {
    double s, c;
    sincospi(-2.0 * (double)hl / 256.0, &s, &c); t.w1 = cmake(c, s);
    sincospi(-4.0 * (double)hl / 256.0, &s, &c); t.w2 = cmake(c, s);
    sincospi(-8.0 * (double)hl / 256.0, &s, &c); t.w4 = cmake(c, s);
    sincospi(-16.0 * (double)hl / 256.0, &s, &c); t.w8 = cmake(c, s);
}

// the same four powers from the engine table tabA[(r - 1) * 16 + k] = exp(-2 pi i r k / 256) (upload_engine_table(16, 16))
__device__ __forceinline__ void h16_load(H16Tw& t, const cd* __restrict__ tabA, int hl) {
    t.w1 = tabA[hl]; t.w2 = tabA[16 + hl]; t.w4 = tabA[48 + hl]; t.w8 = tabA[112 + hl];
}

// v[r] *= w^r, r = 1 .. 15, from the powers w^1, w^2, w^4, w^8: eleven more products, at most two per power
__device__ __forceinline__ void h16_twiddle(cd (&v)[16], cd w1, cd w2, cd w4, cd w8) {
    const cd w3 = cmul(w1, w2);
    v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3); v[4] = cmul(v[4], w4);
    const cd w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
    v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7); v[8] = cmul(v[8], w8);
    v[9] = cmul(v[9], cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
    v[12] = cmul(v[12], cmul(w8, w4)); v[13] = cmul(v[13], cmul(w8, w5)); v[14] = cmul(v[14], cmul(w8, w6));
    v[15] = cmul(v[15], cmul(w8, w7));
}

// sgn = -1 forward, +1 unnormalised inverse.  scratch: H16_SCRATCH elements private to this half warp.
// live = false: a half warp without a job keeps the lockstep (same instruction stream, __syncwarp) but touches no memory.
__device__ __forceinline__ void hfft256(cd (&v)[16], cd* scratch, int hl, const H16Tw& tw, double sgn, bool live = true) {
    butterfly16(v, sgn);
    if (live) {
#pragma unroll
        for (int r = 0; r < 16; ++r) scratch[HPAD(16 * hl + r)] = v[r];
    }
    __syncwarp();
    if (live) {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = scratch[HPAD(hl + 16 * q)];
    }
    __syncwarp();
    h16_twiddle(v, cmake(tw.w1.x, -sgn * tw.w1.y), cmake(tw.w2.x, -sgn * tw.w2.y), cmake(tw.w4.x, -sgn * tw.w4.y), cmake(tw.w8.x, -sgn * tw.w8.y));
    butterfly16(v, sgn);
}
