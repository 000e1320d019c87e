// fft_smem.cuh -- batched in-place complex FFT on planes resident in shared memory.
//
// Stockham index pattern (reads j + r*n/R, writes (j/Ns)*Ns*R + j%Ns + r*Ns) made in-place by
// holding each butterfly in registers across a barrier.  Any length whose prime factors are in
// {2,3,5,7,11,13}; all threads of the CTA cooperate on `nplanes` independent transforms.
// Twiddles come from a table tw[e] = exp(-2 pi i e / n); the inverse uses the conjugate.
#pragma once
#include "common.cuh"

// exp(-2 pi i q / R) for the odd radices handled by the generic O(R^2) butterfly; filled per device at init.
__constant__ double2 c_wgen[4][16];   // rows: radix 5, 7, 11, 13

template <int R> struct GenRow;
template <> struct GenRow<5>  { static const int row = 0; };
template <> struct GenRow<7>  { static const int row = 1; };
template <> struct GenRow<11> { static const int row = 2; };
template <> struct GenRow<13> { static const int row = 3; };

template <int R>
__device__ __forceinline__ void butterfly(cd* v, double sgn);

template <>
__device__ __forceinline__ void butterfly<2>(cd* v, double) {
    cd a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

template <>
__device__ __forceinline__ void butterfly<3>(cd* v, double sgn) {
    const double s3 = 0.86602540378443864676;
    cd t1 = cadd(v[1], v[2]);
    cd t2 = cmake(v[0].x - 0.5 * t1.x, v[0].y - 0.5 * t1.y);
    cd t3 = cmuli(cscale(csub(v[1], v[2]), s3), sgn);
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, t3);
    v[2] = csub(t2, t3);
}

template <>
__device__ __forceinline__ void butterfly<4>(cd* v, double sgn) {
    cd t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    cd t2 = cadd(v[1], v[3]), t3 = cmuli(csub(v[1], v[3]), sgn);
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}

template <>
__device__ __forceinline__ void butterfly<8>(cd* v, double sgn) {
    const double h = 0.70710678118654752440;
    cd e[4] = {v[0], v[2], v[4], v[6]};
    cd o[4] = {v[1], v[3], v[5], v[7]};
    butterfly<4>(e, sgn);
    butterfly<4>(o, sgn);
    // o[k] *= exp(sgn * 2 pi i k / 8)
    cd o1 = cmake(h * (o[1].x - sgn * o[1].y), h * (o[1].y + sgn * o[1].x));
    cd o2 = cmuli(o[2], sgn);
    cd o3 = cmake(h * (-o[3].x - sgn * o[3].y), h * (-o[3].y + sgn * o[3].x));
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
}

template <int R>
__device__ __forceinline__ void butterfly_generic(cd* v, double sgn) {
    cd out[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        cd acc = v[0];
#pragma unroll
        for (int r = 1; r < R; ++r) {
            double2 w = c_wgen[GenRow<R>::row][(q * r) % R];
            cd ww = cmake(w.x, -sgn * w.y);
            cfma(acc, v[r], ww);
        }
        out[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = out[q];
}
template <> __device__ __forceinline__ void butterfly<5>(cd* v, double sgn)  { butterfly_generic<5>(v, sgn); }
template <> __device__ __forceinline__ void butterfly<7>(cd* v, double sgn)  { butterfly_generic<7>(v, sgn); }
template <> __device__ __forceinline__ void butterfly<11>(cd* v, double sgn) { butterfly_generic<11>(v, sgn); }
template <> __device__ __forceinline__ void butterfly<13>(cd* v, double sgn) { butterfly_generic<13>(v, sgn); }

// maximum butterflies one thread may hold when a plane has more butterflies than the CTA has threads
template <int R> struct MaxB { static const int value = (16 / R) > 0 ? (16 / R) : 1; };

template <int R>
__device__ __forceinline__ void fft_load_bfly(cd* v, const cd* plane, int j, int nb, int Ns, int tstep,
                                              const cd* __restrict__ tw, double sgn) {
    const int k = j % Ns;
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = plane[j + r * nb];
    if (Ns > 1) {
#pragma unroll
        for (int r = 1; r < R; ++r) {
            cd w = tw[r * k * tstep];
            v[r] = cmul(v[r], cmake(w.x, -sgn * w.y));
        }
    }
    butterfly<R>(v, sgn);
}

template <int R>
__device__ __forceinline__ void fft_store_bfly(const cd* v, cd* plane, int j, int Ns) {
    const int k = j % Ns;
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) plane[j0 + r * Ns] = v[r];
}

// One stage over all planes.  Entry: all prior writes to buf visible (caller barrier).  Exit: barrier done.
template <int R>
__device__ __noinline__ void fft_stage(cd* buf, int pitch, int nplanes, int n, int Ns,
                                          const cd* __restrict__ tw, double sgn) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int nb = n / R;
    const int tstep = n / (Ns * R);
    if (nb <= nthr) {
        const int ppc = nthr / nb;                 // planes per chunk
        const int pl = tid / nb, j = tid - pl * nb;
        for (int p0 = 0; p0 < nplanes; p0 += ppc) {
            const int p = p0 + pl;
            const bool act = (pl < ppc) && (p < nplanes);
            cd v[R];
            if (act) fft_load_bfly<R>(v, buf + (size_t)p * pitch, j, nb, Ns, tstep, tw, sgn);
            __syncthreads();
            if (act) fft_store_bfly<R>(v, buf + (size_t)p * pitch, j, Ns);
        }
        __syncthreads();
    } else {
        constexpr int MB = MaxB<R>::value;
        for (int p = 0; p < nplanes; ++p) {
            cd v[MB][R];
            cd* plane = buf + (size_t)p * pitch;
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int j = tid + b * nthr;
                if (j < nb) fft_load_bfly<R>(v[b], plane, j, nb, Ns, tstep, tw, sgn);
            }
            __syncthreads();
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                const int j = tid + b * nthr;
                if (j < nb) fft_store_bfly<R>(v[b], plane, j, Ns);
            }
        }
        __syncthreads();
    }
}

// In-place FFT of `nplanes` planes of length d.n at buf + p*pitch.  sgn = -1 forward, +1 unnormalised inverse.
// Requires a barrier-visible buf on entry; returns after a barrier.
__device__ __forceinline__ void fft_planes(cd* buf, int pitch, int nplanes, const FftDesc& d,
                                        const cd* __restrict__ tw, double sgn) {
    int Ns = 1;
    for (int s = 0; s < d.ns; ++s) {
        const int R = d.radix[s];
        switch (R) {
            case 2:  fft_stage<2>(buf, pitch, nplanes, d.n, Ns, tw, sgn); break;
            case 3:  fft_stage<3>(buf, pitch, nplanes, d.n, Ns, tw, sgn); break;
            case 4:  fft_stage<4>(buf, pitch, nplanes, d.n, Ns, tw, sgn); break;
            case 5:  fft_stage<5>(buf, pitch, nplanes, d.n, Ns, tw, sgn); break;
            case 7:  fft_stage<7>(buf, pitch, nplanes, d.n, Ns, tw, sgn); break;
            case 8:  fft_stage<8>(buf, pitch, nplanes, d.n, Ns, tw, sgn); break;
            case 11: fft_stage<11>(buf, pitch, nplanes, d.n, Ns, tw, sgn); break;
            case 13: fft_stage<13>(buf, pitch, nplanes, d.n, Ns, tw, sgn); break;
            default: break;
        }
        Ns *= R;
    }
}
