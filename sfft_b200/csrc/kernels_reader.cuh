// kernels_reader.cuh -- consumers of the Solution vector that the packets call right after GSS
// (sfft/utils/SFFTSolutionReader.py:116-196: Realize_MatchingKernel, Realize_FluxScaling), evaluated on the device so
// that the PureCupy-style path needs no D2H copy of the Solution.
#pragma once
#include "common.cuh"

struct ReaderArgs {
    const double* sol;        // (NEQ) a_ijab | b_pq
    const double* xy;         // (nq, 2) requested coordinates, FortranCoor (pixel centre r,c -> r+1, c+1)
    int nq, N0, N1, L0, L1, DK, Fij, Fab;
    double* kerstack;         // (nq, L0, L1) or null
    double* fscal;            // (nq) or null
};

#ifdef SFFTB_TU_MAIN
// One CTA per requested coordinate.  K_q[a,b] = sum_ij x^i y^j s_ij[a,b] with s = a_ijab / N in the Cartesian-delta
// basis: the centre tap of every (i,j) block becomes 2 s_ij[0,0] - sum_ab s_ij[a,b] (SVKDict_SFFT2ST.convert, :102-114);
// the flux scaling is sum_ij s_ij[0,0] x^i y^j (:160-181).
__global__ void __launch_bounds__(256) realize_kernel(ReaderArgs a)
{
    __shared__ double bas[SFFTB_MAX_PLANES];     // x^i y^j
    __shared__ double blk[SFFTB_MAX_PLANES];     // sum_ab a_ijab of block ij
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const double x = a.xy[2 * q] / a.N0, y = a.xy[2 * q + 1] / a.N1;
    if (tid == 0) {
        int ij = 0;
        for (int i = 0; i <= a.DK; ++i)
            for (int j = 0; j <= a.DK - i; ++j) bas[ij++] = ipow(x, i) * ipow(y, j);
    }
    for (int ij = warp; ij < a.Fij; ij += nw) {
        double s = 0.0;
        for (int ab = lane; ab < a.Fab; ab += 32) s += a.sol[ij * a.Fab + ab];
        s = warp_sum(s);
        if (lane == 0) blk[ij] = s;
    }
    __syncthreads();
    const double inv = 1.0 / ((double)a.N0 * (double)a.N1);
    const int centre = (a.L0 / 2) * a.L1 + a.L1 / 2;
    if (a.kerstack) {
        for (int ab = tid; ab < a.Fab; ab += blockDim.x) {
            double v = 0.0;
            for (int ij = 0; ij < a.Fij; ++ij) {
                double s = a.sol[ij * a.Fab + ab];
                if (ab == centre) s = 2.0 * s - blk[ij];
                v = fma(bas[ij], s * inv, v);
            }
            a.kerstack[(size_t)q * a.Fab + ab] = v;
        }
    }
    if (a.fscal && tid == 0) {
        double v = 0.0;
        for (int ij = 0; ij < a.Fij; ++ij) v = fma(bas[ij], a.sol[ij * a.Fab + centre] * inv, v);
        a.fscal[q] = v;
    }
}
#endif  // SFFTB_TU_MAIN
