// kernels_solve.cuh -- dense fp64 solve of the (diagonally scaled) normal equations.
//
// Replaces LSSolver (cuSOLVER getrf/getrs via cupyx, sfft/sfftcore/SFFTSubtract.py:15-23, 397-408; CPU
// np.linalg.solve :744-747).  LHMAT = D^T D / N is symmetric positive definite, so the primary path is a blocked
// right-looking Cholesky on the matrix augmented with the right-hand side as an extra row (the panel solves then
// deliver the forward substitution for free); a pivoted LU on all SMs is the fallback when a pivot is not positive.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CH_NB 64

#ifdef SFFTB_TU_CHOL
// Factor the kb x kb diagonal block at k0 (every CTA redundantly, in shared memory), then solve the rows below it:
// X <- X L_kk^{-T}.  Matrix is row-major with leading dimension ld, ntot rows (n + 1 with the rhs row), n columns.
__global__ void __launch_bounds__(128) chol_panel_kernel(double* __restrict__ A, int ld, int ntot, int n, int k0, int* __restrict__ info)
{
    __shared__ double D[CH_NB][CH_NB + 1];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int kb = min(CH_NB, n - k0);
    for (int idx = tid; idx < CH_NB * CH_NB; idx += nthr) {
        const int r = idx / CH_NB, c = idx - r * CH_NB;
        double v = (r == c) ? 1.0 : 0.0;
        if (r < kb && c <= r) v = A[(size_t)(k0 + r) * ld + k0 + c];
        D[r][c] = v;
    }
    __syncthreads();
    for (int c = 0; c < kb; ++c) {
        double s = 0.0;
        const int r = tid;
        if (r >= c && r < kb) {
            s = D[r][c];
            for (int k = 0; k < c; ++k) s -= D[r][k] * D[c][k];
            if (r == c) {
                if (!(s > 0.0) || !isfinite(s)) { if (blockIdx.x == 0) atomicCAS(&info[0], 0, k0 + c + 1); s = 1.0; }
                D[c][c] = sqrt(s);
            }
        }
        __syncthreads();
        if (r > c && r < kb) D[r][c] = s / D[c][c];
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        for (int idx = tid; idx < kb * kb; idx += nthr) {
            const int r = idx / kb, c = idx - r * kb;
            if (c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = D[r][c];
        }
    }
    const int g = k0 + kb + blockIdx.x * nthr + tid;
    if (g < ntot) {
        double x[CH_NB];
        double* row = A + (size_t)g * ld + k0;
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) x[c] = (c < kb) ? row[c] : 0.0;
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) {
            double s = x[c];
#pragma unroll
            for (int k = 0; k < c; ++k) s = fma(-x[k], D[c][k], s);
            x[c] = s / D[c][c];
        }
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) if (c < kb) row[c] = x[c];
    }
}

// Trailing update A[i][j] -= sum_k P[i][k] P[j][k] on 64 x 64 tiles of the lower triangle, rows/cols >= k0 + kb.
__global__ void __launch_bounds__(256) chol_update_kernel(double* __restrict__ A, int ld, int ntot, int n, int k0)
{
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    __shared__ double Pi[64][33];
    __shared__ double Pj[64][33];
    const int kb = min(CH_NB, n - k0);
    const int s = k0 + kb;
    const int i0 = s + bi * 64, j0 = s + bj * 64;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = 0.0;
    for (int kk = 0; kk < kb; kk += 32) {
        for (int idx = tid; idx < 64 * 32; idx += 256) {
            const int r = idx >> 5, k = idx & 31;
            Pi[r][k] = (i0 + r < ntot && kk + k < kb) ? A[(size_t)(i0 + r) * ld + k0 + kk + k] : 0.0;
            Pj[r][k] = (j0 + r < ntot && kk + k < kb) ? A[(size_t)(j0 + r) * ld + k0 + kk + k] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) av[ii] = Pi[ty + 16 * ii][k];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) bv[jj] = Pj[tx + 16 * jj][k];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = fma(av[ii], bv[jj], acc[ii][jj]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        const int gi = i0 + ty + 16 * ii;
        if (gi >= ntot) continue;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int gj = j0 + tx + 16 * jj;
            if (gj < n && gj <= gi) A[(size_t)gi * ld + gj] -= acc[ii][jj];
        }
    }
}

// Back substitution L^T x = y (y = row n of the factored augmented matrix), unscale, scatter to the NEQ-long
// solution (Extend_Solution, sfft/sfftcore/SFFTConfigure.py:716-732).  One CTA.
__global__ void __launch_bounds__(1024) chol_backsolve_kernel(const double* __restrict__ A, int ld, int n,
                                                              const double* __restrict__ sc, const int* __restrict__ idx,
                                                              double* __restrict__ sol, int NEQ)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* y = reinterpret_cast<double*>(smem_raw);            // n
    double* xb = y + n;                                         // CH_NB
    double* D = xb + CH_NB;                                     // CH_NB * (CH_NB + 1)
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    for (int c = tid; c < n; c += nthr) y[c] = A[(size_t)n * ld + c];
    for (int c = tid; c < NEQ; c += nthr) sol[c] = 0.0;
    const int nblk = (n + CH_NB - 1) / CH_NB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int k0 = b * CH_NB, kb = min(CH_NB, n - k0);
        __syncthreads();
        for (int i = tid; i < CH_NB * CH_NB; i += nthr) {
            const int r = i / CH_NB, c = i - r * CH_NB;
            D[r * (CH_NB + 1) + c] = (r < kb && c <= r) ? A[(size_t)(k0 + r) * ld + k0 + c] : (r == c ? 1.0 : 0.0);
        }
        __syncthreads();
        if (tid < 32) {
            double y0 = (lane < kb) ? y[k0 + lane] : 0.0;
            double y1 = (32 + lane < kb) ? y[k0 + 32 + lane] : 0.0;
            for (int c = CH_NB - 1; c >= 0; --c) {
                const double val = (c >> 5) ? y1 : y0;
                const double xc = __shfl_sync(0xffffffffu, val, c & 31) / D[c * (CH_NB + 1) + c];
                if (lane < c) y0 = fma(-D[c * (CH_NB + 1) + lane], xc, y0);
                if (32 + lane < c) y1 = fma(-D[c * (CH_NB + 1) + 32 + lane], xc, y1);
                if (lane == (c & 31)) xb[c] = xc;
            }
        }
        __syncthreads();
        for (int r = tid; r < k0; r += nthr) {
            double s = 0.0;
            for (int c = 0; c < kb; ++c) s = fma(A[(size_t)(k0 + c) * ld + r], xb[c], s);
            y[r] -= s;
        }
        if (tid < kb) y[k0 + tid] = xb[tid];
    }
    __syncthreads();
    for (int c = tid; c < n; c += nthr) sol[idx[c]] = y[c] * sc[c];
}

// ---- fallback: LU with partial pivoting on all SMs (cooperative launch, one grid.sync per column) -----------------
// A is the (n+1) x ld augmented matrix refilled by fill_matrix_kernel (row n / column n = rhs).  Columns are owned
// round-robin by CTAs; the rhs column n is eliminated on the fly, so L is never stored.  diagU receives the pivots.
__global__ void __launch_bounds__(512) lu_solve_kernel(double* __restrict__ A, int ld, int n, double* __restrict__ diagU,
                                                       const double* __restrict__ sc, const int* __restrict__ idx,
                                                       double* __restrict__ sol, int NEQ, int* __restrict__ info)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* lcol = reinterpret_cast<double*>(smem_raw);          // n
    __shared__ double s_val[16];
    __shared__ int s_idx[16];
    __shared__ int s_piv;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int nb = gridDim.x, bid = blockIdx.x;
    for (int k = 0; k < n; ++k) {
        // every CTA reads column k (rows k..n-1) and finds the pivot redundantly
        double best = -1.0; int bidx = k;
        for (int r = k + tid; r < n; r += nthr) {
            const double v = A[(size_t)r * ld + k];
            lcol[r] = v;
            const double av = fabs(v);
            if (av > best || (av == best && r < bidx)) { best = av; bidx = r; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bidx; }
        __syncthreads();
        if (tid == 0) {
            double bb = s_val[0]; int bi = s_idx[0];
            for (int w = 1; w < (nthr >> 5); ++w)
                if (s_val[w] > bb || (s_val[w] == bb && s_idx[w] < bi)) { bb = s_val[w]; bi = s_idx[w]; }
            s_piv = bi;
            if (!(bb > 0.0) || !isfinite(bb)) { if (bid == 0) atomicCAS(&info[2], 0, k + 1); }
        }
        __syncthreads();
        const int p = s_piv;
        if (tid == 0) { const double t = lcol[k]; lcol[k] = lcol[p]; lcol[p] = t; }
        __syncthreads();
        const double piv = lcol[k];
        const double rp = (piv != 0.0) ? 1.0 / piv : 0.0;
        if (bid == 0 && tid == 0) diagU[k] = piv;
        // own columns c > k (c == n is the rhs): swap rows k and p, eliminate
        for (int c = k + 1 + bid; c <= n; c += nb) {
            // (row n of A holds the rhs mirrored; the rhs column is column n)
            __shared__ double s_akc;
            if (tid == 0) {
                const double akc = A[(size_t)p * ld + c];
                A[(size_t)p * ld + c] = A[(size_t)k * ld + c];
                A[(size_t)k * ld + c] = akc;
                s_akc = akc;
            }
            __syncthreads();
            const double akc = s_akc;
            for (int r = k + 1 + tid; r < n; r += nthr) A[(size_t)r * ld + c] = fma(-lcol[r] * rp, akc, A[(size_t)r * ld + c]);
            __syncthreads();
        }
        grid.sync();
    }
    if (bid == 0) {
        // back substitution U x = y (y = column n), row-oriented dot products
        __shared__ double s_red[16];
        for (int c = tid; c < NEQ; c += nthr) sol[c] = 0.0;
        for (int r = tid; r < n; r += nthr) lcol[r] = 0.0;
        __syncthreads();
        for (int r = n - 1; r >= 0; --r) {
            double s = 0.0;
            for (int c = r + 1 + tid; c < n; c += nthr) s = fma(A[(size_t)r * ld + c], lcol[c], s);
            s = warp_sum(s);
            if (lane == 0) s_red[warp] = s;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < (nthr >> 5); ++w) t += s_red[w];
                const double d = diagU[r];
                lcol[r] = (d != 0.0) ? (A[(size_t)r * ld + n] - t) / d : 0.0;
            }
            __syncthreads();
        }
        for (int c = tid; c < n; c += nthr) sol[idx[c]] = lcol[c] * sc[c];
    }
}
#endif  // SFFTB_TU_CHOL
