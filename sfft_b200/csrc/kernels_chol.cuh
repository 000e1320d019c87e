// kernels_chol.cuh -- dense fp64 Cholesky solve of the (diagonally scaled) normal equations in ONE persistent
// cooperative kernel.
//
// Replaces LSSolver (cuSOLVER getrf/getrs through cupyx, sfft/sfftcore/SFFTSubtract.py:15-23, 397-408; CPU
// np.linalg.solve :744-747).  LHMAT = D^T D / N is symmetric positive definite; the matrix is augmented with the
// right-hand side as row n so that the panel solves deliver the forward substitution L y = b for free.
//
// Right-looking, 64-wide panels, one CTA per SM, two grid barriers per panel:
//   (a) every CTA: 32-row slabs of the panel below the diagonal,  L_ik = A_ik W_k^T   (W_k = L_kk^{-1}, a GEMM)
//   (b) every CTA: 64 x 64 tiles of the trailing matrix,          A_ij -= L_ik L_jk^T;
//       the CTA that owns tile (k+1, k+1) updates it first and immediately factors it (look-ahead), producing
//       L_{k+1,k+1} and W_{k+1} while the other CTAs finish the update.
// All GEMMs run on the fp64 tensor path (mma.sync m8n8k4, DMMA) out of shared memory.  The back substitution
// L^T x = y runs in the same kernel, one grid barrier per block, x_k = W_k^T y_k.
// A non-positive or non-finite pivot sets info[0]; the host then falls back to the pivoted LU (kernels_solve.cuh).
#pragma once
#include "common.cuh"

#define CC_NB 64
#ifndef CC_CTAS_PER_SM
#define CC_CTAS_PER_SM 1
#endif
#define CC_NT 256
#define CC_PITCH 68          // operand pitch (doubles): conflict-free m8n8k4 fragment loads
#define CC_DP 65             // pitch of the diagonal-block work arrays

struct CholArgs {
    double* A; int ld, n, ntot;      // augmented matrix, (n + 1) x ld row-major; ntot = n + 1
    double* W;                       // nblk x 64 x 64: inverses of the diagonal blocks of L
    double* yv;                      // n: running right-hand side of the back substitution
    double* xs;                      // n: solution of the scaled system
    unsigned* bar;                   // grid barrier counter, zeroed before the launch
    int* info;
    const double* sc; const int* idx; double* sol; int NEQ;
    unsigned long long* dbg;         // optional: globaltimer stamps of panel phases (NULL = off)
    int resolve;                     // 1: A already holds L (and W its inverse diagonal blocks) from an earlier call
                                     //    with the same matrix; yv holds a new scaled right-hand side -> substitutions only
};

#ifdef SFFTB_TU_CHOL
__device__ __forceinline__ unsigned long long cc_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CC_STAMP(slot) do { if (a.dbg && tid == 0) a.dbg[(slot)] = cc_gtime(); } while (0)

__device__ __forceinline__ unsigned cc_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void cc_grid_barrier(unsigned* cnt, unsigned& target, unsigned G) {
    target += G;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(cnt, 1u);
        int spins = 0;
        while (cc_ld_acquire(cnt) < target) { if (++spins > (1 << 25)) __trap(); }   // watchdog: never hang the device
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void cc_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// acc[r][c] (8 x 8 tiles) += As(16 rows) * Bs(8 WC rows)^T over K = 64; As / Bs point at the warp's first row
template <int WC>
__device__ __forceinline__ void cc_warp_gemm_nt(const double* As, const double* Bs, double (&acc)[2][WC][2], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 4
    for (int kk = 0; kk < CC_NB; kk += 4) {
        double a[2], b[WC];
        a[0] = As[g * CC_PITCH + kk + t];
        a[1] = As[(8 + g) * CC_PITCH + kk + t];
#pragma unroll
        for (int c = 0; c < WC; ++c) b[c] = Bs[(8 * c + g) * CC_PITCH + kk + t];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < WC; ++c) cc_dmma(acc[r][c][0], acc[r][c][1], a[r], b[c]);
    }
}

// load rows [row0, row0 + nrows) x cols [col0, col0 + 64) of A into S (pitch CC_PITCH), zero outside [.., rmax) x [.., cmax)
__device__ __forceinline__ void cc_load_tile(double* S, const double* __restrict__ A, int ld, int row0, int nrows, int rmax,
                                             int col0, int cmax) {
    for (int idx = threadIdx.x; idx < nrows * CC_NB; idx += CC_NT) {
        const int r = idx >> 6, c = idx & 63;
        const int gr = row0 + r, gc = col0 + c;
        S[r * CC_PITCH + c] = (gr < rmax && gc < cmax) ? A[(size_t)gr * ld + gc] : 0.0;
    }
}

__device__ __forceinline__ double cc_fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // seed error e ~ 2^-23; 1/x = y (1 + e)(1 + e^2) to fp64 accuracy with a three-deep dependent chain
    const double e = fma(-x, y, 1.0);
    const double e2 = e * e;
    const double y1 = fma(y, e, y);
    return fma(y1, e2, y1);
}

// Factor the diagonal block k (already updated): LDL^T-style elimination, register resident, compact code (the
// column loop is NOT unrolled: a fully unrolled version is ~0.5 MB of SASS and runs out of the instruction cache).
// Every thread plays two roles with 16 registers each, chosen so that no register is ever indexed dynamically:
//   * D role, column owned: thread (j, part) = (tid >> 2, tid & 3) holds D[r][j], r = part + 4 m.  The pivot COLUMN c
//     is needed by everybody; its four owners publish all their slots.
//   * E role, row owned:    thread (r, part) holds E[r][j], j = part + 4 m, E = (unit lower factor)^{-1} built by
//     applying the same row operations to an identity.  The pivot ROW c is needed; its four owners publish all slots.
// One barrier per column; column and row strips are double buffered in shared memory.
// Output: L (scaled) into A, W = L^{-1} into a.W.   Dm: 64 x 65 doubles of scratch, bufs: 2 x 128 doubles.
// strips are stored padded (16 bytes after every 16 entries) so that the four 128-byte blocks a warp reads with one
// LDS.128 land in different banks
#define CC_SP(i) ((i) + 2 * ((i) >> 4))
#define CC_STRIP 72                      // padded length of a 64-entry strip
__device__ void cc_potrf_inv_v1(const CholArgs& a, int k, double* Dm, double* bufs) {
    const int tid = threadIdx.x, own = tid >> 2, part = tid & 3;
    const int eown = (own + 32) & 63;                    // E role owner index: the row strip is published by another warp
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    // slot m of a thread is index 16 part + m (row index in the D role, column index in the E role)
    double Vd[16], Ve[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int r = 16 * part + m;
        double v = (r == own) ? 1.0 : 0.0;
        if (r < kb && own < kb && r >= own) v = a.A[(size_t)(k0 + r) * a.ld + k0 + own];
        Vd[m] = v;
        Ve[m] = (r == eown) ? 1.0 : 0.0;
    }
    double2* b2 = reinterpret_cast<double2*>(bufs);      // strips as double2: [buffer][col strip | row strip][36]
    double* pivs = bufs + 4 * CC_STRIP;                  // 64 pivots, then 64 reciprocal square roots
    if (own == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_SP(16 * part) / 2 + m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
    }
    if (eown == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) b2[CC_STRIP / 2 + CC_SP(16 * part) / 2 + m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC_NB; ++c) {
        const double* cb = bufs + (c & 1) * 2 * CC_STRIP;        // pivot column c of D (padded)
        const double2* cb2 = reinterpret_cast<const double2*>(cb) + CC_SP(16 * part) / 2;
        const double2* rb2 = reinterpret_cast<const double2*>(cb + CC_STRIP) + CC_SP(16 * part) / 2;
        double2* cn2 = b2 + ((c + 1) & 1) * CC_STRIP + CC_SP(16 * part) / 2;
        double2* rn2 = cn2 + CC_STRIP / 2;
        double piv = cb[CC_SP(c)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0 && c < kb) atomicCAS(&a.info[0], 0, k0 + c + 1);
            piv = 1.0;
        }
        if (tid == 0) pivs[c] = piv;
        const double rp = cc_fast_rcp(piv);
        if (own > c) {
            // D role: column own > c,  D[r][own] -= D[r][c] D[own][c] / piv  (slots with r < own are don't-care and are
            // updated too)
            const double lj = cb[CC_SP(own)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
            if (own == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m], Vd[2 * m + 1]);
            }
        }
        if (eown > c) {
            // E role: row eown > c,  E[eown][j] -= (D[eown][c] / piv) E[c][j]  (E[c][j] = 0 for j > c)
            const double le = cb[CC_SP(eown)] * rp;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 y = rb2[m];
                Ve[2 * m] = fma(-le, y.x, Ve[2 * m]); Ve[2 * m + 1] = fma(-le, y.y, Ve[2 * m + 1]);
            }
            if (eown == c + 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) rn2[m] = make_double2(Ve[2 * m], Ve[2 * m + 1]);
            }
        }
        __syncthreads();
    }
    // column `own` of D was last touched at step own - 1, so Vd still holds the unscaled pivot column:
    // L = (unscaled columns) diag(piv)^{-1/2},  W = diag(piv)^{-1/2} E
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
}

// ---- blocked diagonal factorisation: 8 x 8 micro blocks, DMMA for everything off the micro diagonal -----------------
// The pivot-to-pivot chain is what bounds the look-ahead path, and a rank-1 formulation spends ~650 cycles per pivot
// on issue and barrier overheads of a whole CTA.  Here a micro block is factored (and inverted) by a single warp
// with shuffles (lane = row, ~130 cycles per pivot, no barrier), the 8-column panel below it is solved with
// L21 = A21 W8^T and the trailing part of the 64 x 64 tile updated with mma.sync m8n8k4 by all warps; the full
// inverse W = L^{-1} follows from the micro inverses by block forward substitution, one column block per warp.
// D: 64 x CC_PITCH doubles (the tile, becomes L), Wf: 64 x CC_PITCH doubles (becomes W), scr: 8 x 64 doubles.
__device__ void cc_potrf_inv(const CholArgs& a, int k, double* D, double* Wf, double* scr, bool preloaded = false) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    for (int idx = tid; idx < CC_NB * CC_NB; idx += CC_NT) {
        const int r = idx >> 6, c = idx & 63;
        if (!preloaded) {
            double v = (r == c) ? 1.0 : 0.0;
            if (r < kb && c < kb && c <= r) v = a.A[(size_t)(k0 + r) * a.ld + k0 + c];
            D[r * CC_PITCH + c] = v;
        }
        Wf[r * CC_PITCH + c] = 0.0;
    }
    __syncthreads();
    for (int jb = 0; jb < 8; ++jb) {
        const int d0 = 8 * jb;
        // (1) micro block: LDL^T elimination with the row operations mirrored on an identity, one warp, lane = row
        if (warp == 0) {
            const int r = lane & 7;
            double av[8], ev[8], pv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { av[j] = (j <= r) ? D[(d0 + r) * CC_PITCH + d0 + j] : 0.0; ev[j] = (j == r) ? 1.0 : 0.0; }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                // (the pivot check is off the dependent chain: a bad pivot only raises the flag, see below)
                const double piv = __shfl_sync(0xffffffffu, av[c], c, 8);
                pv[c] = piv;
                const double m = (r > c) ? -av[c] * cc_fast_rcp(piv) : 0.0;
#pragma unroll
                for (int j = c + 1; j < 8; ++j) { const double dj = __shfl_sync(0xffffffffu, av[c], j, 8); av[j] = fma(m, dj, av[j]); }
#pragma unroll
                for (int j = 0; j <= c; ++j) { const double ej = __shfl_sync(0xffffffffu, ev[j], c, 8); ev[j] = fma(m, ej, ev[j]); }
            }
            double rs[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const bool bad = !(pv[c] > 0.0) || !isfinite(pv[c]);
                if (bad) { if (lane == 0 && d0 + c < kb) atomicCAS(&a.info[0], 0, k0 + d0 + c + 1); pv[c] = 1.0; }
                rs[c] = rsqrt(pv[c]);
            }
            double rr = rs[0];
#pragma unroll
            for (int c = 1; c < 8; ++c) rr = (r == c) ? rs[c] : rr;
            if (lane < 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j <= r) {
                        D[(d0 + r) * CC_PITCH + d0 + j] = (j == r) ? pv[j] * rs[j] : av[j] * rs[j];
                        Wf[(d0 + r) * CC_PITCH + d0 + j] = ev[j] * rr;
                    } else {
                        D[(d0 + r) * CC_PITCH + d0 + j] = 0.0;
                    }
                }
            }
        }
        __syncthreads();
        // (2) panel below the micro block: L21 = A21 W8^T, one 8-row tile per warp
        if (warp < 7 - jb) {
            const int r0 = d0 + 8 + 8 * warp;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; kk += 4)
                cc_dmma(c0, c1, D[(r0 + g) * CC_PITCH + d0 + kk + t], Wf[(d0 + g) * CC_PITCH + d0 + kk + t]);
            __syncwarp();
            D[(r0 + g) * CC_PITCH + d0 + 2 * t] = c0;
            D[(r0 + g) * CC_PITCH + d0 + 2 * t + 1] = c1;
        }
        __syncthreads();
        // (3) trailing update of the tile: D[ti][tj] -= L[ti][jb] L[tj][jb]^T for jb < tj <= ti
        {
            const int nt = 7 - jb;
            const int ntile = nt * (nt + 1) / 2;
            for (int q = warp; q < ntile; q += 8) {
                int ti = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
                while (ti * (ti + 1) / 2 > q) --ti;
                while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
                const int tj = q - ti * (ti + 1) / 2;
                const int ri = d0 + 8 + 8 * ti, rj = d0 + 8 + 8 * tj;
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 8; kk += 4)
                    cc_dmma(c0, c1, D[(ri + g) * CC_PITCH + d0 + kk + t], D[(rj + g) * CC_PITCH + d0 + kk + t]);
                D[(ri + g) * CC_PITCH + rj + 2 * t] -= c0;
                D[(ri + g) * CC_PITCH + rj + 2 * t + 1] -= c1;
            }
        }
        __syncthreads();
    }
    // (4) W = L^{-1}: column block `warp`; W_ij = -W8_i sum_{k=j}^{i-1} L_ik W_kj  for i > j
    {
        const int j = warp, c0b = 8 * j;
        double* S = scr + warp * 64;
        for (int i = j + 1; i < 8; ++i) {
            double s0 = 0.0, s1 = 0.0;
            for (int kb8 = j; kb8 < i; ++kb8) {
#pragma unroll
                for (int kk = 0; kk < 8; kk += 4)
                    cc_dmma(s0, s1, D[(8 * i + g) * CC_PITCH + 8 * kb8 + kk + t], Wf[(8 * kb8 + kk + t) * CC_PITCH + c0b + g]);
            }
            S[g * 8 + 2 * t] = s0; S[g * 8 + 2 * t + 1] = s1;
            __syncwarp();
            double w0 = 0.0, w1 = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; kk += 4)
                cc_dmma(w0, w1, Wf[(8 * i + g) * CC_PITCH + 8 * i + kk + t], S[(kk + t) * 8 + g]);
            Wf[(8 * i + g) * CC_PITCH + c0b + 2 * t] = -w0;
            Wf[(8 * i + g) * CC_PITCH + c0b + 2 * t + 1] = -w1;
            __syncwarp();
        }
    }
    __syncthreads();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    for (int idx = tid; idx < CC_NB * CC_NB; idx += CC_NT) {
        const int r = idx >> 6, c = idx & 63;
        Wk[idx] = (c <= r) ? Wf[r * CC_PITCH + c] : 0.0;
        if (r < kb && c <= r) a.A[(size_t)(k0 + r) * a.ld + k0 + c] = D[r * CC_PITCH + c];
    }
    __syncthreads();
}

// trailing tile (I, J) of panel k: A[i0.., j0..] -= L[i0.., k] L[j0.., k]^T
__device__ void cc_update_tile(const CholArgs& a, int k0, int kb, int i0, int j0, double* As, double* Bs) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cc_load_tile(As, a.A, a.ld, i0, CC_NB, a.ntot, k0, k0 + kb);
    if (j0 != i0) cc_load_tile(Bs, a.A, a.ld, j0, CC_NB, a.n, k0, k0 + kb);
    __syncthreads();
    const double* Bq = (j0 != i0) ? Bs : As;
    double acc[2][4][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[r][c][0] = 0.0; acc[r][c][1] = 0.0; }
    const int wr = (warp >> 1) * 16, wc = (warp & 1) * 32;
    cc_warp_gemm_nt<4>(As + wr * CC_PITCH, Bq + wc * CC_PITCH, acc, lane);
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int gi = i0 + wr + 8 * r + g;
        if (gi >= a.ntot) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gj = j0 + wc + 8 * c + 2 * t;
            double* p = a.A + (size_t)gi * a.ld + gj;
            if (gj < a.n && gj <= gi) p[0] -= acc[r][c][0];
            if (gj + 1 < a.n && gj + 1 <= gi) p[1] -= acc[r][c][1];
        }
    }
    __syncthreads();
}

// Look-ahead path: update the next diagonal tile (rows / columns r1 ..) with panel k and leave it in shared memory
// (Dout, identity padded, upper triangle zero) for cc_potrf_inv -- no round trip through global memory.
__device__ void cc_update_diag_to_smem(const CholArgs& a, int k0, int kb, int r1, double* As, double* Dout) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = (warp >> 1) * 16, wc = (warp & 1) * 32;
    const int kbn = min(CC_NB, a.n - r1);
    const int lrhs = a.n - r1;                 // local row of the right-hand-side row if it falls into this tile (else >= 64)
    cc_load_tile(As, a.A, a.ld, r1, CC_NB, a.ntot, k0, k0 + kb);
    double cv[2][4][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int lr = wr + 8 * r + g;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int lc = wc + 8 * c + 2 * t;
            const double* p = a.A + (size_t)(r1 + lr) * a.ld + r1 + lc;
            const bool in0 = (lr < kbn && lc <= lr) || (lr == lrhs && lc < kbn);
            const bool in1 = (lr < kbn && lc + 1 <= lr) || (lr == lrhs && lc + 1 < kbn);
            cv[r][c][0] = in0 ? p[0] : ((lr == lc) ? 1.0 : 0.0);
            cv[r][c][1] = in1 ? p[1] : ((lr == lc + 1) ? 1.0 : 0.0);
        }
    }
    __syncthreads();
    double acc[2][4][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[r][c][0] = 0.0; acc[r][c][1] = 0.0; }
    cc_warp_gemm_nt<4>(As + wr * CC_PITCH, As + wc * CC_PITCH, acc, lane);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int lr = wr + 8 * r + g;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int lc = wc + 8 * c + 2 * t;
            if (lr == lrhs) {
                // the right-hand-side row rides along in global memory; in the tile it is identity padding
                double* p = a.A + (size_t)(r1 + lr) * a.ld + r1 + lc;
                if (lc < kbn) p[0] = cv[r][c][0] - acc[r][c][0];
                if (lc + 1 < kbn) p[1] = cv[r][c][1] - acc[r][c][1];
                Dout[lr * CC_PITCH + lc] = (lr == lc) ? 1.0 : 0.0;
                Dout[lr * CC_PITCH + lc + 1] = (lr == lc + 1) ? 1.0 : 0.0;
            } else {
                Dout[lr * CC_PITCH + lc] = (lr < kbn && lc <= lr) ? cv[r][c][0] - acc[r][c][0] : cv[r][c][0];
                Dout[lr * CC_PITCH + lc + 1] = (lr < kbn && lc + 1 <= lr) ? cv[r][c][1] - acc[r][c][1] : cv[r][c][1];
            }
        }
    }
    __syncthreads();
}

// ---- pipelined trailing update of a CTA's tile list ---------------------------------------------------------------
// Operands of tile n + 1 stream into the second shared-memory buffer pair with cp.async (8-byte copies, zero fill out of
// range) while tile n runs on the DMMA path; the C tile is prefetched into registers before the MMAs and written back
// after them, so a tile costs about its MMA time instead of three dependent L2 round trips.
__device__ __forceinline__ void cc_cp8(double* dst, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cc_issue_tile(const CholArgs& a, int k0, int kb, int i0, int j0, double* As, double* Bs) {
    for (int idx = threadIdx.x; idx < CC_NB * CC_NB; idx += CC_NT) {
        const int r = idx >> 6, c = idx & 63;
        const bool vc = c < kb;
        const int gi = i0 + r, gj = j0 + r;
        const bool vi = vc && gi < a.ntot, vj = vc && gj < a.n;
        cc_cp8(As + r * CC_PITCH + c, a.A + (size_t)(vi ? gi : 0) * a.ld + k0 + (vc ? c : 0), vi);
        if (j0 != i0) cc_cp8(Bs + r * CC_PITCH + c, a.A + (size_t)(vj ? gj : 0) * a.ld + k0 + (vc ? c : 0), vj);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cc_tile_of(int qq, int TC, int& I, int& J) {
    I = (int)((sqrtf(8.0f * (float)qq + 1.0f) - 1.0f) * 0.5f);
    while (I * (I + 1) / 2 > qq) --I;
    while ((I + 1) * (I + 2) / 2 <= qq) ++I;
    if (I > TC) I = TC;
    J = qq - I * (I + 1) / 2;
}
// tiles qq = q0, q0 + qs, ... < ntile of panel k (row-major lower-triangle numbering, see chol_coop_kernel)
__device__ void cc_update_tiles(const CholArgs& a, int k0, int kb, int r1, int TC, int q0, int qs, int ntile, double* buf) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = (warp >> 1) * 16, wc = (warp & 1) * 32;
    if (q0 >= ntile) return;
    int I, J;
    cc_tile_of(q0, TC, I, J);
    cc_issue_tile(a, k0, kb, r1 + CC_NB * I, r1 + CC_NB * J, buf, buf + CC_NB * CC_PITCH);
    int cur = 0;
    for (int qq = q0; qq < ntile; qq += qs, cur ^= 1) {
        double* As = buf + cur * 2 * CC_NB * CC_PITCH;
        double* Bs = As + CC_NB * CC_PITCH;
        const int i0 = r1 + CC_NB * I, j0 = r1 + CC_NB * J;
        int In = 0, Jn = 0;
        const bool more = qq + qs < ntile;
        if (more) {
            cc_tile_of(qq + qs, TC, In, Jn);
            double* An = buf + (cur ^ 1) * 2 * CC_NB * CC_PITCH;
            cc_issue_tile(a, k0, kb, r1 + CC_NB * In, r1 + CC_NB * Jn, An, An + CC_NB * CC_PITCH);
        }
        // C tile -> registers (independent loads, in flight during the MMAs)
        double cv[2][4][2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int gi = i0 + wr + 8 * r + g;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int gj = j0 + wc + 8 * c + 2 * t;
                const double* p = a.A + (size_t)gi * a.ld + gj;
                cv[r][c][0] = (gi < a.ntot && gj < a.n && gj <= gi) ? p[0] : 0.0;
                cv[r][c][1] = (gi < a.ntot && gj + 1 < a.n && gj + 1 <= gi) ? p[1] : 0.0;
            }
        }
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double* Bq = (j0 != i0) ? Bs : As;
        double acc[2][4][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[r][c][0] = 0.0; acc[r][c][1] = 0.0; }
        cc_warp_gemm_nt<4>(As + wr * CC_PITCH, Bq + wc * CC_PITCH, acc, lane);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int gi = i0 + wr + 8 * r + g;
            if (gi >= a.ntot) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int gj = j0 + wc + 8 * c + 2 * t;
                double* p = a.A + (size_t)gi * a.ld + gj;
                if (gj < a.n && gj <= gi) p[0] = cv[r][c][0] - acc[r][c][0];
                if (gj + 1 < a.n && gj + 1 <= gi) p[1] = cv[r][c][1] - acc[r][c][1];
            }
        }
        __syncthreads();          // the buffer pair just read is the prefetch target of the next iteration
        I = In; J = Jn;
    }
}

__global__ void __launch_bounds__(CC_NT, CC_CTAS_PER_SM) chol_coop_kernel(CholArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + CC_NB * CC_PITCH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, bid = blockIdx.x;
    const int nblk = (a.n + CC_NB - 1) / CC_NB;
    unsigned target = 0;

    if (bid == 0) {
        for (int c = tid; c < a.NEQ; c += CC_NT) a.sol[c] = 0.0;
        if (!a.resolve) cc_potrf_inv(a, 0, As, Bs, Bs + CC_NB * CC_PITCH);
    }
    cc_grid_barrier(a.bar, target, G);

    if (a.resolve) {
        // ---- forward substitution L y = b with the stored factor, b in yv: y_k = W_k (b_k - sum_{j<k} L_kj y_j) ----
        double* ysf = As;                      // 64 staged entries
        double* ykf = As + 64;                 // y of the block solved last
        for (int k = 0; k < nblk; ++k) {
            // y_{k-1} is final (in xs).  The owner of block k applies it to b_k and solves y_k; the others apply it to
            // the rows below block k.
            const int kp0 = (k - 1) * CC_NB, kpb = (k > 0) ? CC_NB : 0;
            const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
            const bool owner = bid == k % G;
            if (owner) {
                if (tid < CC_NB) {
                    double sacc = 0.0;
                    if (tid < kb) {
                        sacc = a.yv[k0 + tid];
                        const double* Lrow = a.A + (size_t)(k0 + tid) * a.ld + kp0;
                        for (int c = 0; c < kpb; ++c) sacc = fma(-Lrow[c], a.xs[kp0 + c], sacc);
                    }
                    ysf[tid] = sacc;
                }
                __syncthreads();
                if (tid < kb) {
                    const double* Wk = a.W + (size_t)k * CC_NB * CC_NB + (size_t)tid * CC_NB;
                    double sacc = 0.0;
                    for (int c = 0; c <= tid; ++c) sacc = fma(Wk[c], ysf[c], sacc);
                    a.xs[k0 + tid] = sacc;
                }
            }
            if ((!owner || G == 1) && kpb > 0 && k0 + kb < a.n) {
                if (tid < CC_NB) ykf[tid] = a.xs[kp0 + tid];
                __syncthreads();
                const int o = (G > 1) ? (bid - (k % G) - 1 + G) % G : 0;
                const int GO = (G > 1) ? G - 1 : 1;
                for (int r = k0 + kb + o * CC_NT + tid; r < a.n; r += GO * CC_NT) {
                    const double* Lrow = a.A + (size_t)r * a.ld + kp0;
                    double sacc = 0.0;
                    for (int c = 0; c < kpb; ++c) sacc = fma(Lrow[c], ykf[c], sacc);
                    a.yv[r] -= sacc;
                }
            }
            cc_grid_barrier(a.bar, target, G);
        }
        // y (in xs) becomes the right-hand side of the back substitution
        for (int c = bid * CC_NT + tid; c < a.n; c += G * CC_NT) a.yv[c] = a.xs[c];
        cc_grid_barrier(a.bar, target, G);
    }

    for (int k = 0; k < (a.resolve ? 0 : nblk); ++k) {
        const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0), r1 = k0 + kb;
        if (bid == 0) CC_STAMP(8 * k + 0);
        // ---- (a) panel below the diagonal block: L_ik = A_ik W_k^T, 32-row slabs ----
        const int nslab = (a.ntot - r1 + 31) / 32;
        if (bid < nslab) {
            const double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
            for (int idx = tid; idx < CC_NB * CC_NB; idx += CC_NT) Bs[(idx >> 6) * CC_PITCH + (idx & 63)] = Wk[idx];
            for (int s = bid; s < nslab; s += G) {
                const int row0 = r1 + 32 * s;
                __syncthreads();
                cc_load_tile(As, a.A, a.ld, row0, 32, a.ntot, k0, k0 + kb);
                __syncthreads();
                double acc[2][2][2];
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c) { acc[r][c][0] = 0.0; acc[r][c][1] = 0.0; }
                const int wr = (warp >> 2) * 16, wc = (warp & 3) * 16;
                cc_warp_gemm_nt<2>(As + wr * CC_PITCH, Bs + wc * CC_PITCH, acc, lane);
                const int g = lane >> 2, t = lane & 3;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int gi = row0 + wr + 8 * r + g;
                    if (gi >= a.ntot) continue;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int cc = wc + 8 * c + 2 * t;
                        double* p = a.A + (size_t)gi * a.ld + k0 + cc;
                        if (cc < kb) p[0] = acc[r][c][0];
                        if (cc + 1 < kb) p[1] = acc[r][c][1];
                    }
                }
            }
        }
        if (bid == 0) CC_STAMP(8 * k + 1);
        cc_grid_barrier(a.bar, target, G);
        if (bid == 0) CC_STAMP(8 * k + 2);
        // ---- (b) trailing update; the owner of tile (0, 0) factors the next diagonal block right away ----
        const int T = (a.ntot - r1 + CC_NB - 1) / CC_NB, TC = (a.n - r1 + CC_NB - 1) / CC_NB;
        if (TC > 0) {
            const int dsg = (k + 1) % G;
            if (bid == dsg) {
                if (G == 1) {
                    for (int I = 1; I < T; ++I)
                        for (int J = 0; J <= min(I, TC - 1); ++J) cc_update_tile(a, k0, kb, r1 + CC_NB * I, r1 + CC_NB * J, As, Bs);
                }
                double* R2 = As + 2 * CC_NB * CC_PITCH;
                cc_update_diag_to_smem(a, k0, kb, r1, As, R2);
                CC_STAMP(8 * k + 3);
                cc_potrf_inv(a, k + 1, R2, R2 + CC_NB * CC_PITCH, Bs, true);
                CC_STAMP(8 * k + 4);
            } else {
                // tiles in row-major order of the lower triangle, (0, 0) excluded: q' = I (I + 1) / 2 + J for I < TC,
                // the extra block row I = TC (right-hand-side row) holds TC tiles
                const int o = (bid - dsg - 1 + G) % G;          // 0 .. G-2
                const int ntile = TC * (TC + 1) / 2 + (T > TC ? TC : 0);
                cc_update_tiles(a, k0, kb, r1, TC, 1 + o, G - 1, ntile, As);
                if (o == 0) CC_STAMP(8 * k + 5);
            }
        }
        cc_grid_barrier(a.bar, target, G);
        if (bid == 0) CC_STAMP(8 * k + 6);
    }

    // ---- back substitution L^T x = y, y = row n ----
    if (!a.resolve) {
        for (int c = bid * CC_NT + tid; c < a.n; c += G * CC_NT) a.yv[c] = a.A[(size_t)a.n * a.ld + c];
        cc_grid_barrier(a.bar, target, G);
    }
    double* ys = As;                       // 64 staged right-hand-side entries
    double* xk = As + 64;                  // 64 entries of the block solved last
    for (int k = nblk - 1; k >= 0; --k) {
        // x_{k+1} is final (none for k == nblk - 1).  The owner of block k applies it to y_k and solves x_k = W_k^T y_k;
        // everyone else applies it to the rows above block k.
        const int kn0 = (k + 1) * CC_NB, knb = (k + 1 < nblk) ? min(CC_NB, a.n - kn0) : 0;
        const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
        const bool owner = bid == k % G;
        if (owner) {
            if (tid < CC_NB) {
                double s = 0.0;
                if (tid < kb) {
                    s = a.yv[k0 + tid];
                    for (int c = 0; c < knb; ++c) s = fma(-a.A[(size_t)(kn0 + c) * a.ld + k0 + tid], a.xs[kn0 + c], s);
                }
                ys[tid] = s;
            }
            __syncthreads();
            if (tid < kb) {
                const double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
                double s = 0.0;
                for (int r = tid; r < kb; ++r) s = fma(Wk[r * CC_NB + tid], ys[r], s);
                a.xs[k0 + tid] = s;
                a.sol[a.idx[k0 + tid]] = s * a.sc[k0 + tid];
            }
        }
        if ((!owner || G == 1) && knb > 0 && k0 > 0) {
            if (tid < CC_NB) xk[tid] = (tid < knb) ? a.xs[kn0 + tid] : 0.0;
            __syncthreads();
            const int o = (G > 1) ? (bid - (k % G) - 1 + G) % G : 0;
            const int GO = (G > 1) ? G - 1 : 1;
            for (int r = o * CC_NT + tid; r < k0; r += GO * CC_NT) {
                double s = 0.0;
                for (int c = 0; c < knb; ++c) s = fma(a.A[(size_t)(kn0 + c) * a.ld + r], xk[c], s);
                a.yv[r] -= s;
            }
        }
        cc_grid_barrier(a.bar, target, G);
    }
}

// ---- triangular substitutions with a stored factor, dataflow style -------------------------------------------------
// Used when the Cholesky factor of LHMAT is already resident (shared-template batches: LHMAT depends on the masked
// template only).  One CTA per 64-row block (block-cyclic if there are more blocks than CTAs); a block waits for the
// blocks it depends on through epoch-tagged flags in global memory instead of grid-wide barriers, and streams the
// 64 x 64 blocks of L it needs through a cp.async double buffer so that they are already in shared memory when the
// vector they multiply arrives.  Forward: y_j = W_j (b_j - sum_{k<j} L_jk y_k); backward: x_j = W_j^T (y_j - sum_{k>j}
// L_kj^T x_k).  Launch cooperatively (all CTAs must be co-resident: they spin on each other's flags).
struct SubstArgs {
    const double* A; int ld, n;          // factor L in the lower triangle (row-major, leading dimension ld)
    const double* W;                     // inverse diagonal blocks
    double* yv;                          // in: scaled right-hand side b;  work: y (forward result)
    double* xs;                          // out: solution of the scaled system
    unsigned* flags;                     // 2 * nblk epoch tags (forward | backward)
    unsigned epoch;
    const double* sc; const int* idx; double* sol; int NEQ;
};

__device__ __forceinline__ void cs_wait_flag(const unsigned* f, unsigned epoch) {
    if (threadIdx.x == 0) {
        int spins = 0;
        while (cc_ld_acquire(f) != epoch) { if (++spins > (1 << 26)) __trap(); }
    }
    __syncthreads();
}
__device__ __forceinline__ void cs_set_flag(unsigned* f, unsigned epoch) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory"); }
}
// stream block (rb, cb) of L (64 x 64, zero outside the matrix) into S (pitch CC_DP)
__device__ __forceinline__ void cs_issue_block(const SubstArgs& a, int rb, int cb, double* S) {
    for (int idx = threadIdx.x; idx < CC_NB * CC_NB; idx += CC_NT) {
        const int r = idx >> 6, c = idx & 63;
        const int gr = rb * CC_NB + r, gc = cb * CC_NB + c;
        const bool v = gr < a.n && gc < a.n;
        const unsigned d = (unsigned)__cvta_generic_to_shared(S + r * CC_DP + c);
        const int sz = v ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(a.A + (size_t)(v ? gr : 0) * a.ld + (v ? gc : 0)), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(CC_NT, 1) chol_subst_kernel(SubstArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Lb = reinterpret_cast<double*>(smem_raw);       // 2 x 64 x CC_DP
    double* Wb = Lb + 2 * CC_NB * CC_DP;                    // 64 x CC_DP
    double* sv = Wb + CC_NB * CC_DP;                        // 64: running right-hand side of the block
    double* vk = sv + 64;                                   // 64: the vector that just arrived
    double* red = vk + 64;                                  // 256 partial sums
    const int tid = threadIdx.x, G = gridDim.x, bid = blockIdx.x;
    const int nblk = (a.n + CC_NB - 1) / CC_NB;
    const int r = tid >> 2, part = tid & 3;                 // 4 threads per row / column, 16 entries each
    unsigned* fF = a.flags;
    unsigned* fB = a.flags + nblk;
    if (bid == 0) for (int c = tid; c < a.NEQ; c += CC_NT) a.sol[c] = 0.0;

    for (int pass = 0; pass < 2; ++pass) {
        const bool fwd = pass == 0;
        // owned blocks in dependency order: increasing for the forward pass, decreasing for the backward pass
        for (int jj = 0; jj < nblk; ++jj) {
            const int j = fwd ? jj : nblk - 1 - jj;
            if (j % G != bid) continue;
            const int j0 = j * CC_NB, jb = min(CC_NB, a.n - j0);
            const int ndep = fwd ? j : nblk - 1 - j;        // number of blocks this one waits for
            // diagonal inverse and the first dependency block stream in while the right-hand side is fetched
            for (int idx = tid; idx < CC_NB * CC_NB; idx += CC_NT) Wb[(idx >> 6) * CC_DP + (idx & 63)] = a.W[(size_t)j * CC_NB * CC_NB + idx];
            if (ndep > 0) { const int k = fwd ? 0 : nblk - 1; if (fwd) cs_issue_block(a, j, k, Lb); else cs_issue_block(a, k, j, Lb); }
            if (!fwd) cs_wait_flag(fF + j, a.epoch);        // y_j of the forward pass (possibly from another CTA)
            if (tid < CC_NB) sv[tid] = (tid < jb) ? a.yv[j0 + tid] : 0.0;
            __syncthreads();
            for (int d = 0; d < ndep; ++d) {
                const int k = fwd ? d : nblk - 1 - d;
                double* Lc = Lb + (d & 1) * CC_NB * CC_DP;
                if (d + 1 < ndep) {
                    const int kn = fwd ? d + 1 : nblk - 2 - d;
                    if (fwd) cs_issue_block(a, j, kn, Lb + ((d + 1) & 1) * CC_NB * CC_DP); else cs_issue_block(a, kn, j, Lb + ((d + 1) & 1) * CC_NB * CC_DP);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                cs_wait_flag((fwd ? fF : fB) + k, a.epoch);
                if (tid < CC_NB) { const int gk = k * CC_NB + tid; vk[tid] = (gk < a.n) ? (fwd ? a.yv[gk] : a.xs[gk]) : 0.0; }
                __syncthreads();
                // forward: s[r] -= sum_c L_jk[r][c] y_k[c];  backward: s[c] -= sum_r L_kj[r][c] x_k[r]
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int q = part * 16 + m;
                    acc = fma(fwd ? Lc[r * CC_DP + q] : Lc[q * CC_DP + r], vk[q], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                __syncthreads();
                if (part == 0) sv[r] -= acc;
                __syncthreads();
            }
            // y_j = W_j s (forward) / x_j = W_j^T s (backward)
            {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int q = part * 16 + m;
                    acc = fma(fwd ? Wb[r * CC_DP + q] : Wb[q * CC_DP + r], sv[q], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                if (part == 0 && r < jb) {
                    if (fwd) a.yv[j0 + r] = acc;
                    else { a.xs[j0 + r] = acc; a.sol[a.idx[j0 + r]] = acc * a.sc[j0 + r]; }
                }
            }
            cs_set_flag((fwd ? fF : fB) + j, a.epoch);
        }
    }
}


// ---- the same substitutions with the vector and its ready tag in ONE 16-byte message (two 8-byte packets
// {value word | epoch}, each stored atomically): a consumer polls the data itself, so the critical path from "y_k
// computed" to "y_k in the consumer's shared memory" is one store and one load instead of store + fence + flag store
// + flag load + data load.  The running right-hand side of a block lives in registers, the matrix-vector products
// run on four independent accumulators per thread, and there is one CTA barrier per dependency block.
struct SubstArgs2 {
    const double* A; int ld, n;
    const double* W;
    double* yv;                          // in: scaled right-hand side b;  out: y (plain copy, read back by the owner CTA)
    double* xs;                          // out: solution of the scaled system
    ulonglong2* msg;                     // 2 * nblk * 64 messages (forward | backward)
    unsigned epoch;
    const double* sc; const int* idx; double* sol; int NEQ;
};

__device__ __forceinline__ void cs2_send(ulonglong2* m, double v, unsigned epoch) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), e = (unsigned long long)epoch << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(m), "l"((b & 0xffffffffull) | e), "l"((b >> 32) | e) : "memory");
}
__device__ __forceinline__ double cs2_recv(const ulonglong2* m, unsigned epoch) {
    unsigned long long x, y;
    int spins = 0;
    while (true) {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(m) : "memory");
        if ((unsigned)(x >> 32) == epoch && (unsigned)(y >> 32) == epoch) break;
        if (++spins > (1 << 26)) __trap();
    }
    return __longlong_as_double((long long)((x & 0xffffffffull) | (y << 32)));
}
__device__ __forceinline__ void cs2_issue_block(const SubstArgs2& a, int rb, int cb, double* S) {
    for (int idx = threadIdx.x; idx < CC_NB * CC_NB; idx += CC_NT) {
        const int r = idx >> 6, c = idx & 63;
        const int gr = rb * CC_NB + r, gc = cb * CC_NB + c;
        const bool v = gr < a.n && gc < a.n;
        const unsigned d = (unsigned)__cvta_generic_to_shared(S + r * CC_DP + c);
        const int sz = v ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(a.A + (size_t)(v ? gr : 0) * a.ld + (v ? gc : 0)), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(CC_NT, 1) chol_subst2_kernel(SubstArgs2 a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Lb = reinterpret_cast<double*>(smem_raw);       // 2 x 64 x CC_DP
    double* Wb = Lb + 2 * CC_NB * CC_DP;                    // 64 x CC_DP
    double* sv = Wb + CC_NB * CC_DP;                        // 64: right-hand side of the block for the W product
    double* vk = sv + 64;                                   // 2 x 64: the vectors that arrived (double buffered)
    const int tid = threadIdx.x, G = gridDim.x, bid = blockIdx.x;
    const int nblk = (a.n + CC_NB - 1) / CC_NB;
    const int r = tid >> 2, part = tid & 3;                 // 4 threads per row / column, 16 entries each
    if (bid == 0) {
        for (int c = tid; c < a.NEQ; c += CC_NT) a.sol[c] = 0.0;
        __threadfence();                                    // ordered before anything another CTA does after hearing from block 0
    }
    for (int pass = 0; pass < 2; ++pass) {
        const bool fwd = pass == 0;
        ulonglong2* mout = a.msg + (size_t)pass * nblk * CC_NB;
        for (int jj = 0; jj < nblk; ++jj) {
            const int j = fwd ? jj : nblk - 1 - jj;
            if (j % G != bid) continue;
            const int j0 = j * CC_NB, jb = min(CC_NB, a.n - j0);
            const int ndep = fwd ? j : nblk - 1 - j;
            __syncthreads();                                // previous block of this CTA is done with the buffers
            if (ndep > 0) { const int k = fwd ? 0 : nblk - 1; if (fwd) cs2_issue_block(a, j, k, Lb); else cs2_issue_block(a, k, j, Lb); }
            for (int idx = tid; idx < CC_NB * CC_NB; idx += CC_NT) Wb[(idx >> 6) * CC_DP + (idx & 63)] = a.W[(size_t)j * CC_NB * CC_NB + idx];
            double s_reg = (part == 0 && r < jb) ? a.yv[j0 + r] : 0.0;   // b_j (forward) / y_j written by this CTA (backward)
            for (int d = 0; d < ndep; ++d) {
                const int k = fwd ? d : nblk - 1 - d;
                double* Lc = Lb + (d & 1) * CC_NB * CC_DP;
                double* vc = vk + (d & 1) * 64;
                if (tid < CC_NB) { const int gk = k * CC_NB + tid; vc[tid] = (gk < a.n) ? cs2_recv(mout + gk, a.epoch) : 0.0; }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();                            // L block d and vector d are in shared memory; step d - 1 is finished
                if (d + 1 < ndep) {
                    const int kn = fwd ? d + 1 : nblk - 2 - d;
                    if (fwd) cs2_issue_block(a, j, kn, Lb + ((d + 1) & 1) * CC_NB * CC_DP); else cs2_issue_block(a, kn, j, Lb + ((d + 1) & 1) * CC_NB * CC_DP);
                }
                // forward: s[r] -= sum_c L_jk[r][c] y_k[c];  backward: s[c] -= sum_r L_kj[r][c] x_k[r]
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int m = 0; m < 16; m += 4) {
                    const int q = part * 16 + m;
                    a0 = fma(fwd ? Lc[r * CC_DP + q] : Lc[q * CC_DP + r], vc[q], a0);
                    a1 = fma(fwd ? Lc[r * CC_DP + q + 1] : Lc[(q + 1) * CC_DP + r], vc[q + 1], a1);
                    a2 = fma(fwd ? Lc[r * CC_DP + q + 2] : Lc[(q + 2) * CC_DP + r], vc[q + 2], a2);
                    a3 = fma(fwd ? Lc[r * CC_DP + q + 3] : Lc[(q + 3) * CC_DP + r], vc[q + 3], a3);
                }
                double acc = (a0 + a1) + (a2 + a3);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                s_reg -= acc;
            }
            if (part == 0) sv[r] = s_reg;
            __syncthreads();
            // y_j = W_j s (forward) / x_j = W_j^T s (backward)
            {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int m = 0; m < 16; m += 4) {
                    const int q = part * 16 + m;
                    a0 = fma(fwd ? Wb[r * CC_DP + q] : Wb[q * CC_DP + r], sv[q], a0);
                    a1 = fma(fwd ? Wb[r * CC_DP + q + 1] : Wb[(q + 1) * CC_DP + r], sv[q + 1], a1);
                    a2 = fma(fwd ? Wb[r * CC_DP + q + 2] : Wb[(q + 2) * CC_DP + r], sv[q + 2], a2);
                    a3 = fma(fwd ? Wb[r * CC_DP + q + 3] : Wb[(q + 3) * CC_DP + r], sv[q + 3], a3);
                }
                double acc = (a0 + a1) + (a2 + a3);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                if (part == 0 && r < jb) {
                    cs2_send(mout + j0 + r, acc, a.epoch);
                    if (fwd) a.yv[j0 + r] = acc;
                    else { a.xs[j0 + r] = acc; a.sol[a.idx[j0 + r]] = acc * a.sc[j0 + r]; }
                }
            }
        }
    }
}
#endif  // SFFTB_TU_CHOL
