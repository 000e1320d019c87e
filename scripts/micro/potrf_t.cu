// standalone timing of cc_potrf_inv (one CTA), cycles per call
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../sfft_b200/csrc/kernels_chol.cuh"
__device__ void pvt(const CholArgs& a, int k, double* Dm, double* bufs, long long* st) {
    st[0] = clock64();
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
    st[1] = clock64();
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
    st[2] = clock64();
    if (tid < CC_NB) pivs[64 + tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
    st[3] = clock64();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    const double sd = pivs[64 + own], se = pivs[64 + eown];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = 16 * part + m;
        Wk[eown * CC_NB + i] = (i <= eown) ? Ve[m] * se : 0.0;
        if (i < kb && own < kb && i >= own) a.A[(size_t)(k0 + i) * a.ld + k0 + own] = (i == own) ? pivs[own] * sd : Vd[m] * sd;
    }
    __syncthreads();
    st[4] = clock64();
}


__global__ void __launch_bounds__(CC_NT, 1) kbench(CholArgs a, long long* cyc, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + CC_NB * CC_PITCH;
    for (int rep = 0; rep < reps; ++rep) {
        // restore the matrix
        for (int i = threadIdx.x; i < 64 * 64; i += CC_NT) a.A[(i >> 6) * a.ld + (i & 63)] = a.yv[i];
        __syncthreads();
        long long t0 = clock64();
        long long st[5]; pvt(a, 0, As, Bs, st); if (threadIdx.x == 0 && rep == 3) { for (int i = 0; i < 5; ++i) cyc[8 + i] = st[i] - st[0]; }
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[rep] = t1 - t0;
    }
}
int main() {
    const int n = 64, ld = 65;
    std::vector<double> M(64 * 64), A(65 * 65, 0.0);
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k < 64; ++k) s += sin(0.1 * (i + 1) * (k + 1)) * sin(0.1 * (j + 1) * (k + 1)); M[i * 64 + j] = s + (i == j ? 10.0 : 0.0); }
    CholArgs a; memset(&a, 0, sizeof a);
    double *dA, *dW, *dM; int* info; long long* cyc;
    cudaMalloc(&dA, sizeof(double) * 65 * 65); cudaMalloc(&dW, sizeof(double) * 4096); cudaMalloc(&dM, sizeof(double) * 4096);
    cudaMalloc(&info, 16); cudaMemset(info, 0, 16); cudaMalloc(&cyc, 8 * 16);
    cudaMemcpy(dM, M.data(), sizeof(double) * 4096, cudaMemcpyHostToDevice);
    a.A = dA; a.ld = ld; a.n = n; a.ntot = n + 1; a.W = dW; a.yv = dM; a.info = info;
    size_t sm = sizeof(double) * 2 * CC_NB * CC_PITCH;
    cudaFuncSetAttribute(kbench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    kbench<<<1, CC_NT, sm>>>(a, cyc, 8);
    long long h[8]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("status %s; cycles per potrf_inv:", cudaGetErrorString(cudaDeviceSynchronize()));
    for (int i = 0; i < 8; ++i) printf(" %lld", h[i]);
    printf("\n"); long long h2[5]; cudaMemcpy(h2, cyc + 8, sizeof h2, cudaMemcpyDeviceToHost); printf("phases: loaded %lld loop_end %lld pivs %lld end %lld\n", h2[1], h2[2], h2[3], h2[4]);
    // check L L^T = M
    std::vector<double> L(65 * 65), W(4096); cudaMemcpy(L.data(), dA, sizeof(double) * 65 * 65, cudaMemcpyDeviceToHost); cudaMemcpy(W.data(), dW, sizeof(double) * 4096, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < 64; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += L[i * 65 + k] * L[j * 65 + k]; e1 = fmax(e1, fabs(s - M[i * 64 + j])); }
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k < 64; ++k) s += (k <= i ? W[i * 64 + k] : 0.0) * (j <= k ? L[k * 65 + j] : 0.0); e2 = fmax(e2, fabs(s - (i == j))); }
    printf("max |LL^T - M| = %.3e, max |W L - I| = %.3e\n", e1, e2);
}
