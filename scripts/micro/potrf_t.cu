// standalone timing of cc_potrf_inv (one CTA), cycles per call
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../sfft_b200/csrc/kernels_chol.cuh"
__device__ void pvt(const CholArgs& a, int k, double* D, double* Wf, double* scr, long long* st) {
    long long tA = clock64(), t1 = 0, t2 = 0, t3 = 0; st[0] = tA;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int k0 = k * CC_NB, kb = min(CC_NB, a.n - k0);
    for (int idx = tid; idx < CC_NB * CC_NB; idx += CC_NT) {
        const int r = idx >> 6, c = idx & 63;
        double v = (r == c) ? 1.0 : 0.0;
        if (r < kb && c < kb && c <= r) v = a.A[(size_t)(k0 + r) * a.ld + k0 + c];
        D[r * CC_PITCH + c] = v;
        Wf[r * CC_PITCH + c] = 0.0;
    }
    __syncthreads();
    st[1] = clock64() - tA;
    for (int jb = 0; jb < 8; ++jb) {
        const int d0 = 8 * jb; long long q0 = clock64();
        // (1) micro block: LDL^T elimination with the row operations mirrored on an identity, one warp, lane = row
        if (warp == 0) {
            const int r = lane & 7;
            double av[8], ev[8], pv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { av[j] = (j <= r) ? D[(d0 + r) * CC_PITCH + d0 + j] : 0.0; ev[j] = (j == r) ? 1.0 : 0.0; }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double piv = __shfl_sync(0xffffffffu, av[c], c, 8);
                const bool bad = !(piv > 0.0) || !isfinite(piv);
                if (bad) { if (lane == 0 && d0 + c < kb) atomicCAS(&a.info[0], 0, k0 + d0 + c + 1); piv = 1.0; }
                pv[c] = piv;
                const double m = (r > c) ? -av[c] * cc_fast_rcp(piv) : 0.0;
#pragma unroll
                for (int j = c + 1; j < 8; ++j) { const double dj = __shfl_sync(0xffffffffu, av[c], j, 8); av[j] = fma(m, dj, av[j]); }
#pragma unroll
                for (int j = 0; j <= c; ++j) { const double ej = __shfl_sync(0xffffffffu, ev[j], c, 8); ev[j] = fma(m, ej, ev[j]); }
            }
            double rs[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) rs[c] = rsqrt(pv[c]);
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
        __syncthreads(); long long q1 = clock64(); t1 += q1 - q0;
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
        __syncthreads(); long long q2 = clock64(); t2 += q2 - q1;
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
        __syncthreads(); t3 += clock64() - q2;
    }
    st[2] = t1; st[3] = t2; st[4] = t3; long long tB = clock64();
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
    __syncthreads(); st[5] = clock64() - tB; tB = clock64();
    double* Wk = a.W + (size_t)k * CC_NB * CC_NB;
    for (int idx = tid; idx < CC_NB * CC_NB; idx += CC_NT) {
        const int r = idx >> 6, c = idx & 63;
        Wk[idx] = (c <= r) ? Wf[r * CC_PITCH + c] : 0.0;
        if (r < kb && c <= r) a.A[(size_t)(k0 + r) * a.ld + k0 + c] = D[r * CC_PITCH + c];
    }
    __syncthreads(); st[6] = clock64() - tB;
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
        long long st[8]; pvt(a, 0, As, Bs, Bs + CC_NB * CC_PITCH, st); if (threadIdx.x == 0 && rep == 3) { for (int i = 0; i < 7; ++i) cyc[8 + i] = st[i]; }
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
    size_t sm = sizeof(double) * (2 * CC_NB * CC_PITCH + 512);
    cudaFuncSetAttribute(kbench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    kbench<<<1, CC_NT, sm>>>(a, cyc, 8);
    long long h[8]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("status %s; cycles per potrf_inv:", cudaGetErrorString(cudaDeviceSynchronize()));
    for (int i = 0; i < 8; ++i) printf(" %lld", h[i]);
    printf("\n"); long long h2[7]; cudaMemcpy(h2, cyc + 8, sizeof h2, cudaMemcpyDeviceToHost); printf("phases: load %lld | micro %lld | panel %lld | trailing %lld | inverse %lld | store %lld\n", h2[1], h2[2], h2[3], h2[4], h2[5], h2[6]);
    // check L L^T = M
    std::vector<double> L(65 * 65), W(4096); cudaMemcpy(L.data(), dA, sizeof(double) * 65 * 65, cudaMemcpyDeviceToHost); cudaMemcpy(W.data(), dW, sizeof(double) * 4096, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < 64; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += L[i * 65 + k] * L[j * 65 + k]; e1 = fmax(e1, fabs(s - M[i * 64 + j])); }
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k < 64; ++k) s += (k <= i ? W[i * 64 + k] : 0.0) * (j <= k ? L[k * 65 + j] : 0.0); e2 = fmax(e2, fabs(s - (i == j))); }
    printf("max |LL^T - M| = %.3e, max |W L - I| = %.3e\n", e1, e2);
}
