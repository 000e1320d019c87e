// standalone timing of cc_potrf_inv (one CTA), cycles per call
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../sfft_b200/csrc/kernels_chol.cuh"
__global__ void __launch_bounds__(CC_NT, 1) kbench(CholArgs a, long long* cyc, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + CC_NB * CC_PITCH;
    for (int rep = 0; rep < reps; ++rep) {
        // restore the matrix
        for (int i = threadIdx.x; i < 64 * 64; i += CC_NT) a.A[(i >> 6) * a.ld + (i & 63)] = a.yv[i];
        __syncthreads();
        long long t0 = clock64();
        cc_potrf_inv(a, 0, As, Bs, Bs + CC_NB * CC_PITCH);
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
    printf("\n");
    // check L L^T = M
    std::vector<double> L(65 * 65), W(4096); cudaMemcpy(L.data(), dA, sizeof(double) * 65 * 65, cudaMemcpyDeviceToHost); cudaMemcpy(W.data(), dW, sizeof(double) * 4096, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < 64; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += L[i * 65 + k] * L[j * 65 + k]; e1 = fmax(e1, fabs(s - M[i * 64 + j])); }
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k < 64; ++k) s += (k <= i ? W[i * 64 + k] : 0.0) * (j <= k ? L[k * 65 + j] : 0.0); e2 = fmax(e2, fabs(s - (i == j))); }
    printf("max |LL^T - M| = %.3e, max |W L - I| = %.3e\n", e1, e2);
}
