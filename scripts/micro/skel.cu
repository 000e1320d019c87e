#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); y = fma(fma(-x, y, 1.0), y, y); y = fma(fma(-x, y, 1.0), y, y); return y; }
template <int V> __global__ void __launch_bounds__(256, 1) k(double* out, long long* cyc) {
    __shared__ double buf[256];
    const int tid = threadIdx.x;
    buf[tid] = 1.0 + tid * 1e-3;
    __syncthreads();
    double acc = 0.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int c = 0; c < 64; ++c) {
        double piv = buf[c];
        if (V >= 2) { if (!(piv > 0.0) || !isfinite(piv)) piv = 1.0; }
        double rp = piv;
        if (V >= 3) rp = fast_rcp(piv);
        if (V >= 4) { if (tid == 0) buf[128 + c] = piv; }
        acc += rp;
        if (V >= 5) { if ((tid >> 2) == c + 1) buf[64 + (tid & 3) * 16] = acc; }      // a dependent publish by 4 threads
        if (V >= 1) __syncthreads();
    }
    long long t1 = clock64();
    out[tid] = acc;
    if (tid == 0) cyc[V] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 256); cudaMalloc(&cyc, 8 * 8);
    k<0><<<1, 256>>>(out, cyc); k<1><<<1, 256>>>(out, cyc); k<2><<<1, 256>>>(out, cyc); k<3><<<1, 256>>>(out, cyc); k<4><<<1, 256>>>(out, cyc); k<5><<<1, 256>>>(out, cyc);
    long long h[8]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%s: per-iteration cycles: lds only %.1f | +barrier %.1f | +check %.1f | +rcp %.1f | +pivs store %.1f | +publish %.1f\n", cudaGetErrorString(cudaDeviceSynchronize()),
           h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0);
}
