#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); y = fma(fma(-x, y, 1.0), y, y); y = fma(fma(-x, y, 1.0), y, y); return y; }
template <int V> __global__ void __launch_bounds__(256, 1) k(double* out, long long* cyc) {
    __shared__ __align__(16) double buf[2][160];
    const int tid = threadIdx.x, own = tid & 63, part = tid >> 6;
    for (int i = tid; i < 320; i += 256) (&buf[0][0])[i] = 1.0 + i * 1e-3;
    __syncthreads();
    double Vd[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) Vd[m] = tid + m;
    long long t0 = clock64();
#pragma unroll 1
    for (int c = 0; c < 64; ++c) {
        const double* cb = buf[c & 1];
        double* cn = buf[(c + 1) & 1];
        double piv = cb[c];
        if (!(piv > 0.0) || !isfinite(piv)) piv = 1.0;
        const double rp = fast_rcp(piv);
        if (V >= 1 && own > c) {
            const double lj = cb[own] * rp;
            const double2* cb2 = reinterpret_cast<const double2*>(cb) + 8 * part;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double2 x = cb2[m];
                Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]);
            }
        }
        if (V >= 2 && own == c + 1) {
            double2* cn2 = reinterpret_cast<double2*>(cn) + 8 * part;
#pragma unroll
            for (int m = 0; m < 8; ++m) cn2[m] = make_double2(Vd[2 * m] * 1e-3 + 1.0, Vd[2 * m + 1] * 1e-3 + 1.0);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int m = 0; m < 16; ++m) s += Vd[m];
    out[tid] = s;
    if (tid == 0) cyc[V] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 256); cudaMalloc(&cyc, 8 * 8);
    k<0><<<1, 256>>>(out, cyc); k<1><<<1, 256>>>(out, cyc); k<2><<<1, 256>>>(out, cyc);
    long long h[8]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%s: per-iteration cycles: skeleton %.1f | +D update %.1f | +publish %.1f\n", cudaGetErrorString(cudaDeviceSynchronize()), h[0] / 64.0, h[1] / 64.0, h[2] / 64.0);
}
