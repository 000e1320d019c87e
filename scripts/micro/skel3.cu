#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_seed(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
template <int V> __global__ void __launch_bounds__(256, 1) k(double* out, long long* cyc) {
    __shared__ __align__(16) double buf[2][160];
    const int tid = threadIdx.x, own = tid & 63, part = tid >> 6;
    for (int i = tid; i < 320; i += 256) (&buf[0][0])[i] = 1.0 + i * 1e-3;
    __syncthreads();
    double Vd[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) Vd[m] = tid + m;
    double acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int c = 0; c < 64; ++c) {
        const double* cb = buf[c & 1];
        double piv = cb[c];
        double rp = piv;
        if (V == 1) rp = rcp_seed(piv);                                   // MUFU only
        if (V == 2) { double y = rcp_seed(piv); rp = fma(fma(-piv, y, 1.0), y, y); }      // + 1 Newton
        if (V == 3) { double y = piv * 0.5; y = fma(fma(-piv, y, 1.0), y, y); rp = fma(fma(-piv, y, 1.0), y, y); }   // 4 dependent FMA, no MUFU
        if (V == 4 || V == 5) {                                            // update without rcp
            const double lj = cb[own] * rp;
            const double2* cb2 = reinterpret_cast<const double2*>(cb) + 8 * part;
            if (V == 4 || own > c) {
#pragma unroll
                for (int m = 0; m < 8; ++m) { const double2 x = cb2[m]; Vd[2 * m] = fma(-x.x, lj, Vd[2 * m]); Vd[2 * m + 1] = fma(-x.y, lj, Vd[2 * m + 1]); }
            }
        }
        acc += rp;
        __syncthreads();
    }
    long long t1 = clock64();
    double s = acc;
#pragma unroll
    for (int m = 0; m < 16; ++m) s += Vd[m];
    out[tid] = s;
    if (tid == 0) cyc[V] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 256); cudaMalloc(&cyc, 8 * 8);
    k<0><<<1, 256>>>(out, cyc); k<1><<<1, 256>>>(out, cyc); k<2><<<1, 256>>>(out, cyc); k<3><<<1, 256>>>(out, cyc); k<4><<<1, 256>>>(out, cyc); k<5><<<1, 256>>>(out, cyc);
    long long h[8]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%s: cycles/iter: base %.1f | MUFU.RCP64H %.1f | MUFU+1 Newton %.1f | 4 dep FMA %.1f | update all threads %.1f | update (own>c) %.1f\n", cudaGetErrorString(cudaDeviceSynchronize()), h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0);
}
