// conversion throughput on B200: F2F.F64.F32, F2F.F32.F64, I2F.F64.S32
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int OP> __global__ void k(double* out, float a, int b) {
    float xf[8]; double xd[8]; int xi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { xf[i] = threadIdx.x * 1e-3f + i + a; xd[i] = xf[i]; xi[i] = threadIdx.x + i + b; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) { xd[i] = (double)xf[i]; xf[i] = __int_as_float(__float_as_int(xf[i]) + (int)__double2hiint(xd[i])); }       // F2F.F64.F32 + int ops
            if (OP == 1) { xf[i] = (float)xd[i]; xd[i] = __hiloint2double(__double2hiint(xd[i]) + __float_as_int(xf[i]), __double2loint(xd[i])); }
            if (OP == 2) { xd[i] = (double)xi[i]; xi[i] += __double2hiint(xd[i]); }
            if (OP == 3) { xi[i] = xi[i] * 3 + __float_as_int(xf[i]); xf[i] = __int_as_float(xi[i] & 0x3fffffff); }                    // baseline int-only loop
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += xd[i] + xf[i] + xi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, sizeof(double) * nsm * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"F2F.F64.F32", "F2F.F32.F64", "I2F.F64.S32", "int-only"};
    for (int op = 0; op < 4; ++op) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            switch (op) {
                case 0: k<0><<<nsm, 1024>>>(out, 1.f, 1); break;
                case 1: k<1><<<nsm, 1024>>>(out, 1.f, 1); break;
                case 2: k<2><<<nsm, 1024>>>(out, 1.f, 1); break;
                case 3: k<3><<<nsm, 1024>>>(out, 1.f, 1); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double n = 1024.0 * ITER * 8;
        printf("%-12s %8.3f ms  %7.2f conv/clk/SM\n", names[op], best, n / (best * 1e-3 * khz * 1e3));
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
