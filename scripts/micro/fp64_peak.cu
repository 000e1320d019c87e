// fp64 pipe microbenchmark for B200: DFMA, DADD, DMUL, FFMA and DMMA (mma.sync m8n8k4 f64) throughput per SM per clock.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int OP> __global__ void k(double* out, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    float xf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xf[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            if (OP == 1) x[i] = x[i] + a;
            if (OP == 2) x[i] = x[i] * a;
            if (OP == 3) xf[i] = fmaf(xf[i], (float)a, (float)b);
            if (OP == 5) { x[i] = fma(x[i], a, b); xf[i] = fmaf(xf[i], (float)a, (float)b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + xf[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void kmma(double* out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: DMMA + DFMA interleaved
__global__ void kmix(double* out, double a, double b) {
    double c[4][2], x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            x[2 * i] = fma(x[2 * i], a, b); x[2 * i + 1] = fma(x[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"DFMA", "DADD", "DMUL", "FFMA", "DMMA m8n8k4", "DFMA+FFMA", "DMMA+2DFMA"};
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int op = 0; op < 7; ++op) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                dim3 g(nsm), b(warps * 32);
                switch (op) {
                    case 0: k<0><<<g, b>>>(out, 1.0000001, 1e-9); break;
                    case 1: k<1><<<g, b>>>(out, 1.0000001, 1e-9); break;
                    case 2: k<2><<<g, b>>>(out, 1.0000001, 1e-9); break;
                    case 3: k<3><<<g, b>>>(out, 1.0000001, 1e-9); break;
                    case 4: kmma<<<g, b>>>(out, 1.0000001, 1e-9); break;
                    case 5: k<5><<<g, b>>>(out, 1.0000001, 1e-9); break;
                    case 6: kmix<<<g, b>>>(out, 1.0000001, 1e-9); break;
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            double inst = (double)warps * 32 * ITER * 8;            // thread-level ops per SM (op 5: x2 listed separately)
            double fmas = inst;
            if (op == 4) fmas = (double)warps * ITER * 8 * 256;     // m8n8k4 = 256 FMA per warp instr
            if (op == 6) fmas = (double)warps * ITER * (4 * 256 + 8 * 32);
            double clk = best * 1e-3 * khz * 1e3;
            printf("warps/SM %2d %-12s %8.3f ms  %7.2f fma(or op)/clk/SM (at nominal %d MHz)  -> %.1f TFLOP/s\n", warps, names[op], best, fmas / clk, khz / 1000,
                   2 * fmas * nsm / (best * 1e-3) / 1e12);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
