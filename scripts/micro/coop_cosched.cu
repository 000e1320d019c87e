// Are two cooperative grids that cannot be co-resident TOGETHER gang-scheduled (run one after the other) or interleaved
// (deadlock)?  Two streams, each launches a cooperative kernel with one CTA per SM (shared memory sized so that only one
// CTA fits an SM) that crosses a grid barrier many times.  A watchdog turns a deadlock into a trap.
// nvcc -gencode arch=compute_100a,code=sm_100a -o coop_cosched coop_cosched.cu && timeout 120 ./coop_cosched
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin_kernel(unsigned* cnt, int iters, unsigned G, long long* out) {
    extern __shared__ char pad[];
    unsigned target = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        target += G;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(cnt, 1u);
            long long spins = 0;
            while (*(volatile unsigned*)cnt < target) { if (++spins > (1ll << 27)) { printf("deadlock: block %d iter %d\n", blockIdx.x, it); __trap(); } }
            __threadfence();
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
    pad[threadIdx.x] = 0;
}
int main() {
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = 120 * 1024;
    cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned* cnt; long long* out; cudaMalloc(&cnt, 8); cudaMalloc(&out, 16); cudaMemset(cnt, 0, 8);
    cudaStream_t s[2]; cudaStreamCreate(&s[0]); cudaStreamCreate(&s[1]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {           // 0: one after the other on one stream, 1: two streams at once
        cudaMemset(cnt, 0, 8);
        cudaEventRecord(e0, 0);
        for (int k = 0; k < 2; ++k) {
            unsigned* c = cnt + k; int iters = 20000; unsigned G = nsm; long long* o = out + k;
            void* args[] = {&c, &iters, &G, &o};
            cudaError_t e = cudaLaunchCooperativeKernel((void*)spin_kernel, dim3(nsm), dim3(128), args, smem, s[mode ? k : 0]);
            if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaError_t e = cudaDeviceSynchronize();
        cudaEventRecord(e1, 0); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        printf("mode %d (%s): %s, %.2f ms\n", mode, mode ? "two streams" : "one stream", cudaGetErrorString(e), ms);
        if (e != cudaSuccess) return 2;
    }
    printf("RESULT: two full-grid cooperative kernels on two streams completed (gang-scheduled)\n");
    return 0;
}
