// intrinsic cost of one 256-point forward transform by one warp (8-values-per-thread engine), nw warps per CTA
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../sfft_b200/csrc/fft_vpt.cuh"
__constant__ int dummy;
template <int MODE> __global__ void __launch_bounds__(512, 1) k(VTabs vt_g, long long* cyc, double* out, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* planes = reinterpret_cast<cd*>(smem_raw);
    cd* tw8 = planes + 16 * 288; cd* tw64 = tw8 + 56;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 56; i += blockDim.x) tw8[i] = vt_g.t8_8[i];
    for (int i = tid; i < 192; i += blockDim.x) tw64[i] = vt_g.t64_4[i];
    VTabs vt = vt_g; if (MODE == 1) { vt.t8_8 = tw8; vt.t64_4 = tw64; }
    cd* plane = planes + warp * 288;
    for (int i = lane; i < 288; i += 32) plane[i] = cmake(1.0 + i * 1e-3, 0.5 - i * 1e-3);
    __syncthreads();
    long long t0 = clock64();
    cd acc = cmake(0, 0);
    for (int rep = 0; rep < reps; ++rep) {
        cd v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = plane[VPAD(lane + 32 * q)];
        __syncwarp();
        vfft<256>(v, plane, lane, vt, -1.0, 0);
#pragma unroll
        for (int q = 0; q < 8; ++q) plane[VPAD(lane + 32 * q)] = cscale(v[q], 1.0 / 256.0);
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    out[tid] = plane[lane].x + acc.x;
}
static void table(int Ns, int R, cd** out) {
    std::vector<cd> h((size_t)(R - 1) * Ns);
    for (int r = 1; r < R; ++r) for (int kk = 0; kk < Ns; ++kk) { double ang = 2 * M_PI * r * kk / ((double)Ns * R); h[(size_t)(r - 1) * Ns + kk] = cmake(cos(ang), -sin(ang)); }
    cudaMalloc(out, sizeof(cd) * h.size()); cudaMemcpy(*out, h.data(), sizeof(cd) * h.size(), cudaMemcpyHostToDevice);
}
int main() {
    VTabs vt; cd *a, *b, *c, *d, *e; table(8, 8, &a); table(64, 8, &b); table(64, 4, &c); table(256, 4, &d); table(512, 4, &e);
    vt.t8_8 = a; vt.t64_8 = b; vt.t64_4 = c; vt.t256_4 = d; vt.t512_4 = e;
    long long* cyc; double* out; cudaMalloc(&cyc, 8 * 16); cudaMalloc(&out, 8 * 512);
    size_t sm = sizeof(cd) * (16 * 288 + 56 + 192);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    const int reps = 200;
    for (int mode = 0; mode < 2; ++mode)
        for (int nw = 1; nw <= 16; nw *= 2) {
            if (mode == 0) k<0><<<1, 32 * nw, sm>>>(vt, cyc, out, reps); else k<1><<<1, 32 * nw, sm>>>(vt, cyc, out, reps);
            long long h[16]; cudaMemcpy(h, cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost);
            printf("%s tables, %2d warps/CTA: %6.0f cycles per 256-pt transform per warp (%s)\n", mode ? "smem  " : "global", nw, (double)h[0] / reps, cudaGetErrorString(cudaDeviceSynchronize()));
        }
}
