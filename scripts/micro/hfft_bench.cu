// throughput of 256-point forward transforms per SM: half-warp 16-value engine (fft_h16.cuh) vs the warp-wide 8-value engine
// (fft_vpt.cuh); input = staged float2 window in shared memory, output = spectrum plane in shared memory (the fit kernel's job)
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../sfft_b200/csrc/fft_vpt.cuh"
#include "../../sfft_b200/csrc/fft_h16.cuh"
#define PITCH 288
template <int ENGINE> __global__ void __launch_bounds__(512, 1) k(VTabs vt_g, long long* cyc, cd* out, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* planes = reinterpret_cast<cd*>(smem_raw);            // 32 planes
    cd* tw8 = planes + 32 * PITCH; cd* tw64 = tw8 + 56;
    float2* stage = reinterpret_cast<float2*>(tw64 + 192);   // 4 windows of 256
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 56; i += blockDim.x) tw8[i] = vt_g.t8_8[i];
    for (int i = tid; i < 192; i += blockDim.x) tw64[i] = vt_g.t64_4[i];
    for (int i = tid; i < 1024; i += blockDim.x) stage[i] = make_float2(sinf(0.37f * i) + 0.01f * i, cosf(0.11f * i * i));
    VTabs vt = vt_g; vt.t8_8 = tw8; vt.t64_4 = tw64;
    __syncthreads();
    long long t0 = clock64();
    if (ENGINE == 0) {
        cd* plane = planes + warp * PITCH;
        for (int rep = 0; rep < reps; ++rep) {
            const float2* src = stage + 256 * ((rep + warp) & 3);
            cd v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = load_c(src + lane + 32 * q);
            vfft<256>(v, plane, lane, vt, -1.0, 0);
#pragma unroll
            for (int q = 0; q < 8; ++q) plane[VPAD(lane + 32 * q)] = v[q];
            __syncwarp();
        }
    } else {
        const int half = lane >> 4, hl = lane & 15;
        H16Tw tw; h16_init(tw, hl);
        cd* plane = planes + (2 * warp + half) * PITCH;
        for (int rep = 0; rep < reps; ++rep) {
            const float2* src = stage + 256 * ((rep + warp + half) & 3);
            cd v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = load_c(src + hl + 16 * q);
            hfft256(v, plane, hl, tw, -1.0);
#pragma unroll
            for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];
            __syncwarp();
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    // dump the spectrum of window ((reps - 1 + warp [+ half]) & 3) for the correctness check (warp 0 only)
    if (warp == 0) {
        if (ENGINE == 0) { for (int i = lane; i < 256; i += 32) out[i] = planes[VPAD(i)]; }
        else { for (int i = lane; i < 256; i += 32) out[i] = planes[HPAD(i)]; }
    }
}
static void table(int Ns, int R, cd** out) {
    std::vector<cd> h((size_t)(R - 1) * Ns);
    for (int r = 1; r < R; ++r) for (int kk = 0; kk < Ns; ++kk) { double ang = 2 * M_PI * r * kk / ((double)Ns * R); h[(size_t)(r - 1) * Ns + kk] = cmake(cos(ang), -sin(ang)); }
    cudaMalloc(out, sizeof(cd) * h.size()); cudaMemcpy(*out, h.data(), sizeof(cd) * h.size(), cudaMemcpyHostToDevice);
}
int main() {
    VTabs vt; cd *a, *b, *c, *d, *e; table(8, 8, &a); table(64, 8, &b); table(64, 4, &c); table(256, 4, &d); table(512, 4, &e);
    vt.t8_8 = a; vt.t64_8 = b; vt.t64_4 = c; vt.t256_4 = d; vt.t512_4 = e;
    long long* cyc; cd* out; cudaMalloc(&cyc, 8 * 16); cudaMalloc(&out, 16 * 256);
    size_t sm = sizeof(cd) * (32 * PITCH + 56 + 192) + 8 * 1024;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    const int reps = 400;      // reps % 4 == 0: warp 0 (half 0) ends on window 3
    std::vector<cd> r0(256), r1(256);
    for (int eng = 0; eng < 2; ++eng)
        for (int nw = 1; nw <= 16; nw = nw < 4 ? nw * 2 : nw + 2) {
            if (eng == 0) k<0><<<1, 32 * nw, sm>>>(vt, cyc, out, reps); else k<1><<<1, 32 * nw, sm>>>(vt, cyc, out, reps);
            long long h[16]; cudaMemcpy(h, cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            const int per = eng == 0 ? nw : 2 * nw;
            printf("%s engine, %2d warps: %7.0f cycles per transform per warp-slot, %6.1f cycles per transform per SM (%s)\n", eng ? "h16" : "v8 ", nw,
                   (double)mx / reps, (double)mx / reps / per, cudaGetErrorString(cudaDeviceSynchronize()));
            cudaMemcpy(eng ? r1.data() : r0.data(), out, sizeof(cd) * 256, cudaMemcpyDeviceToHost);
        }
    double err = 0, nrm = 0;
    for (int i = 0; i < 256; ++i) { err = fmax(err, hypot(r0[i].x - r1[i].x, r0[i].y - r1[i].y)); nrm = fmax(nrm, hypot(r0[i].x, r0[i].y)); }
    printf("max |v8 - h16| = %.3e (max |X| = %.3e)\n", err, nrm);
}
