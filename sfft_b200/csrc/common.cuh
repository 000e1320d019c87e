// common.cuh -- complex helpers, storage conversion, argument structs shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cd;   // all arithmetic is complex double; HBM storage is float2 or double2

#define SFFTB_MAX_STAGES 16
#define SFFTB_MAX_PLANES 16     // Fij <= 10 (DK <= 3) + J
#define SFFTB_MAX_PAIRS  64     // Fij (Fij + 1) / 2 <= 55

struct FftDesc {
    int n;                          // transform length
    int ns;                         // number of stages
    int radix[SFFTB_MAX_STAGES];
};

__host__ __device__ __forceinline__ cd cmake(double x, double y) { cd r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// conj(a) * b
__device__ __forceinline__ cd cmulcj(cd a, cd b) { return cmake(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ cd cconj(cd a) { return cmake(a.x, -a.y); }
__device__ __forceinline__ cd cscale(cd a, double s) { return cmake(a.x * s, a.y * s); }
__device__ __forceinline__ void cfma(cd& acc, cd a, cd b) {   // acc += a * b
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// multiply by sgn * i  (sgn = -1: forward transform's -i, +1: inverse)
__device__ __forceinline__ cd cmuli(cd a, double sgn) { return cmake(-sgn * a.y, sgn * a.x); }

__device__ __forceinline__ cd load_c(const double2* p) { return *p; }
__device__ __forceinline__ cd load_c(const float2* p) { float2 v = *p; return cmake((double)v.x, (double)v.y); }
__device__ __forceinline__ void store_c(double2* p, cd v) { *p = v; }
__device__ __forceinline__ void store_c(float2* p, cd v) { *p = make_float2((float)v.x, (float)v.y); }

__device__ __forceinline__ double ipow(double x, int e) {
    double r = 1.0;
    for (int k = 0; k < e; ++k) r *= x;
    return r;
}

__device__ __forceinline__ int imod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Everything the column passes need to know about the configuration; passed by value.
struct ColArgs {
    int N0, N1, NH;                 // image shape, half-spectrum width N1/2+1
    int V, M, pitch;                // fold factor, slice length N0/V, smem plane pitch (elements)
    int DK, DB, Fij, Fpq, nj;       // nj = DK + 1 stored row-spectrum planes of I
    int w0, w1;
    int npairs;                     // Fij (Fij + 1) / 2 unordered I-plane pairs
    int nl0, nlj0;                  // 4 w0 + 1 (Omega lags), 2 w0 + 1 (Theta / Lambda lags)
    int PB;                         // pairs transformed per group
    FftDesc fd;                     // plan for length M
    const cd* twM;                  // exp(-2 pi i e / M),  e in [0, M)
    const cd* tw0;                  // exp(-2 pi i e / N0), e in [0, N0)
    unsigned char pl_i[SFFTB_MAX_PLANES], pl_j[SFFTB_MAX_PLANES];   // (i, j) of I-plane A (REF_ij order)
    unsigned char plane_of[4][4];   // inverse map (i, j) -> A
    unsigned char pairA[SFFTB_MAX_PAIRS], pairB[SFFTB_MAX_PAIRS];
};
