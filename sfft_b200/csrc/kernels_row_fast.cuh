// kernels_row_fast.cuh -- INVERSE row pass on the register-resident 16-values-per-thread FFT engine (even N1 with N1/2 in
// {512,...,8192}); the forward pass of this family was replaced by kernels_row_h16.cuh / kernels_row_v8.cuh.
// Same contract as kernels_row.cuh: packed real-to-complex rows with the cy^j factor fused into the load, half
// spectra stored transposed g[j][k1][r]; and the inverse (transposed half spectra -> real rows - background).
// A CTA of 512 threads owns RB = 512 / T rows (T = H / 16 threads per row), so the transposed stores/loads move
// RB consecutive rows per column.
#pragma once
#include "fft_regs.cuh"
#include "kernels_row.cuh"

#define ROWF_NT 512

__device__ __forceinline__ void load2(const float* p, double& a, double& b) { const float2 t = *reinterpret_cast<const float2*>(p); a = t.x; b = t.y; }
__device__ __forceinline__ void load2(const double* p, double& a, double& b) { const double2 t = *reinterpret_cast<const double2*>(p); a = t.x; b = t.y; }
__device__ __forceinline__ void store2(float* p, double a, double b) { *reinterpret_cast<float2*>(p) = make_float2((float)a, (float)b); }
__device__ __forceinline__ void store2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

struct RowFastArgs {
    int N0, N1, NH, H;
    const cd* tabA; const cd* tabB; const cd* tabC;   // engine twiddle tables (global)
    const cd* tw1;                                     // exp(-2 pi i e / N1)
    const double* vtab;                                // general-basis plans: column tables instead of cy^j (or NULL)
};

template <int H>
__device__ __forceinline__ GroupSync row_group_sync(int grp, int lane_in_warp_group) {
    constexpr int T = H / 16;
    GroupSync gs;
    gs.mask = 0xffffffffu;
    gs.bar_id = 1 + grp;
    gs.count = T;
    return gs;
}

struct RowInvFastArgs {
    RowFastArgs r;
    double scale;
    int Fpq;
    int row0;                       // first row of this launch (chunked launches overlap the D2H copy of finished rows)
    unsigned char p_of[16], q_of[16];
};

template <typename TSt, typename TOut, int H>
__global__ void __launch_bounds__(ROWF_NT) row_inv_fast_kernel(RowInvFastArgs ia, const TSt* __restrict__ spec,
                                                               const double* __restrict__ bpq, TOut* __restrict__ out)
{
    constexpr int T = H / 16, RB = ROWF_NT / T, PITCH = H + H / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* buf = reinterpret_cast<cd*>(smem_raw);
    const RowFastArgs& a = ia.r;
    const int tid = threadIdx.x;
    const int grp = tid / T, lane = tid - grp * T;
    const int r0 = ia.row0 + blockIdx.x * RB;
    const int r = r0 + grp;
    cd* scratch = buf + (size_t)grp * PITCH;
    const GroupSync gs = row_group_sync<H>(grp, lane);

    // gather RB rows of the transposed spectrum and build the packed half-length spectrum Z = Ze + i Zo
    for (int idx = tid; idx < RB * H; idx += ROWF_NT) {
        const int k = idx / RB, row = idx - k * RB;
        const int rr = r0 + row;
        cd z = cmake(0.0, 0.0);
        if (rr < a.N0) {
            const cd gk = load_c(spec + (size_t)k * a.N0 + rr);
            const cd gm = cconj(load_c(spec + (size_t)(H - k) * a.N0 + rr));
            const cd ze = cscale(cadd(gk, gm), 0.5);
            const cd zo = cscale(cmul(csub(gk, gm), cconj(a.tw1[k])), 0.5);
            z = cmake(ze.x - zo.y, ze.y + zo.x);
        }
        buf[(size_t)row * PITCH + RPAD(k)] = z;
    }
    __syncthreads();
    cd v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = scratch[RPAD(lane + q * T)];
    __syncthreads();
    reg_fft<H>(v, scratch, lane, a.tabA, a.tabB, a.tabC, +1.0, gs);
    if (r >= a.N0) return;

    const double inv0 = 1.0 / (double)a.N0, inv1 = 1.0 / (double)a.N1;
    const double cx = (r + 1) * inv0;
    // background polynomial of this row as a Horner form in cy: cq[q] = sum_p b_pq cx^p
    double cq[4] = {0.0, 0.0, 0.0, 0.0};
    if (bpq != nullptr) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < ia.Fpq) {
                const double t = bpq[k] * ipow(cx, ia.p_of[k]);
                const int qq = ia.q_of[k];
                cq[0] += (qq == 0) ? t : 0.0; cq[1] += (qq == 1) ? t : 0.0;
                cq[2] += (qq == 2) ? t : 0.0; cq[3] += (qq == 3) ? t : 0.0;
            }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = lane + q * T;
        const double cy0 = (2 * n + 1) * inv1, cy1 = (2 * n + 2) * inv1;
        const double x0 = fma(v[q].x, ia.scale, -fma(fma(fma(cq[3], cy0, cq[2]), cy0, cq[1]), cy0, cq[0]));
        const double x1 = fma(v[q].y, ia.scale, -fma(fma(fma(cq[3], cy1, cq[2]), cy1, cq[1]), cy1, cq[0]));
        store2(out + (size_t)r * a.N1 + 2 * n, x0, x1);
    }
}
