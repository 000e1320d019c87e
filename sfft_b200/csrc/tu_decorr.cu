// tu_decorr.cu -- noise decorrelation: the decorrelation kernel in Fourier or real space and the convolution
// that applies it.
//
// Reference: PureCupy_DeCorrelation_Calculator.PCDC (sfft/utils/PureCupyDeCorrelationCalculator.py:46-125),
// DeCorrelation_Calculator.DCC (sfft/utils/DeCorrelationCalculator.py:11-103), BSpline_DeCorrelation.BDC
// (sfft/BSplineSFFT.py:4757-4868) and PureCupy_FFTKits.FFT_CONVOLVE (sfft/utils/PureCupyFFTKits.py:71-105).  The reference
// pads every (small) match kernel to the image size, takes a full fft2 of each and works on image-sized complex planes.
// A kernel with L0 x L1 taps has the closed-form spectrum  FK[k0, k1] = sum_a W0^{k0 (a - w0)} sum_b K[a, b] W1^{k1 (b - w1)},
// so nothing image-sized is transformed here:
//   1. row tables   P_m[a][k1] = sum_b K_m[a, b] W1^{k1 (b - w1)}                     (dc_rowtab_kernel)
//   2. denominator  DeNo[k0, k1] = sum_J sig^2 |FK_J|^2 / NJ^2 + |FMK|^2 sum_I sig^2 |FK_I|^2 / NI^2,
//                   FK_m[k0, k1] = sum_a P_m[a][k1] W0^{k0 (a - w0)}                   (dc_deno_kernel; running maximum for the clip)
//   3. FKDECO = 1 / sqrt(DeNo) (optionally clipped from below, optionally normalised by its [0, 0] value)
//   4. real output: only the L0 x L1 taps that survive the tail truncation of KERNEL_CSZ_INV / iCSZ are evaluated, as two
//      separable partial inverse DFTs (dc_inv_rows_kernel, dc_inv_cols_kernel).
// The convolution is evaluated directly in real space with the zero (constant) padding and NaN fill of FFT_CONVOLVE.
#define SFFTB_TU_DECORR
#include "plan.h"

#define DC_MAXK 64

struct DcKer { int L0, L1, role, off; double s2; };     // role 0 = J queue, 1 = I queue, 2 = match kernel; s2 = sig^2 / N_role^2

__device__ __forceinline__ cd dc_cis(long long num, int den, double sgn) {
    // exp(sgn 2 pi i num / den), num reduced exactly
    long long e = num % den;
    if (e < 0) e += den;
    double s, c;
    sincospi(2.0 * (double)e / (double)den, &s, &c);
    return cmake(c, sgn * s);
}

__global__ void dc_twiddle_kernel(int n, cd* __restrict__ tw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) tw[e] = dc_cis(e, n, -1.0);
}

// P[off_m + a * N1 + k1]
__global__ void dc_rowtab_kernel(int N1, DcKer k, const double* __restrict__ kdata, cd* __restrict__ P) {
    const int k1 = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (k1 >= N1) return;
    const int w1 = (k.L1 - 1) / 2;
    cd acc = cmake(0.0, 0.0);
    for (int b = 0; b < k.L1; ++b) {
        const cd w = dc_cis((long long)k1 * (b - w1), N1, -1.0);
        const double v = kdata[k.off + a * k.L1 + b];
        acc.x = fma(v, w.x, acc.x); acc.y = fma(v, w.y, acc.y);
    }
    P[(size_t)a * N1 + k1] = acc;
}

struct DcArgs {
    int N0, N1, nker;
    DcKer k[DC_MAXK];
    size_t poff[DC_MAXK];        // offset of P_m in the table buffer
};

__global__ void __launch_bounds__(256) dc_deno_kernel(DcArgs a, const cd* __restrict__ P, const cd* __restrict__ tw0, double* __restrict__ deno,
                                                      unsigned long long* __restrict__ dmax) {
    const int k1 = blockIdx.x * blockDim.x + threadIdx.x, k0 = blockIdx.y;
    double v = 0.0;
    if (k1 < a.N1) {
        double sj = 0.0, si = 0.0, fm2 = 1.0;
        for (int m = 0; m < a.nker; ++m) {
            const DcKer& k = a.k[m];
            const int w0 = (k.L0 - 1) / 2;
            const cd* Pm = P + a.poff[m] + k1;
            cd F = cmake(0.0, 0.0);
            for (int aa = 0; aa < k.L0; ++aa) {
                int e = (int)(((long long)k0 * (aa - w0)) % a.N0);
                if (e < 0) e += a.N0;
                cfma(F, Pm[(size_t)aa * a.N1], tw0[e]);
            }
            const double f2 = F.x * F.x + F.y * F.y;
            if (k.role == 0) sj = fma(k.s2, f2, sj);
            else if (k.role == 1) si = fma(k.s2, f2, si);
            else fm2 = f2;
        }
        v = sj + si * fm2;
        deno[(size_t)k0 * a.N1 + k1] = v;
    }
    // running maximum (positive doubles order like their bit patterns)
    double mx = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(dmax, (unsigned long long)__double_as_longlong(mx));
}

// FKDECO = 1 / sqrt(max(DeNo, max / clip_ratio))   (clip: BDC :4838-4841)
__global__ void dc_finish_kernel(size_t n, double* __restrict__ deno, const unsigned long long* __restrict__ dmax, double clip_ratio) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double thr = clip_ratio > 0.0 ? __longlong_as_double((long long)*dmax) / clip_ratio : 0.0;
    deno[i] = 1.0 / sqrt(fmax(deno[i], thr));
}

// T[k0][b] = sum_k1 F[k0][k1] exp(+2 pi i k1 (b - w1) / N1): one warp per (k0, b)
__global__ void dc_inv_rows_kernel(int N0, int N1, int LO1, const double* __restrict__ F, cd* __restrict__ T) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= N0 * LO1) return;
    const int k0 = gw / LO1, b = gw - k0 * LO1, w1 = (LO1 - 1) / 2;
    double sx = 0.0, sy = 0.0;
    for (int k1 = lane; k1 < N1; k1 += 32) {
        const cd w = dc_cis((long long)k1 * (b - w1), N1, +1.0);
        const double f = F[(size_t)k0 * N1 + k1];
        sx = fma(f, w.x, sx); sy = fma(f, w.y, sy);
    }
    sx = warp_sum(sx); sy = warp_sum(sy);
    if (lane == 0) T[(size_t)k0 * LO1 + b] = cmake(sx, sy);
}

// K[a][b] = Re sum_k0 T[k0][b] exp(+2 pi i k0 (a - w0) / N0) / (N0 N1): one warp per (a, b)
__global__ void dc_inv_cols_kernel(int N0, int N1, int LO0, int LO1, const cd* __restrict__ T, double* __restrict__ K) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= LO0 * LO1) return;
    const int a = gw / LO1, b = gw - a * LO1, w0 = (LO0 - 1) / 2;
    double s = 0.0;
    for (int k0 = lane; k0 < N0; k0 += 32) {
        const cd w = dc_cis((long long)k0 * (a - w0), N0, +1.0);
        const cd t = T[(size_t)k0 * LO1 + b];
        s += t.x * w.x - t.y * w.y;
    }
    s = warp_sum(s);
    if (lane == 0) K[gw] = s / ((double)N0 * (double)N1);
}

// sum and sum of absolute values of n doubles (single block; n is a kernel stamp or a small grid)
__global__ void dc_sums_kernel(int n, const double* __restrict__ x, double* __restrict__ out2) {
    __shared__ double sh[2][32];
    double s = 0.0, sa = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { s += x[i]; sa += fabs(x[i]); }
    s = warp_sum(s); sa = warp_sum(sa);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = sa; }
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[0][threadIdx.x] : 0.0;
        sa = threadIdx.x < (blockDim.x >> 5) ? sh[1][threadIdx.x] : 0.0;
        s = warp_sum(s); sa = warp_sum(sa);
        if (threadIdx.x == 0) { out2[0] = s; out2[1] = sa; }
    }
}
__global__ void dc_scale_kernel(int n, double* __restrict__ x, const double* __restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] /= sums[0];
}

extern "C" int sfftb_decorr(int device, void* cuda_stream, int N0, int N1, int nker, const double* kdata, const int* kshape, const int* role,
                            const double* sig, double clip_ratio, int out_mode, int LO0, int LO1, int normalize, double* out, int out_memkind,
                            double* lost_weight) {
    if (!kdata || !kshape || !role || !sig || !out) return fail(SFFTB_EINVAL, "null argument");
    if (N0 < 1 || N1 < 1 || nker < 1 || nker > DC_MAXK) return fail(SFFTB_EINVAL, "bad grid size or kernel count (at most %d kernels)", DC_MAXK);
    if (out_mode != 0 && out_mode != 1) return fail(SFFTB_EINVAL, "bad output mode");
    if (out_mode == 1 && (LO0 < 1 || LO1 < 1 || LO0 % 2 == 0 || LO1 % 2 == 0 || LO0 > N0 || LO1 > N1))
        return fail(SFFTB_EINVAL, "real-space output size (%d, %d) must be odd and fit the grid", LO0, LO1);
    CK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DcArgs a;
    memset(&a, 0, sizeof a);
    a.N0 = N0; a.N1 = N1; a.nker = nker;
    int nJ = 0, nI = 0, nM = 0, off = 0;
    size_t poff = 0;
    for (int m = 0; m < nker; ++m) {
        const int L0 = kshape[2 * m], L1 = kshape[2 * m + 1];
        if (L0 < 1 || L1 < 1 || L0 % 2 == 0 || L1 % 2 == 0 || L0 > N0 || L1 > N1)
            return fail(SFFTB_EINVAL, "kernel %d of shape (%d, %d): only odd sizes that fit the grid are supported", m, L0, L1);
        if (role[m] < 0 || role[m] > 2) return fail(SFFTB_EINVAL, "bad kernel role");
        nJ += role[m] == 0; nI += role[m] == 1; nM += role[m] == 2;
        a.k[m].L0 = L0; a.k[m].L1 = L1; a.k[m].role = role[m]; a.k[m].off = off;
        a.poff[m] = poff;
        off += L0 * L1; poff += (size_t)L0 * N1;
    }
    if (nJ < 1 || nM > 1) return fail(SFFTB_EINVAL, "at least one J kernel and at most one match kernel are required");
    for (int m = 0; m < nker; ++m)
        a.k[m].s2 = role[m] == 0 ? sig[m] * sig[m] / ((double)nJ * nJ) : (role[m] == 1 ? sig[m] * sig[m] / ((double)nI * nI) : 0.0);
    double *dk = nullptr, *deno = nullptr, *dK = nullptr, *dsums = nullptr;
    cd *P = nullptr, *tw0 = nullptr, *T = nullptr;
    unsigned long long* dmax = nullptr;
    const size_t npix = (size_t)N0 * N1;
    int rc = 0;
#define DCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(SFFTB_ECUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); goto done; } } while (0)
    DCK(cudaMalloc(&dk, sizeof(double) * off));
    DCK(cudaMalloc(&P, sizeof(cd) * poff));
    DCK(cudaMalloc(&tw0, sizeof(cd) * N0));
    DCK(cudaMalloc(&dmax, sizeof(unsigned long long)));
    DCK(cudaMalloc(&dsums, sizeof(double) * 4));
    if (out_mode == 0 && out_memkind == SFFTB_MEM_DEVICE) deno = out;
    else DCK(cudaMalloc(&deno, sizeof(double) * npix));
    DCK(cudaMemcpyAsync(dk, kdata, sizeof(double) * off, cudaMemcpyHostToDevice, st));
    DCK(cudaMemsetAsync(dmax, 0, sizeof(unsigned long long), st));
    dc_twiddle_kernel<<<(N0 + 255) / 256, 256, 0, st>>>(N0, tw0);
    for (int m = 0; m < nker; ++m)
        dc_rowtab_kernel<<<dim3((N1 + 127) / 128, a.k[m].L0), 128, 0, st>>>(N1, a.k[m], dk, P + a.poff[m]);
    dc_deno_kernel<<<dim3((N1 + 255) / 256, N0), 256, 0, st>>>(a, P, tw0, deno, dmax);
    dc_finish_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(npix, deno, dmax, clip_ratio);
    DCK(cudaGetLastError());
    if (out_mode == 0) {
        if (normalize) {
            // divide by FKDECO[0, 0] (PCDC :104-107); the value is saved first because the scaling runs in place
            if (npix > 0x7fffffff) { rc = fail(SFFTB_EINVAL, "grid too large"); goto done; }
            DCK(cudaMemcpyAsync(dsums, deno, sizeof(double), cudaMemcpyDeviceToDevice, st));
            dc_scale_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>((int)npix, deno, dsums);
        }
        if (deno != out) DCK(cudaMemcpyAsync(out, deno, sizeof(double) * npix, cudaMemcpyDeviceToHost, st));
        if (lost_weight) *lost_weight = 0.0;
        DCK(cudaStreamSynchronize(st));
    } else {
        const int nk = LO0 * LO1;
        DCK(cudaMalloc(&T, sizeof(cd) * (size_t)N0 * LO1));
        DCK(cudaMalloc(&dK, sizeof(double) * nk));
        {
            const long long nw = (long long)N0 * LO1;
            dc_inv_rows_kernel<<<(unsigned)((nw * 32 + 255) / 256), 256, 0, st>>>(N0, N1, LO1, deno, T);
            dc_inv_cols_kernel<<<(nk * 32 + 255) / 256, 256, 0, st>>>(N0, N1, LO0, LO1, T, dK);
        }
        dc_sums_kernel<<<1, 256, 0, st>>>(nk, dK, dsums);
        double lw = nan("");
        if (lost_weight && npix <= (size_t)1 << 16) {
            // tail-truncation lost weight (iCSZ / KERNEL_CSZ_INV): needs sum |DeCo| over the whole grid -- small grids only
            double* full = nullptr; cd* Tf = nullptr;
            DCK(cudaMalloc(&full, sizeof(double) * npix));
            DCK(cudaMalloc(&Tf, sizeof(cd) * npix));
            // DeCo on the whole grid through the same partial transforms with an (N0, N1) "stamp": the offsets only permute the
            // grid, and the sum of absolute values does not care
            dc_inv_rows_kernel<<<(unsigned)(((long long)N0 * N1 * 32 + 255) / 256), 256, 0, st>>>(N0, N1, N1, deno, Tf);
            dc_inv_cols_kernel<<<(unsigned)(((long long)npix * 32 + 255) / 256), 256, 0, st>>>(N0, N1, N0, N1, Tf, full);
            dc_sums_kernel<<<1, 256, 0, st>>>((int)npix, full, dsums + 2);
            double h[4];
            DCK(cudaMemcpyAsync(h, dsums, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
            DCK(cudaStreamSynchronize(st));
            lw = 1.0 - h[1] / h[3];
            cudaFree(full); cudaFree(Tf);
        }
        if (lost_weight) *lost_weight = lw;
        if (normalize) dc_scale_kernel<<<(nk + 255) / 256, 256, 0, st>>>(nk, dK, dsums);
        DCK(cudaGetLastError());
        DCK(cudaMemcpyAsync(out, dK, sizeof(double) * nk, out_memkind == SFFTB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
        DCK(cudaStreamSynchronize(st));
    }
done:
    cudaFree(dk); cudaFree(P); cudaFree(tw0); cudaFree(dmax); cudaFree(dsums); cudaFree(T); cudaFree(dK);
    if (deno && deno != out) cudaFree(deno);
    return rc;
#undef DCK
}

// ---- direct convolution with the padding / NaN semantics of FFT_CONVOLVE -------------------------------------------------------
// out[r, c] = sum_{a, b} K[a, b] in[r - (a - w0), c - (b - w1)], samples outside the image = pad_fill, NaN samples = nan_fill
// (nan_fill = NaN keeps them: NAN_FILL_VALUE=None).  32 x 32 output tile per CTA, halo tile and kernel in shared memory.
#define CV_T 32
template <typename T>
__global__ void __launch_bounds__(256) conv_direct_kernel(int N0, int N1, int L0, int L1, const T* __restrict__ in, const double* __restrict__ K,
                                                          double pad_fill, double nan_fill, int fill_nan, double kscale, T* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* ks = reinterpret_cast<double*>(smem_raw);              // L0 * L1, flipped
    const int TH = CV_T + L0 - 1, TW = CV_T + L1 - 1, TP = TW | 1;
    double* tile = ks + L0 * L1;
    const int w0 = (L0 - 1) / 2, w1 = (L1 - 1) / 2;
    const int r0 = blockIdx.y * CV_T, c0 = blockIdx.x * CV_T;
    for (int i = threadIdx.x; i < L0 * L1; i += 256) ks[i] = K[L0 * L1 - 1 - i] * kscale;     // flipped: correlation form below
    for (int i = threadIdx.x; i < TH * TW; i += 256) {
        const int tr = i / TW, tc = i - tr * TW;
        const int r = r0 + tr - (L0 - 1 - w0), c = c0 + tc - (L1 - 1 - w1);
        double v = pad_fill;
        if (r >= 0 && r < N0 && c >= 0 && c < N1) {
            v = (double)in[(size_t)r * N1 + c];
            if (fill_nan && isnan(v)) v = nan_fill;
        }
        tile[tr * TP + tc] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 8 row groups of 4 rows
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int a = 0; a < L0; ++a)
        for (int b = 0; b < L1; ++b) {
            const double kv = ks[a * L1 + b];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fma(kv, tile[(ty * 4 + q + a) * TP + tx + b], acc[q]);
        }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + ty * 4 + q, c = c0 + tx;
        if (r < N0 && c < N1) out[(size_t)r * N1 + c] = (T)acc[q];
    }
}

extern "C" int sfftb_convolve(int device, void* cuda_stream, const void* img, int dtype, int N0, int N1, const double* kernel, int L0, int L1,
                              double pad_fill, double nan_fill, int fill_nan, int normalize_kernel, void* out, int memkind) {
    if (!img || !kernel || !out) return fail(SFFTB_EINVAL, "null argument");
    if (dtype != SFFTB_F64 && dtype != SFFTB_F32) return fail(SFFTB_EINVAL, "bad dtype");
    if (L0 < 1 || L1 < 1 || L0 % 2 == 0 || L1 % 2 == 0) return fail(SFFTB_EINVAL, "only odd-sized kernels are supported");
    CK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t esz = dtype == SFFTB_F64 ? 8 : 4, bytes = (size_t)N0 * N1 * esz;
    const int TW = CV_T + L1 - 1, TP = TW | 1;
    const size_t smem = sizeof(double) * ((size_t)L0 * L1 + (size_t)(CV_T + L0 - 1) * TP);
    int maxsm = 0;
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (smem > (size_t)maxsm) return fail(SFFTB_EINVAL, "kernel of shape (%d, %d) too large for the direct convolution", L0, L1);
    double ksum = 0.0;
    for (int i = 0; i < L0 * L1; ++i) ksum += kernel[i];
    const double kscale = normalize_kernel ? 1.0 / ksum : 1.0;
    void *din = nullptr, *dout = nullptr; double* dk = nullptr;
    int rc = 0;
#define CVK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(SFFTB_ECUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); goto done; } } while (0)
    CVK(cudaMalloc(&dk, sizeof(double) * L0 * L1));
    CVK(cudaMemcpyAsync(dk, kernel, sizeof(double) * L0 * L1, cudaMemcpyHostToDevice, st));
    if (memkind == SFFTB_MEM_HOST) {
        CVK(cudaMalloc(&din, bytes)); CVK(cudaMalloc(&dout, bytes));
        CVK(cudaMemcpyAsync(din, img, bytes, cudaMemcpyHostToDevice, st));
    } else { din = const_cast<void*>(img); dout = out; }
    {
        dim3 grd((N1 + CV_T - 1) / CV_T, (N0 + CV_T - 1) / CV_T);
        if (dtype == SFFTB_F64) {
            if (set_smem(conv_direct_kernel<double>, smem)) { rc = SFFTB_ECUDA; goto done; }
            conv_direct_kernel<double><<<grd, 256, smem, st>>>(N0, N1, L0, L1, (const double*)din, dk, pad_fill, nan_fill, fill_nan, kscale, (double*)dout);
        } else {
            if (set_smem(conv_direct_kernel<float>, smem)) { rc = SFFTB_ECUDA; goto done; }
            conv_direct_kernel<float><<<grd, 256, smem, st>>>(N0, N1, L0, L1, (const float*)din, dk, pad_fill, nan_fill, fill_nan, kscale, (float*)dout);
        }
    }
    CVK(cudaGetLastError());
    if (memkind == SFFTB_MEM_HOST) CVK(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, st));
    CVK(cudaStreamSynchronize(st));
done:
    cudaFree(dk);
    if (memkind == SFFTB_MEM_HOST) { cudaFree(din); cudaFree(dout); }
    return rc;
#undef CVK
}

// ---- grid-wise spatially varying convolution: BSpline_GridConvolve.GSVC_GPU (sfft/BSplineSFFT.py:4870-5010) ------------------------
// Every pixel carries the label of its grid cell (AllocatedL); the output pixel is the image convolved with THAT cell's kernel,
// zero outside the image.  The reference cuts a mini image per cell (the cell extended by w + 1 pixels), convolves it with
// scipy's convolve2d(mode='same', fillvalue=0) and pastes the cell back -- the extension covers the kernel footprint, so this is
// the same sum evaluated per pixel.  32 x 32 output tile per CTA, halo tile in shared memory; the kernel stack stays in global
// memory (a warp usually sits inside one cell, so its tap loads are broadcasts).
template <typename T>
__global__ void __launch_bounds__(256) conv_grid_kernel(int N0, int N1, int L0, int L1, int nseg, const T* __restrict__ in, const int* __restrict__ lab,
                                                        const double* __restrict__ K, double nan_fill, T* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tile = reinterpret_cast<double*>(smem_raw);
    const int TH = CV_T + L0 - 1, TW = CV_T + L1 - 1, TP = TW | 1;
    const int w0 = (L0 - 1) / 2, w1 = (L1 - 1) / 2;
    const int r0 = blockIdx.y * CV_T, c0 = blockIdx.x * CV_T;
    for (int i = threadIdx.x; i < TH * TW; i += 256) {
        const int tr = i / TW, tc = i - tr * TW;
        const int r = r0 + tr - (L0 - 1 - w0), c = c0 + tc - (L1 - 1 - w1);
        double v = 0.0;
        if (r >= 0 && r < N0 && c >= 0 && c < N1) {
            v = (double)in[(size_t)r * N1 + c];
            if (isnan(v)) v = nan_fill;
        }
        tile[tr * TP + tc] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + ty * 4 + q, c = c0 + tx;
        if (r >= N0 || c >= N1) continue;
        const int l = lab[(size_t)r * N1 + c];
        double acc = 0.0;
        if (l >= 0 && l < nseg) {
            const double* k = K + (size_t)l * L0 * L1;
            // out[r, c] = sum_{a, b} K[a, b] in[r + w0 - a, c + w1 - b]: tile row (r - r0) + (L0 - 1 - a)
            for (int a = 0; a < L0; ++a)
                for (int b = 0; b < L1; ++b)
                    acc = fma(__ldg(k + a * L1 + b), tile[(ty * 4 + q + L0 - 1 - a) * TP + tx + L1 - 1 - b], acc);
        }
        out[(size_t)r * N1 + c] = (T)acc;
    }
}

extern "C" int sfftb_convolve_grid(int device, void* cuda_stream, const void* img, int dtype, int N0, int N1, const int* labels, int nseg,
                                   const double* kerstack, int L0, int L1, double nan_fill, int normalize_kernel, void* out, int memkind) {
    if (!img || !labels || !kerstack || !out) return fail(SFFTB_EINVAL, "null argument");
    if (dtype != SFFTB_F64 && dtype != SFFTB_F32) return fail(SFFTB_EINVAL, "bad dtype");
    if (nseg < 1 || L0 < 1 || L1 < 1 || L0 % 2 == 0 || L1 % 2 == 0) return fail(SFFTB_EINVAL, "only odd-sized kernels are supported");
    CK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t esz = dtype == SFFTB_F64 ? 8 : 4, npix = (size_t)N0 * N1, bytes = npix * esz, nk = (size_t)nseg * L0 * L1;
    const int TW = CV_T + L1 - 1, TP = TW | 1;
    const size_t smem = sizeof(double) * (size_t)(CV_T + L0 - 1) * TP;
    int maxsm = 0;
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (smem > (size_t)maxsm) return fail(SFFTB_EINVAL, "kernels of shape (%d, %d) too large for the grid convolution", L0, L1);
    std::vector<double> hk(kerstack, kerstack + nk);
    if (normalize_kernel)
        for (int s = 0; s < nseg; ++s) {
            double sum = 0.0;
            for (int i = 0; i < L0 * L1; ++i) sum += hk[(size_t)s * L0 * L1 + i];
            for (int i = 0; i < L0 * L1; ++i) hk[(size_t)s * L0 * L1 + i] /= sum;
        }
    void *din = nullptr, *dout = nullptr; double* dk = nullptr; int* dl = nullptr;
    int rc = 0;
#define CGK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(SFFTB_ECUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); goto done; } } while (0)
    CGK(cudaMalloc(&dk, sizeof(double) * nk));
    CGK(cudaMemcpyAsync(dk, hk.data(), sizeof(double) * nk, cudaMemcpyHostToDevice, st));
    if (memkind == SFFTB_MEM_HOST) {
        CGK(cudaMalloc(&din, bytes)); CGK(cudaMalloc(&dout, bytes)); CGK(cudaMalloc(&dl, sizeof(int) * npix));
        CGK(cudaMemcpyAsync(din, img, bytes, cudaMemcpyHostToDevice, st));
        CGK(cudaMemcpyAsync(dl, labels, sizeof(int) * npix, cudaMemcpyHostToDevice, st));
    } else { din = const_cast<void*>(img); dout = out; dl = const_cast<int*>(labels); }
    {
        dim3 grd((N1 + CV_T - 1) / CV_T, (N0 + CV_T - 1) / CV_T);
        if (dtype == SFFTB_F64) {
            if (set_smem(conv_grid_kernel<double>, smem)) { rc = SFFTB_ECUDA; goto done; }
            conv_grid_kernel<double><<<grd, 256, smem, st>>>(N0, N1, L0, L1, nseg, (const double*)din, dl, dk, nan_fill, (double*)dout);
        } else {
            if (set_smem(conv_grid_kernel<float>, smem)) { rc = SFFTB_ECUDA; goto done; }
            conv_grid_kernel<float><<<grd, 256, smem, st>>>(N0, N1, L0, L1, nseg, (const float*)din, dl, dk, nan_fill, (float*)dout);
        }
    }
    CGK(cudaGetLastError());
    if (memkind == SFFTB_MEM_HOST) CGK(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, st));
    CGK(cudaStreamSynchronize(st));
done:
    cudaFree(dk);
    if (memkind == SFFTB_MEM_HOST) { cudaFree(din); cudaFree(dout); cudaFree(dl); }
    return rc;
#undef CGK
}
