// kernels_fitsio.cuh -- the I/O edge of the packets on the device: the reference reads every FITS image
// with astropy, transposes it and converts it to float64 on the host (sfft/CustomizedPacket.py:93-112) before the
// upload.  Here the raw big-endian data block goes to the device as it is in the file and one kernel decodes it
// (byte swap, BITPIX conversion, BSCALE / BZERO) straight into the transposed (NAXIS1, NAXIS2) array the plan reads;
// the difference image goes back the same way.  NaN bookkeeping of the packets (:114-126, :183-188) as two small kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t fio_bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ uint64_t fio_bswap64(uint64_t x) {
    return ((uint64_t)fio_bswap32((uint32_t)x) << 32) | (uint64_t)fio_bswap32((uint32_t)(x >> 32));
}

__device__ __forceinline__ double fio_decode(const unsigned char* raw, size_t idx, int bitpix) {
    switch (bitpix) {
        case 8:   return (double)raw[idx];
        case 16:  { uint16_t v = reinterpret_cast<const uint16_t*>(raw)[idx]; return (double)(int16_t)((v >> 8) | (v << 8)); }
        case 32:  return (double)(int32_t)fio_bswap32(reinterpret_cast<const uint32_t*>(raw)[idx]);
        case 64:  return (double)(int64_t)fio_bswap64(reinterpret_cast<const uint64_t*>(raw)[idx]);
        case -32: return (double)__uint_as_float(fio_bswap32(reinterpret_cast<const uint32_t*>(raw)[idx]));
        default:  return __longlong_as_double((long long)fio_bswap64(reinterpret_cast<const uint64_t*>(raw)[idx]));
    }
}

// raw: FITS data block, element (y, x) at y * n1 + x (x = NAXIS1 index fastest);  out[x * n2 + y] = decoded value.
// 32 x 32 tiles through shared memory so that both sides are coalesced.  grid (ceil(n1/32), ceil(n2/32)), block (32, 8).
template <typename TOut>
__global__ void fits_decode_T_kernel(const unsigned char* __restrict__ raw, int bitpix, int n1, int n2, double bscale, double bzero,
                                     TOut* __restrict__ out)
{
    __shared__ double tile[32][33];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int x = x0 + threadIdx.x, y = y0 + j;
        if (x < n1 && y < n2) tile[j][threadIdx.x] = fio_decode(raw, (size_t)y * n1 + x, bitpix) * bscale + bzero;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int x = x0 + j, y = y0 + threadIdx.x;
        if (x < n1 && y < n2) out[(size_t)x * n2 + y] = (TOut)tile[threadIdx.x][j];
    }
}

// img (n1, n2) row-major -> raw FITS data block (n2, n1) big-endian, BITPIX -32 or -64
template <typename TIn>
__global__ void fits_encode_T_kernel(const TIn* __restrict__ img, int bitpix, int n1, int n2, unsigned char* __restrict__ raw)
{
    __shared__ double tile[32][33];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int x = x0 + j, y = y0 + threadIdx.x;
        if (x < n1 && y < n2) tile[j][threadIdx.x] = (double)img[(size_t)x * n2 + y];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int x = x0 + threadIdx.x, y = y0 + j;
        if (x < n1 && y < n2) {
            const double v = tile[threadIdx.x][j];
            const size_t idx = (size_t)y * n1 + x;
            if (bitpix == -32) reinterpret_cast<uint32_t*>(raw)[idx] = fio_bswap32(__float_as_uint((float)v));
            else reinterpret_cast<uint64_t*>(raw)[idx] = fio_bswap64((uint64_t)__double_as_longlong(v));
        }
    }
}

// NaN union of the unmasked pair: where either is NaN both pixels are taken from the masked images (CP :114-126, :165-170)
template <typename T>
__global__ void nan_union_fill_kernel(T* __restrict__ A, T* __restrict__ B, const T* __restrict__ mA, const T* __restrict__ mB, size_t n,
                                      unsigned char* __restrict__ mask, int* __restrict__ flags)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool bad = isnan((double)A[i]) || isnan((double)B[i]);
    mask[i] = bad ? 1 : 0;
    if (bad) { A[i] = mA[i]; B[i] = mB[i]; flags[0] = 1; }
    if (isnan((double)mA[i]) || isnan((double)mB[i])) flags[1] = 1;
}

// DIFF[mask] = NaN, optional sign flip (CP :183-188)
template <typename T>
__global__ void nan_mask_apply_kernel(T* __restrict__ D, const unsigned char* __restrict__ mask, size_t n, double sign)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = (double)D[i] * sign;
    D[i] = (mask && mask[i]) ? (T)nan("") : (T)v;
}
