// Film output stage on the device: float film -> body bytes of the reference's image files
// (gamma_encoding ky.cpp:1548; store_ppm_impl 1669-1681; store_bmp_impl 1719-1733; store_hdr_impl 1750-1779).
//
// HBM-bound byte work: 12 B read per pixel, 3 (gamma8 / bmp) or 4 (rgbe) bytes written.  One thread produces
// one aligned 4-byte word of the body, so a warp writes 128 contiguous bytes and reads the 512 contiguous
// bytes of film they come from (as float4 per thread where the mapping is the identity).
//
// gamma_encoding is a double pow in the reference.  It is monotone on the float grid, so it is evaluated as
// "how many of the 255 byte thresholds are <= clamp01(x)" (kyd_gamma_table.h, bisected from and checked
// against the reference's own function for every float in [0, 1]): a MUFU estimate of the byte, then at most a
// couple of shared-memory comparisons to make it exact.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kyd.h"
#include "kyd_gamma_table.h"
#include "kyd_internal.h"

namespace {

__constant__ uint32_t c_gamma_threshold_bits[256];

// float -> uint8_t as the reference binary does it on x86-64: cvttss2si to int32 ("integer indefinite"
// 0x80000000 for NaN and out-of-range values), then the low byte
__device__ __forceinline__ uint32_t x86_u8_of_f32(float v)
{
    return (v > -2147483904.f && v < 2147483648.f) ? ((uint32_t)__float2int_rz(v) & 0xffu) : 0u;
}

// threshold[] is the table padded with +inf at [256], [257]
__device__ __forceinline__ uint32_t gamma_byte(const float* threshold, float x)
{
    // clamp01 = std::clamp: NaN passes through, and a NaN ends up as byte 0 (every comparison below is false)
    const float xc = x < 0.f ? 0.f : (1.f < x ? 1.f : x);
    // MUFU estimate of pow(x, 1/2.2) * 255 + .5: off by far less than half a byte step, so the true byte is est - 1,
    // est or est + 1; two table comparisons settle it (tests/test_gpu_film_stage.py walks every float of [0, 1])
    const int est = __float2int_rz(__fadd_rn(__fmul_rn(__powf(xc, 0.45454545f), 255.f), 0.5f));
    const int lo = min(max(est - 1, 0), 255);
    return (uint32_t)(lo + (xc >= threshold[lo + 1] ? 1 : 0) + (xc >= threshold[lo + 2] ? 1 : 0));
}

__device__ __forceinline__ void load_thresholds(float* threshold)
{
    for (int i = threadIdx.x; i < 258; i += blockDim.x)
        threshold[i] = i < 256 ? __uint_as_float(c_gamma_threshold_bits[i]) : __int_as_float(0x7f800000);
    __syncthreads();
}

__device__ __forceinline__ uint32_t gamma_word(const float* threshold, float4 f)
{
    return gamma_byte(threshold, f.x) | (gamma_byte(threshold, f.y) << 8) | (gamma_byte(threshold, f.z) << 16) |
           (gamma_byte(threshold, f.w) << 24);
}

// GAMMA8: body byte j <- film float j.  Word k = floats 4k..4k+3; two words in flight per thread.
__global__ void __launch_bounds__(256) k_film_gamma8(const float* __restrict__ film, int64_t n_bytes, uint8_t* __restrict__ out)
{
    __shared__ float threshold[258];
    load_thresholds(threshold);
    const int64_t words = n_bytes >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float4* src = reinterpret_cast<const float4*>(film);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out);
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; k + stride < words; k += 2 * stride)
    {
        const float4 f0 = __ldg(src + k), f1 = __ldg(src + k + stride);
        dst[k] = gamma_word(threshold, f0);
        dst[k + stride] = gamma_word(threshold, f1);
    }
    if (k < words)
        dst[k] = gamma_word(threshold, __ldg(src + k));
    // 0-3 tail bytes
    if (blockIdx.x == 0 && threadIdx.x < (n_bytes & 3))
    {
        const int64_t j = (words << 2) + threadIdx.x;
        out[j] = (uint8_t)gamma_byte(threshold, film[j]);
    }
}

// BMP24: body byte j = row' * 3w + 3x + c  <-  film[((h-1-row') * w + x) * 3 + (2 - c)].  A block walks output lines;
// line r owns the words whose first byte lies in it, so no division is needed to find a word's line; the four bytes
// of a word may straddle the line end (3w is not a multiple of 4 in general).
__global__ void __launch_bounds__(256) k_film_bmp24(const float* __restrict__ film, int width, int height, uint8_t* __restrict__ out)
{
    __shared__ float threshold[258];
    load_thresholds(threshold);
    const int row_bytes = 3 * width;
    const int64_t n_bytes = (int64_t)row_bytes * height;
    const int64_t words = n_bytes >> 2;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out);
    for (int r = blockIdx.x; r < height; r += gridDim.x)
    {
        const int64_t line = (int64_t)r * row_bytes;
        const int64_t k0 = (line + 3) >> 2;
        const int64_t k1 = min((line + row_bytes + 3) >> 2, words);
        const float* src0 = film + (int64_t)(height - 1 - r) * row_bytes;
        for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x)
        {
            int within = (int)((k << 2) - line);
            const float* src = src0;
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                if (within == row_bytes) { within = 0; src -= row_bytes; }
                const int x = within / 3, c = within - 3 * x;
                w |= gamma_byte(threshold, __ldg(src + 3 * x + (2 - c))) << (8 * i);
                ++within;
            }
            dst[k] = w;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n_bytes & 3))
    {
        const int64_t j = (words << 2) + threadIdx.x;
        const int64_t row = j / row_bytes;
        const int within = (int)(j - row * row_bytes);
        const int x = within / 3, c = within - 3 * x;
        out[j] = (uint8_t)gamma_byte(threshold, __ldg(film + ((int64_t)(height - 1) - row) * row_bytes + 3 * x + (2 - c)));
    }
}

// RGBE: one pixel per thread (ky.cpp:1752-1779).  frexp(v) * 256 / v is exactly 2^(8 - e) for a normal v, so
// the reference's float m is built from the exponent field; v = inf gives m = NaN there and bytes 0,0,0,128.
__device__ __forceinline__ uint32_t rgbe_word(float r, float g, float b)
{
    float v = r;            // std::max({r, g, b}): a later element replaces only if it compares greater
    if (v < g) v = g;
    if (v < b) v = b;
    uint32_t w = 0;
    if (v >= 1e-32f)
    {
        const uint32_t bits = __float_as_uint(v);
        const int biased = (int)((bits >> 23) & 0xffu);
        if (biased == 0xff)
            w = 128u << 24;                                       // inf: frexp -> (inf, e = 0), m = NaN
        else
        {
            const int e = biased - 126;                           // v = mant * 2^e, mant in [0.5, 1)
            const float m = __uint_as_float((uint32_t)(127 + 8 - e) << 23);
            w = x86_u8_of_f32(__fmul_rn(r, m)) | (x86_u8_of_f32(__fmul_rn(g, m)) << 8) |
                (x86_u8_of_f32(__fmul_rn(b, m)) << 16) | ((uint32_t)((e + 128) & 0xff) << 24);
        }
    }
    return w;
}

__global__ void __launch_bounds__(256) k_film_rgbe(const float* __restrict__ film, int64_t pixels, uint8_t* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out);
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; p + stride < pixels; p += 2 * stride)
    {
        const int64_t q = p + stride;
        const float r0 = __ldg(film + 3 * p), g0 = __ldg(film + 3 * p + 1), b0 = __ldg(film + 3 * p + 2);
        const float r1 = __ldg(film + 3 * q), g1 = __ldg(film + 3 * q + 1), b1 = __ldg(film + 3 * q + 2);
        dst[p] = rgbe_word(r0, g0, b0);
        dst[q] = rgbe_word(r1, g1, b1);
    }
    if (p < pixels)
        dst[p] = rgbe_word(__ldg(film + 3 * p), __ldg(film + 3 * p + 1), __ldg(film + 3 * p + 2));
}

int g_table_device_ready[64] = {};

} // namespace

namespace kyd {

cudaError_t launch_film_encode(int device, int sm_count, const float* film_dev, int width, int height, int format, uint8_t* out_dev, cudaStream_t stream)
{
    if (device >= 0 && device < 64 && !g_table_device_ready[device])
    {
        cudaError_t e = cudaMemcpyToSymbol(c_gamma_threshold_bits, kyd_gamma_threshold_bits, sizeof(kyd_gamma_threshold_bits));
        if (e != cudaSuccess) return e;
        g_table_device_ready[device] = 1;
    }
    const int64_t pixels = (int64_t)width * height;
    const int threads = 256;
    const int64_t work = format == KYD_FILM_RGBE ? pixels : (pixels * 3 + 3) / 4;
    int64_t blocks = (work + threads - 1) / threads;
    if (format == KYD_FILM_BMP24) blocks = height;      // k_film_bmp24 walks output lines
    const int64_t cap = (int64_t)sm_count * 8 * 2;      // grid-stride: two waves of resident blocks
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (format == KYD_FILM_GAMMA8)
        k_film_gamma8<<<(unsigned)blocks, threads, 0, stream>>>(film_dev, pixels * 3, out_dev);
    else if (format == KYD_FILM_BMP24)
        k_film_bmp24<<<(unsigned)blocks, threads, 0, stream>>>(film_dev, width, height, out_dev);
    else
        k_film_rgbe<<<(unsigned)blocks, threads, 0, stream>>>(film_dev, pixels, out_dev);
    return cudaGetLastError();
}

} // namespace kyd
