// kyd_device.cuh -- device-side building blocks of the sm_100a rendering core.
//
// Every function restates, operation by operation, the reference function it cites (reference
// ky.cpp) under the numerical contract of DESIGN.md:
//   * FP32 add / mul / div / sqrt are IEEE round-to-nearest and never contracted: this file must be
//     compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false;
//   * vec3 normalize multiplies by (float)(1.0 / sqrt((double)|v|^2)) because ky.cpp:314 resolves to
//     ::sqrt(double);
//   * float transcendentals are the correctly rounded values: CUDA's double sin/cos/pow/acos (<= 2 ulp
//     in double, explicit FMAs inside libdevice, not affected by -fmad=false) rounded once to float;
//   * std::max / std::clamp / operator precedence / argument evaluation order are g++'s.
// The kernels in kyd_kernels.cu only compose these blocks.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "kyd.h"
#include "kyd_scene.h"

// The device code is compiled twice: as namespace kyd for scenes in constant memory (kyd_kernels.cu) and, with
// KYD_BIG_SCENE, as namespace kyd_big for scenes under a bounding-volume hierarchy in global memory
// (kyd_kernels_big.cu) -- two sets of kernels, so that the headline kernels carry no trace of the other path (run-time
// hooks inside the traversal cost them 7 %, profiles/r01_ab_variants.txt).
#ifndef KYD_BIG_SCENE
#define KYD_BIG_SCENE 0
#endif
#if KYD_BIG_SCENE
#define KYD_KERNEL_NS kyd_big
namespace kyd_big { using namespace kyd; }
#else
#define KYD_KERNEL_NS kyd
#endif

namespace KYD_KERNEL_NS {

#define KYD_DEV __device__ __forceinline__
// the double-precision libm bodies are hundreds of instructions each and have many call sites: kept out of
// line so that the shading kernels stay closer to the instruction cache (KYD_MATH_INLINE=1 inlines them)
#if defined(KYD_MATH_INLINE) && KYD_MATH_INLINE
#define KYD_MATH __device__ __forceinline__
#else
#define KYD_MATH __device__ __noinline__
#endif

// unroll factor of the traversal loops: 1 keeps the shade kernels' code small (they are instruction-fetch bound: 6 KB L0 /
// 32 KB L1.5 instruction caches against 50-60 KB of kernel; profiles/r02_ab_variants.txt)
#ifndef KYD_TRAVERSAL_UNROLL
#define KYD_TRAVERSAL_UNROLL 1
#endif
#define KYD_PRAGMA__(x) _Pragma(#x)
#define KYD_PRAGMA_(x) KYD_PRAGMA__(x)
#define KYD_UNROLL_TRAVERSAL KYD_PRAGMA_(unroll KYD_TRAVERSAL_UNROLL)

// the scene of the context that launched the kernel.  This header is included by exactly one
// translation unit (kyd_kernels.cu), which therefore owns the symbol.
__constant__ DevScene c_scene;

// ---- constants (ky.cpp:180-188) --------------------------------------------------------------------
#define KYD_PI 3.14159274101257324e+00f        /* (float)pi            */
#define KYD_INV_PI 3.18309873342514038e-01f    /* (float)(1/pi)        */
#define KYD_2PI (2.f * KYD_PI)
#define KYD_PI_OVER2 (KYD_PI / 2.f)
#define KYD_PI_OVER4 (KYD_PI / 4.f)
#define KYD_INV_2PI (KYD_INV_PI / 2.f)
#define KYD_FLT_EPSILON 1.1920928955078125e-7f
#define KYD_SHAPE_EPSILON 0.001f               /* shape_t::epsilon, ky.cpp:1093 */
#define KYD_INF CUDART_INF_F

// std::max(a, b) = (a < b) ? b : a ; std::clamp
KYD_DEV float max_std(float a, float b) { return (a < b) ? b : a; }
KYD_DEV float clamp_std(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }

// A NaN radiance (the reference produces them too: 0 * inf in a throughput, about one sample in 10^8 of the Cornell box)
// leaves the device with the bit pattern the reference's x86-64 build gives it: SSE's default NaN 0xffc00000 -- every NaN
// on this path is born from an invalid operation and then only propagated -- where this GPU would write 0x7fffffff.
KYD_DEV float film_value(float v) { return v != v ? __int_as_float((int)0xffc00000u) : v; }

// libm contract
KYD_MATH float cr_sin(float x) { return __double2float_rn(sin((double)x)); }
KYD_MATH float cr_cos(float x) { return __double2float_rn(cos((double)x)); }
KYD_MATH void cr_sincos(float x, float* s, float* c)
{
    double ds, dc;
    sincos((double)x, &ds, &dc);
    *s = __double2float_rn(ds);
    *c = __double2float_rn(dc);
}
KYD_MATH float cr_acos(float x) { return __double2float_rn(acos((double)x)); }
// powf's contract value: the double pow rounded once (the definition; slow path of cr_pow)
KYD_MATH float cr_pow_reference(float x, float y) { return __double2float_rn(pow((double)x, (double)y)); }

// The same float for the cases Phong shading produces (x > 0), at a third of the cost: v = exp2(y * log2(x)) in
// double is within 2^-45 of x^y when |y log2 x| < 120 (log2 and exp2 are 1-ulp functions; the product's error is
// 2^-52 |y log2 x|), the definition's double pow is within 2^-52 of it; if v sits at least 2^-41 (relative) inside
// its float rounding interval both round to the same float.  Results below 2^-160 round to +0 either way.
// A negative base with an integer exponent -- Phong's eval takes powf of a negative cosine (ky.cpp:2496-2499; its exponents
// 30 / 90 / 5000 are integers) -- is (-1)^y |x|^y, exactly and with symmetric rounding, so it takes the same path.
// Everything else -- zero, non-finite or huge arguments, denormal results, values too close to a rounding
// boundary (2^-15 of calls) -- evaluates the definition.  tests: kyd_selftest(KYD_SELFTEST_POW).
#if defined(KYD_POW_INLINE) && KYD_POW_INLINE
KYD_DEV float cr_pow(float x, float y)
#else
// (out of line: the Phong kernels call it at three sites, ~150 instructions each, and are instruction-fetch bound)
__device__ __noinline__ float cr_pow(float x, float y)
#endif
{
#if !defined(KYD_FAST_POW) || KYD_FAST_POW
    float ax = x;
    bool negate = false;
    if (x < 0.f && fabsf(y) < 1e6f && y == truncf(y))
    {
        ax = -x;
        negate = (__float2int_rz(y) & 1) != 0;
    }
    if (ax > 0x1p-100f && ax < 0x1p100f && fabsf(y) < 1e6f)
    {
        const double t = (double)y * log2((double)ax);
        if (t < -160.0)
            return negate ? -0.f : 0.f;
        if (fabs(t) < 120.0)
        {
            const double v = exp2(t);
            const float f = __double2float_rn(v);
            const double d = v - (double)f;
            const unsigned fb = __float_as_uint(f);
            const double half_ulp = (double)__uint_as_float((fb & 0x7f800000u) - (24u << 23));
            const double hi = half_ulp * (1.0 - 0x1p-16);
            const double lo = (fb & 0x007fffffu) != 0u ? -hi : -0.5 * hi;
            if (d < hi && d > lo)
                return negate ? -f : f;
        }
    }
#endif
    return cr_pow_reference(x, y);
}

// ---- vectors (ky.cpp:226-388) -----------------------------------------------------------------------
KYD_DEV float3 V3(float x, float y, float z) { return make_float3(x, y, z); }
KYD_DEV float3 add(float3 a, float3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
KYD_DEV float3 sub(float3 a, float3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
KYD_DEV float3 mul(float3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
KYD_DEV float3 neg(float3 a) { return V3(-a.x, -a.y, -a.z); }
KYD_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
KYD_DEV float msq(float3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
KYD_DEV float mag(float3 a) { return __fsqrt_rn(msq(a)); } // (float)sqrt((double)s) == sqrtf(s)
KYD_DEV float3 cross(float3 a, float3 v)
{
    return V3(a.y * v.z - a.z * v.y, a.z * v.x - a.x * v.z, a.x * v.y - a.y * v.x);
}
// ky.cpp:314: (float)(1.0 / sqrt((double)s)) -- the definition, used as the slow path
KYD_MATH float rsqrt_ky_reference(float s) { return __double2float_rn(__drcp_rn(__dsqrt_rn((double)s))); }

// Slow path of rsqrt_ky (out of line: the shading kernels are instruction-fetch bound and inline rsqrt_ky ~10 times).
__device__ __noinline__ float rsqrt_ky_slow(float s)
{
    // What rsqrt_ky's FP32 check cannot decide in FP32 is, in practice, the square of an almost-unit vector: s = 4^m (1 - k 2^-24)
    // with k = 2 mod 4 has s^-1/2 = 2^-m (1 + k 2^-25 + 3/8 k^2 2^-48 + ...), a float rounding boundary plus a sliver -- which
    // decides the rounding: the sliver exceeds both doubles' rounding errors (>= 1.5 x 2^-48 against 2^-52) and stays below a
    // quarter of an ulp for k < 4729, so the definition's value is 2^-m (1 + ((k + 2) >> 2) 2^-23) for every k <= 4096
    // (checked against the definition for all of them on the CPU, and by the all-floats self-test on the device).
    {
        const unsigned sb = __float_as_uint(s);
        const unsigned k = 0x00800000u - (sb & 0x007fffffu);
        if (k <= 4096u && (sb & 0x00800000u) == 0u && s > 0x1p-60f && s < 0x1p60f)
        {
            const int m = ((int)(sb >> 23) - 126) >> 1;
            return __uint_as_float(((unsigned)(127 - m) << 23) + ((k + 2u) >> 2));
        }
    }
    return __double2float_rn(__drcp_rn(__dsqrt_rn((double)s)));
}

// The same value without FP64 on the fast path.  Let y* = s^-1/2 exactly; the definition above is
// RN32(p) with |p / y* - 1| <= 2^-51.9 (two correctly rounded double operations).
//   y0  = MUFU seed, |y0 / y* - 1| <= 2^-22.9 (PTX rsqrt.approx.f32)
//   r2  = 1 - s y0^2 to ~2^-46 absolute: t + tl = s y0 exactly (FMA), then two more FMAs
//   yt  = y0 (1 + r2 / 2): one Newton step; |yt / y* - 1| <= 3/8 r2^2 + 2^-45 < 2^-44
//   yh  = RN32(yt), rho = yt - yh (|rho| <= half an ulp of yh, computed with one FMA)
// If rho stays 2^-14 of a half-ulp clear of the rounding boundary (2^-39 relative, 32x the error budget),
// RN32(p) == yh.  Below a power of two the float spacing halves, so there the lower boundary is half as far
// (axis-aligned normals normalise to exactly 1.0: the common case in a Cornell box).  Otherwise (6e-5 of
// calls), and for zero / denormal / huge / non-finite s, the definition is evaluated.  Verified for every float bit
// pattern by kyd_selftest (tests/test_gpu_parity.py::test_fast_rsqrt_is_exact_for_every_float).
KYD_DEV bool rsqrt_ky_accept(float s, float yh, float rho)
{
    const unsigned yb = __float_as_uint(yh);
    const float half_ulp = __uint_as_float((yb & 0x7f800000u) - (24u << 23));
    const float hi = __fmul_rn(half_ulp, 0.99993896484375f);
    const float lo = (yb & 0x007fffffu) != 0u ? -hi : __fmul_rn(-0.5f, hi);
    return s > 0x1p-60f && s < 0x1p60f && rho < hi && rho > lo;
}

#if defined(KYD_RSQRT_NOINLINE) && KYD_RSQRT_NOINLINE
__device__ __noinline__ float rsqrt_ky(float s)
#else
KYD_DEV float rsqrt_ky(float s)
#endif
{
#if defined(KYD_FAST_RSQRT) && !KYD_FAST_RSQRT
    return __double2float_rn(__drcp_rn(__dsqrt_rn((double)s)));
#endif
    float y0;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(s));
    const float t = __fmul_rn(s, y0);
    const float tl = __fmaf_rn(s, y0, -t);
    const float r = __fmaf_rn(-t, y0, 1.0f);
    const float r2 = __fmaf_rn(-tl, y0, r);
    const float h = __fmul_rn(0.5f, r2);
    const float yh = __fmaf_rn(y0, h, y0);
    const float rho = __fmaf_rn(y0, h, __fsub_rn(y0, yh));
    if (rsqrt_ky_accept(s, yh, rho))
        return yh;
    return rsqrt_ky_slow(s);
}
KYD_DEV float3 normalize(float3 a) { return mul(a, rsqrt_ky(a.x * a.x + a.y * a.y + a.z * a.z)); }
KYD_DEV float abs_dot(float3 a, float3 b) { return fabsf(dot(a, b)); }
KYD_DEV float distance_sq(float3 a, float3 b) { return msq(sub(a, b)); }
KYD_DEV float distance(float3 a, float3 b) { return mag(sub(a, b)); }

// colors share float3: r,g,b = x,y,z
KYD_DEV float3 cdiv(float3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
KYD_DEV float3 cmulc(float3 a, float3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
KYD_DEV bool is_black(float3 c) { return (c.x <= 0) && (c.y <= 0) && (c.z <= 0); }
KYD_DEV float max_component(float3 c)
{
    float m = c.x;
    if (m < c.y) m = c.y;
    if (m < c.z) m = c.z;
    return m;
}
#define KYD_BLACK make_float3(0.f, 0.f, 0.f)

// frame_t(normal) ky.cpp:537-541, 566-571
struct Frame { float3 s, t, n; };
KYD_DEV Frame frame_from_z(float3 normal)
{
    Frame f;
    f.n = normalize(normal);
    float3 tmp = (fabsf(f.n.x) > 0.99f) ? V3(0, 1, 0) : V3(1, 0, 0);
    f.t = normalize(cross(f.n, tmp));
    f.s = normalize(cross(f.t, f.n));
    return f;
}
KYD_DEV float3 to_local(const Frame& f, float3 w) { return V3(dot(f.s, w), dot(f.t, w), dot(f.n, w)); }
KYD_DEV float3 to_world(const Frame& f, float3 l) { return add(add(mul(f.s, l.x), mul(f.t, l.y)), mul(f.n, l.z)); }

// offset_ray_origin ky.cpp:614-620
KYD_DEV float3 offset_ray_origin(float3 position, float3 normal, float3 direction)
{
    float3 offset = mul(normal, 0.01f); // (float)1e-2
    if (dot(normal, direction) < 0)
        offset = neg(offset);
    return add(position, offset);
}

// ---- sampling warps ky.cpp:710-808 ---------------------------------------------------------------------
KYD_DEV float2 concentric_disk_sample(float2 u)
{
    float rx = 2.f * u.x - 1, ry = 2.f * u.y - 1;
    if (rx == 0 && ry == 0)
        return make_float2(0.f, 0.f);
    float radius, theta;
    if (fabsf(rx) > fabsf(ry))
    {
        radius = rx;
        theta = KYD_PI_OVER4 * (ry / rx);
    }
    else
    {
        radius = ry;
        theta = KYD_PI_OVER2 - KYD_PI_OVER4 * (rx / ry);
    }
    float s, c;
    cr_sincos(theta, &s, &c);
    return make_float2(c * radius, s * radius);
}

KYD_DEV float3 cosine_hemisphere_sample(float2 u)
{
    float2 p = concentric_disk_sample(u);
    float z = __fsqrt_rn(max_std(0.f, 1 - p.x * p.x - p.y * p.y));
    return V3(p.x, p.y, z);
}

KYD_DEV float3 uniform_sphere_sample(float2 u)
{
    float z = 1 - 2 * u.x;
    float radius = __fsqrt_rn(max_std(0.f, 1.f - z * z));
    float phi = 2 * KYD_PI * u.y;
    float s, c;
    cr_sincos(phi, &s, &c);
    return V3(radius * c, radius * s, z);
}

// ---- sampler: counter-seeded LCG48 / debug sampler ------------------------------------------------------
KYD_DEV unsigned long long mix64(unsigned long long z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

#define KYD_LCG_A 0x5DEECE66Dull
#define KYD_LCG_C 0xBull
#define KYD_LCG_MASK 0xFFFFFFFFFFFFull

// k LCG steps are one: X -> (A^k X + C (A^(k-1) + ... + 1)) mod 2^48.  Row l = the jump over the 4 l draws the light loop
// spends on the lights before light l (ky.cpp:3866-3868); generated by scripts, checked against the
// step-by-step loop by tests/test_gpu_kat.py through every film of a multi-light scene.
__constant__ unsigned long long c_lcg_jump4[17][2] = {
    { 0x000000000001ull, 0x000000000000ull },
    { 0x32EB772C5F11ull, 0x2D3873C4CD04ull },
    { 0x75489F259F21ull, 0x7CBA449AE648ull },
    { 0x199C3838D031ull, 0xD4CF89E2CFCCull },
    { 0x6DC260740241ull, 0x0D0352014D90ull },
    { 0xB05B0EB64551ull, 0x4C56A6636394ull },
    { 0x35692EBFA961ull, 0x83F34BC255D8ull },
    { 0x230D6E413E71ull, 0x9213C2A7A85Cull },
    { 0xFAC6CAED1481ull, 0x1E4C4C311F20ull },
    { 0x8633F1863B91ull, 0x40D8F714BE24ull },
    { 0x18F17DF0C3A1ull, 0xAE50F8E4C968ull },
    { 0xDCE22C41BCB1ull, 0x205FD793C4ECull },
    { 0x8BEF0ACF36C1ull, 0x7F29273874B0ull },
    { 0x231EBD4041D1ull, 0x5E03E011DCB4ull },
    { 0x5FC3E09CEDE1ull, 0x6D8690CB40F8ull },
    { 0xE973A05E4AF1ull, 0xD4ADF100257Cull },
    { 0xAB768C7E6901ull, 0xF77B98004E40ull },
};

struct Sampler
{
    unsigned long long state;
    bool debug;

    KYD_DEV void start(int kind, unsigned long long seed, int x, int y, int sample_index)
    {
        debug = (kind == KYD_SAMPLER_DEBUG);
        unsigned long long key = (unsigned long long)sample_index | ((unsigned long long)x << 24) | ((unsigned long long)y << 40);
        state = mix64(seed * 0x9E3779B97F4A7C15ull + key) >> 16;
    }
    KYD_DEV float get_float()
    {
        if (debug) return 0.5f;
        state = (state * KYD_LCG_A + KYD_LCG_C) & KYD_LCG_MASK;
        return (float)(unsigned)(state >> 24) * 0x1p-24f; // 24-bit integer: exact
    }
    KYD_DEV float2 get_float2()
    {
        float x = get_float();
        float y = get_float();
        return make_float2(x, y);
    }
    KYD_DEV void skip(int n)
    {
        if (debug) return;
        for (int i = 0; i < n; ++i)
            state = (state * KYD_LCG_A + KYD_LCG_C) & KYD_LCG_MASK;
    }
    // skip(4 * lights), lights <= 16, in one step
    KYD_DEV void skip_lights(int lights)
    {
        if (debug) return;
        state = (state * c_lcg_jump4[lights][0] + c_lcg_jump4[lights][1]) & KYD_LCG_MASK;
    }
    // get_camera_sample's offset inside the pixel (ky.cpp:966-974): two draws, or -- KYD_SAMPLER_TRAPEZOIDAL -- smallpt's
    // tent filter on a 2x2 sub-pixel grid (smallpt2pbrt/smallpt_rewrite.cpp:449-466 in float; sub-pixel = sample / (spp / 4))
    KYD_DEV float2 camera_jitter(int kind, int spp, int sample_index)
    {
        if (kind != KYD_SAMPLER_TRAPEZOIDAL)
            return get_float2();
        const int sub_pixel = sample_index / (spp / 4);
        const int sub_x = sub_pixel % 2, sub_y = sub_pixel / 2;
        const float random1 = 2 * get_float();
        const float random2 = 2 * get_float();
        const float delta_x = random1 < 1 ? __fsqrt_rn(random1) - 1 : 1 - __fsqrt_rn(2 - random1);
        const float delta_y = random2 < 1 ? __fsqrt_rn(random2) - 1 : 1 - __fsqrt_rn(2 - random2);
        return make_float2(((float)sub_x + delta_x + 0.5f) / 2, ((float)sub_y + delta_y + 0.5f) / 2);
    }
};

// patch P5 of oracle/ref/build_ref.sh: stateless plastic lobe draw
KYD_DEV float plastic_random(float3 p, float3 wo)
{
    unsigned long long h = mix64((unsigned long long)__float_as_uint(wo.y) | ((unsigned long long)__float_as_uint(wo.z) << 32));
    h = mix64(((unsigned long long)__float_as_uint(p.z) | ((unsigned long long)__float_as_uint(wo.x) << 32)) ^ h);
    h = mix64(((unsigned long long)__float_as_uint(p.x) | ((unsigned long long)__float_as_uint(p.y) << 32)) ^ h);
    return (float)(unsigned)(h >> 40) * 0x1p-24f;
}

// Scene traits: which light kinds / light shapes can occur, known when the scene is uploaded.  Kernels instantiated
// for a restricted scene drop the code of the other kinds (the shade kernels are instruction-fetch sensitive:
// profiles/r01_ab_variants.txt).  0 = anything; 1 = area lights on rectangles only (Cornell); 2 = area lights on
// spheres only (Veach).  TRAITS_AREA_RECTANGLE additionally means: exactly ONE light (no light loop).
enum { TRAITS_ANY = 0, TRAITS_AREA_RECTANGLE = 1, TRAITS_AREA_SPHERE = 2 };
template <int TRAITS> KYD_DEV int light_kind(const DevLight& l) { return TRAITS == TRAITS_ANY ? l.kind : KYD_LIGHT_AREA; }
template <int TRAITS> KYD_DEV int light_shape_kind(const DevShape& s)
{
    return TRAITS == TRAITS_AREA_RECTANGLE ? KYD_SHAPE_RECTANGLE : TRAITS == TRAITS_AREA_SPHERE ? KYD_SHAPE_SPHERE : s.kind;
}

// ---- shapes ky.cpp:1009-1519 ------------------------------------------------------------------------------
struct Ray { float3 o, d; float tmax; };
struct HitGeom { float3 position, normal, wo; };

KYD_DEV float3 ray_at(const Ray& r, float t) { return add(r.o, mul(r.d, t)); }

KYD_DEV bool is_equal0(float x) // is_equal(x, 0) ky.cpp:213-220
{
    float m = 1.f;
    if (m < fabsf(x)) m = fabsf(x);
    return fabsf(x - 0.f) <= KYD_FLT_EPSILON * m;
}

// the hit test of shape_t::intersect WITHOUT filling the isect: returns true and the distance when the
// shape is hit inside (epsilon, tmax).  The isect is a pure function of (ray, distance, shape) and is
// filled once for the final hit by shape_hit_geom().
KYD_DEV bool shape_hit_distance_kind(const DevShape& s, int kind, const Ray& r, float tmax, float* out_t)
{
    switch (kind)
    {
    case KYD_SHAPE_SPHERE: // ky.cpp:1365-1383
    {
        float3 oc = sub(s.p0, r.o);
        float neg_b = dot(oc, r.d);
        float discr = neg_b * neg_b - dot(oc, oc) + s.radius_sq;
        if (discr >= 0)
        {
            float sqrt_discr = __fsqrt_rn(discr);
            float t = neg_b - sqrt_discr;
            if (t > KYD_SHAPE_EPSILON && t < tmax) { *out_t = t; return true; }
            t = neg_b + sqrt_discr;
            if (t > KYD_SHAPE_EPSILON && t < tmax) { *out_t = t; return true; }
        }
        return false;
    }
    case KYD_SHAPE_RECTANGLE: // ky.cpp:1265-1285
    {
        float3 oa = sub(s.p0, r.o), ob = sub(s.p1, r.o), oc = sub(s.p2, r.o), od = sub(s.p3, r.o);
        float v0d = dot(cross(oc, ob), r.d);
        float v1d = dot(cross(ob, oa), r.d);
        float v2d = dot(cross(oa, od), r.d);
        float v3d = dot(cross(od, oc), r.d);
        if (((v0d < 0.f) && (v1d < 0.f) && (v2d < 0.f) && (v3d < 0.f)) ||
            ((v0d >= 0.f) && (v1d >= 0.f) && (v2d >= 0.f) && (v3d >= 0.f)))
        {
            float t = dot(s.n, oa) / dot(s.n, r.d);
            if ((t > KYD_SHAPE_EPSILON) && (t < tmax)) { *out_t = t; return true; }
        }
        return false;
    }
    case KYD_SHAPE_TRIANGLE: // ky.cpp:1183-1204
    {
        float3 oa = sub(s.p0, r.o), ob = sub(s.p1, r.o), oc = sub(s.p2, r.o);
        float v0d = dot(cross(oc, ob), r.d);
        float v1d = dot(cross(ob, oa), r.d);
        float v2d = dot(cross(oa, oc), r.d);
        if (((v0d < 0.f) && (v1d < 0.f) && (v2d < 0.f)) || ((v0d >= 0.f) && (v1d >= 0.f) && (v2d >= 0.f)))
        {
            float t = dot(s.n, oa) / dot(s.n, r.d);
            if ((t > KYD_SHAPE_EPSILON) && (t < tmax)) { *out_t = t; return true; }
        }
        return false;
    }
    default: // disk ky.cpp:1113-1128
    {
        if (is_equal0(dot(r.d, s.n)))
            return false;
        float3 op = sub(s.p0, r.o);
        float t = dot(s.n, op) / dot(s.n, r.d);
        if ((t > KYD_SHAPE_EPSILON) && (t < tmax))
        {
            float3 p = ray_at(r, t);
            if (distance(s.p0, p) <= s.radius) { *out_t = t; return true; }
        }
        return false;
    }
    }
}

KYD_DEV bool shape_hit_distance(const DevShape& s, const Ray& r, float tmax, float* out_t)
{
    return shape_hit_distance_kind(s, s.kind, r, tmax, out_t);
}

// the isect_t a shape builds for a hit at distance t (ky.cpp:1125, 1208, 1288-1290, 1388-1389)
KYD_DEV HitGeom shape_hit_geom_kind(const DevShape& s, int kind, const Ray& r, float t)
{
    HitGeom g;
    g.position = ray_at(r, t);
    g.wo = neg(r.d);
    switch (kind)
    {
    case KYD_SHAPE_SPHERE: g.normal = normalize(sub(g.position, s.p0)); break;
    case KYD_SHAPE_RECTANGLE: g.normal = dot(s.n, r.d) <= 0 ? s.n : neg(s.n); break;
    default: g.normal = s.n; break;
    }
    return g;
}

KYD_DEV HitGeom shape_hit_geom(const DevShape& s, const Ray& r, float t) { return shape_hit_geom_kind(s, s.kind, r, t); }

// shape_t::sample_position ky.cpp:1144, 1225, 1307, 1404
template <int TRAITS = TRAITS_ANY>
KYD_DEV void shape_sample_position(const DevShape& s, float2 u, float3* lp, float3* ln, float* area_pdf)
{
    switch (light_shape_kind<TRAITS>(s))
    {
    case KYD_SHAPE_SPHERE:
    {
        float3 direction = uniform_sphere_sample(u);
        *lp = add(s.p0, mul(direction, s.radius));
        *ln = normalize(direction);
        break;
    }
    case KYD_SHAPE_RECTANGLE:
        *lp = add(add(s.p1, mul(sub(s.p0, s.p1), u.x)), mul(sub(s.p2, s.p1), u.y));
        *ln = normalize(s.n);
        break;
    case KYD_SHAPE_TRIANGLE:
    {
        float su0 = __fsqrt_rn(u.x);
        float bx = 1 - su0, by = u.y * su0;
        *lp = add(add(mul(s.p0, bx), mul(s.p1, by)), mul(s.p2, 1 - bx - by));
        *ln = s.n;
        break;
    }
    default:
    {
        Frame f = frame_from_z(s.n);
        float2 sp = concentric_disk_sample(u);
        *lp = add(s.p0, mul(add(mul(f.s, sp.x), mul(f.t, sp.y)), s.radius));
        *ln = normalize(s.n);
        break;
    }
    }
    *area_pdf = 1 / s.area;
}

// sphere_t::sample_direction from a point inside the sphere (ky.cpp:1422-1444): uniform point on the sphere, area pdf converted
// with the SHADING normal -- the same statements as the inline branch below, behind a call
struct SphereInsideSample { float3 lp, ln; float pdf; };
__device__ __noinline__ SphereInsideSample sphere_sample_from_inside(float3 center, float radius, float area, float3 p, float3 n_shade, float2 u)
{
    SphereInsideSample r;
    float3 direction = uniform_sphere_sample(u);
    r.lp = add(center, mul(direction, radius));
    r.ln = normalize(direction);
    const float area_pdf = 1 / area;
    float3 wi = sub(r.lp, p);
    if (msq(wi) == 0)
        r.pdf = 0;
    else
    {
        wi = normalize(wi);
        r.pdf = area_pdf * distance_sq(r.lp, p) / abs_dot(n_shade, neg(wi));
    }
    if (isinf(r.pdf))
        r.pdf = 0.f;
    return r;
}

// shape_t::sample_direction ky.cpp:1028-1051 and sphere_t's override ky.cpp:1419-1501
template <int TRAITS = TRAITS_ANY>
KYD_DEV void shape_sample_direction(const DevShape& s, float3 p, float3 n_shade, float2 u, float3* lp, float3* ln, float* pdf)
{
    if (light_shape_kind<TRAITS>(s) == KYD_SHAPE_SPHERE)
    {
        float3 center = s.p0;
        float radius = s.radius;
        if (distance_sq(p, center) <= radius * radius)
        {
            if (TRAITS == TRAITS_AREA_SPHERE)
            {
                // (cold in the scenes the sphere-light kernels are built for, and those kernels are instruction-fetch bound)
                const SphereInsideSample in = sphere_sample_from_inside(s.p0, s.radius, s.area, p, n_shade, u);
                *lp = in.lp; *ln = in.ln; *pdf = in.pdf;
                return;
            }
            float area_pdf;
            shape_sample_position<TRAITS>(s, u, lp, ln, &area_pdf);
            float3 wi = sub(*lp, p);
            if (msq(wi) == 0)
                *pdf = 0;
            else
            {
                wi = normalize(wi);
                *pdf = area_pdf * distance_sq(*lp, p) / abs_dot(n_shade, neg(wi));
            }
            if (isinf(*pdf))
                *pdf = 0.f;
            return;
        }

        float dist = distance(p, center);
        float inv_dist = 1 / dist;
        float sin_theta_max = radius * inv_dist;
        float sin_theta_max_sq = sin_theta_max * sin_theta_max;
        float inv_sin_theta_max = 1 / sin_theta_max;
        float cos_theta_max = __fsqrt_rn(max_std(0.f, 1 - sin_theta_max_sq));

        float cos_theta = (cos_theta_max - 1) * u.x + 1;
        float sin_theta_sq = 1 - cos_theta * cos_theta;
        if (sin_theta_max_sq < 0.00068523f)
        {
            sin_theta_sq = sin_theta_max_sq * u.x;
            cos_theta = __fsqrt_rn(1 - sin_theta_sq);
        }

        float cos_alpha = sin_theta_sq * inv_sin_theta_max +
            cos_theta * __fsqrt_rn(max_std(0.f, 1.f - sin_theta_sq * inv_sin_theta_max * inv_sin_theta_max));
        float sin_alpha = __fsqrt_rn(max_std(0.f, 1.f - cos_alpha * cos_alpha));
        float phi = u.y * 2 * KYD_PI;

        float3 normal = mul(sub(center, p), inv_dist);
        Frame f = frame_from_z(normal);

        float sp, cp;
        cr_sincos(phi, &sp, &cp);
        float3 wn = add(add(mul(neg(f.s), sin_alpha * cp), mul(neg(f.t), sin_alpha * sp)), mul(neg(f.n), cos_alpha));
        *lp = add(center, mul(wn, radius));
        *ln = wn;
        *pdf = 1 / (2 * KYD_PI * (1 - cos_theta_max));
        return;
    }

    float area_pdf;
    shape_sample_position<TRAITS>(s, u, lp, ln, &area_pdf);
    float3 wi = sub(*lp, p);
    if (msq(wi) == 0)
        *pdf = 0;
    else
    {
        wi = normalize(wi);
        *pdf = area_pdf * distance_sq(*lp, p) / abs_dot(*ln, neg(wi));
        if (isinf(*pdf))
            *pdf = 0.f;
    }
}

// shape_t::pdf_direction (ky.cpp:1055-1090) for a sphere seen from inside: the general branch below with the kind fixed
__device__ __noinline__ float sphere_pdf_from_inside(const DevShape& s, float3 p, float3 n_shade, float3 wi)
{
    Ray r;
    r.o = offset_ray_origin(p, n_shade, wi);
    r.d = wi;
    r.tmax = KYD_INF;
    float t;
    if (!shape_hit_distance_kind(s, KYD_SHAPE_SPHERE, r, r.tmax, &t))
        return 0.f;
    HitGeom g = shape_hit_geom_kind(s, KYD_SHAPE_SPHERE, r, t);
    float pdf = distance_sq(p, g.position) / (abs_dot(g.normal, neg(wi)) * s.area);
    if (isinf(pdf))
        pdf = 0.f;
    return pdf;
}

// shape_t::pdf_direction ky.cpp:1055-1090 and sphere_t's override ky.cpp:1503-1513
template <int TRAITS = TRAITS_ANY>
KYD_DEV float shape_pdf_direction(const DevShape& s, float3 p, float3 n_shade, float3 wi)
{
    if (light_shape_kind<TRAITS>(s) == KYD_SHAPE_SPHERE)
    {
        if (!(distance_sq(p, s.p0) <= s.radius * s.radius))
        {
            float sin_theta_max_sq = s.radius * s.radius / distance_sq(p, s.p0);
            float cos_theta_max = __fsqrt_rn(max_std(0.f, 1 - sin_theta_max_sq));
            return 1 / (2 * KYD_PI * (1 - cos_theta_max));
        }
    }
    if (TRAITS == TRAITS_AREA_SPHERE)
        return sphere_pdf_from_inside(s, p, n_shade, wi);   // (cold, behind a call: see sphere_sample_from_inside)
    Ray r;
    r.o = offset_ray_origin(p, n_shade, wi);
    r.d = wi;
    r.tmax = KYD_INF;
    float t;
    if (!shape_hit_distance_kind(s, light_shape_kind<TRAITS>(s), r, r.tmax, &t))
        return 0.f;
    HitGeom g = shape_hit_geom_kind(s, light_shape_kind<TRAITS>(s), r, t);
    float pdf = distance_sq(p, g.position) / (abs_dot(g.normal, neg(wi)) * s.area);
    if (isinf(pdf))
        pdf = 0.f;
    return pdf;
}

// ---- BSDFs ky.cpp:1918-2555, materials ky.cpp:2579-2682 -----------------------------------------------------
enum { LOBE_LAMBERT = 0, LOBE_MIRROR = 1, LOBE_FRESNEL = 2, LOBE_PHONG = 3 };
enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16 };

struct Bsdf
{
    Frame f;
    float3 a;       // lambert albedo / mirror + fresnel reflectance / phong specular reflectance
    float3 t;       // fresnel transmittance
    float eta_t;    // eta_i is 1 (ky.cpp:2630)
    float exponent;
    int lobe;
};

struct BsdfSample { float3 f, wi; float pdf; int type; };

KYD_DEV bool bsdf_is_delta(int lobe) { return lobe == LOBE_MIRROR || lobe == LOBE_FRESNEL; }
KYD_DEV bool same_hemisphere(float3 w, float3 wp) { return w.z * wp.z > 0; }

KYD_DEV float3 reflect_z(float3 wo) // reflect(wo, (0,0,1)) ky.cpp:1923-1928
{
    float3 n = V3(0, 0, 1);
    return add(neg(wo), mul(n, 2 * dot(wo, n)));
}

KYD_DEV float fresnel_dielectric(float cos_theta_i, float eta_i, float eta_t) // ky.cpp:1963-1996
{
    cos_theta_i = clamp_std(cos_theta_i, -1.f, 1.f);
    if (!(cos_theta_i > 0.f))
    {
        float tmp = eta_i; eta_i = eta_t; eta_t = tmp;
        cos_theta_i = fabsf(cos_theta_i);
    }
    float sin_theta_i = __fsqrt_rn(max_std(0.f, 1 - cos_theta_i * cos_theta_i));
    float sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1)
        return 1;
    float cos_theta_t = __fsqrt_rn(max_std(0.f, 1 - sin_theta_t * sin_theta_t));
    float r_para = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
    float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
    return (r_para * r_para + r_perp * r_perp) / 2;
}

KYD_DEV bool refract(float3 wi, float3 normal, float eta_ratio, float3* wt) // ky.cpp:1931-1957
{
    float cos_theta_i = dot(normal, wi);
    float sin_theta_i_sq = max_std(0.f, 1 - cos_theta_i * cos_theta_i);
    float sin_theta_t_sq = eta_ratio * eta_ratio * sin_theta_i_sq;
    if (sin_theta_t_sq >= 1)
        return false;
    float cos_theta_t = __fsqrt_rn(1 - sin_theta_t_sq);
    *wt = add(mul(neg(wi), eta_ratio), mul(normal, eta_ratio * cos_theta_i - cos_theta_t));
    return true;
}

KYD_DEV float3 bsdf_eval_local(const Bsdf& b, float3 wo, float3 wi) // ky.cpp:2227, 2289, 2352, 2489
{
    if (b.lobe == LOBE_LAMBERT)
    {
        if (!same_hemisphere(wo, wi))
            return KYD_BLACK;
        return mul(b.a, KYD_INV_PI);
    }
    if (b.lobe == LOBE_PHONG)
    {
        if (!same_hemisphere(wo, wi))
            return KYD_BLACK;
        float3 wr = reflect_z(wo);
        float cos_alpha = dot(wr, wi);
        float3 rho = mul(mul(b.a, b.exponent + 2.f), KYD_INV_2PI);
        return mul(rho, cr_pow(cos_alpha, b.exponent));
    }
    return KYD_BLACK;
}

KYD_DEV float bsdf_pdf_local(const Bsdf& b, float3 wo, float3 wi) // ky.cpp:2237, 2290, 2353, 2502
{
    if (b.lobe == LOBE_LAMBERT)
        return same_hemisphere(wo, wi) ? fabsf(wi.z) * KYD_INV_PI : 0;
    if (b.lobe == LOBE_PHONG)
    {
        float3 wr = reflect_z(wo);
        float cos_theta = max_std(0.f, dot(wr, wi));
        return (b.exponent + 1.f) * cr_pow(cos_theta, b.exponent) * KYD_INV_2PI;
    }
    return 0;
}

// eval_ and pdf_ of the same (wo, wi).  Phong raises cos_alpha (eval_, ky.cpp:2499) and max(0, cos_alpha)
// (pdf_, ky.cpp:2548) to the same exponent: for cos_alpha > 0 that is one pow() instead of two -- the same
// value by definition; every other case (zero, negative, NaN) takes the two separate calls.
KYD_DEV void bsdf_eval_pdf_local(const Bsdf& b, float3 wo, float3 wi, float3* f, float* pdf)
{
    if (b.lobe == LOBE_PHONG)
    {
        float3 wr = reflect_z(wo);
        float cos_alpha = dot(wr, wi);
        if (cos_alpha > 0.f)
        {
            const float pw = cr_pow(cos_alpha, b.exponent);
            float3 rho = mul(mul(b.a, b.exponent + 2.f), KYD_INV_2PI);
            *f = same_hemisphere(wo, wi) ? mul(rho, pw) : KYD_BLACK;
            *pdf = (b.exponent + 1.f) * pw * KYD_INV_2PI;
            return;
        }
    }
    *f = bsdf_eval_local(b, wo, wi);
    *pdf = bsdf_pdf_local(b, wo, wi);
}

KYD_DEV float3 bsdf_eval(const Bsdf& b, float3 world_wo, float3 world_wi) // ky.cpp:2162
{
    return bsdf_eval_local(b, to_local(b.f, world_wo), to_local(b.f, world_wi));
}
KYD_DEV float bsdf_pdf(const Bsdf& b, float3 world_wo, float3 world_wi) // ky.cpp:2167
{
    return bsdf_pdf_local(b, to_local(b.f, world_wo), to_local(b.f, world_wi));
}

// bsdf_t::sample ky.cpp:2173 with sample_ of ky.cpp:2242, 2292, 2355, 2510
KYD_DEV BsdfSample bsdf_sample(const Bsdf& b, float3 world_wo, float2 u)
{
    BsdfSample s;
    s.f = KYD_BLACK;
    s.wi = V3(0, 0, 0);
    s.pdf = 0;
    s.type = 0;
    float3 wo = to_local(b.f, world_wo);

    if (b.lobe == LOBE_LAMBERT)
    {
        s.wi = cosine_hemisphere_sample(u);
        if (wo.z < 0)
            s.wi.z *= -1;
        s.f = bsdf_eval_local(b, wo, s.wi);
        s.pdf = bsdf_pdf_local(b, wo, s.wi);
        s.type = BSDF_REFLECTION | BSDF_DIFFUSE;
    }
    else if (b.lobe == LOBE_MIRROR)
    {
        s.wi = V3(-wo.x, -wo.y, wo.z);
        s.f = cdiv(b.a, fabsf(s.wi.z));
        s.pdf = 1;
        s.type = BSDF_REFLECTION | BSDF_SPECULAR;
    }
    else if (b.lobe == LOBE_FRESNEL)
    {
        float reflect_percent = fresnel_dielectric(wo.z, 1.f, b.eta_t);
        float refract_percent = 1 - reflect_percent;
        if (u.x < reflect_percent)
        {
            s.wi = V3(-wo.x, -wo.y, wo.z);
            s.pdf = reflect_percent;
            s.f = cdiv(mul(b.a, reflect_percent), fabsf(s.wi.z));
            s.type = BSDF_REFLECTION | BSDF_SPECULAR;
        }
        else
        {
            float3 normal = V3(0, 0, 1);
            bool into = dot(normal, wo) > 0;
            float3 wo_normal = into ? normal : mul(normal, -1.f);
            float eta = into ? 1.f / b.eta_t : b.eta_t / 1.f;
            if (refract(wo, wo_normal, eta, &s.wi))
            {
                s.pdf = refract_percent;
                s.f = cdiv(mul(b.t, refract_percent), fabsf(s.wi.z));
                s.type = BSDF_TRANSMISSION | BSDF_SPECULAR;
            }
        }
    }
    else // LOBE_PHONG
    {
        float phi = 2.f * KYD_PI * u.x;
        float cos_theta = cr_pow(u.y, 1.f / (b.exponent + 1.f));
        float sin_theta = __fsqrt_rn(1.f - cos_theta * cos_theta);
        float sp, cp;
        cr_sincos(phi, &sp, &cp);
        float3 local = V3(cp * sin_theta, sp * sin_theta, cos_theta);
        Frame fr = frame_from_z(reflect_z(wo));
        s.wi = to_world(fr, local);
        if (wo.z < 0)
            s.wi.z *= -1;
        bsdf_eval_pdf_local(b, wo, s.wi, &s.f, &s.pdf);
        s.type = BSDF_REFLECTION | BSDF_GLOSSY;
    }

    s.wi = to_world(b.f, s.wi);
    return s;
}

// material_t::scattering ky.cpp:2587, 2604, 2628, 2661
KYD_DEV void material_scattering(const DevMaterial& m, const HitGeom& g, Bsdf* b)
{
    b->f = frame_from_z(g.normal);
    b->t = KYD_BLACK;
    b->eta_t = 1.f;
    b->exponent = 0.f;
    if (m.kind == KYD_MAT_MATTE)
    {
        b->lobe = LOBE_LAMBERT;
        b->a = m.diffuse;
    }
    else if (m.kind == KYD_MAT_MIRROR)
    {
        b->lobe = LOBE_MIRROR;
        b->a = m.specular;
    }
    else if (m.kind == KYD_MAT_GLASS)
    {
        b->lobe = LOBE_FRESNEL;
        b->eta_t = m.eta;
        b->a = m.specular;
        b->t = m.transmission;
    }
    else
    {
        float random = plastic_random(g.position, g.wo);
        if (random < m.p_specular)
        {
            b->lobe = LOBE_PHONG;
            b->a = m.plastic_phong;
            b->exponent = m.exponent;
        }
        else
        {
            b->lobe = LOBE_LAMBERT;
            b->a = m.plastic_lambert;
        }
    }
}

// ---- per-surface data: constant memory, or global memory for a large scene -----------------------------------
// (KYD_BIG_SCENE: this translation unit is the large-scene build of the kernels, kyd_kernels_big.cu)
#if KYD_BIG_SCENE
KYD_DEV const DevShape& surface_shape(int i) { return c_scene.big_shape[i]; }
KYD_DEV int surface_material(int i) { return c_scene.big_material[i]; }
KYD_DEV int surface_light(int i) { return c_scene.big_light[i]; }
#else
KYD_DEV const DevShape& surface_shape(int i) { return c_scene.surf_shape[i]; }
KYD_DEV int surface_material(int i) { return c_scene.surf_material[i]; }
KYD_DEV int surface_light(int i) { return c_scene.surf_light[i]; }
#endif

// Per-surface data of the vertex being shaded, in shared memory (the one-light headline kernels; KYD_SURFACE_SMEM=1).  A vertex reads its surface's
// normal / centre, material parameters and light index by the surface index of ITS path: in constant memory that is a load
// with up to a dozen different addresses in a warp, replayed once per address (what made k_nee's per-light loads cost 12 % of
// that kernel, profiles/r02_ab_variants.txt); shared memory serves different addresses in one pass.  The material is copied
// per surface, which also removes the surface -> material indirection.
struct SurfaceTable
{
    DevShape shape[KYD_MAX_SURFACES];
    DevMaterial material[KYD_MAX_SURFACES];   // materials[surf_material[i]]
    int light[KYD_MAX_SURFACES];
};
// (measured: C5 2638 vs 2644-2666 Msamples/s without it -- a dozen loads per vertex, not on the critical path: left off)
#ifndef KYD_SURFACE_SMEM
#define KYD_SURFACE_SMEM 0
#endif
#if !KYD_BIG_SCENE && KYD_SURFACE_SMEM
#define KYD_HAS_SURFACE_TABLE 1
KYD_DEV SurfaceTable& surface_table()
{
    __shared__ SurfaceTable t;
    return t;
}
// every thread of the block, once, before the first vertex
KYD_DEV void stage_surfaces()
{
    SurfaceTable& t = surface_table();
    const int n = c_scene.n_surfaces;
    constexpr int WS = (int)(sizeof(DevShape) / 4), WM = (int)(sizeof(DevMaterial) / 4);
    unsigned* shape_words = reinterpret_cast<unsigned*>(t.shape);
    const unsigned* shape_src = reinterpret_cast<const unsigned*>(c_scene.surf_shape);
    for (int i = threadIdx.x; i < n * WS; i += blockDim.x)
        shape_words[i] = shape_src[i];
    unsigned* material_words = reinterpret_cast<unsigned*>(t.material);
    for (int i = threadIdx.x; i < n * WM; i += blockDim.x)
    {
        const int k = i / WM, f = i - k * WM;
        material_words[i] = reinterpret_cast<const unsigned*>(&c_scene.materials[c_scene.surf_material[k]])[f];
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        t.light[i] = c_scene.surf_light[i];
    __syncthreads();
}
template <bool TABLE> KYD_DEV const DevShape& surface_shape_of(int i) { return TABLE ? surface_table().shape[i] : surface_shape(i); }
template <bool TABLE> KYD_DEV const DevMaterial& surface_material_of(int i) { return TABLE ? surface_table().material[i] : c_scene.materials[surface_material(i)]; }
template <bool TABLE> KYD_DEV int surface_light_of(int i) { return TABLE ? surface_table().light[i] : surface_light(i); }
#else
#define KYD_HAS_SURFACE_TABLE 0
KYD_DEV void stage_surfaces() {}
template <bool TABLE> KYD_DEV const DevShape& surface_shape_of(int i) { return surface_shape(i); }
template <bool TABLE> KYD_DEV const DevMaterial& surface_material_of(int i) { return c_scene.materials[surface_material(i)]; }
template <bool TABLE> KYD_DEV int surface_light_of(int i) { return surface_light(i); }
#endif

// ---- scene traversal ky.cpp:3077-3088, 3172-3206 ------------------------------------------------------------

// The distance at which shape_t::intersect would report a hit if tmax were unbounded: the same arithmetic and the
// same `distance > epsilon` test as shape_hit_distance_kind(), with `distance < ray.distance()` left to the caller.
// (Sphere: the far root is tried only when the near one is not beyond epsilon; a near root rejected by tmax implies
// the far one is rejected too, so picking the candidate first is equivalent to ky.cpp:1375-1383.)
template <int KIND>
KYD_DEV bool shape_hit_candidate(const DevShape& s, const Ray& r, float* out_t)
{
    if (KIND == KYD_SHAPE_SPHERE)
    {
        float3 oc = sub(s.p0, r.o);
        float neg_b = dot(oc, r.d);
        float discr = neg_b * neg_b - dot(oc, oc) + s.radius_sq;
        if (!(discr >= 0))
            return false;
        float sqrt_discr = __fsqrt_rn(discr);
        float t = neg_b - sqrt_discr;
        if (!(t > KYD_SHAPE_EPSILON))
            t = neg_b + sqrt_discr;
        *out_t = t;
        return t > KYD_SHAPE_EPSILON;
    }
    else if (KIND == KYD_SHAPE_RECTANGLE)
    {
        float3 oa = sub(s.p0, r.o), ob = sub(s.p1, r.o), oc = sub(s.p2, r.o), od = sub(s.p3, r.o);
        float v0d = dot(cross(oc, ob), r.d);
        float v1d = dot(cross(ob, oa), r.d);
        float v2d = dot(cross(oa, od), r.d);
        float v3d = dot(cross(od, oc), r.d);
        if (!(((v0d < 0.f) && (v1d < 0.f) && (v2d < 0.f) && (v3d < 0.f)) ||
              ((v0d >= 0.f) && (v1d >= 0.f) && (v2d >= 0.f) && (v3d >= 0.f))))
            return false;
        float t = dot(s.n, oa) / dot(s.n, r.d);
        *out_t = t;
        return t > KYD_SHAPE_EPSILON;
    }
    else if (KIND == KYD_SHAPE_TRIANGLE)
    {
        float3 oa = sub(s.p0, r.o), ob = sub(s.p1, r.o), oc = sub(s.p2, r.o);
        float v0d = dot(cross(oc, ob), r.d);
        float v1d = dot(cross(ob, oa), r.d);
        float v2d = dot(cross(oa, oc), r.d);
        if (!(((v0d < 0.f) && (v1d < 0.f) && (v2d < 0.f)) || ((v0d >= 0.f) && (v1d >= 0.f) && (v2d >= 0.f))))
            return false;
        float t = dot(s.n, oa) / dot(s.n, r.d);
        *out_t = t;
        return t > KYD_SHAPE_EPSILON;
    }
    else
    {
        if (is_equal0(dot(r.d, s.n)))
            return false;
        float3 op = sub(s.p0, r.o);
        float t = dot(s.n, op) / dot(s.n, r.d);
        *out_t = t;
        if (!(t > KYD_SHAPE_EPSILON))
            return false;
        return distance(s.p0, ray_at(r, t)) <= s.radius;
    }
}

#if KYD_BIG_SCENE
// ---- large scenes: the same three queries over a bounding-volume hierarchy in global memory ----------------------
// Any visiting order gives the linear walk's answer as long as no box is skipped that holds a qualifying hit and ties
// follow the list order: boxes are padded at upload, the slab test rejects only intervals that are empty by a margin,
// and a node is skipped for distance only if it starts strictly beyond the current best (equal distances must be seen:
// the lower surface index wins).
KYD_DEV bool shape_hit_candidate_any_kind(const DevShape& s, const Ray& r, float* out_t)
{
    switch (s.kind)
    {
    case KYD_SHAPE_SPHERE: return shape_hit_candidate<KYD_SHAPE_SPHERE>(s, r, out_t);
    case KYD_SHAPE_RECTANGLE: return shape_hit_candidate<KYD_SHAPE_RECTANGLE>(s, r, out_t);
    case KYD_SHAPE_TRIANGLE: return shape_hit_candidate<KYD_SHAPE_TRIANGLE>(s, r, out_t);
    default: return shape_hit_candidate<KYD_SHAPE_DISK>(s, r, out_t);
    }
}

// entry distance of the ray into the node's box, or a negative value if it misses the box within [0, limit]
KYD_DEV float bvh_box_entry(const BvhNode& n, float3 o, float3 inv_d, float limit)
{
    // fminf / fmaxf drop NaNs (0 * inf when the origin lies on a slab plane of an axis the ray is parallel to)
    float t0 = 0.f, t1 = limit;
    const float ax = (n.bmin[0] - o.x) * inv_d.x, bx = (n.bmax[0] - o.x) * inv_d.x;
    t0 = fmaxf(t0, fminf(ax, bx)); t1 = fminf(t1, fmaxf(ax, bx));
    const float ay = (n.bmin[1] - o.y) * inv_d.y, by = (n.bmax[1] - o.y) * inv_d.y;
    t0 = fmaxf(t0, fminf(ay, by)); t1 = fminf(t1, fmaxf(ay, by));
    const float az = (n.bmin[2] - o.z) * inv_d.z, bz = (n.bmax[2] - o.z) * inv_d.z;
    t0 = fmaxf(t0, fminf(az, bz)); t1 = fminf(t1, fmaxf(az, bz));
    return t0 <= t1 * 1.0001f + 1e-4f ? t0 : -1.f;
}

#define KYD_BVH_STACK 48

// MODE 0: closest hit (returns the surface, *out_t the distance); MODE 1: any hit inside (epsilon, r.tmax) (returns 0 / -1);
// MODE 2: any surface other than `light_surface` before it, r.tmax = its distance (returns 0 / -1)
template <int MODE>
__device__ __noinline__ int bvh_query(Ray r, int light_surface, float* out_t)
{
    const float3 inv_d = V3(1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z);
    float tmax = r.tmax;
    int best = -1;
    int stack[KYD_BVH_STACK];
    int top = 0;
    stack[top++] = 0;
    while (top > 0)
    {
        const BvhNode node = c_scene.bvh_nodes[stack[--top]];
        const float entry = bvh_box_entry(node, r.o, inv_d, tmax);
        if (entry < 0.f)
            continue;
        if (node.count == 0)
        {
            if (top + 2 <= KYD_BVH_STACK)
            {
                stack[top++] = node.left + 1;
                stack[top++] = node.left;
            }
            continue;
        }
        for (int k = 0; k < node.count; ++k)
        {
            const int surface = c_scene.bvh_prims[node.left + k];
            float t;
            if (!shape_hit_candidate_any_kind(c_scene.big_shape[surface], r, &t))
                continue;
            if (MODE == 0)
            {
                if (t < tmax || (t == tmax && surface < best))
                {
                    tmax = t;
                    best = surface;
                }
            }
            else if (MODE == 1)
            {
                if (t < r.tmax)
                    return 0;
            }
            else
            {
                if (surface != light_surface && (t < r.tmax || (t == r.tmax && surface < light_surface)))
                    return 0;
            }
        }
    }
    if (MODE == 0)
        *out_t = tmax;
    return best;
}

#endif // KYD_BIG_SCENE

// One loop per shape kind over the kind-sorted copy of the surface list: no per-surface dispatch, and the loop
// counter, the bounds and the shape data stay warp-uniform (uniform-datapath constant loads).  The reference walks
// the list in order with a strict `t < tmax` (ky.cpp:3172-3184), i.e. the closest hit, the LOWEST surface index among
// equal distances; visiting in another order needs that tie rule spelled out.
template <int GROUP, int KIND>
KYD_DEV void scene_closest_kind(const Ray& r, float& tmax, int& best)
{
    const int end = c_scene.kind_end[GROUP];
    KYD_UNROLL_TRAVERSAL
    for (int k = GROUP == 0 ? 0 : c_scene.kind_end[GROUP == 0 ? 0 : GROUP - 1]; k < end; ++k)
    {
        float t;
        if (shape_hit_candidate<KIND>(c_scene.sorted_shape[k], r, &t))
        {
            const int surface = c_scene.sorted_surface[k];
            if (t < tmax || (t == tmax && surface < best))
            {
                tmax = t;
                best = surface;
            }
        }
    }
}

KYD_DEV int scene_closest(const Ray& r, float* out_t)
{
#if KYD_BIG_SCENE
    return bvh_query<0>(r, -1, out_t);
#endif
    float tmax = r.tmax;
    int best = -1;
    scene_closest_kind<0, KYD_SHAPE_RECTANGLE>(r, tmax, best);
    scene_closest_kind<1, KYD_SHAPE_SPHERE>(r, tmax, best);
    scene_closest_kind<2, KYD_SHAPE_TRIANGLE>(r, tmax, best);
    scene_closest_kind<3, KYD_SHAPE_DISK>(r, tmax, best);
    *out_t = tmax;
    return best;
}

template <int GROUP, int KIND>
KYD_DEV bool scene_any_hit_kind(const Ray& r)
{
    const int end = c_scene.kind_end[GROUP];
    KYD_UNROLL_TRAVERSAL
    for (int k = GROUP == 0 ? 0 : c_scene.kind_end[GROUP == 0 ? 0 : GROUP - 1]; k < end; ++k)
    {
        float t;
        if (shape_hit_candidate<KIND>(c_scene.sorted_shape[k], r, &t) && t < r.tmax)
            return true;
    }
    return false;
}

// scene_t::occluded's question (ky.cpp:3187-3206): is any surface hit inside (epsilon, tmax)?  Order-independent.
KYD_DEV bool scene_any_hit(const Ray& r)
{
#if KYD_BIG_SCENE
    return r.tmax > KYD_SHAPE_EPSILON && bvh_query<1>(r, -1, nullptr) == 0;
#endif
    return scene_any_hit_kind<0, KYD_SHAPE_RECTANGLE>(r) || scene_any_hit_kind<1, KYD_SHAPE_SPHERE>(r) ||
           scene_any_hit_kind<2, KYD_SHAPE_TRIANGLE>(r) || scene_any_hit_kind<3, KYD_SHAPE_DISK>(r);
}

// Occlusion form of the BSDF-sampled query: is any surface other than `light_surface` hit before it?  "Before" is the
// list-order closest-hit rule: a smaller distance, or the same distance and a lower surface index.  r.tmax = the distance
// of light_surface along r.
template <int GROUP, int KIND>
KYD_DEV bool scene_blocked_before_kind(const Ray& r, int light_surface)
{
    const int end = c_scene.kind_end[GROUP];
    KYD_UNROLL_TRAVERSAL
    for (int k = GROUP == 0 ? 0 : c_scene.kind_end[GROUP == 0 ? 0 : GROUP - 1]; k < end; ++k)
    {
        float t;
        if (shape_hit_candidate<KIND>(c_scene.sorted_shape[k], r, &t))
        {
            const int surface = c_scene.sorted_surface[k];
            if (surface != light_surface && (t < r.tmax || (t == r.tmax && surface < light_surface)))
                return true;
        }
    }
    return false;
}

KYD_DEV bool scene_blocked_before(const Ray& r, int light_surface)
{
#if KYD_BIG_SCENE
    return bvh_query<2>(r, light_surface, nullptr) == 0;
#endif
    return scene_blocked_before_kind<0, KYD_SHAPE_RECTANGLE>(r, light_surface) || scene_blocked_before_kind<1, KYD_SHAPE_SPHERE>(r, light_surface) ||
           scene_blocked_before_kind<2, KYD_SHAPE_TRIANGLE>(r, light_surface) || scene_blocked_before_kind<3, KYD_SHAPE_DISK>(r, light_surface);
}

// warp-uniform form (every lane calls it; lanes without a query pass light_surface < 0)
template <int GROUP, int KIND>
KYD_DEV bool scene_blocked_before_kind_uniform(const Ray& r, int light_surface, bool done)
{
    const int end = c_scene.kind_end[GROUP];
    KYD_UNROLL_TRAVERSAL
    for (int k = GROUP == 0 ? 0 : c_scene.kind_end[GROUP == 0 ? 0 : GROUP - 1]; k < end; ++k)
    {
        float t;
        if (shape_hit_candidate<KIND>(c_scene.sorted_shape[k], r, &t))
        {
            const int surface = c_scene.sorted_surface[k];
            if (surface != light_surface && (t < r.tmax || (t == r.tmax && surface < light_surface)))
                done = true;
        }
    }
    return done;
}

KYD_DEV bool scene_blocked_before_uniform(const Ray& r, int light_surface)
{
#if KYD_BIG_SCENE
    return light_surface >= 0 && bvh_query<2>(r, light_surface, nullptr) == 0;
#endif
    const bool idle = light_surface < 0;
    bool done = idle;
    done = scene_blocked_before_kind_uniform<0, KYD_SHAPE_RECTANGLE>(r, light_surface, done);
    if (__all_sync(0xffffffffu, done)) return !idle;
    done = scene_blocked_before_kind_uniform<1, KYD_SHAPE_SPHERE>(r, light_surface, done);
    if (__all_sync(0xffffffffu, done)) return !idle;
    done = scene_blocked_before_kind_uniform<2, KYD_SHAPE_TRIANGLE>(r, light_surface, done);
    done = scene_blocked_before_kind_uniform<3, KYD_SHAPE_DISK>(r, light_surface, done);
    return done && !idle;
}

// The same question asked by all 32 lanes of a warp together (lanes without a query pass tmax < 0): no per-lane early
// exit -- in SIMT it saves nothing while any lane is still looking -- so the loops stay warp-uniform; the warp leaves
// between kind groups once every lane has its answer.
template <int GROUP, int KIND>
KYD_DEV bool scene_any_hit_kind_uniform(const Ray& r, bool hit)
{
    const int end = c_scene.kind_end[GROUP];
    KYD_UNROLL_TRAVERSAL
    for (int k = GROUP == 0 ? 0 : c_scene.kind_end[GROUP == 0 ? 0 : GROUP - 1]; k < end; ++k)
    {
        float t;
        if (shape_hit_candidate<KIND>(c_scene.sorted_shape[k], r, &t) && t < r.tmax)
            hit = true;
    }
    return hit;
}

KYD_DEV bool scene_any_hit_uniform(const Ray& r)
{
#if KYD_BIG_SCENE
    return r.tmax > KYD_SHAPE_EPSILON && bvh_query<1>(r, -1, nullptr) == 0;
#endif
    bool hit = !(r.tmax > KYD_SHAPE_EPSILON);   // nothing lies in (epsilon, tmax): answered (the caller ignores it)
    if (__all_sync(0xffffffffu, hit)) return false;
    hit = scene_any_hit_kind_uniform<0, KYD_SHAPE_RECTANGLE>(r, hit);
    if (__all_sync(0xffffffffu, hit)) return r.tmax > KYD_SHAPE_EPSILON;
    hit = scene_any_hit_kind_uniform<1, KYD_SHAPE_SPHERE>(r, hit);
    if (__all_sync(0xffffffffu, hit)) return r.tmax > KYD_SHAPE_EPSILON;
    hit = scene_any_hit_kind_uniform<2, KYD_SHAPE_TRIANGLE>(r, hit);
    hit = scene_any_hit_kind_uniform<3, KYD_SHAPE_DISK>(r, hit);
    return hit && r.tmax > KYD_SHAPE_EPSILON;
}

// ---- two-phase traversal (wavefront kernels, scenes in constant memory) ---------------------------------------------------
// The reference tests a ray against a rectangle with four edge functions E = ((v_i - o) x (v_j - o)) . d whose signs must
// agree, then computes t = n.(p0 - o) / n.d and accepts eps < t < tmax (ky.cpp:1261-1297): ~100 instructions, for every
// rectangle and every ray, although a ray hits about one of them.  Both halves are pure functions, so they may be
// evaluated in any order and only where they can matter:
//   phase 1 (warp-uniform loop, ~40 FMA-pipe instructions per rectangle, no division): an approximate hit point in the
//     rectangle's own (u, v) coordinates CLASSIFIES the ray as certainly outside, certainly inside, or within a band around
//     the edges; likewise its approximate distance against (eps, tmax);
//   phase 2 (per lane, over the lane's own candidates read from shared memory): the reference's exact t for candidates, and
//     the reference's four edge functions only for rays in the band.
// Why the classification is safe.  With p the exact hit point in the plane, E_ij = (d.n) |v_i - v_j| dist(p, line ij), i.e.
// (d.n) * Area * min(u, 1-u, v, 1-v) for the nearest edge of the parallelogram.  The float evaluation of E differs from that
// by at most 8.5 eps M^2 (M = largest |v_i - o|, eps = 2^-24; products, differences and the rounded inputs counted), about
// 2^-21 M^2.  A ray is called inside (outside) only when every edge function is at least 2^-15 M1^2 in magnitude with the
// sign of an interior (for one edge: exterior) point, M1 >= M being the L1 bound computed below: 60x the error bound, so
// the float signs are the exact signs and the reference's test would say the same.  The classifier's own arithmetic
// (approximate reciprocal, FMA, rounded constants) moves u and v by at most 15 eps M |e| / (Area |d.n|) <= 0.12 of that
// margin.  The distance test uses tau = 2^-15 M1 / |d.n| + 2^-18 |t| against an error of 2^-20 M / |d.n| + 2^-21 |t|.
// Anything not certainly outside becomes a candidate, NaNs and infinities included (every comparison with them is false).
// kyd_selftest(KYD_SELFTEST_TRAVERSAL) compares the three queries with the list walk on adversarial rays (through edges and
// corners, grazing, starting on surfaces) for the uploaded scene; the film tests compare whole renders with the oracle.
#ifndef KYD_TWO_PHASE
#define KYD_TWO_PHASE 1
#endif
#define KYD_RECT_SMEM_STRIDE 17   // 15 floats of geometry + surface index, odd stride: lanes on different rectangles hit different banks

#if !KYD_BIG_SCENE
KYD_DEV float* rect_smem()
{
    __shared__ float s[32 * KYD_RECT_SMEM_STRIDE];
    return s;
}

// every thread of the block, once, before the first query: exact rectangle data for the per-lane phase, in mask-bit order
KYD_DEV void stage_rects()
{
#if KYD_TWO_PHASE
    float* s = rect_smem();
    const int n = c_scene.n_rect_cull;
    for (int i = threadIdx.x; i < n * 16; i += blockDim.x)
    {
        const int k = i >> 4, j = i & 15;
        const int sorted = c_scene.rect_order[k];
        const float* shape = reinterpret_cast<const float*>(&c_scene.sorted_shape[sorted]);   // p0, p1, p2, p3, n: 15 floats
        s[k * KYD_RECT_SMEM_STRIDE + j] = j < 15 ? shape[j] : __int_as_float(c_scene.sorted_surface[sorted]);
    }
    __syncthreads();
#endif
}

// per-ray constants of the classification: M1 >= |v - o| for every vertex v of the classified rectangles
struct RayBound { float m1_scaled, m1_sq; };   // 2^-15 M1, M1^2
KYD_DEV RayBound ray_bound(const Ray& r)
{
    const float3 c = c_scene.bound_center;
    const float M1 = (fabsf(r.o.x - c.x) + fabsf(r.o.y - c.y)) + (fabsf(r.o.z - c.z) + c_scene.bound_l1);
    RayBound b;
    b.m1_scaled = 0x1p-15f * M1;
    b.m1_sq = M1 * M1;
    return b;
}

template <int AXIS> KYD_DEV float comp(float3 v) { return AXIS == 0 ? v.x : AXIS == 1 ? v.y : v.z; }

// The decision from the approximate plane distance t, the approximate in-plane coordinates (u, v), |1 / (d.n)| and the
// rectangle's 2^-15 / area: sets bit `bit` of cand (phase 2 needed), inside (edge functions settled) and -- CERTAIN --
// certain (hit inside (eps, t_hi) without phase 2).  NaN / infinite inputs make every comparison false: candidate.
template <bool CERTAIN>
KYD_DEV void rect_decide(float t, float u, float v, float ard, float c_area, const RayBound& rb, float t_hi, unsigned bit,
                         unsigned& cand, unsigned& inside, unsigned& certain)
{
    const float m = (c_area * rb.m1_sq) * ard;
    const float tau = __fmaf_rn(0x1p-18f, fabsf(t), rb.m1_scaled * ard);
    const float lo = fminf(fminf(u, v), fminf(1.f - u, 1.f - v));
    // (fminf drops a NaN operand; a NaN among u, v needs an infinite intermediate, which makes m infinite or NaN as well)
    const bool out = (t < KYD_SHAPE_EPSILON - tau) || (t > t_hi + tau) || (lo < -m);
    const bool in = lo > m;
    if (CERTAIN)
    {
        const bool sure = in && (t > KYD_SHAPE_EPSILON + tau) && (t < t_hi - tau);
        certain |= sure ? bit : 0u;
        cand |= (out || sure) ? 0u : bit;
    }
    else
        cand |= out ? 0u : bit;
    inside |= in ? bit : 0u;
}

template <int AXIS, bool CERTAIN>
KYD_DEV void rects_phase1_aligned(const Ray& r, const RayBound& rb, float t_hi, int begin, int end, unsigned& cand, unsigned& inside, unsigned& certain)
{
    constexpr int B = (AXIS + 1) % 3, C = (AXIS + 2) % 3;
    const float oa = comp<AXIS>(r.o), ob = comp<B>(r.o), oc = comp<C>(r.o);
    const float db = comp<B>(r.d), dc = comp<C>(r.d);
    float rd;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rd) : "f"(comp<AXIS>(r.d)));
    const float ard = fabsf(rd);
    KYD_UNROLL_TRAVERSAL
    for (int k = begin; k < end; ++k)
    {
        const RectAligned& c = c_scene.rect_aligned[k];
        const float t = (c.pa - oa) * rd;
        const float u = c.gb * __fmaf_rn(t, db, ob - c.lo_b);
        const float v = c.gc * __fmaf_rn(t, dc, oc - c.lo_c);
        rect_decide<CERTAIN>(t, u, v, ard, c.c_area, rb, t_hi, 1u << k, cand, inside, certain);
    }
}

template <bool CERTAIN>
KYD_DEV void rects_phase1_general(const Ray& r, const RayBound& rb, float t_hi, int first_bit, int count, unsigned& cand, unsigned& inside, unsigned& certain)
{
    KYD_UNROLL_TRAVERSAL
    for (int j = 0; j < count; ++j)
    {
        const RectCull& c = c_scene.rect_general[j];
        const float obx = c.b.x - r.o.x, oby = c.b.y - r.o.y, obz = c.b.z - r.o.z;
        const float N = __fmaf_rn(c.n.z, obz, __fmaf_rn(c.n.y, oby, c.n.x * obx));
        const float D = __fmaf_rn(c.n.z, r.d.z, __fmaf_rn(c.n.y, r.d.y, c.n.x * r.d.x));
        float rd;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rd) : "f"(D));
        const float t = N * rd;
        const float qx = __fmaf_rn(t, r.d.x, -obx), qy = __fmaf_rn(t, r.d.y, -oby), qz = __fmaf_rn(t, r.d.z, -obz);
        const float u = __fmaf_rn(c.gu.z, qz, __fmaf_rn(c.gu.y, qy, c.gu.x * qx));
        const float v = __fmaf_rn(c.gv.z, qz, __fmaf_rn(c.gv.y, qy, c.gv.x * qx));
        rect_decide<CERTAIN>(t, u, v, fabsf(rd), c.c_area, rb, t_hi, 1u << (first_bit + j), cand, inside, certain);
    }
}

// the reference's edge functions on the staged copy of a rectangle (ky.cpp:1265-1281).  Out of line: only rays in the band
// around an edge get here, and the shade kernels are instruction-fetch bound -- cold code stays out of their hot path.
__device__ __noinline__ bool rect_edges_exact(const float* s, float3 o, float3 d)
{
    const float3 oa = sub(V3(s[0], s[1], s[2]), o), ob = sub(V3(s[3], s[4], s[5]), o);
    const float3 oc = sub(V3(s[6], s[7], s[8]), o), od = sub(V3(s[9], s[10], s[11]), o);
    const float v0d = dot(cross(oc, ob), d);
    const float v1d = dot(cross(ob, oa), d);
    const float v2d = dot(cross(oa, od), d);
    const float v3d = dot(cross(od, oc), d);
    return ((v0d < 0.f) && (v1d < 0.f) && (v2d < 0.f) && (v3d < 0.f)) || ((v0d >= 0.f) && (v1d >= 0.f) && (v2d >= 0.f) && (v3d >= 0.f));
}
KYD_DEV bool rect_edges_exact(const float* s, const Ray& r) { return rect_edges_exact(s, r.o, r.d); }

KYD_DEV float rect_t_exact(const float* s, const Ray& r) // ky.cpp:1283
{
    const float3 n = V3(s[12], s[13], s[14]);
    return dot(n, sub(V3(s[0], s[1], s[2]), r.o)) / dot(n, r.d);
}

// phase 1 over the classified rectangles: bit k of `cand` = rectangle k needs phase 2, of `inside` = its edge functions are
// settled, of `certain` (CERTAIN only) = it is certainly hit inside (eps, t_hi) and needs no phase 2
template <bool CERTAIN>
KYD_DEV void rects_phase1(const Ray& r, float t_hi, unsigned& cand, unsigned& inside, unsigned& certain)
{
    cand = inside = certain = 0u;
    const RayBound rb = ray_bound(r);
    const int e0 = c_scene.rect_aligned_end[0], e1 = c_scene.rect_aligned_end[1], e2 = c_scene.rect_aligned_end[2];
    rects_phase1_aligned<0, CERTAIN>(r, rb, t_hi, 0, e0, cand, inside, certain);
    rects_phase1_aligned<1, CERTAIN>(r, rb, t_hi, e0, e1, cand, inside, certain);
    rects_phase1_aligned<2, CERTAIN>(r, rb, t_hi, e1, e2, cand, inside, certain);
    rects_phase1_general<CERTAIN>(r, rb, t_hi, e2, c_scene.n_rect_general, cand, inside, certain);
}

// Everything the classifiers do not cover -- rectangles beyond the first 32, triangles, disks -- by the list walk's own tests,
// out of line: no reference scene has any (the `shapes` test scene does), so the call is skipped by a uniform branch.
KYD_DEV bool scene_has_rest() { return c_scene.kind_end[3] > c_scene.kind_end[1] || c_scene.kind_end[0] > 32; }

struct ClosestHit { float t; int surface; };
__device__ __noinline__ ClosestHit scene_rest_closest(float3 o, float3 d, float tmax, int best)
{
    Ray r;
    r.o = o; r.d = d; r.tmax = tmax;
    const int end = c_scene.kind_end[0];
    for (int k = 32; k < end; ++k)   // (the classifiers cover the first 32 rectangles of the sorted copy)
    {
        float t;
        if (shape_hit_candidate<KYD_SHAPE_RECTANGLE>(c_scene.sorted_shape[k], r, &t))
        {
            const int surface = c_scene.sorted_surface[k];
            if (t < tmax || (t == tmax && surface < best)) { tmax = t; best = surface; }
        }
    }
    scene_closest_kind<2, KYD_SHAPE_TRIANGLE>(r, tmax, best);
    scene_closest_kind<3, KYD_SHAPE_DISK>(r, tmax, best);
    ClosestHit h;
    h.t = tmax; h.surface = best;
    return h;
}

__device__ __noinline__ bool scene_rest_blocked(float3 o, float3 d, float tmax, int exclude)
{
    Ray r;
    r.o = o; r.d = d; r.tmax = tmax;
    bool found = false;
    const int end = c_scene.kind_end[0];
    for (int k = 32; k < end; ++k)
    {
        float t;
        if (shape_hit_candidate<KYD_SHAPE_RECTANGLE>(c_scene.sorted_shape[k], r, &t))
        {
            const int surface = c_scene.sorted_surface[k];
            if (surface != exclude && (t < r.tmax || (t == r.tmax && surface < exclude))) found = true;
        }
    }
    found = scene_blocked_before_kind_uniform<2, KYD_SHAPE_TRIANGLE>(r, exclude, found);
    found = scene_blocked_before_kind_uniform<3, KYD_SHAPE_DISK>(r, exclude, found);
    return found;
}

// (KYD_WF_NOINLINE=1 keeps ONE copy of each query per kernel behind a call instead of inlining it into every use)
#if defined(KYD_WF_NOINLINE) && KYD_WF_NOINLINE
#define KYD_WF __device__ __noinline__
#else
#define KYD_WF __device__ __forceinline__
#endif

// scene_t::intersect (ky.cpp:3172-3184): closest hit inside (epsilon, r.tmax), lowest surface index among equal distances.
// `best0` generalises the starting state of the walk: with best0 = s >= 0 and r.tmax = the distance at which the ray hits
// surface s, the result differs from s exactly when the list walk prefers another surface to that hit (closer, or as close
// with a lower index) -- the occlusion form of a BSDF-sampled light query; with best0 = -1 and a finite r.tmax the result is
// >= 0 exactly when scene_t::occluded says yes.  One traversal thus serves every query of a path vertex.
KYD_WF int scene_closest_2p(const Ray& r, int best0, float* out_t)
{
    float tmax = r.tmax;
    int best = best0;
    scene_closest_kind<1, KYD_SHAPE_SPHERE>(r, tmax, best);
    unsigned cand, inside, certain;
    rects_phase1<false>(r, tmax, cand, inside, certain);
    if (!(r.tmax > KYD_SHAPE_EPSILON))
        cand = 0u;   // a null ray (idle lane): nothing to resolve
    const float* sm = rect_smem();
    while (cand != 0u)
    {
        const int k = __ffs(cand) - 1;
        cand &= cand - 1u;
        const float* s = sm + k * KYD_RECT_SMEM_STRIDE;
        const float t = rect_t_exact(s, r);
        const int surface = __float_as_int(s[15]);
        if ((t > KYD_SHAPE_EPSILON) && (t < tmax || (t == tmax && surface < best)))
        {
            bool hit = ((inside >> k) & 1u) != 0u;
            if (!hit)
                hit = rect_edges_exact(s, r);
            if (hit)
            {
                tmax = t;
                best = surface;
            }
        }
    }
    if (scene_has_rest())
    {
        const ClosestHit h = scene_rest_closest(r.o, r.d, tmax, best);
        tmax = h.t;
        best = h.surface;
    }
    *out_t = tmax;
    return best;
}

// The two boolean queries of the light loop as one: is there a surface other than `exclude` that the list walk prefers to
// a hit of surface `exclude` at distance r.tmax -- closer, or as close with a lower index?
//   exclude >= 0: occlusion form of the BSDF-sampled query (scene_blocked_before), r.tmax = distance of the light's surface;
//   exclude = -1: scene_t::occluded (ky.cpp:3187-3206), any surface inside (epsilon, r.tmax): no index is lower than -1, so the
//                 tie clause never fires and the test is the reference's strict t < tmax.
// Every lane may call it; a lane without a query passes active = false.
KYD_WF bool scene_blocked_2p(const Ray& r, int exclude, bool active)
{
    bool blocked = scene_blocked_before_kind_uniform<1, KYD_SHAPE_SPHERE>(r, exclude, false);
    unsigned cand, inside, certain;
    rects_phase1<true>(r, r.tmax, cand, inside, certain);
    const float* sm = rect_smem();
    // certainly hit before r.tmax: blocks unless it is the excluded surface itself (whose t equals r.tmax: never "certain", but
    // the check costs nothing)
    for (unsigned c = certain; c != 0u && active && !blocked; c &= c - 1u)
        if (__float_as_int(sm[(__ffs(c) - 1) * KYD_RECT_SMEM_STRIDE + 15]) != exclude)
            blocked = true;
    if (!active || blocked)
        cand = 0u;
    while (cand != 0u)
    {
        const int k = __ffs(cand) - 1;
        cand &= cand - 1u;
        const float* s = sm + k * KYD_RECT_SMEM_STRIDE;
        const int surface = __float_as_int(s[15]);
        if (surface == exclude)
            continue;
        const float t = rect_t_exact(s, r);
        if ((t > KYD_SHAPE_EPSILON) && (t < r.tmax || (t == r.tmax && surface < exclude)) &&
            ((((inside >> k) & 1u) != 0u) || rect_edges_exact(s, r)))
        {
            blocked = true;
            cand = 0u;
        }
    }
    if (scene_has_rest() && active && !blocked)
        blocked = scene_rest_blocked(r.o, r.d, r.tmax, exclude);
    return blocked && active;
}

KYD_DEV bool scene_any_hit_2p(const Ray& r) { return scene_blocked_2p(r, -1, r.tmax > KYD_SHAPE_EPSILON); }
KYD_DEV bool scene_blocked_before_2p(const Ray& r, int light_surface) { return scene_blocked_2p(r, light_surface, light_surface >= 0); }
#else
KYD_DEV void stage_rects() {}
#endif // !KYD_BIG_SCENE

// the queries of the wavefront kernels: two-phase traversal for scenes in constant memory, else the list walk / hierarchy
KYD_DEV int wf_closest(const Ray& r, float* out_t)
{
#if KYD_TWO_PHASE && !KYD_BIG_SCENE
    return scene_closest_2p(r, -1, out_t);
#else
    return scene_closest(r, out_t);
#endif
}
// closest-hit walk from a starting state (scene_closest_2p); the list walk / hierarchy builds answer the two boolean forms
KYD_DEV int wf_closest_from(const Ray& r, int best0, float* out_t)
{
#if KYD_TWO_PHASE && !KYD_BIG_SCENE
    return scene_closest_2p(r, best0, out_t);
#else
    if (best0 >= 0)
        return scene_blocked_before(r, best0) ? -1 : best0;   // (any other surface stands for "blocked")
    return scene_closest(r, out_t);
#endif
}
KYD_DEV bool wf_any_hit(const Ray& r)            // callable by all lanes (r.tmax < 0: no query)
{
#if KYD_TWO_PHASE && !KYD_BIG_SCENE
    return scene_any_hit_2p(r);
#else
    return scene_any_hit_uniform(r);
#endif
}
KYD_DEV bool wf_blocked(const Ray& r, int exclude, bool active)   // both boolean queries as one (scene_blocked_2p)
{
#if KYD_TWO_PHASE && !KYD_BIG_SCENE
    return scene_blocked_2p(r, exclude, active);
#else
    if (!active)
        return false;
    return exclude >= 0 ? scene_blocked_before(r, exclude) : scene_any_hit(r);
#endif
}
KYD_DEV bool wf_blocked_before(const Ray& r, int light_surface)   // callable by all lanes (light_surface < 0: no query)
{
#if KYD_TWO_PHASE && !KYD_BIG_SCENE
    return scene_blocked_before_2p(r, light_surface);
#else
    return scene_blocked_before_uniform(r, light_surface);
#endif
}

KYD_DEV float3 areal_radiance(const DevLight& l, float3 light_normal, float3 wo) // ky.cpp:2957-2960
{
    return (dot(light_normal, wo) > 0) ? l.color : KYD_BLACK;
}

template <bool TABLE = false>
KYD_DEV float3 surface_emission(int surface, const HitGeom& g) // ky.cpp:3084
{
    int li = surface_light_of<TABLE>(surface);
    return li >= 0 ? areal_radiance(c_scene.lights[li], g.normal, g.wo) : KYD_BLACK;
}

KYD_DEV float3 environment_lighting() // ky.cpp:3231-3237
{
    return c_scene.env_light >= 0 ? c_scene.lights[c_scene.env_light].color : KYD_BLACK;
}

// scene_t::occluded(isect, point) ky.cpp:3187-3201: the shadow ray
KYD_DEV Ray shadow_ray(const HitGeom& from, float3 target)
{
    float3 direction = normalize(sub(target, from.position));
    float dist = distance(from.position, target);
    Ray r;
    r.o = offset_ray_origin(from.position, from.normal, direction);
    r.d = direction;
    r.tmax = dist - 2e-3f;
    return r;
}

KYD_DEV Ray spawn_ray(const HitGeom& g, float3 direction) // ky.cpp:665-668
{
    Ray r;
    r.o = offset_ray_origin(g.position, g.normal, direction);
    r.d = direction;
    r.tmax = KYD_INF;
    return r;
}

// ---- lights ky.cpp:2810-3062 ----------------------------------------------------------------------------------
struct LightSample { float3 position, wi; float pdf; float3 Li; };

KYD_DEV bool light_is_delta(int kind) { return kind == KYD_LIGHT_POINT || kind == KYD_LIGHT_DIRECTION; }
KYD_DEV float spherical_theta(float3 v) { return cr_acos(clamp_std(v.z, -1.f, 1.f)); } // ky.cpp:410

template <int TRAITS = TRAITS_ANY>
KYD_DEV LightSample light_sample_Li(int light_index, const HitGeom& g, float2 u)
{
    const DevLight& l = c_scene.lights[light_index];
    LightSample s;
    s.position = V3(0, 0, 0);
    s.wi = V3(0, 0, 0);
    s.pdf = 0;
    s.Li = KYD_BLACK;
    if (light_kind<TRAITS>(l) == KYD_LIGHT_POINT) // ky.cpp:2825-2853
    {
        s.position = l.position;
        s.wi = normalize(sub(l.position, g.position));
        s.pdf = 1.f;
        s.Li = cdiv(l.color, distance_sq(l.position, g.position));
    }
    else if (light_kind<TRAITS>(l) == KYD_LIGHT_DIRECTION) // ky.cpp:2891-2901
    {
        s.wi = neg(l.direction);
        s.position = add(g.position, mul(mul(s.wi, 2.f), l.world_radius));
        s.pdf = 1;
        s.Li = l.color;
    }
    else if (light_kind<TRAITS>(l) == KYD_LIGHT_AREA) // ky.cpp:2964-2981
    {
        float3 lp, ln;
        shape_sample_direction<TRAITS>(c_scene.light_shape[light_index], g.position, g.normal, u, &lp, &ln, &s.pdf);
        s.position = lp;
        if (!(s.pdf == 0 || msq(sub(lp, g.position)) == 0))
        {
            s.wi = normalize(sub(lp, g.position));
            s.Li = areal_radiance(l, ln, neg(s.wi));
        }
    }
    else // environment ky.cpp:3026-3041
    {
        s.wi = uniform_sphere_sample(u);
        s.position = add(g.position, mul(mul(s.wi, 2.f), l.world_radius));
        float sin_theta = cr_sin(spherical_theta(s.wi));
        s.pdf = 1 / (2 * KYD_PI * KYD_PI * sin_theta);
        if (sin_theta == 0)
            s.pdf = 0;
        s.Li = l.color;
    }
    return s;
}

template <int TRAITS = TRAITS_ANY>
KYD_DEV float light_pdf_Li(int light_index, const HitGeom& g, float3 wi)
{
    const DevLight& l = c_scene.lights[light_index];
    if (light_kind<TRAITS>(l) == KYD_LIGHT_AREA) // ky.cpp:2984-2988
        return shape_pdf_direction<TRAITS>(c_scene.light_shape[light_index], g.position, g.normal, wi);
    if (light_kind<TRAITS>(l) == KYD_LIGHT_ENVIRONMENT) // ky.cpp:3043-3053
    {
        float sin_theta = cr_sin(spherical_theta(wi));
        if (sin_theta == 0)
            return 0;
        return 1 / (2 * KYD_PI * KYD_PI * sin_theta);
    }
    return 0;
}

// ---- direct lighting ky.cpp:3834-4088 ---------------------------------------------------------------------------
// Each estimator is split in two halves so that the ray query in the middle can be a separate
// wavefront stage: *_setup() computes the ray and the value the estimator returns IF the ray sees what
// it has to see (everything in the reference after the query is a pure function of data known before
// it, because areal_radiance is either the light's radiance or black); *_resolve() applies the query.

struct NeeRay
{
    Ray ray;        // tmax = inf: closest-hit query (BSDF-sampled); finite: occlusion query (light-sampled)
    float3 value;   // contribution if the query succeeds
    int light;      // closest-hit query: the light the hit surface must carry (or be missed for the environment light)
    int light_surface; // >= 0: the BSDF-sampled query in its occlusion form -- ray.tmax is the distance at which the ray
                    // hits this surface (the only one carrying the light, lit side) and the query succeeds unless another
                    // surface comes first (scene_blocked_before)
    bool active;    // the query can change the result and has to be traced
    bool ref_query; // the reference issues this scene query (it traces before it knows the result is black)
};

// BSDF-sampled half: estimate_direct_lighting_by_bsdf (ky.cpp:3889-3930) when mis == false,
// estimate_direct_lighting_by_bsdf_mis (ky.cpp:3968-4033) when mis == true
// (the BSDF sample itself is an argument so that a caller that also draws the path's continuation can share one copy of
// bsdf_sample's code between the two)
template <int TRAITS = TRAITS_ANY>
KYD_DEV NeeRay nee_bsdf_from_sample(const HitGeom& g, const Bsdf& b, int light_index, const BsdfSample& bs, bool mis)
{
    NeeRay q;
    q.active = false;
    q.ref_query = false;
    q.light = light_index;
    q.light_surface = -1;
    q.value = KYD_BLACK;
    const DevLight& l = c_scene.lights[light_index];
    if (bsdf_is_delta(b.lobe) || light_is_delta(light_kind<TRAITS>(l)))
        return q;
    float3 f_cos = mul(bs.f, abs_dot(bs.wi, g.normal));
    if (is_black(f_cos) || (mis ? (bs.pdf <= 0) : (bs.pdf == 0)))
        return q;
    q.ref_query = true;
    q.ray = spawn_ray(g, bs.wi);
    // Li is the light's radiance when the query succeeds (area: one-sided test in nee_bsdf_resolve)
    float3 Li = l.color;
    if (is_black(Li))
        return q;
    // The reference traces the ray and then asks whether the closest hit carries this light (ky.cpp:3905-3919,
    // 3987-4001).  "The closest hit is surface s" == "s is hit, and no other surface is hit before it": the first half
    // needs one shape test and settles most queries right here (a sphere light's pdf_Li is positive for EVERY direction,
    // ky.cpp:1509-1512, so without this each of them costs a full traversal); the second half is an occlusion query.
    // (upload enables this per light where it pays, kyd_api.cu)
    const int ls = c_scene.light_surface[light_index];
    if (TRAITS == TRAITS_AREA_RECTANGLE)
    {
        // The single-rectangle-light kernels: the host selects them only if exactly one surface carries the light and that
        // surface's shape IS the light's shape (bit for bit).  The ray pdf_Li re-intersects the light's shape with
        // (ky.cpp:1055-1090) is then this very ray against this very rectangle -- offset_ray_origin(p, n, wi) is what spawn_ray
        // computes -- so one hit test serves the occlusion form and the pdf.
        if (ls < 0)
            return q;
        float t_light;
        const DevShape& rect = c_scene.light_shape[light_index];
        if (!shape_hit_distance_kind(rect, KYD_SHAPE_RECTANGLE, q.ray, KYD_INF, &t_light))
            return q;   // pdf_Li = 0, and the closest hit cannot be the light
        const HitGeom lg = shape_hit_geom_kind(rect, KYD_SHAPE_RECTANGLE, q.ray, t_light);
        float light_pdf = distance_sq(g.position, lg.position) / (abs_dot(lg.normal, neg(bs.wi)) * rect.area);
        if (isinf(light_pdf))
            light_pdf = 0.f;
        // (the reference traces before it asks for the pdf: a zero pdf still counts as a query there, ref_query is set above)
        if (!(dot(lg.normal, lg.wo) > 0))
            return q;   // areal_radiance is one-sided (ky.cpp:2957-2960)
        if (!mis)
            q.value = cdiv(cmulc(f_cos, Li), bs.pdf);
        else
        {
            if (!(light_pdf > 0))
                return q;
            q.value = cdiv(mul(cmulc(f_cos, Li), 2.f), bs.pdf + light_pdf);
        }
        q.light_surface = ls;
        q.ray.tmax = t_light;
        q.active = true;
        return q;
    }
    if (light_kind<TRAITS>(l) == KYD_LIGHT_AREA && ls != -2)
    {
        if (ls < 0)
            return q;   // no surface carries the light: whatever the ray hits, it is not this light
        float t_light;
        const DevShape& light_surface_shape = surface_shape(ls);
        // (the sphere-light kernels are selected only if the surfaces that carry the lights are spheres as well)
        const int surface_kind = TRAITS == TRAITS_AREA_SPHERE ? (int)KYD_SHAPE_SPHERE : light_surface_shape.kind;
        if (!shape_hit_distance_kind(light_surface_shape, surface_kind, q.ray, KYD_INF, &t_light))
            return q;
        HitGeom lg = shape_hit_geom_kind(light_surface_shape, surface_kind, q.ray, t_light);
        if (!(dot(lg.normal, lg.wo) > 0))
            return q;   // areal_radiance is one-sided (ky.cpp:2957-2960)
        q.light_surface = ls;
        q.ray.tmax = t_light;
    }
    if (!mis)
        q.value = cdiv(cmulc(f_cos, Li), bs.pdf);
    else
    {
        float light_pdf = light_pdf_Li<TRAITS>(light_index, g, bs.wi);
        if (!(light_pdf > 0))
            return q;
        q.value = cdiv(mul(cmulc(f_cos, Li), 2.f), bs.pdf + light_pdf);
    }
    q.active = true;
    return q;
}

// Conservative early-out of a BSDF-sampled light query against a SPHERE light (multi-light scenes: Veach).  The reference
// samples the BSDF once per light and asks whether that ray's closest hit carries the light (ky.cpp:3968-4001); for a small
// sphere the answer is "no" for all but a sliver of the samples, yet finding the exact direction costs a double-precision
// sincos, IEEE divisions and square roots.  Here the direction is first estimated with the fast FP32 intrinsics (absolute
// error < 1e-5) and tested against the light's surface sphere grown by 1e-3 (1 + distance): a certain miss needs nothing
// else -- the estimator's value is 0, and the reference's own ray is counted (ref_query) because its conditions (f cos > 0,
// pdf > 0) are decided with margins two orders above the estimate's error.  Anything near a decision boundary -- grazing
// directions, a ray that might touch the grown sphere, black albedo, a zero draw -- returns false and takes the exact path.
template <int TRAITS>
KYD_DEV bool bsdf_query_certainly_misses(const HitGeom& g, const Bsdf& b, int light_index, float2 u)
{
    if (TRAITS != TRAITS_AREA_SPHERE)
        return false;
    const int ls = c_scene.light_surface[light_index];
    if (ls < 0 || !(max_component(b.a) > 0.f))
        return false;
    const float3 wo = to_local(b.f, g.wo);
    if (!(fabsf(wo.z) > 1e-4f))
        return false;
    float3 wi;   // estimated sample direction, local frame
    if (b.lobe == LOBE_LAMBERT)
    {
        const float rx = 2.f * u.x - 1.f, ry = 2.f * u.y - 1.f;
        const bool x_major = fabsf(rx) > fabsf(ry);
        const float radius = x_major ? rx : ry;
        const float theta = x_major ? KYD_PI_OVER4 * __fdividef(ry, rx) : KYD_PI_OVER2 - KYD_PI_OVER4 * __fdividef(rx, ry);
        float st, ct;
        __sincosf(theta, &st, &ct);
        const float px = ct * radius, py = st * radius;
        const float z = sqrtf(fmaxf(0.f, 1.f - px * px - py * py));
        if (!(z > 1e-2f) || !(fabsf(radius) > 0.f))
            return false;
        wi = V3(px, py, wo.z < 0.f ? -z : z);
    }
    else if (b.lobe == LOBE_PHONG)
    {
        if (!(u.y > 0.f) || !(b.exponent > 0.f))
            return false;
        const float ct = __powf(u.y, __fdividef(1.f, b.exponent + 1.f));
        const float st = sqrtf(fmaxf(0.f, 1.f - ct * ct));
        float sp, cp;
        __sincosf(2.f * KYD_PI * u.x, &sp, &cp);
        // frame around the mirror direction (frame_t(reflect(wo, z)), ky.cpp:2539), estimated
        const float3 wr = V3(-wo.x, -wo.y, wo.z);
        const float inv = rsqrtf(msq(wr));
        const float3 n = mul(wr, inv);
        const float3 tmp = fabsf(n.x) > 0.99f ? V3(0.f, 1.f, 0.f) : V3(1.f, 0.f, 0.f);
        float3 t = cross(n, tmp);
        t = mul(t, rsqrtf(msq(t)));
        const float3 sx = cross(t, n);
        wi = add(add(mul(sx, cp * st), mul(t, sp * st)), mul(n, ct));
        if (wo.z < 0.f)
            wi.z = -wi.z;
        if (!(wo.z * wi.z > 0.f) || !(fabsf(wi.z) > 1e-2f))
            return false;   // below the surface (f = 0, the reference traces nothing) or too close to call
    }
    else
        return false;
    const float3 w = to_world(b.f, wi);
    const float side = dot(g.normal, w);
    if (!(fabsf(side) > 1e-3f))
        return false;
    const float3 o = add(g.position, mul(g.normal, side > 0.f ? 0.01f : -0.01f));   // offset_ray_origin, ky.cpp:614-620
    const DevShape& sphere = surface_shape(ls);
    const float3 oc = sub(sphere.p0, o);
    const float d2 = msq(oc), bq = dot(oc, w);
    const float grown = sphere.radius + 1e-3f * (1.f + fabsf(oc.x) + fabsf(oc.y) + fabsf(oc.z));
    const float g2 = grown * grown;
    return (d2 - bq * bq > g2) || (bq < 0.f && d2 > g2);
}

template <int TRAITS = TRAITS_ANY, bool CULL = false>
KYD_DEV NeeRay nee_bsdf_setup(const HitGeom& g, const Bsdf& b, int light_index, float2 random_bsdf, bool mis)
{
    const DevLight& l = c_scene.lights[light_index];
    if (CULL && !bsdf_is_delta(b.lobe) && bsdf_query_certainly_misses<TRAITS>(g, b, light_index, random_bsdf))
    {
        NeeRay q;
        q.active = false;
        q.ref_query = true;    // the reference traces this ray and finds that it does not end on the light
        q.light = light_index;
        q.light_surface = -1;
        q.value = KYD_BLACK;
        return q;
    }
    if (bsdf_is_delta(b.lobe) || light_is_delta(light_kind<TRAITS>(l)))
    {
        NeeRay q;
        q.active = false;
        q.ref_query = false;
        q.light = light_index;
        q.light_surface = -1;
        q.value = KYD_BLACK;
        return q;
    }
    return nee_bsdf_from_sample<TRAITS>(g, b, light_index, bsdf_sample(b, g.wo, random_bsdf), mis);
}

// ky.cpp:3905-3919 / 3987-4001: what the BSDF-sampled ray found
KYD_DEV float3 nee_bsdf_resolve(const NeeRay& q, int hit_surface, float hit_t)
{
    const DevLight& l = c_scene.lights[q.light];
    if (hit_surface >= 0)
    {
        if (surface_light(hit_surface) != q.light)
            return KYD_BLACK;
        // emission of the hit: areal_radiance(light_isect, light_isect.wo) with the hit's normal
        HitGeom lg = shape_hit_geom(surface_shape(hit_surface), q.ray, hit_t);
        return (dot(lg.normal, lg.wo) > 0) ? q.value : KYD_BLACK;
    }
    return l.kind == KYD_LIGHT_ENVIRONMENT ? q.value : KYD_BLACK;
}

// traces an active BSDF-sampled query in whichever form nee_bsdf_setup left it
KYD_DEV float3 nee_bsdf_trace(const NeeRay& q)
{
    if (q.light_surface >= 0)
        return scene_blocked_before(q.ray, q.light_surface) ? KYD_BLACK : q.value;
    float t;
    const int s = scene_closest(q.ray, &t);
    return nee_bsdf_resolve(q, s, t);
}

// light-sampled half: estimate_direct_lighting_by_emitter (ky.cpp:3933-3962) when mis == false,
// estimate_direct_lighting_by_emitter_mis (ky.cpp:4035-4074) when mis == true
// Conservative early-out of a light-sampled query of a PHONG lobe against a sphere light seen from outside (Veach: exponent
// 5000).  The estimator multiplies by f = rho pow(wr . wi, n) (ky.cpp:2496-2499; the base is not clamped, so an even n also
// lights the lobe around -wr); every direction into the light's cone makes an angle of at most asin(R / dist) with the axis
// to its centre, hence |wr . wi| <= |wr . axis| + 2 R / dist.  If that bound stays below 2^(-170 / n) the power underflows to
// exactly 0 (cr_pow returns 0 below 2^-160), f cos is black and the query contributes nothing -- without sampling the cone.
// The reference's own shadow ray is still counted (it traces before it evaluates the BSDF, ky.cpp:4047-4052): the sample's
// radiance is the light's (interior of the cone, u.x < 0.99: the sampled point faces the shading point) and its pdf positive.
template <int TRAITS>
KYD_DEV bool light_query_certainly_black(const HitGeom& g, const Bsdf& b, int light_index, float2 u)
{
    if (TRAITS != TRAITS_AREA_SPHERE || b.lobe != LOBE_PHONG)
        return false;
    if (is_black(c_scene.lights[light_index].color) || !(b.exponent >= 1.f) || !(u.x < 0.99f))
        return false;
    const DevShape& sphere = c_scene.light_shape[light_index];
    const float3 pc = sub(sphere.p0, g.position);
    const float d2 = msq(pc);
    if (!(d2 > 1.02f * sphere.radius * sphere.radius) || !(d2 < 1e30f))
        return false;
    const float inv = rsqrtf(d2);
    const float spread = 2.f * sphere.radius * inv;
    const float3 wr = sub(mul(b.f.n, 2.f * dot(b.f.n, g.wo)), g.wo);     // mirror direction of wo about the shading normal
    const float ca = dot(wr, pc) * inv * rsqrtf(fmaxf(msq(wr), 1e-30f));
    const float limit = exp2f(__fdividef(-170.f, b.exponent)) - 0.01f;
    return (ca + spread < limit) && (ca - spread > -limit);
}

template <int TRAITS = TRAITS_ANY, bool CULL = false>
KYD_DEV NeeRay nee_light_setup(const HitGeom& g, const Bsdf& b, int light_index, float2 random_light, bool mis)
{
    NeeRay q;
    q.active = false;
    q.ref_query = false;
    q.light = light_index;
    q.light_surface = -1;
    q.value = KYD_BLACK;
    const DevLight& l = c_scene.lights[light_index];
    if (bsdf_is_delta(b.lobe))
        return q;
    if (CULL && light_query_certainly_black<TRAITS>(g, b, light_index, random_light))
    {
        q.ref_query = true;
        return q;
    }
    LightSample ls = light_sample_Li<TRAITS>(light_index, g, random_light);
    if (is_black(ls.Li) || (mis ? (ls.pdf <= 0) : (ls.pdf == 0)))
        return q;
    q.ref_query = true;
    q.ray = shadow_ray(g, ls.position);
    float3 wo_l = to_local(b.f, g.wo), wi_l = to_local(b.f, ls.wi);
    float3 f_eval;
    float bsdf_pdf_v;
    const bool need_pdf = mis && !light_is_delta(light_kind<TRAITS>(l));
    if (need_pdf)
        bsdf_eval_pdf_local(b, wo_l, wi_l, &f_eval, &bsdf_pdf_v);
    else
    {
        f_eval = bsdf_eval_local(b, wo_l, wi_l);
        bsdf_pdf_v = 0.f;
    }
    float3 f_cos = mul(f_eval, abs_dot(ls.wi, g.normal));
    if (is_black(f_cos))
        return q;
    if (!need_pdf)
        q.value = cdiv(cmulc(f_cos, ls.Li), ls.pdf);
    else
    {
        q.value = cdiv(mul(cmulc(f_cos, ls.Li), 2.f), ls.pdf + bsdf_pdf_v);
    }
    q.active = true;
    return q;
}

// camera_t::generate_ray ky.cpp:1884-1892 (+ origin push of smallpt_rewrite.cpp:676)
KYD_DEV Ray generate_ray(float px, float py)
{
    const DevCamera& cam = c_scene.camera;
    float3 direction = add(add(cam.front, mul(cam.right, px / cam.res_x - 0.5f)), mul(cam.up, 0.5f - py / cam.res_y));
    Ray r;
    r.o = cam.position;
    if (cam.push != 0)
        r.o = add(r.o, mul(direction, cam.push));
    r.d = normalize(direction);
    r.tmax = KYD_INF;
    return r;
}

} // namespace KYD_KERNEL_NS
