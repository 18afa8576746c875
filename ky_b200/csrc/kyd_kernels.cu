// kyd_kernels.cu -- sm_100a kernels of the rendering core: the per-pixel path (all integrators) and
// the clamp kernel.  The wavefront stages live in kyd_wavefront.cu.
//
// Compile with -fmad=false (see kyd_device.cuh).
#include <cstdlib>
#include <cstring>
#include "kyd_internal.h"
#include "kyd_device.cuh"
#include "kyd_wavefront.cuh"

namespace KYD_KERNEL_NS {

void upload_scene_constant(const DevScene& scene, cudaStream_t stream)
{
    cudaMemcpyToSymbolAsync(c_scene, &scene, sizeof(DevScene), 0, cudaMemcpyHostToDevice, stream);
}

// ---- integrators as per-thread loops -----------------------------------------------------------------
struct Counters { unsigned rays, traced; };

struct Isect
{
    HitGeom g;
    Bsdf b;
    float3 emission;
    int surface;
};

// scene_t::intersect + surface_t::intersect ky.cpp:3172-3184, 3077-3088
KYD_DEV bool scene_intersect(const Ray& r, Isect& is, Counters& c)
{
    c.rays++;
    c.traced++;
    float t;
    int s = scene_closest(r, &t);
    if (s < 0)
        return false;
    is.surface = s;
    is.g = shape_hit_geom(surface_shape(s), r, t);
    material_scattering(c_scene.materials[surface_material(s)], is.g, &is.b);
    is.emission = surface_emission(s, is.g);
    return true;
}

// emission seen along a ray, without building a BSDF (ky.cpp:4345-4349, 4358-4372)
KYD_DEV float3 trace_emission(const Ray& r, Counters& c)
{
    c.rays++;
    c.traced++;
    float t;
    int s = scene_closest(r, &t);
    if (s < 0)
        return environment_lighting();
    HitGeom g = shape_hit_geom(surface_shape(s), r, t);
    return surface_emission(s, g);
}

KYD_DEV float3 nee_trace(const NeeRay& q, bool closest, Counters& c)
{
    if (q.ref_query)
        c.rays++;
    if (!q.active)
        return KYD_BLACK;
    c.traced++;
    if (closest)
        return nee_bsdf_trace(q);
    return scene_any_hit(q.ray) ? KYD_BLACK : q.value;
}

// sample_all_light ky.cpp:3834-3872 (g++ draws random_bsdf before random_light)
KYD_DEV float3 sample_all_light(const HitGeom& g, const Bsdf& b, Sampler& smp, int direct_sample, Counters& c)
{
    float3 Ld = KYD_BLACK;
    const int n = c_scene.n_lights;
    for (int l = 0; l < n; ++l)
    {
        float2 random_bsdf = smp.get_float2();
        float2 random_light = smp.get_float2();
        float3 e = KYD_BLACK;
        switch (direct_sample)
        {
        case KYD_DS_BSDF:
            if (!light_is_delta(c_scene.lights[l].kind))
                e = nee_trace(nee_bsdf_setup(g, b, l, smp.get_float2(), false), true, c); // a third pair, ky.cpp:3900
            break;
        case KYD_DS_LIGHT:
            e = nee_trace(nee_light_setup(g, b, l, random_light, false), false, c);
            break;
        case KYD_DS_BSDF_MIS:
            e = nee_trace(nee_bsdf_setup(g, b, l, random_bsdf, true), true, c);
            break;
        case KYD_DS_LIGHT_MIS:
            e = nee_trace(nee_light_setup(g, b, l, random_light, true), false, c);
            break;
        case KYD_DS_BOTH_MIS: // ky.cpp:4076-4088
        {
            float3 Lb = nee_trace(nee_bsdf_setup(g, b, l, random_bsdf, true), true, c);
            float3 Ll = nee_trace(nee_light_setup(g, b, l, random_light, true), false, c);
            e = add(mul(Lb, 0.5f), mul(Ll, 0.5f));
            break;
        }
        default:
            break;
        }
        Ld = add(Ld, e);
    }
    return Ld;
}

KYD_DEV float3 li_debug(const Ray& r, int integrator, Counters& c) // ky.cpp:4105-4122
{
    Isect is;
    if (scene_intersect(r, is, c))
    {
        if (integrator == KYD_INT_POSITION) return normalize(is.g.position);
        if (integrator == KYD_INT_NORMAL) return normalize(is.g.normal);
        if (integrator == KYD_INT_BASECOLOR) return bsdf_eval(is.b, is.g.wo, is.g.normal);
    }
    return KYD_BLACK;
}

KYD_DEV float3 li_direct(const Ray& r, Sampler& smp, int direct_sample, Counters& c) // ky.cpp:4136-4154
{
    Isect is;
    if (!scene_intersect(r, is, c))
        return environment_lighting();
    float3 Lo = is.emission;
    if (!bsdf_is_delta(is.b.lobe))
        Lo = add(Lo, sample_all_light(is.g, is.b, smp, direct_sample, c));
    return Lo;
}

KYD_DEV float3 li_path_iteration(Ray r, Sampler& smp, int max_depth, int direct_sample, Counters& c) // ky.cpp:4529-4617
{
    float3 Lo = KYD_BLACK;
    float3 beta = V3(1, 1, 1);
    bool is_prev_specular = false;

    for (int bounces = 0;; ++bounces)
    {
        Isect is;
        bool hit = scene_intersect(r, is, c);

        if (bounces == 0 || is_prev_specular)
            Lo = add(Lo, cmulc(beta, hit ? is.emission : environment_lighting()));

        if (!hit || bounces >= max_depth)
            break;

        if (!bsdf_is_delta(is.b.lobe))
        {
            float3 Ld = cmulc(beta, sample_all_light(is.g, is.b, smp, direct_sample, c));
            Lo = add(Lo, Ld);
        }

        BsdfSample bs = bsdf_sample(is.b, is.g.wo, smp.get_float2());
        if (is_black(bs.f) || bs.pdf == 0.f)
            break;

        beta = cmulc(beta, cdiv(mul(bs.f, abs_dot(bs.wi, is.g.normal)), bs.pdf));
        is_prev_specular = (bs.type & BSDF_SPECULAR) != 0;
        r = spawn_ray(is.g, bs.wi);

        if (bounces > 3)
        {
            float q = max_std(0.05f, 1 - max_component(beta));
            if (smp.get_float() < q)
                break;
            beta = mul(beta, 1 / (1 - q));
        }
    }
    return Lo;
}

// ---- the three recursive integrators (ky.cpp:4191-4514) as loops with an explicit stack ----------------
// A level returns  Lo_level + ((f * Li_deeper) * |cos|) / pdf  (ky.cpp:4233, 4400, 4512): the forward
// loop records (Lo_level, f, |cos|, pdf) per level and the unwind loop applies the same expression
// from the deepest level outwards, so the FP32 result is the recursion's.
struct Level { float3 Lo, f; float a, p; };

// russian roulette shared by the three (ky.cpp:4219-4226, 4389-4397, 4501-4509)
KYD_DEV bool recursion_roulette(Sampler& smp, int* depth, BsdfSample* bs)
{
    if (++*depth > 3)
    {
        float m = max_component(bs->f);
        if (smp.get_float() < m)
            bs->f = mul(bs->f, 1 / m);
        else
            return false;
    }
    return true;
}

KYD_DEV float3 unwind(const Level* stack, int n, float3 result)
{
    while (n-- > 0)
        result = add(stack[n].Lo, cdiv(mul(cmulc(stack[n].f, result), stack[n].a), stack[n].p));
    return result;
}

KYD_DEV float3 li_simple_recursion(Ray r, Sampler& smp, int max_depth, Counters& c) // ky.cpp:4201-4237
{
    Level stack[KYD_MAX_RECURSION];
    int n = 0, depth = 0;
    float3 result;
    for (;;)
    {
        Isect is;
        if (!scene_intersect(r, is, c)) { result = environment_lighting(); break; }
        if (depth >= max_depth) { result = is.emission; break; }
        BsdfSample bs = bsdf_sample(is.b, is.g.wo, smp.get_float2());
        if (is_black(bs.f) || bs.pdf == 0.f) { result = is.emission; break; }
        if (!recursion_roulette(smp, &depth, &bs)) { result = is.emission; break; }
        stack[n].Lo = is.emission;
        stack[n].f = bs.f;
        stack[n].a = abs_dot(bs.wi, is.g.normal);
        stack[n].p = bs.pdf;
        ++n;
        r.o = is.g.position; // no origin offset, ky.cpp:4232
        r.d = bs.wi;
        r.tmax = KYD_INF;
    }
    return unwind(stack, n, result);
}

// ky.cpp:4321-4401 (deferred == false) and ky.cpp:4440-4513 (deferred == true, with the lighting filter
// of render_lighting_enum as defined in DESIGN.md)
template <bool DEFERRED>
KYD_DEV float3 li_recursion(Ray r, Sampler& smp, int max_depth, int direct_sample, int lighting, Counters& c)
{
    Level stack[KYD_MAX_RECURSION];
    int n = 0, depth = 0;
    bool prev_specular = false;
    float3 result;
    for (;;)
    {
        float3 Lo = KYD_BLACK;
        Isect is;
        bool hit = scene_intersect(r, is, c);

        if (depth == 0 || (DEFERRED && prev_specular))
        {
            float3 Le = hit ? is.emission : environment_lighting();
            bool keep = !DEFERRED || (depth == 0 ? (lighting & KYD_LIGHTING_EMIT) : (lighting & KYD_LIGHTING_INDIRECT));
            Lo = add(Lo, keep ? Le : KYD_BLACK);
        }

        bool recurse = false;
        if (hit && depth < max_depth)
        {
            bool delta = bsdf_is_delta(is.b.lobe);
            if (!delta)
            {
                float3 Ld = sample_all_light(is.g, is.b, smp, direct_sample, c);
                bool keep = !DEFERRED || (depth == 0 ? (lighting & KYD_LIGHTING_DIRECT) : (lighting & KYD_LIGHTING_INDIRECT));
                Lo = add(Lo, keep ? Ld : KYD_BLACK);
            }
            else if (!DEFERRED)
            {
                BsdfSample bs = bsdf_sample(is.b, is.g.wo, smp.get_float2());
                Ray wi_ray;
                wi_ray.o = is.g.position; // no origin offset, ky.cpp:4343
                wi_ray.d = bs.wi;
                wi_ray.tmax = KYD_INF;
                float3 Le = trace_emission(wi_ray, c);
                Lo = add(Lo, cdiv(mul(cmulc(bs.f, Le), abs_dot(bs.wi, is.g.normal)), bs.pdf));
            }

            BsdfSample bs = bsdf_sample(is.b, is.g.wo, smp.get_float2());
            int d = depth;
            if (!(is_black(bs.f) || bs.pdf == 0.f) && recursion_roulette(smp, &d, &bs))
            {
                stack[n].Lo = Lo;
                stack[n].f = bs.f;
                stack[n].a = abs_dot(bs.wi, is.g.normal);
                stack[n].p = bs.pdf;
                ++n;
                if (DEFERRED)
                {
                    r.o = is.g.position; // no origin offset, ky.cpp:4511
                    r.d = bs.wi;
                    r.tmax = KYD_INF;
                }
                else
                    r = spawn_ray(is.g, bs.wi); // ky.cpp:4399
                prev_specular = delta;
                depth = d;
                recurse = true;
            }
            else
                Lo = add(Lo, KYD_BLACK); // Lo += color_t{} (ky.cpp:4352, 4462)
        }
        if (!recurse) { result = Lo; break; }
    }
    return unwind(stack, n, result);
}

// ---- integrator_t::render ky.cpp:3689-3729: one thread per pixel, samples in order --------------------
enum { IC_PATH = 0, IC_DIRECT = 1, IC_DEBUG = 2, IC_RECURSION = 3 };

template <int IC>
__global__ void __launch_bounds__(128) k_render_pixels(RenderParams rp, float* __restrict__ film, DevCounters* __restrict__ counters)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int npix = rp.width * rp.height;
    Counters cnt = { 0u, 0u };
    if (idx < npix)
    {
        const int x = idx % rp.width, y = idx / rp.width;
        float* o = film + 3 * (size_t)idx;
        // continuing a job (KYD_FLAG_ACCUMULATE) continues the pixel's running FP32 sum in sample order
        float3 L = (rp.flags & KYD_FLAG_ACCUMULATE) ? V3(o[0], o[1], o[2]) : KYD_BLACK;
        for (int s = rp.sample_begin; s < rp.sample_end; ++s)
        {
            Sampler smp;
            smp.start(rp.sampler, rp.seed, x, y, s);
            float2 jitter = smp.camera_jitter(rp.sampler, rp.spp, s); // get_camera_sample ky.cpp:3714, 971-974
            Ray r = generate_ray((float)x + jitter.x, (float)y + jitter.y);
            float3 Li;
            if (IC == IC_PATH)
                Li = li_path_iteration(r, smp, rp.max_depth, rp.direct_sample, cnt);
            else if (IC == IC_DIRECT)
                Li = li_direct(r, smp, rp.direct_sample, cnt);
            else if (IC == IC_DEBUG)
                Li = li_debug(r, rp.integrator, cnt);
            else
            {
                if (rp.integrator == KYD_INT_SIMPLE_PT_RECURSION)
                    Li = li_simple_recursion(r, smp, rp.max_depth, cnt);
                else if (rp.integrator == KYD_INT_PT_RECURSION)
                    Li = li_recursion<false>(r, smp, rp.max_depth, rp.direct_sample, rp.lighting, cnt);
                else
                    Li = li_recursion<true>(r, smp, rp.max_depth, rp.direct_sample, rp.lighting, cnt);
            }
            L = add(L, mul(Li, rp.weight)); // L = L + Li * (1. / spp), ky.cpp:3717-3721
        }
        if (rp.flags & KYD_FLAG_CLAMP)
            L = V3(clamp_std(L.x, 0.f, 1.f), clamp_std(L.y, 0.f, 1.f), clamp_std(L.z, 0.f, 1.f));
        o[0] = film_value(L.x); o[1] = film_value(L.y); o[2] = film_value(L.z);
    }
    unsigned rays = __reduce_add_sync(0xffffffffu, cnt.rays);
    unsigned traced = __reduce_add_sync(0xffffffffu, cnt.traced);
    if ((threadIdx.x & 31) == 0 && (rays | traced))
    {
        atomicAdd(&counters->rays, (unsigned long long)rays);
        atomicAdd(&counters->rays_traced, (unsigned long long)traced);
    }
}

void launch_render_pixels(const RenderParams& rp, float* film_dev, DevCounters* counters, cudaStream_t stream)
{
    const int npix = rp.width * rp.height;
    const int block = 128;
    const int grid = (npix + block - 1) / block;
    switch (rp.integrator)
    {
    case KYD_INT_PT_ITERATION:
        k_render_pixels<IC_PATH><<<grid, block, 0, stream>>>(rp, film_dev, counters);
        break;
    case KYD_INT_DIRECT_LIGHTING:
        k_render_pixels<IC_DIRECT><<<grid, block, 0, stream>>>(rp, film_dev, counters);
        break;
    case KYD_INT_POSITION:
    case KYD_INT_NORMAL:
    case KYD_INT_BASECOLOR:
        k_render_pixels<IC_DEBUG><<<grid, block, 0, stream>>>(rp, film_dev, counters);
        break;
    default:
        k_render_pixels<IC_RECURSION><<<grid, block, 0, stream>>>(rp, film_dev, counters);
        break;
    }
}

#if !KYD_BIG_SCENE // host-side pieces shared by both builds of the kernels live in the small-scene build only
__global__ void k_clamp(float* __restrict__ film, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        film[i] = film_value(clamp_std(film[i], 0.f, 1.f));
}

void launch_clamp(float* film_dev, int64_t n, cudaStream_t stream)
{
    const int block = 256;
    k_clamp<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(film_dev, n);
}

// ---- multi-GPU: the partial films of the other ranks added in rank order, clamp after the sum (ky.cpp:3721-3726) ----
// One pass over rank 0's film; the peers' films are loaded in place through their peer mappings (16-byte loads: the
// NVLink transfer IS the kernel's load stream, nothing is staged), so the reduce and the clamp cost one kernel.
struct PartTable { const float* p[KYD_MAX_MULTI]; };

__global__ void __launch_bounds__(256) k_sum_partials(float* __restrict__ film, PartTable parts, int n_parts, int64_t n, int64_t n4, int clamp)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;   // n4: float4 units handled by the vector loop (0 for unaligned films)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    {
        float4 v = reinterpret_cast<float4*>(film)[i];
        for (int r = 0; r < n_parts; ++r)
        {
            const float4 a = __ldg(reinterpret_cast<const float4*>(parts.p[r]) + i);
            v.x = v.x + a.x; v.y = v.y + a.y; v.z = v.z + a.z; v.w = v.w + a.w;
        }
        if (clamp)
        {
            v.x = clamp_std(v.x, 0.f, 1.f); v.y = clamp_std(v.y, 0.f, 1.f); v.z = clamp_std(v.z, 0.f, 1.f); v.w = clamp_std(v.w, 0.f, 1.f);
        }
        reinterpret_cast<float4*>(film)[i] = make_float4(film_value(v.x), film_value(v.y), film_value(v.z), film_value(v.w));
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        float v = film[i];
        for (int r = 0; r < n_parts; ++r)
            v = v + parts.p[r][i];
        if (clamp)
            v = clamp_std(v, 0.f, 1.f);
        film[i] = film_value(v);
    }
}

void launch_sum_partials(float* film_dev, const float* const* parts, int n_parts, int64_t n, bool clamp, int sm_count, cudaStream_t stream)
{
    if (n_parts == 0 && !clamp)
        return;
    PartTable t{};
    bool aligned = ((uintptr_t)film_dev & 15u) == 0;
    for (int r = 0; r < n_parts; ++r)
    {
        t.p[r] = parts[r];
        aligned = aligned && ((uintptr_t)parts[r] & 15u) == 0;
    }
    // (cudaMalloc'ed films are 256-byte aligned; a caller's odd pointer falls back to the scalar loop)
    k_sum_partials<<<sm_count * 8, 256, 0, stream>>>(film_dev, t, n_parts, n, aligned ? n >> 2 : 0, clamp ? 1 : 0);
}

// ---- self-tests of the exactness-critical fast paths ---------------------------------------------------
// out[0] = bit patterns whose fast result differs from the definition, out[1] = patterns that took the slow path
__global__ void k_selftest_rsqrt(unsigned long long first, unsigned long long count, unsigned long long* __restrict__ out)
{
    unsigned long long bad = 0, slow = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        const float s = __uint_as_float((unsigned)(first + i));
        const float want = rsqrt_ky_reference(s);
        const float got = rsqrt_ky(s);
        if (__float_as_uint(want) != __float_as_uint(got) && !(want != want && got != got))
            ++bad;
        // recompute the acceptance test's outcome: the slow path is taken exactly when the seed-based value is rejected
        float y0;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(s));
        const float t = __fmul_rn(s, y0), tl = __fmaf_rn(s, y0, -t);
        const float r2 = __fmaf_rn(-tl, y0, __fmaf_rn(-t, y0, 1.0f));
        const float h = __fmul_rn(0.5f, r2), yh = __fmaf_rn(y0, h, y0);
        const float rho = __fmaf_rn(y0, h, __fsub_rn(y0, yh));
        if (!rsqrt_ky_accept(s, yh, rho))
            ++slow;
    }
    if (bad) atomicAdd(&out[0], bad);
    if (slow) atomicAdd(&out[1], slow);
}

// pow: (x, y) pairs derived from the index: the exponents Phong shading uses (n and 1/(n+1) for n = 30, 90, 5000),
// random exponents, bases over (0, 4) with extra density next to 1
__global__ void k_selftest_pow(unsigned long long first, unsigned long long count, unsigned long long* __restrict__ out)
{
    unsigned long long bad = 0, slow = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        const unsigned long long h = mix64(first + i), h2 = mix64(h ^ 0x9E3779B97F4A7C15ull);
        const float u = (float)(unsigned)(h >> 40) * 0x1p-24f, v = (float)(unsigned)(h2 >> 40) * 0x1p-24f;
        const int k = (int)(h & 15);
        float x, y;
        const float ys[6] = { 90.f, 5000.f, 30.f, 1.f / 91.f, 1.f / 5001.f, 1.f / 31.f };
        if (k < 6) { y = ys[k]; x = k < 3 ? 1.f - u * v * 0.05f : u; }
        else if (k < 12) { y = ys[k - 6]; x = k < 9 ? 1.f - u * u * u : u * v; }
        else if (k < 14) { y = (v * 2.f - 1.f) * 100.f; x = u * 4.f; }
        else { y = ys[(h >> 8) % 3]; x = k == 14 ? -(u * v) : -(1.f - u * v * 0.05f); }   // negative base, integer exponent
        const float want = cr_pow_reference(x, y);
        const float got = cr_pow(x, y);
        if (__float_as_uint(want) != __float_as_uint(got) && !(want != want && got != got))
            ++bad;
        // a second evaluation of the fast path's conditions to count how often the definition was needed
        bool fast = false;
        const float ax = (x < 0.f && y == truncf(y)) ? -x : x;
        if (ax > 0x1p-100f && ax < 0x1p100f && fabsf(y) < 1e6f)
        {
            const double t = (double)y * log2((double)ax);
            fast = t < -160.0 || fabs(t) < 120.0;
        }
        if (!fast) ++slow;
    }
    if (bad) atomicAdd(&out[0], bad);
    if (slow) atomicAdd(&out[1], slow);
}

// ---- known-answer harness: one device function of the path per launch, inputs and outputs in the layouts of the oracle's
// kyo_* test entry points (oracle/kyo.c), so that tests/golden/golden_kat.npz -- generated from the reference build -- pins
// the DEVICE functions one by one, including branches no film reaches (a shading point inside a sphere light, total internal
// reflection, the disk's parallel-ray reject, the inf -> 0 pdf guards)
template <int TRAITS>
KYD_DEV void kat_light(int light, const float* a, float* o)
{
    HitGeom g;
    g.position = V3(a[0], a[1], a[2]);
    g.normal = V3(a[3], a[4], a[5]);
    g.wo = V3(0.f, 0.f, 1.f);
    const LightSample ls = light_sample_Li<TRAITS>(light, g, make_float2(a[6], a[7]));
    o[0] = ls.position.x; o[1] = ls.position.y; o[2] = ls.position.z;
    o[3] = ls.wi.x; o[4] = ls.wi.y; o[5] = ls.wi.z;
    o[6] = ls.pdf; o[7] = ls.Li.x; o[8] = ls.Li.y; o[9] = ls.Li.z;
    o[10] = light_pdf_Li<TRAITS>(light, g, V3(a[8], a[9], a[10]));
}

__global__ void k_kat(int which, DevShape shape, DevMaterial material, int index, int traits, int n, const float* __restrict__ in, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    switch (which)
    {
    case KYD_KAT_SHAPE_INTERSECT: // shape_t::intersect ky.cpp:1111-1393; in {o, d, tmax}, out {hit, tmax, position, normal}
    {
        const float* a = in + 7 * i;
        Ray r;
        r.o = V3(a[0], a[1], a[2]); r.d = V3(a[3], a[4], a[5]); r.tmax = a[6];
        float t;
        const bool hit = shape_hit_distance(shape, r, r.tmax, &t);
        float* o = out + 8 * i;
        o[0] = hit ? 1.f : 0.f;
        o[1] = hit ? t : r.tmax;
        HitGeom g;
        g.position = g.normal = V3(0.f, 0.f, 0.f);
        if (hit)
            g = shape_hit_geom(shape, r, t);
        o[2] = g.position.x; o[3] = g.position.y; o[4] = g.position.z;
        o[5] = g.normal.x; o[6] = g.normal.y; o[7] = g.normal.z;
        break;
    }
    case KYD_KAT_SHAPE_SAMPLE_DIRECTION: // ky.cpp:1028-1051, 1419-1501; in {p, n, u}, out {lp, ln, pdf}
    {
        const float* a = in + 8 * i;
        float3 lp = V3(0.f, 0.f, 0.f), ln = lp;
        float pdf = 0.f;
        const float3 p = V3(a[0], a[1], a[2]), ns = V3(a[3], a[4], a[5]);
        const float2 u = make_float2(a[6], a[7]);
        if (traits == TRAITS_AREA_RECTANGLE) shape_sample_direction<TRAITS_AREA_RECTANGLE>(shape, p, ns, u, &lp, &ln, &pdf);
        else if (traits == TRAITS_AREA_SPHERE) shape_sample_direction<TRAITS_AREA_SPHERE>(shape, p, ns, u, &lp, &ln, &pdf);
        else shape_sample_direction<TRAITS_ANY>(shape, p, ns, u, &lp, &ln, &pdf);
        float* o = out + 7 * i;
        o[0] = lp.x; o[1] = lp.y; o[2] = lp.z; o[3] = ln.x; o[4] = ln.y; o[5] = ln.z; o[6] = pdf;
        break;
    }
    case KYD_KAT_SHAPE_PDF_DIRECTION: // ky.cpp:1055-1090, 1503-1513; in {p, n, wi}, out pdf
    {
        const float* a = in + 9 * i;
        const float3 p = V3(a[0], a[1], a[2]), ns = V3(a[3], a[4], a[5]), wi = V3(a[6], a[7], a[8]);
        out[i] = traits == TRAITS_AREA_RECTANGLE ? shape_pdf_direction<TRAITS_AREA_RECTANGLE>(shape, p, ns, wi)
               : traits == TRAITS_AREA_SPHERE ? shape_pdf_direction<TRAITS_AREA_SPHERE>(shape, p, ns, wi)
               : shape_pdf_direction<TRAITS_ANY>(shape, p, ns, wi);
        break;
    }
    case KYD_KAT_MATERIAL_BSDF: // material_t::scattering + bsdf_t::sample / eval / pdf, ky.cpp:2147-2682
    {                           // in {p, n, wo, wi, u}, out {sample.f, sample.wi, sample.pdf, sample.type, eval, pdf, is_delta}
        const float* a = in + 14 * i;
        HitGeom g;
        g.position = V3(a[0], a[1], a[2]); g.normal = V3(a[3], a[4], a[5]); g.wo = V3(a[6], a[7], a[8]);
        Bsdf b;
        material_scattering(material, g, &b);
        const BsdfSample bs = bsdf_sample(b, g.wo, make_float2(a[12], a[13]));
        const float3 wi = V3(a[9], a[10], a[11]);
        const float3 f = bsdf_eval(b, g.wo, wi);
        const float pdf = bsdf_pdf(b, g.wo, wi);
        float* o = out + 13 * i;
        o[0] = bs.f.x; o[1] = bs.f.y; o[2] = bs.f.z; o[3] = bs.wi.x; o[4] = bs.wi.y; o[5] = bs.wi.z;
        o[6] = bs.pdf; o[7] = (float)bs.type; o[8] = f.x; o[9] = f.y; o[10] = f.z; o[11] = pdf;
        o[12] = bsdf_is_delta(b.lobe) ? 1.f : 0.f;
        break;
    }
    case KYD_KAT_CAMERA_RAYS: // camera_t::generate_ray ky.cpp:1884-1892 of the uploaded scene; in {px, py}, out {o, d}
    {
        const Ray r = generate_ray(in[2 * i], in[2 * i + 1]);
        float* o = out + 6 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
        break;
    }
    case KYD_KAT_LIGHT_SAMPLE: // light_t::sample_Li / pdf_Li of light `index` of the uploaded scene, ky.cpp:2810-3062
    {                          // in {p, n, u, wi}, out {position, wi, pdf, Li, pdf_Li(wi)}
        if (traits == TRAITS_AREA_RECTANGLE) kat_light<TRAITS_AREA_RECTANGLE>(index, in + 11 * i, out + 11 * i);
        else if (traits == TRAITS_AREA_SPHERE) kat_light<TRAITS_AREA_SPHERE>(index, in + 11 * i, out + 11 * i);
        else kat_light<TRAITS_ANY>(index, in + 11 * i, out + 11 * i);
        break;
    }
    case KYD_KAT_SAMPLER: // lcg48 stream of (seed = material.kind bits.., x, y, sample): in {x, y, sample, count<=16}, out 16 floats
    {
        const float* a = in + 4 * i;
        Sampler smp;
        smp.start(KYD_SAMPLER_LCG48, (unsigned long long)(unsigned)index, (int)a[0], (int)a[1], (int)a[2]);
        for (int k = 0; k < 16; ++k)
            out[16 * i + k] = smp.get_float();
        break;
    }
    default:
        break;
    }
}

void launch_kat(int which, const DevShape& shape, const DevMaterial& material, int index, int traits, int n, const float* in_dev, float* out_dev, cudaStream_t stream)
{
    k_kat<<<(n + 127) / 128, 128, 0, stream>>>(which, shape, material, index, traits, n, in_dev, out_dev);
}

// two-phase traversal against the list walk on the uploaded scene: adversarial rays -- between points of two surfaces, the
// points snapped to within 2^-5 .. 2^-30 of edges and corners half of the time, origins on or just off their surface, random
// and grazing directions -- and for each the three queries with limits that sit exactly on hit distances.
// out[0] = queries whose answers differ, out[1] = rays whose closest-hit query hit something (the test is not vacuous)
KYD_DEV float3 selftest_surface_point(int surface, unsigned long long h, float3* normal)
{
    const DevShape& sh = c_scene.surf_shape[surface];
    float u = (float)(unsigned)(h >> 40) * 0x1p-24f, v = (float)(unsigned)((h >> 16) & 0xffffffu) * 0x1p-24f;
    const unsigned snap = (unsigned)h & 15u;
    if (snap & 1u) { const float e = ldexpf((snap & 4u) ? 1.f : -1.f, -5 - (int)((h >> 8) % 26u)); u = ((h >> 4) & 1u) ? 1.f + e : e; }
    if (snap & 2u) { const float e = ldexpf((snap & 8u) ? 1.f : -1.f, -5 - (int)((h >> 12) % 26u)); v = ((h >> 5) & 1u) ? 1.f + e : e; }
    *normal = sh.n;
    if (sh.kind == KYD_SHAPE_SPHERE)
    {
        const float3 dir = uniform_sphere_sample(make_float2(fminf(fmaxf(u, 0.f), 1.f), fminf(fmaxf(v, 0.f), 1.f)));
        *normal = dir;
        return add(sh.p0, mul(dir, sh.radius));
    }
    if (sh.kind == KYD_SHAPE_DISK)
    {
        const Frame f = frame_from_z(sh.n);
        return add(sh.p0, mul(add(mul(f.s, 2.f * u - 1.f), mul(f.t, 2.f * v - 1.f)), sh.radius * 0.75f));
    }
    // rectangle: the parallelogram p1 + u (p0 - p1) + v (p2 - p1); triangle: the same, folded
    if (sh.kind == KYD_SHAPE_TRIANGLE && u + v > 1.f) { u = 1.f - u; v = 1.f - v; }
    return add(add(sh.p1, mul(sub(sh.p0, sh.p1), u)), mul(sub(sh.p2, sh.p1), v));
}

__global__ void __launch_bounds__(256) k_selftest_traversal(unsigned long long first, unsigned long long count, unsigned long long* __restrict__ out)
{
    stage_rects();
    unsigned long long bad = 0, hits = 0;
    const int n = c_scene.n_surfaces;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count && n > 0; i += stride)
    {
        const unsigned long long h0 = mix64(first + i), h1 = mix64(h0 ^ 0x9E3779B97F4A7C15ull), h2 = mix64(h1 + 0x632BE59BD9B4E019ull);
        const int s1 = (int)((h0 >> 8) % (unsigned)n), s2 = (int)((h0 >> 32) % (unsigned)n);
        float3 n1, n2;
        const float3 a = selftest_surface_point(s1, h1, &n1);
        const float3 b = selftest_surface_point(s2, h2, &n2);
        Ray r;
        const unsigned mode = (unsigned)h0 & 7u;
        float3 dir = sub(b, a);
        if (mode == 6u)
            dir = uniform_sphere_sample(make_float2((float)(unsigned)(h2 >> 40) * 0x1p-24f, (float)(unsigned)(h1 >> 40) * 0x1p-24f));
        if (mode == 7u)   // grazing: almost inside the plane of the first surface
            dir = add(cross(n1, sub(b, a)), mul(n1, ldexpf(1.f, -3 - (int)((h0 >> 56) % 24u))));
        if (!(msq(dir) > 0.f))
            dir = V3(0.f, 0.f, 1.f);
        r.d = normalize(dir);
        r.o = (mode & 1u) ? offset_ray_origin(a, normalize(n1), r.d) : a;
        r.tmax = KYD_INF;

        float t0, t1;
        const int c0 = scene_closest(r, &t0), c1 = scene_closest_2p(r, -1, &t1);
        if (c0 != c1 || (c0 >= 0 && __float_as_uint(t0) != __float_as_uint(t1)))
            ++bad;
        if (c0 >= 0)
            ++hits;
        // occlusion queries: the shadow-ray limit, and limits on and next to the closest hit's distance
        const float limits[4] = { distance(r.o, b) - 2e-3f, t0, c0 >= 0 ? __uint_as_float(__float_as_uint(t0) + 1u) : 1.f,
                                  c0 >= 0 ? t0 * ((float)(unsigned)(h2 & 0xffffu) * 0x1p-16f) : 0.5f };
        for (int k = 0; k < 4; ++k)
        {
            Ray q = r;
            q.tmax = limits[k];
            if (scene_any_hit(q) != scene_any_hit_2p(q))
                ++bad;
            float tq;
            if (q.tmax > KYD_SHAPE_EPSILON && scene_any_hit(q) != (scene_closest_2p(q, -1, &tq) >= 0))   // occlusion as a closest-hit walk
                ++bad;
        }
        // occlusion form of a BSDF-sampled query: towards the closest hit's surface, and towards the second surface if the ray hits it
        if (c0 >= 0)
        {
            Ray q = r;
            q.tmax = t0;
            if (scene_blocked_before(q, c0) != scene_blocked_before_2p(q, c0))
                ++bad;
            float tq;
            if (scene_blocked_before(q, c0) != (scene_closest_2p(q, c0, &tq) != c0))   // ... and from the light surface's own hit
                ++bad;
        }
        float t2;
        if (shape_hit_distance(c_scene.surf_shape[s2], r, KYD_INF, &t2))
        {
            Ray q = r;
            q.tmax = t2;
            if (scene_blocked_before(q, s2) != scene_blocked_before_2p(q, s2))
                ++bad;
            float tq;
            if (scene_blocked_before(q, s2) != (scene_closest_2p(q, s2, &tq) != s2))
                ++bad;
        }
    }
    if (bad) atomicAdd(&out[0], bad);
    if (hits) atomicAdd(&out[1], hits);
}

void launch_selftest_traversal(unsigned long long first, unsigned long long count, unsigned long long* out_dev, cudaStream_t stream)
{
    k_selftest_traversal<<<148 * 8, 256, 0, stream>>>(first, count, out_dev);
}

void launch_selftest_pow(unsigned long long first, unsigned long long count, unsigned long long* out_dev, cudaStream_t stream)
{
    k_selftest_pow<<<148 * 16, 256, 0, stream>>>(first, count, out_dev);
}

void launch_selftest_rsqrt(unsigned long long first, unsigned long long count, unsigned long long* out_dev, cudaStream_t stream)
{
    k_selftest_rsqrt<<<148 * 16, 256, 0, stream>>>(first, count, out_dev);
}

// ---- wavefront orchestration ---------------------------------------------------------------------------
void StageTimer::begin(int stage)
{
    if (used + 2 > events.size())
    {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        events.push_back(a);
        events.push_back(b);
    }
    stages.resize(events.size() / 2);
    stages[used / 2] = stage;
    cudaEventRecord(events[used], stream);
}
void StageTimer::end()
{
    cudaEventRecord(events[used + 1], stream);
    used += 2;
}
void StageTimer::collect(double* stage_ms)
{
    if (used == 0) return;
    cudaEventSynchronize(events[used - 1]);
    for (size_t k = 0; k < used; k += 2)
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, events[k], events[k + 1]);
        stage_ms[stages[k / 2]] += ms;
    }
    used = 0;
}
StageTimer::~StageTimer()
{
    for (cudaEvent_t e : events) cudaEventDestroy(e);
}

template <class T>
static cudaError_t alloc_array(T** p, size_t count)
{
    return cudaMalloc((void**)p, count * sizeof(T));
}

void free_wave_buffers(WaveBuffers& w)
{
    void* ptrs[] = { w.path, w.vertex, w.nee, w.levels, w.queue_a, w.queue_b, w.queue_lobe[0][0], w.queue_lobe[0][1], w.queue_lobe[0][2], w.queue_lobe[0][3],
                     w.queue_lobe[1][0], w.queue_lobe[1][1], w.queue_lobe[1][2], w.queue_lobe[1][3],
                     w.queue_nee[0], w.queue_nee[1], w.queue_pair[0], w.queue_pair[1] };
    for (void* p : ptrs)
        if (p) cudaFree(p);
    w = WaveBuffers{};
}

int ensure_wave_buffers(WaveBuffers& w, int64_t capacity, int nee_lights, int nee_units, bool vertex, int levels)
{
    if (w.capacity >= capacity && w.max_lights >= nee_lights && (nee_lights == 0 || w.nee_units >= nee_units) && (w.has_vertex || !vertex) &&
        w.max_levels >= levels)
        return cudaSuccess;
    if (capacity < w.capacity) capacity = w.capacity;
    if (levels < w.max_levels) levels = w.max_levels;
    if (nee_lights < w.max_lights) nee_lights = w.max_lights;
    if (nee_units < w.nee_units) nee_units = w.nee_units;
    vertex = vertex || w.has_vertex;
    free_wave_buffers(w);
    const size_t P = (size_t)capacity, L = (size_t)nee_lights;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    ok(alloc_array(&w.path, 4 * P));
    if (vertex) ok(alloc_array(&w.vertex, 6 * P));
    if (L > 0) ok(alloc_array(&w.nee, (size_t)nee_units * L * P));
    if (levels > 0) ok(alloc_array(&w.levels, 2 * (size_t)levels * P));
    ok(alloc_array(&w.queue_a, P)); ok(alloc_array(&w.queue_b, P));
    for (int c = 0; c < 8; ++c) ok(alloc_array(&w.queue_lobe[c >> 2][c & 3], P));
    for (int c = 0; c < 2; ++c) ok(alloc_array(&w.queue_nee[c], P));
    if (L > 0 && nee_units > 1)
        for (int c = 0; c < 2; ++c) ok(alloc_array(&w.queue_pair[c], L * P));
    if (e != cudaSuccess)
    {
        free_wave_buffers(w);
        return e;
    }
    w.capacity = capacity;
    w.max_lights = nee_lights;
    w.nee_units = L > 0 ? nee_units : 0;
    w.has_vertex = vertex;
    w.max_levels = levels;
    return cudaSuccess;
}

#endif // !KYD_BIG_SCENE

// the mirror and glass kernels sample no lights: one instantiation serves both light counts
template <int TRAITS, bool HOT, int NL, bool FUSE>
static void launch_shade_lobes(int grid, cudaStream_t stream, const WaveParams& wp, const WaveBuffers& w, DevCounters* counters, int bounce)
{
    constexpr int NL_SPECULAR = HOT ? NL_ONE : NL_ANY;
    k_shade<LOBE_LAMBERT, TRAITS, HOT, NL, FUSE><<<grid, SHADE_THREADS, 0, stream>>>(wp, w, counters, bounce);
    k_shade<LOBE_PHONG, TRAITS, HOT, NL, FUSE><<<grid, SHADE_THREADS, 0, stream>>>(wp, w, counters, bounce);
    k_shade<LOBE_MIRROR, TRAITS, HOT, NL_SPECULAR, FUSE><<<grid, SHADE_THREADS, 0, stream>>>(wp, w, counters, bounce);
    k_shade<LOBE_FRESNEL, TRAITS, HOT, NL_SPECULAR, FUSE><<<grid, SHADE_THREADS, 0, stream>>>(wp, w, counters, bounce);
}

template <int TRAITS>
static void launch_shade_one_light(bool fused, int grid, cudaStream_t stream, const WaveParams& wp, const WaveBuffers& w, DevCounters* counters, int bounce)
{
    if (fused) launch_shade_lobes<TRAITS, true, NL_ONE, true>(grid, stream, wp, w, counters, bounce);
    else launch_shade_lobes<TRAITS, true, NL_ONE, false>(grid, stream, wp, w, counters, bounce);
}

template <int INTEGRATOR>
static void launch_shade_rec(int grid, cudaStream_t stream, const WaveParams& wp, const WaveBuffers& w, DevCounters* counters, int bounce)
{
    k_shade_rec<LOBE_LAMBERT, INTEGRATOR><<<grid, SHADE_THREADS, 0, stream>>>(wp, w, counters, bounce);
    k_shade_rec<LOBE_PHONG, INTEGRATOR><<<grid, SHADE_THREADS, 0, stream>>>(wp, w, counters, bounce);
    k_shade_rec<LOBE_MIRROR, INTEGRATOR><<<grid, SHADE_THREADS, 0, stream>>>(wp, w, counters, bounce);
    k_shade_rec<LOBE_FRESNEL, INTEGRATOR><<<grid, SHADE_THREADS, 0, stream>>>(wp, w, counters, bounce);
}

static void launch_shade(int traits, bool hot, bool one_light, bool fused, int grid, cudaStream_t stream, const WaveParams& wp, const WaveBuffers& w,
                         DevCounters* counters, int bounce)
{
    if (!hot) launch_shade_lobes<TRAITS_ANY, false, NL_ANY, false>(grid, stream, wp, w, counters, bounce);
#if !KYD_BIG_SCENE
    else if (traits == TRAITS_AREA_RECTANGLE) launch_shade_one_light<TRAITS_AREA_RECTANGLE>(fused, grid, stream, wp, w, counters, bounce);
    else if (traits == TRAITS_AREA_SPHERE)
    {
        if (one_light) launch_shade_one_light<TRAITS_AREA_SPHERE>(fused, grid, stream, wp, w, counters, bounce);
        else if (fused) launch_shade_lobes<TRAITS_AREA_SPHERE, true, NL_MANY, true>(grid, stream, wp, w, counters, bounce);
        else launch_shade_lobes<TRAITS_AREA_SPHERE, true, NL_MANY, false>(grid, stream, wp, w, counters, bounce);
    }
    else
    {
        if (one_light) launch_shade_one_light<TRAITS_ANY>(fused, grid, stream, wp, w, counters, bounce);
        else if (fused) launch_shade_lobes<TRAITS_ANY, true, NL_MANY, true>(grid, stream, wp, w, counters, bounce);
        else launch_shade_lobes<TRAITS_ANY, true, NL_MANY, false>(grid, stream, wp, w, counters, bounce);
    }
#endif
}

void launch_render_wavefront(const RenderParams& rp, const DevScene& scene, WaveBuffers& w, int64_t wave_paths, float* film_dev, DevCounters* counters,
                             cudaStream_t stream, int sm_count, uint64_t* launches, StageTimer* timer)
{
    auto T = [&](int stage) { if (timer) { if (stage >= 0) timer->begin(stage); else timer->end(); } };
    const long long npix_total = (long long)rp.width * rp.height;
    const int nsamples = rp.sample_end - rp.sample_begin;
    const bool direct_only = rp.integrator == KYD_INT_DIRECT_LIGHTING;
    const int last_bounce = direct_only ? 0 : rp.max_depth;
    // the headline kernels trace a single light's queries inside shade: no light-sampling lines, no shadow stage
    const WavefrontPlan plan = wavefront_plan(rp, scene);
    const bool hot = plan.hot, nee = plan.nee;

    // scene traits and the compiled-out headline configuration select the shade instantiation (kyd_wavefront.cuh)
    int traits = TRAITS_ANY;
    {
        bool all_rect = scene.n_lights > 0, all_sphere = scene.n_lights > 0;
        for (int l = 0; l < scene.n_lights; ++l)
        {
            const bool area = scene.lights[l].kind == KYD_LIGHT_AREA;
            all_rect = all_rect && area && scene.light_shape[l].kind == KYD_SHAPE_RECTANGLE;
            all_sphere = all_sphere && area && scene.light_shape[l].kind == KYD_SHAPE_SPHERE;
        }
        // what the specialised kernels assume beyond the light kinds (kyd_device.cuh, nee_bsdf_from_sample): the BSDF-sampled
        // query is in its occlusion form (no light carried by several surfaces), the surfaces that carry sphere lights are
        // spheres, and the single rectangle light's surface IS the light's rectangle
        bool unique = true, sphere_surfaces = true, same_rect = false;
        for (int l = 0; l < scene.n_lights; ++l)
        {
            const int ls = scene.light_surface[l];
            unique = unique && ls != -2;
            if (ls >= 0 && scene.bvh_nodes == nullptr)
                sphere_surfaces = sphere_surfaces && scene.surf_shape[ls].kind == KYD_SHAPE_SPHERE;
        }
        if (all_rect && scene.n_lights == 1 && scene.light_surface[0] >= 0 && scene.bvh_nodes == nullptr)
            same_rect = memcmp(&scene.surf_shape[scene.light_surface[0]], &scene.light_shape[0], sizeof(DevShape)) == 0;
        traits = (all_rect && scene.n_lights == 1 && same_rect) ? TRAITS_AREA_RECTANGLE
               : (all_sphere && unique && sphere_surfaces) ? TRAITS_AREA_SPHERE : TRAITS_ANY;
    }

    // persistent-style grids: enough blocks to fill every SM several times over, grid-stride inside
    // (blocks per SM; measured flat between 8 and 24, profiles/r01_ab_variants.txt; KYD_GRID256 / KYD_GRID128 override them for experiments)
    static const int grid256_per_sm = getenv("KYD_GRID256") ? atoi(getenv("KYD_GRID256")) : 16;
    static const int grid128_per_sm = getenv("KYD_GRID128") ? atoi(getenv("KYD_GRID128")) : 24;
    // The shade kernels (launch bounds: four 128-thread blocks per SM) run fastest as ONE resident wave of persistent blocks --
    // every further block pays the shared-memory staging and the queue-count loads again, and the lockstep barrier works on
    // blocks that start together (C5 shade 80.9 ms at 4 blocks per SM, 84.2 at 24, 87.1 at 48; C3 shade 53.8 / 58.9 / 63.9);
    // k_nee (six resident blocks per SM) wants several waves (profiles/r02_ab_variants.txt).  KYD_GRID_SHADE overrides.
    static const int grid_shade_per_sm = getenv("KYD_GRID_SHADE") ? atoi(getenv("KYD_GRID_SHADE")) : KYD_SHADE_MIN_BLOCKS;
    static const int grid_nee_per_sm = getenv("KYD_GRID_NEE") ? atoi(getenv("KYD_GRID_NEE")) : 36;
    const int grid256 = sm_count * grid256_per_sm, grid128 = sm_count * grid128_per_sm;
    const int grid_shade = sm_count * grid_shade_per_sm, grid_nee = sm_count * grid_nee_per_sm;

    if (!(rp.flags & KYD_FLAG_ACCUMULATE))
    {
        k_zero<<<sm_count * 8, 256, 0, stream>>>(film_dev, npix_total * 3);
        ++*launches;
    }

    // tile the film so that a wave fits the buffers; several samples per wave when the film is small
    const long long cap = wave_paths < w.capacity ? wave_paths : w.capacity;
    const long long tile = npix_total < cap ? npix_total : cap;
    const int spp_per_wave = (int)((cap / tile) < 1 ? 1 : (cap / tile));

    for (long long p0 = 0; p0 < npix_total; p0 += tile)
    {
        const int npix = (int)((npix_total - p0) < tile ? (npix_total - p0) : tile);
        for (int s0 = 0; s0 < nsamples; s0 += spp_per_wave)
        {
            WaveParams wp{};
            wp.rp = rp;
            wp.pixel_begin = (int)p0;
            wp.npix = npix;
            wp.sample_begin = rp.sample_begin + s0;
            wp.nspp = (nsamples - s0) < spp_per_wave ? (nsamples - s0) : spp_per_wave;
            wp.nslots = npix * wp.nspp;
            wp.plane = w.capacity;
            wp.direct_only = direct_only ? 1 : 0;
            wp.split_light_sample = (rp.flags & KYD_FLAG_SPLIT_LIGHT_SAMPLE) ? 1 : 0;
            wp.no_pending = nee ? 0 : 1;
            // (2: the sphere-light k_nee also leaves a per-vertex summary behind its results, kyd_wavefront.cuh add_pending)
            wp.pair_kernel = plan.pair_kernel ? ((KYD_NEE_DEFER && KYD_NEE_LIGHT_MAJOR && KYD_NEE_SUMMARY && !KYD_BIG_SCENE && traits == TRAITS_AREA_SPHERE) ? 2 : 1) : 0;
            wp.recursion = plan.recursion ? rp.integrator : 0;

            // queue tails start at zero; the camera rays are generated inside the first intersect launch
            cudaMemsetAsync(counters->queue, 0, sizeof(counters->queue), stream);
            for (int bounce = 0; bounce <= last_bounce; ++bounce)
            {
                if (bounce == 0 || !plan.fused)
                {
                    T(StageTimer::INTERSECT);
                    if (bounce == 0)
                        k_intersect<true><<<grid256, 256, 0, stream>>>(wp, w, counters, bounce);
                    else
                        k_intersect<false><<<grid256, 256, 0, stream>>>(wp, w, counters, bounce);
                    T(-1);
                    ++*launches;
                }
                if (plan.fused)   // the lobe queues the shade kernels of this bounce fill were consumed one bounce ago
                {
                    cudaMemsetAsync(&counters->queue[Q_LOBE0 + 4 * ((bounce & 1) ^ 1)], 0, 4 * sizeof(unsigned long long), stream);
                    if (plan.pair_kernel && bounce > 0)   // (and k_nee's vertex queues by the previous bounce's k_nee)
                        cudaMemsetAsync(&counters->queue[Q_NEE0], 0, 2 * sizeof(unsigned long long), stream);
                }
                T(StageTimer::SHADE);
#if !KYD_BIG_SCENE
                if (plan.recursion)
                {
                    if (rp.integrator == KYD_INT_SIMPLE_PT_RECURSION) launch_shade_rec<KYD_INT_SIMPLE_PT_RECURSION>(grid128, stream, wp, w, counters, bounce);
                    else if (rp.integrator == KYD_INT_PT_RECURSION) launch_shade_rec<KYD_INT_PT_RECURSION>(grid128, stream, wp, w, counters, bounce);
                    else launch_shade_rec<KYD_INT_PT_RECURSION_DEFERED>(grid128, stream, wp, w, counters, bounce);
                }
                else
#endif
                    launch_shade(traits, hot, plan.inline_queries, plan.fused, grid_shade, stream, wp, w, counters, bounce);
                T(-1);
                *launches += 4;
                if (plan.pair_kernel)
                {
                    // the light loop over (vertex, light) pairs, set-up and scene queries in one kernel per lobe
                    T(StageTimer::LIGHT_SAMPLE);
#if !KYD_BIG_SCENE
                    if (traits == TRAITS_AREA_SPHERE)
                    {
#if KYD_NEE_DEFER
                        k_nee<LOBE_LAMBERT, TRAITS_AREA_SPHERE><<<grid_nee, 128, 0, stream>>>(wp, w, counters);
                        k_nee<LOBE_PHONG, TRAITS_AREA_SPHERE><<<grid_nee, 128, 0, stream>>>(wp, w, counters);
#else
                        k_nee_serial<LOBE_LAMBERT, TRAITS_AREA_SPHERE><<<grid_nee, 128, 0, stream>>>(wp, w, counters);
                        k_nee_serial<LOBE_PHONG, TRAITS_AREA_SPHERE><<<grid_nee, 128, 0, stream>>>(wp, w, counters);
#endif
                    }
                    else
#endif
                    {
                        k_nee_serial<LOBE_LAMBERT, TRAITS_ANY><<<grid_nee, 128, 0, stream>>>(wp, w, counters);
                        k_nee_serial<LOBE_PHONG, TRAITS_ANY><<<grid_nee, 128, 0, stream>>>(wp, w, counters);
                    }
                    T(-1);
                    *launches += 2;
                }
                else if (nee)
                {
                    if (wp.split_light_sample)
                    {
                        T(StageTimer::LIGHT_SAMPLE);
                        k_light_sample<<<grid128, 128, 0, stream>>>(wp, w, counters);
                        T(-1);
                        ++*launches;
                    }
                    T(StageTimer::SHADOW);
                    if (wp.split_light_sample)
                        k_shadow<false><<<grid256, 256, 0, stream>>>(wp, w, counters);
                    else
                        k_shadow<true><<<grid256, 256, 0, stream>>>(wp, w, counters);
                    T(-1);
                    ++*launches;
                }
            }
            T(StageTimer::ACCUMULATE);
            if (plan.recursion)
            {
                k_unwind<<<grid256, 256, 0, stream>>>(wp, w);   // the recursion's return path, into the slot k_accumulate reads
                ++*launches;
            }
            k_accumulate<<<grid256, 256, 0, stream>>>(wp, w, film_dev);
            T(-1);
            ++*launches;
        }
    }
    if (rp.flags & KYD_FLAG_CLAMP)
    {
        launch_clamp(film_dev, npix_total * 3, stream);
        ++*launches;
    }
}

} // namespace KYD_KERNEL_NS
