// kyd_wavefront.cuh -- the wavefront organisation of path_tracing_iteration_t / direct_lighting_t
// (reference ky.cpp:4523-4618, 4125-4155): separate raygen, intersect, shade (BSDF sample), light-sample
// (NEE + MIS set-up), shadow (NEE ray queries) and accumulate kernels over SoA path state in HBM, with
// warp-ballot compacted queues of path slots between them.  Included by kyd_kernels.cu only.
//
// A wave is a tile of `npix` consecutive pixels times `nspp` consecutive sample indices; path slot
// = s_local * npix + pixel_local.  Per-slot results do not depend on how the film is cut into waves.
//
// Order of FP32 additions into a path's radiance Lo is the reference's: emitted light of a vertex, then
// beta * Ld of that vertex (ky.cpp:4553-4576).  The Ld of a vertex becomes known one stage later than the
// vertex is shaded, so it is added ("pending") at the start of the path's next shade, or by the accumulate
// kernel if the path ended -- before anything else is added in both cases.
#pragma once

#include "kyd_device.cuh"
#include "kyd_internal.h"

namespace kyd {

enum { Q_CUR = 0, Q_NEXT = 1, Q_NEE = 2 };
enum { FLAG_PREV_SPECULAR = 1 };

struct WaveParams
{
    RenderParams rp;
    int pixel_begin, npix;    // tile of the film (linear pixel indices)
    int sample_begin, nspp;   // sample indices of this wave
    int nslots;               // npix * nspp
    long long plane;          // stride between per-light planes of the NEE buffers (= capacity)
    int direct_only;          // direct_lighting_t: stop after the first vertex' light loop
};

KYD_DEV void flush_counters(unsigned rays, unsigned traced, DevCounters* counters)
{
    rays = __reduce_add_sync(0xffffffffu, rays);
    traced = __reduce_add_sync(0xffffffffu, traced);
    if ((threadIdx.x & 31) == 0 && (rays | traced))
    {
        atomicAdd(&counters->rays, (unsigned long long)rays);
        atomicAdd(&counters->rays_traced, (unsigned long long)traced);
    }
}

// warp-aggregated push: one atomic per warp, ballot + popc prefix for the lane offsets
KYD_DEV void queue_push(bool pred, int value, int* __restrict__ queue, unsigned long long* __restrict__ tail)
{
    unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0)
        return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader)
        base = atomicAdd(tail, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred)
        queue[base + __popc(mask & ((1u << lane) - 1))] = value;
}

KYD_DEV uint2 pack_rng(unsigned long long s) { return make_uint2((unsigned)s, (unsigned)(s >> 32)); }
KYD_DEV unsigned long long unpack_rng(uint2 v) { return (unsigned long long)v.x | ((unsigned long long)v.y << 32); }

// ---- raygen: camera_t::generate_ray for every slot of the wave (ky.cpp:3714-3715) ----------------------
__global__ void __launch_bounds__(256) k_raygen(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters)
{
    const int stride = gridDim.x * blockDim.x;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < wp.nslots; slot += stride)
    {
        const int pixel = wp.pixel_begin + slot % wp.npix;
        const int s = wp.sample_begin + slot / wp.npix;
        const int x = pixel % wp.rp.width, y = pixel / wp.rp.width;
        Sampler smp;
        smp.start(wp.rp.sampler, wp.rp.seed, x, y, s);
        float2 jitter = smp.get_float2();
        Ray r = generate_ray((float)x + jitter.x, (float)y + jitter.y);
        w.ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, r.tmax);
        w.ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, 0.f);
        w.beta[slot] = make_float4(1.f, 1.f, 1.f, __int_as_float(0)); // w: flags | bounce << 8
        w.radiance[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
        w.vx_beta[slot] = make_float4(0.f, 0.f, 0.f, 0.f);            // w: pending light count
        w.rng[slot] = pack_rng(smp.state);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        counters->queue[Q_CUR] = (unsigned long long)wp.nslots;
        counters->queue[Q_NEXT] = 0;
        counters->queue[Q_NEE] = 0;
    }
}

// ---- intersect: scene_t::intersect closest-hit query for every queued path (ky.cpp:3172-3184) -------------
// identity == true: the queue is 0..n-1 (first bounce of a wave)
template <bool IDENTITY>
__global__ void __launch_bounds__(256) k_intersect(WaveBuffers w, const int* __restrict__ queue, DevCounters* __restrict__ counters, int qsel, int qnext)
{
    const int n = (int)counters->queue[qsel];
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        // the tails this bounce's shade will push to (their previous contents were consumed by earlier kernels)
        counters->queue[qnext] = 0;
        counters->queue[Q_NEE] = 0;
    }
    const int stride = gridDim.x * blockDim.x;
    unsigned rays = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const int slot = IDENTITY ? i : queue[i];
        float4 o = w.ray_o[slot], d = w.ray_d[slot];
        Ray r;
        r.o = V3(o.x, o.y, o.z);
        r.d = V3(d.x, d.y, d.z);
        r.tmax = o.w;
        float t;
        int s = scene_closest(r, &t);
        w.hit[slot] = make_float2(t, __int_as_float(s));
        rays++;
    }
    flush_counters(rays, rays, counters);
}

// ---- shade: one path vertex (ky.cpp:4545-4613 without the light loop) ----------------------------------
template <bool IDENTITY>
__global__ void __launch_bounds__(128) k_shade(WaveParams wp, WaveBuffers w, const int* __restrict__ queue, int* __restrict__ next_queue,
                                               int* __restrict__ nee_queue, DevCounters* __restrict__ counters, int qsel, int qnext, int bounce)
{
    const int n = (int)counters->queue[qsel];
    const int stride = gridDim.x * blockDim.x;
    const int n_lights = c_scene.n_lights;
    const int base_i = blockIdx.x * blockDim.x + threadIdx.x;
    // whole warps iterate together so that the ballots in queue_push are convergent
    for (int i0 = base_i - (threadIdx.x & 31); i0 < n; i0 += stride)
    {
        const int i = i0 + (threadIdx.x & 31);
        bool alive = false, wants_nee = false;
        int slot = 0;
        if (i < n)
        {
            slot = IDENTITY ? i : queue[i];
            float4 o4 = w.ray_o[slot], d4 = w.ray_d[slot], b4 = w.beta[slot], L4 = w.radiance[slot];
            float2 h = w.hit[slot];
            Ray r;
            r.o = V3(o4.x, o4.y, o4.z);
            r.d = V3(d4.x, d4.y, d4.z);
            r.tmax = o4.w;
            float3 beta = V3(b4.x, b4.y, b4.z), Lo = V3(L4.x, L4.y, L4.z);
            const int flags = __float_as_int(b4.w);
            const int surface = __float_as_int(h.y);
            const bool hit = surface >= 0;

            // light gathered at the previous vertex (see file header)
            float4 vb = w.vx_beta[slot];
            const int pending = __float_as_int(vb.w);
            if (pending > 0)
            {
                float3 Ld = KYD_BLACK;
                for (int l = 0; l < pending; ++l)
                {
                    float4 e = w.nee_result[(long long)l * wp.plane + slot];
                    Ld = add(Ld, V3(e.x, e.y, e.z));
                }
                Lo = add(Lo, cmulc(V3(vb.x, vb.y, vb.z), Ld));
            }
            int new_pending = 0;

            HitGeom g;
            if (hit)
                g = shape_hit_geom(c_scene.surf_shape[surface], r, h.x);

            if (bounce == 0 || (flags & FLAG_PREV_SPECULAR))
                Lo = add(Lo, cmulc(beta, hit ? surface_emission(surface, g) : environment_lighting()));

            if (hit && bounce < wp.rp.max_depth + wp.direct_only)
            {
                Bsdf b;
                material_scattering(c_scene.materials[c_scene.surf_material[surface]], g, &b);
                Sampler smp;
                smp.debug = (wp.rp.sampler == KYD_SAMPLER_DEBUG);
                smp.state = unpack_rng(w.rng[slot]);

                if (!bsdf_is_delta(b.lobe))
                {
                    if (wp.rp.direct_sample != KYD_DS_IDLE && n_lights > 0)
                    {
                        // vertex record for the light-sample stage
                        w.vx_position[slot] = make_float4(g.position.x, g.position.y, g.position.z, __int_as_float(b.lobe));
                        w.vx_normal[slot] = make_float4(g.normal.x, g.normal.y, g.normal.z, b.exponent);
                        w.vx_wo[slot] = make_float4(g.wo.x, g.wo.y, g.wo.z, 0.f);
                        w.vx_color[slot] = make_float4(b.a.x, b.a.y, b.a.z, 0.f);
                        w.vx_rng[slot] = pack_rng(smp.state);
                        new_pending = n_lights;
                        wants_nee = true;
                    }
                    // sample_all_light draws 4 floats per light, plus 2 per non-delta light in `bsdf` mode
                    smp.skip(4 * n_lights + (wp.rp.direct_sample == KYD_DS_BSDF ? 2 * c_scene.n_nondelta_lights : 0));
                }

                if (!wp.direct_only)
                {
                    BsdfSample bs = bsdf_sample(b, g.wo, smp.get_float2());
                    if (!(is_black(bs.f) || bs.pdf == 0.f))
                    {
                        float3 vertex_beta = beta;
                        beta = cmulc(beta, cdiv(mul(bs.f, abs_dot(bs.wi, g.normal)), bs.pdf));
                        Ray nr = spawn_ray(g, bs.wi);
                        alive = true;
                        if (bounce > 3)
                        {
                            float q = max_std(0.05f, 1 - max_component(beta));
                            if (smp.get_float() < q)
                                alive = false;
                            else
                                beta = mul(beta, 1 / (1 - q));
                        }
                        if (alive)
                        {
                            w.ray_o[slot] = make_float4(nr.o.x, nr.o.y, nr.o.z, nr.tmax);
                            w.ray_d[slot] = make_float4(nr.d.x, nr.d.y, nr.d.z, 0.f);
                            w.beta[slot] = make_float4(beta.x, beta.y, beta.z, __int_as_float((bs.type & BSDF_SPECULAR) ? FLAG_PREV_SPECULAR : 0));
                            w.rng[slot] = pack_rng(smp.state);
                        }
                        beta = vertex_beta;
                    }
                }
            }
            // beta of THIS vertex multiplies its Ld later
            if (new_pending > 0 || pending > 0)
                w.vx_beta[slot] = make_float4(beta.x, beta.y, beta.z, __int_as_float(new_pending));
            w.radiance[slot] = make_float4(Lo.x, Lo.y, Lo.z, 0.f);
        }
        queue_push(alive, slot, next_queue, &counters->queue[qnext]);
        queue_push(wants_nee, slot, nee_queue, &counters->queue[Q_NEE]);
    }
}

// draws consumed by sample_all_light before light l (ky.cpp:3864-3869, 3900)
KYD_DEV int light_draw_offset(int l, int direct_sample)
{
    int n = 4 * l;
    if (direct_sample == KYD_DS_BSDF)
        for (int j = 0; j < l; ++j)
            n += light_is_delta(c_scene.lights[j].kind) ? 0 : 2;
    return n;
}

KYD_DEV void store_nee(const WaveBuffers& w, long long at, const NeeRay& q)
{
    // tmax < 0 marks "no query"; w of nee_d: bit 0 = the reference issues this query (ray statistics)
    w.nee_o[at] = make_float4(q.ray.o.x, q.ray.o.y, q.ray.o.z, q.active ? q.ray.tmax : -1.f);
    w.nee_d[at] = make_float4(q.ray.d.x, q.ray.d.y, q.ray.d.z, __int_as_float(q.ref_query ? 1 : 0));
    w.nee_value[at] = make_float4(q.value.x, q.value.y, q.value.z, 0.f);
}

// ---- light-sample: NEE + MIS set-up for every (vertex, light) (ky.cpp:3864-3869 + first halves of 3889-4074)
__global__ void __launch_bounds__(128) k_light_sample(WaveParams wp, WaveBuffers w, const int* __restrict__ nee_queue, DevCounters* __restrict__ counters)
{
    const int n = (int)counters->queue[Q_NEE];
    const int n_lights = c_scene.n_lights;
    const long long total = (long long)n * n_lights;
    const int stride = gridDim.x * blockDim.x;
    const int ds = wp.rp.direct_sample;
    for (long long idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
    {
        const int l = (int)(idx / n);            // light-major: a warp works on one light
        const int slot = nee_queue[idx - (long long)l * n];
        float4 p4 = w.vx_position[slot], n4 = w.vx_normal[slot], wo4 = w.vx_wo[slot], c4 = w.vx_color[slot];
        HitGeom g;
        g.position = V3(p4.x, p4.y, p4.z);
        g.normal = V3(n4.x, n4.y, n4.z);
        g.wo = V3(wo4.x, wo4.y, wo4.z);
        Bsdf b;
        b.f = frame_from_z(g.normal);
        b.a = V3(c4.x, c4.y, c4.z);
        b.t = KYD_BLACK;
        b.eta_t = 1.f;
        b.exponent = n4.w;
        b.lobe = __float_as_int(p4.w);

        Sampler smp;
        smp.debug = (wp.rp.sampler == KYD_SAMPLER_DEBUG);
        smp.state = unpack_rng(w.vx_rng[slot]);
        smp.skip(light_draw_offset(l, ds));
        float2 random_bsdf = smp.get_float2();
        float2 random_light = smp.get_float2();

        NeeRay qb, ql;
        qb.active = ql.active = false;
        qb.ref_query = ql.ref_query = false;
        qb.value = ql.value = KYD_BLACK;
        qb.ray.o = qb.ray.d = ql.ray.o = ql.ray.d = V3(0, 0, 0);
        qb.ray.tmax = ql.ray.tmax = -1.f;
        if (ds == KYD_DS_BSDF)
        {
            if (!light_is_delta(c_scene.lights[l].kind))
                qb = nee_bsdf_setup(g, b, l, smp.get_float2(), false);
        }
        else if (ds == KYD_DS_BSDF_MIS || ds == KYD_DS_BOTH_MIS)
            qb = nee_bsdf_setup(g, b, l, random_bsdf, true);
        if (ds == KYD_DS_LIGHT)
            ql = nee_light_setup(g, b, l, random_light, false);
        else if (ds == KYD_DS_LIGHT_MIS || ds == KYD_DS_BOTH_MIS)
            ql = nee_light_setup(g, b, l, random_light, true);

        const long long at = (long long)(2 * l) * wp.plane + slot;
        store_nee(w, at, qb);
        store_nee(w, at + wp.plane, ql);
    }
}

// ---- shadow: the scene queries of the light loop and the estimators' second halves ------------------------
// closest-hit query for the BSDF-sampled direction, occlusion query for the light-sampled point; writes
// the estimator value of (vertex, light): Lb, Ll or 0.5 Lb + 0.5 Ll (ky.cpp:4083)
__global__ void __launch_bounds__(256) k_shadow(WaveParams wp, WaveBuffers w, const int* __restrict__ nee_queue, DevCounters* __restrict__ counters)
{
    const int n = (int)counters->queue[Q_NEE];
    const int n_lights = c_scene.n_lights;
    const long long total = (long long)n * n_lights;
    const int stride = gridDim.x * blockDim.x;
    const int ds = wp.rp.direct_sample;
    unsigned rays = 0, traced = 0;
    for (long long idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
    {
        const int l = (int)(idx / n);
        const int slot = nee_queue[idx - (long long)l * n];
        const long long at = (long long)(2 * l) * wp.plane + slot;

        float3 Lb = KYD_BLACK, Ll = KYD_BLACK;
        {
            float4 o = w.nee_o[at], d = w.nee_d[at];
            rays += (unsigned)(__float_as_int(d.w) & 1);
            if (o.w >= 0.f)
            {
                NeeRay q;
                q.ray.o = V3(o.x, o.y, o.z);
                q.ray.d = V3(d.x, d.y, d.z);
                q.ray.tmax = o.w;
                float4 v = w.nee_value[at];
                q.value = V3(v.x, v.y, v.z);
                q.light = l;
                float t;
                int s = scene_closest(q.ray, &t);
                Lb = nee_bsdf_resolve(q, s, t);
                traced++;
            }
        }
        {
            float4 o = w.nee_o[at + wp.plane], d = w.nee_d[at + wp.plane];
            rays += (unsigned)(__float_as_int(d.w) & 1);
            if (o.w >= 0.f)
            {
                Ray r;
                r.o = V3(o.x, o.y, o.z);
                r.d = V3(d.x, d.y, d.z);
                r.tmax = o.w;
                float4 v = w.nee_value[at + wp.plane];
                Ll = scene_any_hit(r) ? KYD_BLACK : V3(v.x, v.y, v.z);
                traced++;
            }
        }
        float3 e;
        if (ds == KYD_DS_BOTH_MIS)
            e = add(mul(Lb, 0.5f), mul(Ll, 0.5f));
        else if (ds == KYD_DS_BSDF || ds == KYD_DS_BSDF_MIS)
            e = Lb;
        else
            e = Ll;
        w.nee_result[(long long)l * wp.plane + slot] = make_float4(e.x, e.y, e.z, 0.f);
    }
    flush_counters(rays, traced, counters);
}

// ---- accumulate: film_t::add_color order -- L = L + Li * (1/spp) sample after sample (ky.cpp:3717-3721) ------
__global__ void __launch_bounds__(256) k_accumulate(WaveParams wp, WaveBuffers w, float* __restrict__ film)
{
    const int stride = gridDim.x * blockDim.x;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < wp.npix; p += stride)
    {
        float* o = film + 3 * (size_t)(wp.pixel_begin + p);
        float3 L = V3(o[0], o[1], o[2]);
        for (int s = 0; s < wp.nspp; ++s)
        {
            const int slot = s * wp.npix + p;
            float4 L4 = w.radiance[slot];
            float3 Li = V3(L4.x, L4.y, L4.z);
            float4 vb = w.vx_beta[slot];
            const int pending = __float_as_int(vb.w);
            if (pending > 0)
            {
                float3 Ld = KYD_BLACK;
                for (int l = 0; l < pending; ++l)
                {
                    float4 e = w.nee_result[(long long)l * wp.plane + slot];
                    Ld = add(Ld, V3(e.x, e.y, e.z));
                }
                Li = add(Li, cmulc(V3(vb.x, vb.y, vb.z), Ld));
            }
            L = add(L, mul(Li, wp.rp.weight));
        }
        o[0] = L.x; o[1] = L.y; o[2] = L.z;
    }
}

__global__ void k_zero(float* __restrict__ p, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        p[i] = 0.f;
}

} // namespace kyd
