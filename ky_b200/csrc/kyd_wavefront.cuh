// kyd_wavefront.cuh -- the wavefront organisation of path_tracing_iteration_t / direct_lighting_t
// (reference ky.cpp:4523-4618, 4125-4155): raygen, intersect, shade (BSDF sample), light-sample (NEE + MIS
// set-up), shadow (NEE ray queries) and accumulate kernels over path state in HBM, with warp-ballot
// compacted queues of path slots between them.  Included by kyd_kernels.cu only.
//
// A wave is a tile of `npix` consecutive pixels times `nspp` consecutive sample indices; path slot
// = s_local * npix + pixel_local.  Per-slot results do not depend on how the film is cut into waves.
//
// Divergence control (profiles/r01_wave1M_*: the first version ran shade at 11 of 32 lanes): the intersect
// stage sorts the paths it hit by the BSDF lobe they will be shaded with (Lambert / Phong / mirror / glass;
// a plastic surface picks its lobe with a hash of the hit, so the lobe is known there) into one queue per
// lobe.  Shade and light-sample walk the queues one lobe after the other with lobe-specialised code, so a
// warp never mixes lobes.  Paths that miss the scene are finished inside intersect.
//
// Data layout (profiles/r01_wave_lobe_soa_*: lobe-sorted queues turn the state accesses into gathers, and
// 16-byte SoA elements then waste half of every 32-byte DRAM sector): the state of a path is ONE 128-byte
// line, fields grouped by the sector the stages touch together:
//     sector 0   ray origin.xyz, tmax | ray direction.xyz, flags          intersect reads, shade rewrites
//     sector 1   beta.rgb, - | Lo.rgb, -                                   shade
//     sector 2   beta of the vertex that sampled lights .rgb, pending | rng state, -     shade, accumulate
//     sector 3   hit distance, surface, -, - | -                          intersect writes, shade reads
// and the light-sampling record of a (light, path) pair is one line: the BSDF-sampled query (origin, tmax |
// direction, flag | value), the light-sampled query (same three), and the estimator's result.
// Every stage reads and writes whole sectors.
//
// Order of FP32 additions into a path's radiance Lo is the reference's: emitted light of a vertex, then
// beta * Ld of that vertex (ky.cpp:4553-4576).  The Ld of a vertex becomes known one stage later than the
// vertex is shaded, so it is added ("pending") at the start of the path's next shade (or when the path
// misses, or by the accumulate kernel if the path ended) -- before anything else is added in every case.
#pragma once

#include "kyd_device.cuh"
#include "kyd_internal.h"

namespace kyd {

// queue tails in DevCounters::queue
enum { Q_RAY0 = 0, Q_RAY1 = 1, Q_NEE0 = 2 /* +lobe (Lambert, Phong) */, Q_LOBE0 = 4 /* + 4 * parity + lobe */ };
enum { FLAG_PREV_SPECULAR = 1 };

#ifndef KYD_SHADE_MIN_BLOCKS
#define KYD_SHADE_MIN_BLOCKS 4
#endif
#define SHADE_THREADS 128

// float4 units of a path line / a light-sampling line
enum { P_ORIGIN = 0, P_DIRECTION = 1, P_BETA = 2, P_RADIANCE = 3, P_VERTEX_BETA = 4, P_RNG = 5, P_HIT = 6, P_HIT_PAD = 7, PATH_UNITS = 8 };
enum { N_BSDF_O = 0, N_BSDF_D = 1, N_BSDF_VALUE = 2, N_LIGHT_O = 3, N_LIGHT_D = 4, N_LIGHT_VALUE = 5, N_RESULT = 6, N_RESULT_PAD = 7, NEE_UNITS = 8 };
enum { V_POSITION = 0, V_NORMAL = 1, V_WO = 2, V_COLOR = 3, V_RNG = 4, V_PAD = 5, VERTEX_UNITS = 6 };

struct WaveParams
{
    RenderParams rp;
    int pixel_begin, npix;    // tile of the film (linear pixel indices)
    int sample_begin, nspp;   // sample indices of this wave
    int nslots;               // npix * nspp
    long long plane;          // lines between the per-light planes of the light-sampling buffer (= capacity)
    int direct_only;          // direct_lighting_t: stop after the first vertex' light loop
    int split_light_sample;   // light-sample as its own kernel (KYD_FLAG_SPLIT_LIGHT_SAMPLE) instead of inside shade
};

KYD_DEV float4* path_line(const WaveBuffers& w, int slot) { return w.path + (size_t)slot * PATH_UNITS; }
KYD_DEV float4* nee_line(const WaveBuffers& w, long long plane, int light, int slot) { return w.nee + ((size_t)light * plane + slot) * NEE_UNITS; }
KYD_DEV float4* vertex_line(const WaveBuffers& w, int slot) { return w.vertex + (size_t)slot * VERTEX_UNITS; }

KYD_DEV void flush_counters(unsigned rays, unsigned traced, DevCounters* counters)
{
    __syncwarp();
    rays = __reduce_add_sync(0xffffffffu, rays);
    traced = __reduce_add_sync(0xffffffffu, traced);
    if ((threadIdx.x & 31) == 0 && (rays | traced))
    {
        atomicAdd(&counters->rays, (unsigned long long)rays);
        atomicAdd(&counters->rays_traced, (unsigned long long)traced);
    }
}

// warp-aggregated push: one atomic per warp, ballot + popc prefix for the lane offsets.  Must be reached
// by all 32 lanes; the __syncwarp() makes them arrive together (without it the lanes of a diverged warp
// execute the ballot / shuffle group by group: 18 % of the first version's shade instructions).
KYD_DEV void queue_push(bool pred, int value, int* __restrict__ queue, unsigned long long* __restrict__ tail)
{
    __syncwarp();
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0)
        return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader)
        base = atomicAdd(tail, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred)
        queue[base + __popc(mask & ((1u << lane) - 1))] = value;
}

// Pushes into up to NQ queues with ONE round of atomics per warp and iteration (lane q issues queue q's
// atomic, so the NQ atomics are in flight together) whose latency is hidden: reserve() issues them and
// commit(), called one loop iteration later, reads the returned bases and writes the entries.  In the first
// per-lobe version 61 % of k_intersect's stall samples were lanes waiting for these atomics
// (profiles/r01_wave_lines_stalls.txt).
template <int NQ>
struct WarpPush
{
    unsigned masks[NQ];        // ballots of the iteration whose entries are not written yet
    unsigned long long base;   // lane q < NQ: first index reserved in queue q
    int value;                 // this lane's entry
    unsigned preds;            // bit q: this lane pushes `value` into queue q
    bool pending;

    KYD_DEV void init() { pending = false; preds = 0; value = 0; base = 0; }

    KYD_DEV void reserve(unsigned preds_, int value_, unsigned long long* const (&tails)[NQ])
    {
        __syncwarp();
        preds = preds_;
        value = value_;
        const int lane = threadIdx.x & 31;
        base = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
        {
            masks[q] = __ballot_sync(0xffffffffu, (preds_ >> q) & 1u);
            if (lane == q && masks[q] != 0)
                base = atomicAdd(tails[q], (unsigned long long)__popc(masks[q]));
        }
        pending = true;
    }

    KYD_DEV void commit(int* const (&queues)[NQ])
    {
        if (!pending) // warp-uniform
            return;
        __syncwarp();
        const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
        {
            const unsigned long long b = __shfl_sync(0xffffffffu, base, q);
            if ((preds >> q) & 1u)
                queues[q][b + __popc(masks[q] & lt)] = value;
        }
        pending = false;
    }
};

KYD_DEV unsigned long long unpack_rng(float4 v) { return (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32); }
KYD_DEV float4 pack_rng(unsigned long long s) { return make_float4(__uint_as_float((unsigned)s), __uint_as_float((unsigned)(s >> 32)), 0.f, 0.f); }

// Lo += beta_vertex * (sum over lights of the vertex' estimator values), ky.cpp:4575-4576.
// first: the result of light 0 when the caller already fetched it (shade prefetches it with the path line)
KYD_DEV float3 add_pending(const WaveBuffers& w, long long plane, int slot, float4 vb, float3 Lo, const float4* first = nullptr)
{
    const int pending = __float_as_int(vb.w);
    if (pending > 0)
    {
        float3 Ld = KYD_BLACK;
        for (int l = 0; l < pending; ++l)
        {
            float4 e = (l == 0 && first) ? *first : nee_line(w, plane, l, slot)[N_RESULT];
            Ld = add(Ld, V3(e.x, e.y, e.z));
        }
        Lo = add(Lo, cmulc(V3(vb.x, vb.y, vb.z), Ld));
    }
    return Lo;
}

// Ampere-style asynchronous 16-byte global -> shared copies (LDGSTS): the gather of the NEXT path's line
// is in flight while the current path is shaded, at no register cost
KYD_DEV void cp_async16(void* smem, const void* gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
KYD_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
KYD_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

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
        float4* p = path_line(w, slot);
        p[P_ORIGIN] = make_float4(r.o.x, r.o.y, r.o.z, r.tmax);
        p[P_DIRECTION] = make_float4(r.d.x, r.d.y, r.d.z, __int_as_float(0));
        p[P_BETA] = make_float4(1.f, 1.f, 1.f, 0.f);
        p[P_RADIANCE] = make_float4(0.f, 0.f, 0.f, 0.f);
        p[P_VERTEX_BETA] = make_float4(0.f, 0.f, 0.f, __int_as_float(0)); // w: pending light count
        p[P_RNG] = pack_rng(smp.state);
    }
    if (blockIdx.x == 0 && threadIdx.x < 16)
        counters->queue[threadIdx.x] = threadIdx.x == Q_RAY0 ? (unsigned long long)wp.nslots : 0ull;
}

// the lobe material_t::scattering will build for this hit (ky.cpp:2587-2671); pure function of the hit
KYD_DEV int classify_lobe(int surface, const Ray& r, float t)
{
    const DevMaterial& m = c_scene.materials[c_scene.surf_material[surface]];
    if (m.kind == KYD_MAT_MATTE) return LOBE_LAMBERT;
    if (m.kind == KYD_MAT_MIRROR) return LOBE_MIRROR;
    if (m.kind == KYD_MAT_GLASS) return LOBE_FRESNEL;
    return plastic_random(ray_at(r, t), neg(r.d)) < m.p_specular ? LOBE_PHONG : LOBE_LAMBERT;
}

// ---- intersect: scene_t::intersect closest-hit query for every queued path (ky.cpp:3172-3184) -------------
// IDENTITY: the queue is 0..n-1 (first bounce of a wave).  Hits go to the queue of their lobe; a miss ends
// the path here (ky.cpp:4555-4563).
template <bool IDENTITY>
__global__ void __launch_bounds__(256) k_intersect(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters, int bounce)
{
    const int parity = bounce & 1;
    const int n = (int)counters->queue[Q_RAY0 + parity];
    if (blockIdx.x == 0 && threadIdx.x < 7)
    {
        // tails that later kernels of this bounce push to; their previous contents were consumed by earlier kernels
        const int which[7] = { Q_RAY0 + (parity ^ 1), Q_NEE0, Q_NEE0 + 1, Q_LOBE0 + 4 * (parity ^ 1), Q_LOBE0 + 4 * (parity ^ 1) + 1,
                               Q_LOBE0 + 4 * (parity ^ 1) + 2, Q_LOBE0 + 4 * (parity ^ 1) + 3 };
        counters->queue[which[threadIdx.x]] = 0;
    }
    const int* __restrict__ queue = parity ? w.queue_b : w.queue_a;
    unsigned long long* const tails[4] = { &counters->queue[Q_LOBE0 + 4 * parity], &counters->queue[Q_LOBE0 + 4 * parity + 1],
                                           &counters->queue[Q_LOBE0 + 4 * parity + 2], &counters->queue[Q_LOBE0 + 4 * parity + 3] };
    int* const lobe_queues[4] = { w.queue_lobe[0], w.queue_lobe[1], w.queue_lobe[2], w.queue_lobe[3] };
    WarpPush<4> push;
    push.init();
    const bool has_env = c_scene.env_light >= 0;
    const int stride = gridDim.x * blockDim.x;
    unsigned rays = 0;
    const int base_i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i0 = base_i - (threadIdx.x & 31); i0 < n; i0 += stride)
    {
        const int i = i0 + (threadIdx.x & 31);
        int lobe = -1, slot = 0;
        if (i < n)
        {
            slot = IDENTITY ? i : queue[i];
            float4* p = path_line(w, slot);
            float4 o = p[P_ORIGIN], d = p[P_DIRECTION];
            Ray r;
            r.o = V3(o.x, o.y, o.z);
            r.d = V3(d.x, d.y, d.z);
            r.tmax = o.w;
            float t;
            const int s = scene_closest(r, &t);
            rays++;
            if (s >= 0)
            {
                p[P_HIT] = make_float4(t, __int_as_float(s), 0.f, 0.f);
                p[P_HIT_PAD] = make_float4(0.f, 0.f, 0.f, 0.f);
                lobe = classify_lobe(s, r, t);
            }
            else if (has_env && (bounce == 0 || (__float_as_int(d.w) & FLAG_PREV_SPECULAR)))
            {
                // Lo += beta * environment_lighting at the camera vertex or after a specular bounce
                float4 b4 = p[P_BETA], L4 = p[P_RADIANCE], vb = p[P_VERTEX_BETA];
                float3 Lo = add_pending(w, wp.plane, slot, vb, V3(L4.x, L4.y, L4.z));
                Lo = add(Lo, cmulc(V3(b4.x, b4.y, b4.z), environment_lighting()));
                p[P_BETA] = b4;
                p[P_RADIANCE] = make_float4(Lo.x, Lo.y, Lo.z, 0.f);
                if (__float_as_int(vb.w) > 0)
                {
                    p[P_VERTEX_BETA] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
                    p[P_RNG] = p[P_RNG];
                }
            }
        }
        push.commit(lobe_queues);                                  // the previous iteration's entries
        push.reserve(lobe >= 0 ? (1u << lobe) : 0u, slot, tails);  // this iteration's atomics, consumed next time round
    }
    push.commit(lobe_queues);
    flush_counters(rays, rays, counters);
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

KYD_DEV void store_nee(float4* line, int first_unit, const NeeRay& q)
{
    // tmax < 0 marks "no query"; w of the direction: bit 0 = the reference issues this query (ray statistics)
    line[first_unit] = make_float4(q.ray.o.x, q.ray.o.y, q.ray.o.z, q.active ? q.ray.tmax : -1.f);
    line[first_unit + 1] = make_float4(q.ray.d.x, q.ray.d.y, q.ray.d.z, __int_as_float(q.ref_query ? 1 : 0));
    line[first_unit + 2] = make_float4(q.value.x, q.value.y, q.value.z, 0.f);
}

// NEE + MIS set-up of one (vertex, light): the two draws, the first halves of the estimators
// (ky.cpp:3864-3869, 3889-4074), and the queries written to the pair's line
KYD_DEV void light_sample_pair(const WaveParams& wp, const WaveBuffers& w, const HitGeom& g, const Bsdf& b, int l, int slot, Sampler smp)
{
    const int ds = wp.rp.direct_sample;
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

    float4* line = nee_line(w, wp.plane, l, slot);
    store_nee(line, N_BSDF_O, qb);
    store_nee(line, N_LIGHT_O, ql);
}

// ---- shade: one path vertex (ky.cpp:4545-4613), specialised by lobe ---------------------------------------
// pre / PRE_STRIDE: where the path line's units are read from -- the line itself (stride 1) or this thread's
// column of the shared-memory prefetch buffer; pre_result0: light 0's pending estimator value
template <int LOBE, int PRE_STRIDE>
KYD_DEV void shade_vertex(const WaveParams& wp, const WaveBuffers& w, const float4* pre, const float4* pre_result0, int slot, int bounce, int n_lights,
                          bool* out_alive, bool* out_nee)
{
    float4* p = path_line(w, slot);
    float4 o4 = pre[P_ORIGIN * PRE_STRIDE], d4 = pre[P_DIRECTION * PRE_STRIDE], b4 = pre[P_BETA * PRE_STRIDE], L4 = pre[P_RADIANCE * PRE_STRIDE];
    float4 vb = pre[P_VERTEX_BETA * PRE_STRIDE], rng4 = pre[P_RNG * PRE_STRIDE], h = pre[P_HIT * PRE_STRIDE], res0 = *pre_result0;
    Ray r;
    r.o = V3(o4.x, o4.y, o4.z);
    r.d = V3(d4.x, d4.y, d4.z);
    r.tmax = o4.w;
    float3 beta = V3(b4.x, b4.y, b4.z);
    const int flags = __float_as_int(d4.w);
    const int surface = __float_as_int(h.y);

    // light gathered at the previous vertex (see file header)
    float3 Lo = add_pending(w, wp.plane, slot, vb, V3(L4.x, L4.y, L4.z), &res0);
    int new_pending = 0;

    HitGeom g = shape_hit_geom(c_scene.surf_shape[surface], r, h.x);

    if (bounce == 0 || (flags & FLAG_PREV_SPECULAR))
        Lo = add(Lo, cmulc(beta, surface_emission(surface, g)));

    float3 next_beta = beta;
    unsigned long long rng_state = unpack_rng(rng4);
    if (bounce < wp.rp.max_depth + wp.direct_only)
    {
        const DevMaterial& m = c_scene.materials[c_scene.surf_material[surface]];
        Bsdf b;
        b.f = frame_from_z(g.normal);
        b.t = KYD_BLACK;
        b.eta_t = 1.f;
        b.exponent = 0.f;
        b.lobe = LOBE;
        if (LOBE == LOBE_LAMBERT) b.a = m.kind == KYD_MAT_PLASTIC ? m.plastic_lambert : m.diffuse;
        else if (LOBE == LOBE_PHONG) { b.a = m.plastic_phong; b.exponent = m.exponent; }
        else if (LOBE == LOBE_MIRROR) b.a = m.specular;
        else { b.a = m.specular; b.t = m.transmission; b.eta_t = m.eta; }

        Sampler smp;
        smp.debug = (wp.rp.sampler == KYD_SAMPLER_DEBUG);
        smp.state = rng_state;

        if (LOBE == LOBE_LAMBERT || LOBE == LOBE_PHONG)
        {
            if (wp.rp.direct_sample != KYD_DS_IDLE && n_lights > 0)
            {
                if (wp.split_light_sample)
                {
                    // vertex record for the light-sample stage
                    float4* v = vertex_line(w, slot);
                    v[V_POSITION] = make_float4(g.position.x, g.position.y, g.position.z, 0.f);
                    v[V_NORMAL] = make_float4(g.normal.x, g.normal.y, g.normal.z, b.exponent);
                    v[V_WO] = make_float4(g.wo.x, g.wo.y, g.wo.z, 0.f);
                    v[V_COLOR] = make_float4(b.a.x, b.a.y, b.a.z, 0.f);
                    v[V_RNG] = pack_rng(smp.state);
                    v[V_PAD] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                else
                {
                    Sampler ls = smp;
                    for (int l = 0; l < n_lights; ++l)
                    {
                        light_sample_pair(wp, w, g, b, l, slot, ls);
                        ls.skip(4 + ((wp.rp.direct_sample == KYD_DS_BSDF && !light_is_delta(c_scene.lights[l].kind)) ? 2 : 0));
                    }
                }
                new_pending = n_lights;
                *out_nee = true;
            }
            // sample_all_light draws 4 floats per light, plus 2 per non-delta light in `bsdf` mode
            smp.skip(4 * n_lights + (wp.rp.direct_sample == KYD_DS_BSDF ? 2 * c_scene.n_nondelta_lights : 0));
        }

        if (!wp.direct_only)
        {
            BsdfSample bs = bsdf_sample(b, g.wo, smp.get_float2());
            if (!(is_black(bs.f) || bs.pdf == 0.f))
            {
                float3 nb = cmulc(beta, cdiv(mul(bs.f, abs_dot(bs.wi, g.normal)), bs.pdf));
                Ray nr = spawn_ray(g, bs.wi);
                bool alive = true;
                if (bounce > 3)
                {
                    float q = max_std(0.05f, 1 - max_component(nb));
                    if (smp.get_float() < q)
                        alive = false;
                    else
                        nb = mul(nb, 1 / (1 - q));
                }
                if (alive)
                {
                    o4 = make_float4(nr.o.x, nr.o.y, nr.o.z, nr.tmax);
                    d4 = make_float4(nr.d.x, nr.d.y, nr.d.z, __int_as_float((bs.type & BSDF_SPECULAR) ? FLAG_PREV_SPECULAR : 0));
                    next_beta = nb;
                    rng_state = smp.state;
                    *out_alive = true;
                }
            }
        }
    }
    // whole sectors go back: the ray, (beta, Lo), (beta of this vertex for its pending Ld, rng)
    p[P_ORIGIN] = o4;
    p[P_DIRECTION] = d4;
    p[P_BETA] = make_float4(next_beta.x, next_beta.y, next_beta.z, 0.f);
    p[P_RADIANCE] = make_float4(Lo.x, Lo.y, Lo.z, 0.f);
    p[P_VERTEX_BETA] = make_float4(beta.x, beta.y, beta.z, __int_as_float(new_pending));
    p[P_RNG] = pack_rng(rng_state);
}

// gathers one path line (units 0..6) and light 0's pending result into this thread's prefetch column
KYD_DEV void prefetch_path(const WaveParams& wp, const WaveBuffers& w, float4* column, int slot)
{
    const float4* p = path_line(w, slot);
#pragma unroll
    for (int k = 0; k < 7; ++k)
        cp_async16(column + k * SHADE_THREADS, p + k);
    cp_async16(column + 7 * SHADE_THREADS, nee_line(w, wp.plane, 0, slot) + N_RESULT);
}

template <int LOBE>
KYD_DEV void shade_queue(const WaveParams& wp, const WaveBuffers& w, DevCounters* __restrict__ counters, int bounce)
{
    const int parity = bounce & 1;
    const int n = (int)counters->queue[Q_LOBE0 + 4 * parity + LOBE];
    const int* __restrict__ queue = w.queue_lobe[LOBE];
    int* __restrict__ next_queue = parity ? w.queue_a : w.queue_b;
    const int stride = gridDim.x * blockDim.x;
    const int n_lights = c_scene.n_lights;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    // queue 0: the next bounce's rays; queue 1: vertices whose light queries the shadow stage resolves
    unsigned long long* const tails[2] = { &counters->queue[Q_RAY0 + (parity ^ 1)], &counters->queue[Q_NEE0 + (LOBE == LOBE_PHONG)] };
    int* const out_queues[2] = { next_queue, w.queue_nee[LOBE == LOBE_PHONG] };
    WarpPush<2> push;
    push.init();

#if !defined(KYD_PREFETCH) || !KYD_PREFETCH
    // default: the line is copied synchronously (KYD_PREFETCH=1 selects the cp.async double-buffered gather
    // below, which measured 6 % slower: profiles/r01_ab_variants.txt)
    for (int i0 = i - (threadIdx.x & 31); i0 < n; i0 += stride, i += stride)
    {
        bool alive = false, wants_nee = false;
        int slot = 0;
        if (i < n)
        {
            slot = queue[i];
            shade_vertex<LOBE, 1>(wp, w, path_line(w, slot), nee_line(w, wp.plane, 0, slot) + N_RESULT, slot, bounce, n_lights, &alive, &wants_nee);
        }
        push.commit(out_queues);
        push.reserve((alive ? 1u : 0u) | (wants_nee ? 2u : 0u), slot, tails);
    }
    push.commit(out_queues);
    return;
#else
    // [buffer][unit][thread]: consecutive threads touch consecutive 16-byte words (no bank conflicts); a
    // thread only ever reads the column it filled itself, so cp.async.wait_group is all the ordering needed
    __shared__ float4 s_pre[2][8][SHADE_THREADS];
    int slot_cur = i < n ? queue[i] : -1;
    int slot_next = (i + stride < n && i + stride >= 0) ? queue[i + stride] : -1;
    if (slot_cur >= 0)
        prefetch_path(wp, w, &s_pre[0][0][threadIdx.x], slot_cur);
    cp_async_commit();

    // whole warps iterate together so that the ballots in queue_push are convergent
    int buf = 0;
    for (int i0 = i - (threadIdx.x & 31); i0 < n; i0 += stride, i += stride, buf ^= 1)
    {
        if (slot_next >= 0)
            prefetch_path(wp, w, &s_pre[buf ^ 1][0][threadIdx.x], slot_next);
        cp_async_commit();
        const long long i2 = (long long)i + 2ll * stride;
        const int slot_next2 = i2 < n ? queue[i2] : -1;
        cp_async_wait<1>(); // everything but the newest group has landed: this iteration's line is in shared memory

        bool alive = false, wants_nee = false;
        const int slot = slot_cur < 0 ? 0 : slot_cur;
        if (slot_cur >= 0)
            shade_vertex<LOBE, SHADE_THREADS>(wp, w, &s_pre[buf][0][threadIdx.x], &s_pre[buf][7][threadIdx.x], slot, bounce, n_lights, &alive, &wants_nee);
        push.commit(out_queues);
        push.reserve((alive ? 1u : 0u) | (wants_nee ? 2u : 0u), slot, tails);
        slot_cur = slot_next;
        slot_next = slot_next2;
    }
    push.commit(out_queues);
    cp_async_wait<0>();
#endif
}

// one kernel per lobe: each gets the register allocation its own code needs (the Lambert kernel, which
// shades most vertices, does not pay for Phong's pow() or the dielectric's Fresnel terms)
template <int LOBE>
__global__ void __launch_bounds__(SHADE_THREADS, KYD_SHADE_MIN_BLOCKS) k_shade(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters, int bounce)
{
    shade_queue<LOBE>(wp, w, counters, bounce);
}

// ---- light-sample as its own stage (KYD_FLAG_SPLIT_LIGHT_SAMPLE): one thread per (vertex, light) -------------
template <int LOBE>
KYD_DEV void light_sample_queue(const WaveParams& wp, const WaveBuffers& w, DevCounters* __restrict__ counters)
{
    const int n = (int)counters->queue[Q_NEE0 + (LOBE == LOBE_PHONG)];
    const int* __restrict__ nee_queue = w.queue_nee[LOBE == LOBE_PHONG];
    const int n_lights = c_scene.n_lights;
    const long long total = (long long)n * n_lights;
    const int stride = gridDim.x * blockDim.x;
    for (long long idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
    {
        const int l = (int)(idx / n);            // light-major: a warp works on one light
        const int slot = nee_queue[idx - (long long)l * n];
        const float4* v = vertex_line(w, slot);
        float4 p4 = v[V_POSITION], n4 = v[V_NORMAL], wo4 = v[V_WO], c4 = v[V_COLOR], rng4 = v[V_RNG];
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
        b.lobe = LOBE;

        Sampler smp;
        smp.debug = (wp.rp.sampler == KYD_SAMPLER_DEBUG);
        smp.state = unpack_rng(rng4);
        smp.skip(light_draw_offset(l, wp.rp.direct_sample));
        light_sample_pair(wp, w, g, b, l, slot, smp);
    }
}

__global__ void __launch_bounds__(128) k_light_sample(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters)
{
    light_sample_queue<LOBE_LAMBERT>(wp, w, counters);
    light_sample_queue<LOBE_PHONG>(wp, w, counters);
}

// ---- shadow: the scene queries of the light loop and the estimators' second halves ------------------------
// closest-hit query for the BSDF-sampled direction, occlusion query for the light-sampled point; writes
// the estimator value of (vertex, light): Lb, Ll or 0.5 Lb + 0.5 Ll (ky.cpp:4083)
__global__ void __launch_bounds__(256) k_shadow(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters)
{
    const int n_lights = c_scene.n_lights;
    const int stride = gridDim.x * blockDim.x;
    const int ds = wp.rp.direct_sample;
    unsigned rays = 0, traced = 0;
    for (int c = 0; c < 2; ++c)
    {
        const int n = (int)counters->queue[Q_NEE0 + c];
        const int* __restrict__ nee_queue = w.queue_nee[c];
        const long long total = (long long)n * n_lights;
        for (long long idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
        {
            const int l = (int)(idx / n);
            const int slot = nee_queue[idx - (long long)l * n];
            float4* line = nee_line(w, wp.plane, l, slot);

            float3 Lb = KYD_BLACK, Ll = KYD_BLACK;
            {
                float4 o = line[N_BSDF_O], d = line[N_BSDF_D], v = line[N_BSDF_VALUE];
                rays += (unsigned)(__float_as_int(d.w) & 1);
                if (o.w >= 0.f)
                {
                    NeeRay q;
                    q.ray.o = V3(o.x, o.y, o.z);
                    q.ray.d = V3(d.x, d.y, d.z);
                    q.ray.tmax = o.w;
                    q.value = V3(v.x, v.y, v.z);
                    q.light = l;
                    float t;
                    int s = scene_closest(q.ray, &t);
                    Lb = nee_bsdf_resolve(q, s, t);
                    traced++;
                }
            }
            {
                float4 o = line[N_LIGHT_O], d = line[N_LIGHT_D], v = line[N_LIGHT_VALUE];
                rays += (unsigned)(__float_as_int(d.w) & 1);
                if (o.w >= 0.f)
                {
                    Ray r;
                    r.o = V3(o.x, o.y, o.z);
                    r.d = V3(d.x, d.y, d.z);
                    r.tmax = o.w;
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
            line[N_RESULT] = make_float4(e.x, e.y, e.z, 0.f);
            line[N_RESULT_PAD] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    flush_counters(rays, traced, counters);
}

// ---- accumulate: film_t::add_color order -- L = L + Li * (1/spp) sample after sample (ky.cpp:3717-3721) ------
__global__ void __launch_bounds__(256) k_accumulate(WaveParams wp, WaveBuffers w, float* __restrict__ film)
{
    const int stride = gridDim.x * blockDim.x;
    for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < wp.npix; px += stride)
    {
        float* o = film + 3 * (size_t)(wp.pixel_begin + px);
        float3 L = V3(o[0], o[1], o[2]);
        for (int s = 0; s < wp.nspp; ++s)
        {
            const int slot = s * wp.npix + px;
            const float4* p = path_line(w, slot);
            float4 L4 = p[P_RADIANCE];
            float3 Li = add_pending(w, wp.plane, slot, p[P_VERTEX_BETA], V3(L4.x, L4.y, L4.z));
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
