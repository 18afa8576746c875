// kyd_wavefront.cuh -- the wavefront organisation of path_tracing_iteration_t / direct_lighting_t
// (reference ky.cpp:4523-4618, 4125-4155): intersect (the first one generates the camera rays), shade (BSDF sample), light-sample (NEE + MIS
// set-up), shadow (NEE ray queries) and accumulate kernels over path state in HBM, with warp-ballot
// compacted queues of path slots between them.  Included by kyd_kernels.cu only.
//
// A wave is a tile of `npix` consecutive pixels times `nspp` consecutive sample indices; path slot
// = s_local * npix + pixel_local.  Per-slot results do not depend on how the film is cut into waves.
//
// Divergence control (profiles/r01_wave1M_*: the first version ran shade at 11 of 32 lanes): the intersect
// stage sorts the paths it hit by the BSDF lobe they will be shaded with (Lambert / Phong / mirror / glass;
// a plastic surface picks its lobe with a hash of the hit, so the lobe is known there) into one queue per
// lobe.  Shade runs once per lobe with lobe-specialised code, so a warp never mixes lobes.  Paths that miss
// the scene are finished inside intersect.
//
// Data layout.  Lobe-sorted queues turn the state accesses into gathers over gigabytes; with 128-byte records the
// shade stage was bound by scattered-DRAM throughput (~2.4 TB/s whether it wrote light-sampling records, 352 B per
// vertex, 103 ms per step, or not, 224 B, 64.5 ms -- profiles/r01_ab_variants.txt).  So the records are as small
// as whole 32-byte sectors allow:
//   path line, 64 bytes:
//     sector 0   origin.xyz, hit distance | direction.xyz, flags (previous vertex specular, pending light
//                count, hit surface)                         intersect reads and rewrites it, shade rewrites it
//     sector 1   beta.rgb, Lo.rgb, rng state (2 words)        shade
//   light-sampling line of a (light, path) pair, 128 bytes of which 64 are normally touched:
//     sector 0   light-sampled query: origin.xyz, tmax | direction.xyz, flags
//     sector 1   its value.rgb, vertex beta.r | vertex beta.gb, BSDF-sampled query's value.rg
//     sector 2   BSDF-sampled query: origin.xyz, tmax | direction.xyz, value.b    (only when that query is
//                live: ~5 % of Cornell vertices, the rest have a zero light pdf)
//     sector 3   the estimator's value.rgb | vertex beta.rgb                      (shadow writes, next shade reads)
//   A vertex none of whose queries can contribute writes no light-sampling line and is not queued for shadow.
//
// Order of FP32 additions into a path's radiance Lo is the reference's: emitted light of a vertex, then
// beta * Ld of that vertex (ky.cpp:4553-4576).  Two ways to get Ld:
//   - several lights (or a non-headline configuration): one light-sampling line per (vertex, light), resolved by the
//     shadow stage with one thread per line.  Ld becomes known one stage later than the vertex is shaded, so it
//     is added ("pending") at the start of the path's next shade (or when the path misses, or by the accumulate
//     kernel if the path ended) -- before anything else is added in every case.
//   - one light in the headline configuration (HOT kernels): the vertex' two scene queries are traced inside
//     shade and Ld is added on the spot; no line, no shadow stage, no pending state (profiles/r01_ab_variants.txt:
//     same shade + shadow time for Cornell, but the line traffic and the accumulate/miss re-reads go away; with
//     five lights the per-line threads of the shadow stage win, so Veach keeps the deferred form).
#pragma once

#include "kyd_device.cuh"
#include "kyd_internal.h"

namespace KYD_KERNEL_NS {

// queue tails in DevCounters::queue
enum { Q_RAY0 = 0, Q_RAY1 = 1, Q_NEE0 = 2 /* +lobe (Lambert, Phong): vertices, split light-sample stage only */,
       Q_LOBE0 = 4 /* + 4 * parity + lobe */, Q_PAIR0 = 12 /* +lobe: (vertex, light) pairs with a live light query */ };

#ifndef KYD_SHADE_MIN_BLOCKS
#define KYD_SHADE_MIN_BLOCKS 4
#endif
#ifndef SHADE_THREADS
#define SHADE_THREADS 128
#endif

// number of lights as a compile-time property of the headline shade kernels: one light = its queries are traced inside
// shade, several = one light-sampling line per (vertex, light) for the shadow stage; NL_ANY decides at run time
enum { NL_ANY = 0, NL_ONE = 1, NL_MANY = 2 };

// float4 units of the records
enum { P_ORIGIN = 0, P_DIRECTION = 1, P_BETA = 2, P_TAIL = 3, PATH_UNITS = 4 };
enum { N_LIGHT_O = 0, N_LIGHT_D = 1, N_LIGHT_VALUE = 2, N_MIXED = 3, N_BSDF_O = 4, N_BSDF_D = 5, N_RESULT = 6, N_VERTEX_BETA = 7, NEE_UNITS = 8 };
enum { V_POSITION = 0, V_NORMAL = 1, V_WO = 2, V_COLOR = 3, V_RNG = 4, V_BETA = 5, VERTEX_UNITS = 6 };

// flags word of a path (w of the direction unit)
// bit 0: previous vertex was specular; bits 4-15: 1 + hit surface; bits 16-31: lights whose estimator value is pending
// (a light-sampling line was written for them at the previous vertex and the shadow stage resolves it)
enum { FLAG_PREV_SPECULAR = 1, FLAG_SURFACE_SHIFT = 4, FLAG_SURFACE_MASK = 0xfff << 4, FLAG_PENDING_SHIFT = 16, FLAG_PENDING_MASK = (int)0xffff0000u };
// entry of the pair queues: path slot | light << 24 (a wave holds at most 2^24 paths, a scene at most 16 lights)
enum { PAIR_LIGHT_SHIFT = 24, PAIR_SLOT_MASK = (1 << 24) - 1 };
// The recursive integrators (wp.recursion) trace their light queries inside shade, so no light is ever pending: the 16 bits
// hold the recursion depth of the vertex (bits 16-23) and, once the path has ended, its number of levels (bits 24-31).
enum { FLAG_DEPTH_SHIFT = 16, FLAG_DEPTH_MASK = 0xff << 16, FLAG_LEVELS_SHIFT = 24 };

// flags word of a light-sampling line (w of the light query's direction unit)
// (bits 8..: 1 + the light's surface when the BSDF-sampled query is in its occlusion form, NeeRay::light_surface)
// (NEE_LIGHT_LIVE: the light-sampled query has to be traced.  Its tmax cannot say so: a light point closer than 2e-3 -- a
// sphere light sampled from a point on its own surface -- gives a NEGATIVE tmax = distance - 2e-3 (ky.cpp:3193), which the
// reference answers "not occluded")
enum { NEE_REF_BSDF = 1, NEE_REF_LIGHT = 2, NEE_BSDF_LIVE = 4, NEE_LIGHT_LIVE = 8, NEE_LIGHT_SURFACE_SHIFT = 8 };

struct WaveParams
{
    RenderParams rp;
    int pixel_begin, npix;    // tile of the film (linear pixel indices)
    int sample_begin, nspp;   // sample indices of this wave
    int nslots;               // npix * nspp
    long long plane;          // lines between the per-light planes of the light-sampling buffer (= capacity)
    int direct_only;          // direct_lighting_t: stop after the first vertex' light loop
    int split_light_sample;   // light-sample as its own kernel (KYD_FLAG_SPLIT_LIGHT_SAMPLE) instead of inside shade
    int no_pending;           // light queries are traced inside shade (one light, headline kernels): no path ever carries pending Ld
    int recursion;            // 0, or the recursive integrator being rendered (KYD_INT_*_RECURSION*): per-level records + k_unwind
    int pair_kernel;          // headline configuration with several lights: the light loop runs in k_nee, one thread per (vertex, light);
                              // results are 16 bytes per pair at [slot * n_lights + light] of the light-sampling buffer, every light of
                              // a vertex has one, and the vertex' beta is read from its vertex record
};

KYD_DEV float4* path_line(const WaveBuffers& w, int slot) { return w.path + (size_t)slot * PATH_UNITS; }
KYD_DEV float4* nee_line(const WaveBuffers& w, long long plane, int light, int slot) { return w.nee + ((size_t)light * plane + slot) * NEE_UNITS; }
KYD_DEV float4* vertex_line(const WaveBuffers& w, int slot) { return w.vertex + (size_t)slot * VERTEX_UNITS; }
// k_nee's per-vertex summaries {beta * sum over lights, complete}: behind the n_lights results of every slot (wp.pair_kernel == 2)
KYD_DEV float4* nee_summary(const WaveBuffers& w, long long plane, int n_lights) { return w.nee + (size_t)plane * n_lights; }
// record of one recursion level of a path (recursive integrators): {Lo.rgb, |cos|}, {f.rgb, pdf}; level-major planes
KYD_DEV float4* level_line(const WaveBuffers& w, long long plane, int level, int slot) { return w.levels + ((size_t)level * plane + slot) * 2; }

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

// Pushes into up to NQ queues with ONE atomic instruction per warp and iteration (lane q reserves queue q's
// range) whose latency is hidden: reserve() issues it and commit(), called one loop iteration later, reads
// the returned bases and writes the entries.  History (profiles/r01_wave_lines_stalls.txt): one atomic per
// queue with an immediate shuffle -> 61 % of k_intersect's stall samples were lanes waiting for atomics; NQ
// predicated atomics into one destination register still serialise on that register; per-lane addresses in
// one instruction do not.
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
        __syncwarp(); // all 32 lanes arrive together: a diverged warp would execute the ballots group by group
        preds = preds_;
        value = value_;
        const int lane = threadIdx.x & 31;
        unsigned my_count = 0;
        unsigned long long* my_tail = tails[0];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
        {
            masks[q] = __ballot_sync(0xffffffffu, (preds_ >> q) & 1u);
            if (lane == q)
            {
                my_count = __popc(masks[q]);
                my_tail = tails[q];
            }
        }
        base = 0;
        if (lane < NQ && my_count != 0)
            base = atomicAdd(my_tail, (unsigned long long)my_count);
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

// Pushes a variable number of (vertex, light) pairs per lane -- one entry per set bit of `mask` -- with one atomic per
// warp and iteration, deferred like WarpPush: a warp scan of the per-lane counts, lane 0 reserves the warp's range.
struct PairPush
{
    unsigned mask;             // this lane's lights of the iteration whose entries are not written yet
    int slot;
    unsigned offset;           // entries of lower lanes
    unsigned long long base;   // lane 0: first index reserved
    bool pending;              // warp-uniform

    KYD_DEV void init() { pending = false; mask = 0; slot = 0; offset = 0; base = 0; }

    KYD_DEV void reserve(unsigned mask_, int slot_, unsigned long long* tail)
    {
        __syncwarp();
        pending = __any_sync(0xffffffffu, mask_ != 0u);
        if (!pending)
            return;
        mask = mask_;
        slot = slot_;
        const int lane = threadIdx.x & 31;
        const unsigned count = __popc(mask_);
        unsigned inclusive = count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const unsigned v = __shfl_up_sync(0xffffffffu, inclusive, d);
            if (lane >= d) inclusive += v;
        }
        offset = inclusive - count;
        const unsigned total = __shfl_sync(0xffffffffu, inclusive, 31);
        base = 0;
        if (lane == 0)
            base = atomicAdd(tail, (unsigned long long)total);
    }

    KYD_DEV void commit(int* queue)
    {
        if (!pending)
            return;
        __syncwarp();
        unsigned long long idx = __shfl_sync(0xffffffffu, base, 0) + offset;
        for (unsigned m = mask; m != 0u; m &= m - 1u)
            queue[idx++] = slot | ((__ffs(m) - 1) << PAIR_LIGHT_SHIFT);
        pending = false;
    }
};

// ---- path record ------------------------------------------------------------------------------------------
struct PathState
{
    float3 o, d;
    float t;                    // hit distance (after intersect)
    int flags;
    float3 beta, Lo;
    unsigned long long rng;

    KYD_DEV unsigned pending() const { return (unsigned)flags >> FLAG_PENDING_SHIFT; }
    KYD_DEV int surface() const { return ((flags & FLAG_SURFACE_MASK) >> FLAG_SURFACE_SHIFT) - 1; }
};

KYD_DEV void unpack_path(PathState& s, float4 u0, float4 u1, float4 u2, float4 u3)
{
    s.o = V3(u0.x, u0.y, u0.z); s.t = u0.w;
    s.d = V3(u1.x, u1.y, u1.z); s.flags = __float_as_int(u1.w);
    s.beta = V3(u2.x, u2.y, u2.z);
    s.Lo = V3(u2.w, u3.x, u3.y);
    s.rng = (unsigned long long)__float_as_uint(u3.z) | ((unsigned long long)__float_as_uint(u3.w) << 32);
}

KYD_DEV void store_path_ray(float4* p, float3 o, float t, float3 d, int flags)
{
    p[P_ORIGIN] = make_float4(o.x, o.y, o.z, t);
    p[P_DIRECTION] = make_float4(d.x, d.y, d.z, __int_as_float(flags));
}

KYD_DEV void store_path_tail(float4* p, float3 beta, float3 Lo, unsigned long long rng)
{
    p[P_BETA] = make_float4(beta.x, beta.y, beta.z, Lo.x);
    p[P_TAIL] = make_float4(Lo.y, Lo.z, __uint_as_float((unsigned)rng), __uint_as_float((unsigned)(rng >> 32)));
}

// Lo += beta_vertex * (sum over lights of the vertex' estimator values), ky.cpp:4575-4576.  `pending` = the lights that
// got a light-sampling line; the others' values are exactly +0 (no query of theirs could contribute) and adding +0 to
// the running sum, which starts at +0 and therefore is never -0, changes nothing -- so they are skipped, in light order.
KYD_DEV float3 add_pending(const WaveBuffers& w, long long plane, int slot, unsigned pending, float3 Lo, int pair_kernel = 0)
{
    if (pending != 0 && pair_kernel == 2)
    {
        // k_nee (light-major form) leaves beta_vertex * (the sum below) behind for every vertex whose pairs it finished in light
        // order: one 32-byte sector instead of the lights' values and the vertex record's beta
        const float4 sum = nee_summary(w, plane, c_scene.n_lights)[slot];
        if (sum.w != 0.f)
            return add(Lo, V3(sum.x, sum.y, sum.z));
    }
    if (pending != 0 && pair_kernel)
    {
        // k_nee's results: every light of the vertex, in light order -- sample_all_light's own sum (ky.cpp:3864-3869)
        const int n_lights = c_scene.n_lights;
        const float4* res = w.nee + (size_t)slot * n_lights;
        const float4 vb = vertex_line(w, slot)[V_BETA];
        float3 Ld = KYD_BLACK;
        for (int l = 0; l < n_lights; ++l)
        {
            const float4 e = res[l];
            Ld = add(Ld, V3(e.x, e.y, e.z));
        }
        return add(Lo, cmulc(V3(vb.x, vb.y, vb.z), Ld));
    }
    if (pending != 0)
    {
        const float4* line0 = nee_line(w, plane, __ffs(pending) - 1, slot);
        float4 e0 = line0[N_RESULT], vb = line0[N_VERTEX_BETA];
        float3 Ld = add(KYD_BLACK, V3(e0.x, e0.y, e0.z));
        for (unsigned rest = pending & (pending - 1); rest != 0; rest &= rest - 1)
        {
            float4 e = nee_line(w, plane, __ffs(rest) - 1, slot)[N_RESULT];
            Ld = add(Ld, V3(e.x, e.y, e.z));
        }
        Lo = add(Lo, cmulc(V3(vb.x, vb.y, vb.z), Ld));
    }
    return Lo;
}

// the lobe material_t::scattering will build for this hit (ky.cpp:2587-2671); pure function of the hit
template <bool TABLE = false>
KYD_DEV int classify_lobe(int surface, const Ray& r, float t)
{
    const DevMaterial& m = surface_material_of<TABLE>(surface);
    if (m.kind == KYD_MAT_MATTE) return LOBE_LAMBERT;
    if (m.kind == KYD_MAT_MIRROR) return LOBE_MIRROR;
    if (m.kind == KYD_MAT_GLASS) return LOBE_FRESNEL;
    return plastic_random(ray_at(r, t), neg(r.d)) < m.p_specular ? LOBE_PHONG : LOBE_LAMBERT;
}

// ---- intersect: scene_t::intersect closest-hit query for every queued path (ky.cpp:3172-3184) -------------
// Hits go to the queue of their lobe; a miss ends the path here (ky.cpp:4555-4563).
// CAMERA: first bounce of a wave.  The queue is 0..nslots-1 and the rays are generated right here (camera_t::generate_ray,
// ky.cpp:3714-3715, 1884-1892) instead of being written by a raygen kernel and read back; both sectors of the record
// are written once, with the hit.  The host zeroes the queue tails before this launch.
#ifndef KYD_INTERSECT_MIN_BLOCKS
#define KYD_INTERSECT_MIN_BLOCKS 3   // (80 registers: 24 warps per SM; left to itself the compiler takes 144 and one block per SM)
#endif
template <bool CAMERA>
__global__ void __launch_bounds__(256, KYD_INTERSECT_MIN_BLOCKS) k_intersect(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters, int bounce)
{
    stage_rects();
    const int parity = bounce & 1;
    const int n = CAMERA ? wp.nslots : (int)counters->queue[Q_RAY0 + parity];
    if (!CAMERA && blockIdx.x == 0 && threadIdx.x < 9)
    {
        // tails that later kernels of this bounce push to; their previous contents were consumed by earlier kernels
        const int which[9] = { Q_RAY0 + (parity ^ 1), Q_NEE0, Q_NEE0 + 1, Q_LOBE0 + 4 * (parity ^ 1), Q_LOBE0 + 4 * (parity ^ 1) + 1,
                               Q_LOBE0 + 4 * (parity ^ 1) + 2, Q_LOBE0 + 4 * (parity ^ 1) + 3, Q_PAIR0, Q_PAIR0 + 1 };
        counters->queue[which[threadIdx.x]] = 0;
    }
    const int* __restrict__ queue = parity ? w.queue_b : w.queue_a;
    unsigned long long* const tails[4] = { &counters->queue[Q_LOBE0 + 4 * parity], &counters->queue[Q_LOBE0 + 4 * parity + 1],
                                           &counters->queue[Q_LOBE0 + 4 * parity + 2], &counters->queue[Q_LOBE0 + 4 * parity + 3] };
    int* const lobe_queues[4] = { w.queue_lobe[parity][0], w.queue_lobe[parity][1], w.queue_lobe[parity][2], w.queue_lobe[parity][3] };
    WarpPush<4> push;
    push.init();
    const bool has_env = c_scene.env_light >= 0;
    const int stride = gridDim.x * blockDim.x;
    unsigned rays = 0;
    const int base_i = blockIdx.x * blockDim.x + threadIdx.x;
    // software pipeline of the gather: the queue entry two iterations ahead and the ray one iteration ahead are
    // loaded before the current ray is traversed, so their latency hides behind ~1000 instructions of traversal
    long long ia = base_i;
    int slot_cur = ia < n ? (CAMERA ? (int)ia : queue[ia]) : -1;
    int slot_next = ia + stride < n ? (CAMERA ? (int)(ia + stride) : queue[ia + stride]) : -1;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = o;
    if (!CAMERA && slot_cur >= 0)
    {
        const float4* p0 = path_line(w, slot_cur);
        o = p0[P_ORIGIN];
        d = p0[P_DIRECTION];
    }
    // block-uniform trip count (the tail costs at most one idle iteration per warp): keeps the loop itself uniform
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += stride, ia += stride)
    {
        int lobe = -1;
        const int slot = slot_cur < 0 ? 0 : slot_cur;
        const long long i2 = ia + 2ll * stride;
        const int slot_next2 = i2 < n ? (CAMERA ? (int)i2 : queue[i2]) : -1;
        float4 o_next = make_float4(0.f, 0.f, 0.f, 0.f), d_next = o_next;
        if (!CAMERA && slot_next >= 0)
        {
            const float4* pn = path_line(w, slot_next);
            o_next = pn[P_ORIGIN];
            d_next = pn[P_DIRECTION];
        }
        unsigned long long rng_state = 0;
        if (CAMERA && slot_cur >= 0)
        {
            const int pixel = wp.pixel_begin + slot % wp.npix;
            const int sample = wp.sample_begin + slot / wp.npix;
            const int x = pixel % wp.rp.width, y = pixel / wp.rp.width;
            Sampler smp;
            smp.start(wp.rp.sampler, wp.rp.seed, x, y, sample);
            const float2 jitter = smp.camera_jitter(wp.rp.sampler, wp.rp.spp, sample);
            const Ray cam = generate_ray((float)x + jitter.x, (float)y + jitter.y);
            o = make_float4(cam.o.x, cam.o.y, cam.o.z, cam.tmax);
            d = make_float4(cam.d.x, cam.d.y, cam.d.z, __int_as_float(0));
            rng_state = smp.state;
        }
        // the traversal runs for the whole warp, outside any divergent branch (idle lanes trace a null ray that hits
        // nothing): its loop counters and shape loads then stay in the uniform datapath
        Ray r;
        r.o = V3(o.x, o.y, o.z);
        r.d = V3(d.x, d.y, d.z);
        r.tmax = slot_cur >= 0 ? KYD_INF : -1.f; // extension rays are unbounded (ky.cpp:585, 665-668)
        float t;
        const int s = wf_closest(r, &t);
        if (slot_cur >= 0)
        {
            float4* p = path_line(w, slot);
            const int flags = __float_as_int(d.w);
            rays++;
            float3 Lo_camera = KYD_BLACK;
            if (s >= 0)
            {
                // sector 0 goes back whole, now carrying the hit
                store_path_ray(p, r.o, t, r.d, (flags & (FLAG_PREV_SPECULAR | FLAG_PENDING_MASK)) | ((s + 1) << FLAG_SURFACE_SHIFT));
                lobe = classify_lobe(s, r, t);
            }
            else if (wp.recursion)
            {
                // a ray of a recursive integrator that leaves the scene ends the recursion: this level returns the environment's
                // radiance -- always (simple_path_tracing_recursion_t, ky.cpp:4207-4208), or where the level adds emitted
                // light at all (depth 0, or after a specular bounce in the deferred form; ky.cpp:4326-4336, 4445-4455)
                const int depth = CAMERA ? 0 : (flags & FLAG_DEPTH_MASK) >> FLAG_DEPTH_SHIFT;
                float3 Lo_level = environment_lighting();
                if (wp.recursion != KYD_INT_SIMPLE_PT_RECURSION)
                {
                    const bool deferred = wp.recursion == KYD_INT_PT_RECURSION_DEFERED;
                    const bool adds = depth == 0 || (deferred && (flags & FLAG_PREV_SPECULAR));
                    const bool keep = !deferred || (depth == 0 ? (wp.rp.lighting & KYD_LIGHTING_EMIT) : (wp.rp.lighting & KYD_LIGHTING_INDIRECT)) != 0;
                    Lo_level = adds ? add(KYD_BLACK, keep ? Lo_level : KYD_BLACK) : KYD_BLACK;
                }
                float4* lv = level_line(w, wp.plane, depth, slot);
                lv[0] = make_float4(Lo_level.x, Lo_level.y, Lo_level.z, 0.f);
                lv[1] = make_float4(0.f, 0.f, 0.f, 1.f);
                store_path_ray(p, r.o, o.w, r.d, (depth + 1) << FLAG_LEVELS_SHIFT);
            }
            else if (CAMERA)
            {
                // a camera ray that leaves the scene: the sample is the environment's radiance (beta = 1) or black
                if (has_env)
                    Lo_camera = add(KYD_BLACK, cmulc(V3(1.f, 1.f, 1.f), environment_lighting()));
                store_path_ray(p, r.o, o.w, r.d, 0);
            }
            else if (has_env && (flags & FLAG_PREV_SPECULAR))
            {
                // Lo += beta * environment_lighting after a specular bounce
                PathState st;
                unpack_path(st, o, d, p[P_BETA], p[P_TAIL]);
                float3 Lo = add_pending(w, wp.plane, slot, st.pending(), st.Lo, wp.pair_kernel);
                Lo = add(Lo, cmulc(st.beta, environment_lighting()));
                store_path_ray(p, r.o, o.w, r.d, flags & FLAG_PREV_SPECULAR); // pending consumed
                store_path_tail(p, st.beta, Lo, st.rng);
            }
            if (CAMERA)
                store_path_tail(p, V3(1.f, 1.f, 1.f), Lo_camera, rng_state);
        }
        push.commit(lobe_queues);                                  // the previous iteration's entries
        push.reserve(lobe >= 0 ? (1u << lobe) : 0u, slot, tails);  // this iteration's atomic, consumed next time round
        slot_cur = slot_next;
        slot_next = slot_next2;
        o = o_next;
        d = d_next;
    }
    push.commit(lobe_queues);
    flush_counters(rays, rays, counters);
    rays = __reduce_add_sync(0xffffffffu, rays);
    if ((threadIdx.x & 31) == 0 && rays)
        atomicAdd(&counters->intersect_rays, (unsigned long long)rays);
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

// NEE + MIS set-up of one (vertex, light): the two draws and the first halves of the estimators
// (ky.cpp:3864-3869, 3889-4074)
template <int TRAITS>
KYD_DEV void light_sample_pair(int ds, const HitGeom& g, const Bsdf& b, int l, Sampler smp, NeeRay* qb, NeeRay* ql)
{
    float2 random_bsdf = smp.get_float2();
    float2 random_light = smp.get_float2();

    qb->active = ql->active = false;
    qb->ref_query = ql->ref_query = false;
    qb->value = ql->value = KYD_BLACK;
    qb->ray.o = qb->ray.d = ql->ray.o = ql->ray.d = V3(0, 0, 0);
    qb->ray.tmax = ql->ray.tmax = -1.f;
    qb->light = ql->light = l;
    qb->light_surface = ql->light_surface = -1;
    if (ds == KYD_DS_BSDF)
    {
        if (!light_is_delta(light_kind<TRAITS>(c_scene.lights[l])))
            *qb = nee_bsdf_setup<TRAITS>(g, b, l, smp.get_float2(), false);
    }
    else if (ds == KYD_DS_BSDF_MIS || ds == KYD_DS_BOTH_MIS)
        *qb = nee_bsdf_setup<TRAITS>(g, b, l, random_bsdf, true);
    if (ds == KYD_DS_LIGHT)
        *ql = nee_light_setup<TRAITS>(g, b, l, random_light, false);
    else if (ds == KYD_DS_LIGHT_MIS || ds == KYD_DS_BOTH_MIS)
        *ql = nee_light_setup<TRAITS>(g, b, l, random_light, true);
}

// writes the light-sampling line of (light, path): sectors 0-1 always, sector 2 only for a live BSDF-sampled query
KYD_DEV void store_nee_line(float4* line, const NeeRay& qb, const NeeRay& ql, float3 vertex_beta)
{
    const int flags = (qb.ref_query ? NEE_REF_BSDF : 0) | (ql.ref_query ? NEE_REF_LIGHT : 0) | (qb.active ? NEE_BSDF_LIVE : 0) |
                      (ql.active ? NEE_LIGHT_LIVE : 0) | ((qb.light_surface + 1) << NEE_LIGHT_SURFACE_SHIFT);
    line[N_LIGHT_O] = make_float4(ql.ray.o.x, ql.ray.o.y, ql.ray.o.z, ql.ray.tmax);
    line[N_LIGHT_D] = make_float4(ql.ray.d.x, ql.ray.d.y, ql.ray.d.z, __int_as_float(flags));
    line[N_LIGHT_VALUE] = make_float4(ql.value.x, ql.value.y, ql.value.z, vertex_beta.x);
    line[N_MIXED] = make_float4(vertex_beta.y, vertex_beta.z, qb.value.x, qb.value.y);
    if (qb.active)
    {
        line[N_BSDF_O] = make_float4(qb.ray.o.x, qb.ray.o.y, qb.ray.o.z, qb.ray.tmax);
        line[N_BSDF_D] = make_float4(qb.ray.d.x, qb.ray.d.y, qb.ray.d.z, qb.value.z);
    }
}

// what shade did, for the statistics: reference-equivalent scene queries it accounted for
struct ShadeCounts { unsigned ref_rays, traced; };


// the scene queries of one (vertex, light) and the estimator's second half, inside shade: closest-hit query for the
// BSDF-sampled direction, occlusion query for the light-sampled point; Lb, Ll or 0.5 Lb + 0.5 Ll (ky.cpp:4083)
KYD_DEV float3 nee_resolve_pair(int ds, const NeeRay& qb, const NeeRay& ql, ShadeCounts* counts)
{
    float3 Lb = KYD_BLACK, Ll = KYD_BLACK;
    if (qb.active)
    {
        if (qb.light_surface >= 0)
            Lb = wf_blocked_before(qb.ray, qb.light_surface) ? KYD_BLACK : qb.value;
        else
        {
            float t;
            const int s = wf_closest(qb.ray, &t);
            Lb = nee_bsdf_resolve(qb, s, t);
        }
        counts->traced++;
    }
    if (ql.active)
    {
        Ll = wf_any_hit(ql.ray) ? KYD_BLACK : ql.value;
        counts->traced++;
    }
    if (ds == KYD_DS_BOTH_MIS)
        return add(mul(Lb, 0.5f), mul(Ll, 0.5f));
    return (ds == KYD_DS_BSDF || ds == KYD_DS_BSDF_MIS) ? Lb : Ll;
}

// what shading a vertex leaves behind: the path's state after it (shade_queue writes it back, and -- fused configuration --
// first traces the next ray)
struct VertexOut
{
    bool alive;                 // the path continues with the ray (o, d)
    float3 o, d;
    int flags;                  // FLAG_PREV_SPECULAR | pending lights of this vertex
    float3 beta, Lo;
    unsigned long long rng;
    unsigned pairs;             // lights that got a light-sampling line (deferred queries)
    bool split_vertex;          // a vertex record was written for the stand-alone light-sample stage
    bool traced;                // the next ray's closest hit is already known (hit_surface, hit_t)
    int hit_surface;
    float hit_t;
};

// ---- shade: one path vertex (ky.cpp:4545-4613), specialised by lobe ---------------------------------------
// HOT: the headline configuration (path_tracing_iteration_t, both_mis, LCG48 sampler, light-sample inside shade)
// with those run-time switches compiled out; !HOT reads them from the parameters
template <int LOBE, int TRAITS, bool HOT, int NL>
KYD_DEV void shade_vertex(const WaveParams& wp, const WaveBuffers& w, int slot, int bounce, int n_lights, VertexOut* out,
                          ShadeCounts* counts, float4 rec0, float4 rec1, float4 rec2, float4 rec3)
{
    bool alive_flag = false, split_flag = false;
    unsigned pairs_value = 0;
    bool* out_alive = &alive_flag;
    bool* out_split_vertex = &split_flag;
    unsigned* out_pairs = &pairs_value;
    const int ds = HOT ? (int)KYD_DS_BOTH_MIS : wp.rp.direct_sample;
    const bool direct_only = HOT ? false : wp.direct_only != 0;
    // the light loop as a kernel of its own over (vertex, light) pairs: the measurement mode KYD_FLAG_SPLIT_LIGHT_SAMPLE, and
    // -- always -- the headline configuration with several lights (k_nee)
    const bool split_light_sample = HOT ? (NL == NL_MANY) : wp.split_light_sample != 0;
    const bool debug_sampler = HOT ? false : wp.rp.sampler == KYD_SAMPLER_DEBUG;
    PathState st;
    unpack_path(st, rec0, rec1, rec2, rec3);
    Ray r;
    r.o = st.o;
    r.d = st.d;
    r.tmax = KYD_INF;
    const float3 beta = st.beta;
    const int surface = st.surface();

    // light gathered at the previous vertex (see file header)
    float3 Lo = add_pending(w, wp.plane, slot, st.pending(), st.Lo, wp.pair_kernel);
    unsigned new_pending = 0;   // lights that get a light-sampling line at this vertex

    HitGeom g = shape_hit_geom(surface_shape(surface), r, st.t);

    if (bounce == 0 || (st.flags & FLAG_PREV_SPECULAR))
        Lo = add(Lo, cmulc(beta, surface_emission(surface, g)));

    float3 next_o = st.o, next_d = st.d, next_beta = beta;
    int next_flags = 0;
    unsigned long long rng_state = st.rng;
    if (bounce < wp.rp.max_depth + (direct_only ? 1 : 0))
    {
        const DevMaterial& m = c_scene.materials[surface_material(surface)];
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
        smp.debug = debug_sampler;
        smp.state = rng_state;

        if (LOBE == LOBE_LAMBERT || LOBE == LOBE_PHONG)
        {
            if (ds != KYD_DS_IDLE && n_lights > 0)
            {
                if (split_light_sample)
                {
                    // vertex record for the light-sample stage
                    float4* v = vertex_line(w, slot);
                    // (the spare words carry the shading frame's s and t axes: k_nee's five threads per vertex need not rebuild them)
                    v[V_POSITION] = make_float4(g.position.x, g.position.y, g.position.z, b.f.s.x);
                    v[V_NORMAL] = make_float4(g.normal.x, g.normal.y, g.normal.z, b.exponent);
                    v[V_WO] = make_float4(g.wo.x, g.wo.y, g.wo.z, b.f.s.y);
                    v[V_COLOR] = make_float4(b.a.x, b.a.y, b.a.z, b.f.s.z);
                    v[V_RNG] = make_float4(__uint_as_float((unsigned)smp.state), __uint_as_float((unsigned)(smp.state >> 32)), b.f.t.x, b.f.t.y);
                    v[V_BETA] = make_float4(beta.x, beta.y, beta.z, b.f.t.z);
                    new_pending = (1u << n_lights) - 1u;   // the light-sample kernel writes a line for every light
                    *out_split_vertex = true;
                }
                else if (NL == NL_ONE || (NL == NL_ANY && (TRAITS == TRAITS_AREA_RECTANGLE || n_lights == 1)))
                {
                    // the common case keeps both queries in registers and writes nothing when neither can contribute
                    NeeRay qb, ql;
                    light_sample_pair<TRAITS>(ds, g, b, 0, smp, &qb, &ql);
                    counts->ref_rays += (qb.ref_query ? 1u : 0u) + (ql.ref_query ? 1u : 0u);
                    // (beta * 0 is 0 only for finite beta: a non-finite throughput keeps the reference's NaN)
                    const bool finite_beta = isfinite(beta.x) && isfinite(beta.y) && isfinite(beta.z);
                    if (qb.active || ql.active || !finite_beta)
                    {
                        // HOT: the vertex' two scene queries are traced right here (no light-sampling line, no shadow
                        // stage, no deferred addition): L += beta * Ld as in ky.cpp:4575-4576
                        if (HOT)
                            Lo = add(Lo, cmulc(beta, add(KYD_BLACK, nee_resolve_pair(ds, qb, ql, counts))));
                        else
                        {
                            qb.ref_query = ql.ref_query = false; // counted here
                            store_nee_line(nee_line(w, wp.plane, 0, slot), qb, ql, beta);
                            new_pending = 1u;
                        }
                    }
                }
                else if (NL != NL_ONE && TRAITS != TRAITS_AREA_RECTANGLE)
                {
                    // several lights: one line per (vertex, light) with a query that can contribute; the shadow stage
                    // gives every such pair its own thread
                    Sampler ls = smp;
                    for (int l = 0; l < n_lights; ++l)
                    {
                        NeeRay qb, ql;
                        light_sample_pair<TRAITS>(ds, g, b, l, ls, &qb, &ql);
                        counts->ref_rays += (qb.ref_query ? 1u : 0u) + (ql.ref_query ? 1u : 0u);
                        if (qb.active || ql.active)
                        {
                            qb.ref_query = ql.ref_query = false; // counted here
                            store_nee_line(nee_line(w, wp.plane, l, slot), qb, ql, beta);
                            new_pending |= 1u << l;
                        }
                        ls.skip(4 + ((ds == KYD_DS_BSDF && !light_is_delta(c_scene.lights[l].kind)) ? 2 : 0));
                    }
                    // (beta * 0 is 0 only for finite beta: a non-finite throughput keeps the reference's NaN through an
                    // empty line of light 0, whose value the shadow stage resolves to 0)
                    if (new_pending == 0 && !(isfinite(beta.x) && isfinite(beta.y) && isfinite(beta.z)))
                    {
                        NeeRay none;
                        none.active = none.ref_query = false;
                        none.value = KYD_BLACK;
                        none.ray.o = none.ray.d = V3(0, 0, 0);
                        none.ray.tmax = -1.f;
                        none.light = 0;
                        none.light_surface = -1;
                        store_nee_line(nee_line(w, wp.plane, 0, slot), none, none, beta);
                        new_pending = 1u;
                    }
                }
                if (!split_light_sample)
                    *out_pairs = new_pending;
            }
            // sample_all_light draws 4 floats per light, plus 2 per non-delta light in `bsdf` mode
            if (HOT) smp.skip_lights(n_lights);   // (both_mis: 4 per light, in one step)
            else smp.skip(4 * n_lights + (ds == KYD_DS_BSDF ? 2 * c_scene.n_nondelta_lights : 0));
        }

        if (!direct_only)
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
                    next_o = nr.o;
                    next_d = nr.d;
                    next_flags = (bs.type & BSDF_SPECULAR) ? FLAG_PREV_SPECULAR : 0;
                    next_beta = nb;
                    rng_state = smp.state;
                    *out_alive = true;
                }
            }
        }
    }
    out->alive = alive_flag;
    out->o = next_o;
    out->d = next_d;
    out->flags = next_flags | (int)(new_pending << FLAG_PENDING_SHIFT);
    out->beta = next_beta;
    out->Lo = Lo;
    out->rng = rng_state;
    out->pairs = pairs_value;
    out->split_vertex = split_flag;
    out->traced = false;
    out->hit_surface = -1;
    out->hit_t = KYD_INF;
}

// ---- shade of the headline kernels with one light (Lambert / Phong lobes): the same vertex as shade_vertex<.., HOT, NL_ONE>,
// arranged for code size.  These kernels are instruction-fetch bound (a vertex is ~900 warp instructions of mostly
// straight-line code against a 6 KB L0 / 32 KB L1.5 instruction cache; profiles/r02_ab_variants.txt), so what a vertex does more
// than once runs through ONE copy of its code:
//   - bsdf_t::sample: once for the light loop's BSDF-sampled query (ky.cpp:3977), once for the path's continuation (ky.cpp:4586);
//   - the scene queries: the BSDF-sampled light query (in its occlusion form: the host selects the specialised traits only
//     then), the light-sampled occlusion query and -- FUSE -- the closest hit of the path's next ray are all walks of
//     scene_closest_2p from different starting states.
// A three-trip loop: trip 0 = BSDF-sampled light query, trip 1 = light-sampled query (and L += beta * Ld), trip 2 = continuation.
// The draws are the reference's: random_bsdf, random_light (ky.cpp:3866-3868, g++ order), then the continuation's pair.
template <int LOBE, int TRAITS, bool FUSE>
KYD_DEV void shade_vertex_hot_one(const WaveParams& wp, int bounce, VertexOut* out, ShadeCounts* counts,
                                  float4 rec0, float4 rec1, float4 rec2, float4 rec3, bool whole_block = false)
{
    constexpr bool OCC = TRAITS != TRAITS_ANY;
    PathState st;
    unpack_path(st, rec0, rec1, rec2, rec3);
    Ray r;
    r.o = st.o;
    r.d = st.d;
    r.tmax = KYD_INF;
    const float3 beta = st.beta;
    const int surface = st.surface();
    float3 Lo = st.Lo;   // (no light is ever pending in this configuration)

    constexpr bool TABLE = KYD_HAS_SURFACE_TABLE != 0;   // (shade_queue staged it)
    HitGeom g = shape_hit_geom(surface_shape_of<TABLE>(surface), r, st.t);
    if (bounce == 0 || (st.flags & FLAG_PREV_SPECULAR))
        Lo = add(Lo, cmulc(beta, surface_emission<TABLE>(surface, g)));

    out->alive = false;
    out->o = st.o;
    out->d = st.d;
    out->flags = 0;
    out->beta = beta;
    out->rng = st.rng;
    out->pairs = 0u;
    out->split_vertex = false;
    out->traced = false;
    out->hit_surface = -1;
    out->hit_t = KYD_INF;
    if (bounce < wp.rp.max_depth)
    {
        const DevMaterial& m = surface_material_of<TABLE>(surface);
        Bsdf b;
        b.f = frame_from_z(g.normal);
        b.t = KYD_BLACK;
        b.eta_t = 1.f;
        b.exponent = 0.f;
        b.lobe = LOBE;
        if (LOBE == LOBE_LAMBERT) b.a = m.kind == KYD_MAT_PLASTIC ? m.plastic_lambert : m.diffuse;
        else { b.a = m.plastic_phong; b.exponent = m.exponent; }

        Sampler smp;
        smp.debug = false;
        smp.state = st.rng;
        const float2 random_bsdf = smp.get_float2();
        const float2 random_light = smp.get_float2();
        const float2 random_next = smp.get_float2();
        float3 Lb = KYD_BLACK, Ll = KYD_BLACK;
        bool any_active = false;
#pragma unroll 1
        for (int trip = 0; trip < 3; ++trip)
        {
#if KYD_SHADE_LOCKSTEP & 4
            if (whole_block && trip > 0)
                __syncthreads();   // (every thread of the block shades a vertex in this iteration: see shade_queue)
#endif
            // the ray of this trip, the starting state of its walk, and what its answer is worth
            NeeRay q;
            q.active = false;
            q.ref_query = false;
            q.value = KYD_BLACK;
            q.light = 0;
            q.light_surface = -1;
            q.ray.o = q.ray.d = V3(0.f, 0.f, 0.f);
            q.ray.tmax = -1.f;
            if (trip != 1)
            {
                const BsdfSample bs = bsdf_sample(b, g.wo, trip == 0 ? random_bsdf : random_next);
                if (trip == 0)
                    q = nee_bsdf_from_sample<TRAITS>(g, b, 0, bs, true);      // estimate_direct_lighting_by_bsdf_mis, ky.cpp:3968-4033
                else if (!(is_black(bs.f) || bs.pdf == 0.f))
                {
                    // the path's next vertex (ky.cpp:4586-4613)
                    float3 nb = cmulc(beta, cdiv(mul(bs.f, abs_dot(bs.wi, g.normal)), bs.pdf));
                    bool alive = true;
                    if (bounce > 3)
                    {
                        const float rr = max_std(0.05f, 1 - max_component(nb));
                        if (smp.get_float() < rr)
                            alive = false;
                        else
                            nb = mul(nb, 1 / (1 - rr));
                    }
                    if (alive)
                    {
                        q.ray = spawn_ray(g, bs.wi);
                        out->alive = true;
                        out->o = q.ray.o;
                        out->d = q.ray.d;
                        out->flags = (bs.type & BSDF_SPECULAR) ? FLAG_PREV_SPECULAR : 0;
                        out->beta = nb;
                        out->rng = smp.state;
                        q.active = FUSE;   // (not fused: the intersect kernel of the next bounce traces it)
                    }
                }
            }
            else
                q = nee_light_setup<TRAITS>(g, b, 0, random_light, true);     // estimate_direct_lighting_by_emitter_mis, ky.cpp:4035-4074
            if (trip < 2)
                counts->ref_rays += q.ref_query ? 1u : 0u;

            if (q.active)
            {
                counts->traced++;
                if (!OCC && trip == 0 && q.light_surface < 0)
                {
                    // closest-hit form (a light carried by several surfaces, or the environment light)
                    float t;
                    const int s = wf_closest(q.ray, &t);
                    Lb = nee_bsdf_resolve(q, s, t);
                    any_active = true;
                }
                else
                {
                    float t;
                    const int s = wf_closest_from(q.ray, q.light_surface, &t);
                    if (trip == 0) { any_active = true; if (s == q.light_surface) Lb = q.value; }
                    else if (trip == 1) { any_active = true; if (s < 0) Ll = q.value; }
                    else
                    {
                        counts->ref_rays++;
                        out->traced = true;
                        out->hit_surface = s;
                        out->hit_t = t;
                    }
                }
            }
            if (trip == 1)
            {
                // sample_all_light's sum with its one light: 0.5 Lb + 0.5 Ll (ky.cpp:4083), then L += beta * Ld (ky.cpp:4575-4576)
                // (beta * 0 is 0 only for finite beta: a non-finite throughput keeps the reference's NaN)
                const bool finite_beta = isfinite(beta.x) && isfinite(beta.y) && isfinite(beta.z);
                if (any_active || !finite_beta)
                    Lo = add(Lo, cmulc(beta, add(KYD_BLACK, add(mul(Lb, 0.5f), mul(Ll, 0.5f)))));
            }
        }
    }
    out->Lo = Lo;
}

// FUSE (headline configurations; wavefront_plan().fused): shade also traces the path's next ray -- closest hit,
// lobe classification, the miss -- and pushes the path straight into the NEXT bounce's lobe queue, so only camera rays go
// through k_intersect.  That kernel had become a pure gather / scatter of path records once the two-phase traversal made its
// arithmetic cheap (long-scoreboard bound, 2 TB/s of 32-byte sectors; profiles/r02_*): fused, its 64 B per ray never move.
template <int LOBE, int TRAITS, bool HOT, int NL, bool FUSE>
KYD_DEV void shade_queue(const WaveParams& wp, const WaveBuffers& w, DevCounters* __restrict__ counters, int bounce)
{
    constexpr bool HOT_ONE = HOT && NL == NL_ONE && (LOBE == LOBE_LAMBERT || LOBE == LOBE_PHONG);
    if (HOT_ONE)
        stage_surfaces();
    if (HOT_ONE || FUSE)
        stage_rects();   // this kernel traces scene queries itself
    const int parity = bounce & 1;
    const int n = (int)counters->queue[Q_LOBE0 + 4 * parity + LOBE];
    const int* __restrict__ queue = w.queue_lobe[parity][LOBE];
    int* __restrict__ next_queue = parity ? w.queue_a : w.queue_b;
    const int stride = gridDim.x * blockDim.x;
    const int n_lights = c_scene.n_lights;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // not fused -- queue 0: the next bounce's rays; queue 1: vertices for the stand-alone light-sample stage (split mode only);
    // pair queue: (vertex, light) pairs whose light-sampling line the shadow stage resolves
    unsigned long long* const tails[2] = { &counters->queue[Q_RAY0 + (parity ^ 1)], &counters->queue[Q_NEE0 + (LOBE == LOBE_PHONG)] };
    int* const out_queues[2] = { next_queue, w.queue_nee[LOBE == LOBE_PHONG] };
    // fused: the next bounce's lobe queues
    unsigned long long* const lobe_tails[4] = { &counters->queue[Q_LOBE0 + 4 * (parity ^ 1)], &counters->queue[Q_LOBE0 + 4 * (parity ^ 1) + 1],
                                                &counters->queue[Q_LOBE0 + 4 * (parity ^ 1) + 2], &counters->queue[Q_LOBE0 + 4 * (parity ^ 1) + 3] };
    int* const lobe_queues[4] = { w.queue_lobe[parity ^ 1][0], w.queue_lobe[parity ^ 1][1], w.queue_lobe[parity ^ 1][2], w.queue_lobe[parity ^ 1][3] };
    unsigned long long* const pair_tail = &counters->queue[Q_PAIR0 + (LOBE == LOBE_PHONG)];
    int* const pair_queue = w.queue_pair[LOBE == LOBE_PHONG];
    // (vertex, light) pairs for the shadow stage: the general kernels only (headline kernels trace one light's queries
    // themselves and hand several lights' to k_nee as whole vertices)
    constexpr bool DEFERS = !HOT && (LOBE == LOBE_LAMBERT || LOBE == LOBE_PHONG);
    WarpPush<2> push;
    push.init();
    WarpPush<4> lobe_push;
    lobe_push.init();
    PairPush pairs;
    pairs.init();
    ShadeCounts counts = { 0u, 0u };
    const bool has_env = c_scene.env_light >= 0;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&counters->shade_vertices, (unsigned long long)n); // traffic model of bench.py

    // whole warps iterate together so that the ballots of the push are convergent; the queue entry is read one
    // iteration ahead
    // (KYD_SHADE_PREFETCH: the single-light kernels may also load the next vertex' record while this one is shaded)
#ifndef KYD_SHADE_PREFETCH_MANY
#define KYD_SHADE_PREFETCH_MANY 1
#endif
#ifndef KYD_SHADE_LOCKSTEP
#define KYD_SHADE_LOCKSTEP 3   // (C5: 0 -> 2655, 2 -> 2685, 3 -> 2736-2755, 7 = also between the trips -> 2705 Msamples/s)
#endif
#ifndef KYD_PENDING_PREFETCH
#define KYD_PENDING_PREFETCH 0   // (L2 prefetch of the previous vertex' light values one iteration ahead: shade 68.9 vs 68.0 ms on C3, no gain)
#endif
#ifndef KYD_SHADE_PREFETCH_SPECULAR
#define KYD_SHADE_PREFETCH_SPECULAR 0   // (measured: 2543 vs 2544 Msamples/s on C5, no effect)
#endif
#ifndef KYD_SHADE_PREFETCH
#define KYD_SHADE_PREFETCH 0   // (1: +16 registers for the next record, spills in the one-light kernels; A/B in profiles/r02_ab_variants.txt)
#endif
    // (the multi-light headline kernels hand their light loop to k_nee: what is left is short and waits for its record --
    // long-scoreboard stalls 14 per issue in profiles/r02_c3_*; they have the registers to load the next one early)
    // (and the mirror / glass kernels: short, 93-105 registers, long-scoreboard 2.7-5 stalls per issue without it)
    constexpr bool SPECULAR_LOBE = LOBE == LOBE_MIRROR || LOBE == LOBE_FRESNEL;
    constexpr bool PREFETCH = HOT && ((NL == NL_ONE && (KYD_SHADE_PREFETCH != 0 || (SPECULAR_LOBE && KYD_SHADE_PREFETCH_SPECULAR != 0))) ||
                                      (NL == NL_MANY && KYD_SHADE_PREFETCH_MANY != 0));
    long long ia = i;
    int slot_cur = ia < n ? queue[ia] : -1;
    int slot_next = ia + stride < n ? queue[ia + stride] : -1;
    float4 rec0 = make_float4(0.f, 0.f, 0.f, 0.f), rec1 = rec0, rec2 = rec0, rec3 = rec0;
    if (PREFETCH && slot_cur >= 0)
    {
        const float4* p0 = path_line(w, slot_cur);
        rec0 = p0[P_ORIGIN]; rec1 = p0[P_DIRECTION]; rec2 = p0[P_BETA]; rec3 = p0[P_TAIL];
    }
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += stride, ia += stride) // block-uniform trip count
    {
#if KYD_SHADE_LOCKSTEP
        // (experiment: the block's four warps enter every vertex together, so that they fetch the same instruction lines at
        // about the same time -- the one-light kernels are instruction-fetch bound; bit 0: Lambert, bit 1: Phong)
        if (HOT_ONE && (((KYD_SHADE_LOCKSTEP & 1) && LOBE == LOBE_LAMBERT) || ((KYD_SHADE_LOCKSTEP & 2) && LOBE == LOBE_PHONG)))
            __syncthreads();
#endif
        const long long i2 = ia + (PREFETCH ? 2ll : 1ll) * stride;
        const int slot_ahead = i2 < n ? queue[i2] : -1;
        float4 nx0 = make_float4(0.f, 0.f, 0.f, 0.f), nx1 = nx0, nx2 = nx0, nx3 = nx0;
        if (PREFETCH && slot_next >= 0)
        {
            const float4* pn = path_line(w, slot_next);
            nx0 = pn[P_ORIGIN]; nx1 = pn[P_DIRECTION]; nx2 = pn[P_BETA]; nx3 = pn[P_TAIL];
        }
#if KYD_PENDING_PREFETCH
        if (HOT && NL == NL_MANY && bounce > 0 && slot_next >= 0)
        {
            // The previous vertex' light values (k_nee's results, 16 B per light, and that vertex' beta) are read as soon as this
            // path's record says they are pending -- a second DRAM round trip behind the record's, 60 % of this kernel's stall
            // cycles (profiles/r02q_prof_c3_line_stalls.txt).  Their address only needs the slot, known one iteration ahead:
            // start them towards L2 now (no registers held).
            const char* res = reinterpret_cast<const char*>(w.nee + (size_t)slot_next * n_lights);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(res));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(res + n_lights * 16 - 1));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vertex_line(w, slot_next) + V_BETA));
        }
#endif
        bool alive = false, split_vertex = false;
        unsigned pair_mask = 0;
        int next_lobe = -1;
        const int slot = slot_cur < 0 ? 0 : slot_cur;
        if (slot_cur >= 0)
        {
            float4* p = path_line(w, slot);
            if (!PREFETCH)
            {
                rec0 = p[P_ORIGIN]; rec1 = p[P_DIRECTION]; rec2 = p[P_BETA]; rec3 = p[P_TAIL];
            }
            VertexOut v;
            if (HOT_ONE)
                shade_vertex_hot_one<LOBE, TRAITS, FUSE>(wp, bounce, &v, &counts, rec0, rec1, rec2, rec3, i0 + (int)blockDim.x <= n);
            else
                shade_vertex<LOBE, TRAITS, HOT, NL>(wp, w, slot, bounce, n_lights, &v, &counts, rec0, rec1, rec2, rec3);
            pair_mask = v.pairs;
            split_vertex = v.split_vertex;
            float hit_t = KYD_INF;
            if (FUSE && v.alive)
            {
                // scene_t::intersect for the next loop iteration of path_tracing_iteration_t::Li (ky.cpp:4542)
                // (the one-light headline kernels trace every continuation themselves, in the third trip of their loop: without
                // the constant, a second, never executed copy of the walk sits in the middle of an instruction-fetch bound kernel)
                if (!HOT_ONE && !v.traced)
                {
                    Ray nr;
                    nr.o = v.o; nr.d = v.d; nr.tmax = KYD_INF;
                    v.hit_surface = wf_closest(nr, &v.hit_t);
                    counts.ref_rays++;
                    counts.traced++;
                }
                if (v.hit_surface >= 0)
                {
                    Ray nr;
                    nr.o = v.o; nr.d = v.d; nr.tmax = KYD_INF;
                    next_lobe = classify_lobe<HOT_ONE && KYD_HAS_SURFACE_TABLE != 0>(v.hit_surface, nr, v.hit_t);
                    v.flags |= (v.hit_surface + 1) << FLAG_SURFACE_SHIFT;
                    hit_t = v.hit_t;
                }
                else
                {
                    // the path leaves the scene: Lo += beta * environment_lighting after a specular bounce (ky.cpp:4548-4563)
                    if (has_env && (v.flags & FLAG_PREV_SPECULAR))
                        v.Lo = add(v.Lo, cmulc(v.beta, environment_lighting()));
                    v.alive = false;
                }
            }
            alive = v.alive;
            // both sectors go back whole; a path that ends here is read again only by k_accumulate, which needs Lo (sector 1)
            // and -- only where light queries are deferred -- the pending count in sector 0
            if (v.alive || !wp.no_pending)
                store_path_ray(p, v.o, hit_t, v.d, v.flags);
            store_path_tail(p, v.beta, v.Lo, v.rng);
        }
        if (FUSE)
        {
            lobe_push.commit(lobe_queues);
            lobe_push.reserve(next_lobe >= 0 ? (1u << next_lobe) : 0u, slot, lobe_tails);
            if (NL == NL_MANY)
            {
                // several lights: the vertex still goes to k_nee's queue (the host zeroes that tail before every bounce's shade)
                push.commit(out_queues);
                push.reserve(split_vertex ? 2u : 0u, slot, tails);
            }
        }
        else
        {
            push.commit(out_queues);
            push.reserve((alive ? 1u : 0u) | (split_vertex ? 2u : 0u), slot, tails);
        }
        if (DEFERS)
        {
            pairs.commit(pair_queue);
            pairs.reserve(pair_mask, slot, pair_tail);
        }
        if (PREFETCH)
        {
            slot_cur = slot_next;
            slot_next = slot_ahead;
            rec0 = nx0; rec1 = nx1; rec2 = nx2; rec3 = nx3;
        }
        else
            slot_cur = slot_ahead;
    }
    if (FUSE)
        lobe_push.commit(lobe_queues);
    if (!FUSE || NL == NL_MANY)
        push.commit(out_queues);
    if (DEFERS)
        pairs.commit(pair_queue);
    flush_counters(counts.ref_rays, counts.traced, counters);
}

// blocks per SM the register allocation aims at.  The mirror / glass kernels and the multi-light Lambert / Phong kernels (which
// hand their light loop to k_nee) are short and wait for their records (long-scoreboard 2.6-6 stalls per issue at 16 warps
// per SM, profiles/r02_c5_stalls.txt, r02_c3_stalls.txt): they can trade registers for warps in flight.
#ifndef KYD_SHADE_MIN_BLOCKS_SPECULAR
#define KYD_SHADE_MIN_BLOCKS_SPECULAR KYD_SHADE_MIN_BLOCKS
#endif
#ifndef KYD_SHADE_MIN_BLOCKS_MANY
#define KYD_SHADE_MIN_BLOCKS_MANY KYD_SHADE_MIN_BLOCKS
#endif
constexpr int shade_min_blocks(int lobe, bool hot, int nl)
{
    return (hot && (lobe == LOBE_MIRROR || lobe == LOBE_FRESNEL)) ? KYD_SHADE_MIN_BLOCKS_SPECULAR
         : (hot && nl == NL_MANY) ? KYD_SHADE_MIN_BLOCKS_MANY : KYD_SHADE_MIN_BLOCKS;
}

// one kernel per lobe: each gets the register allocation its own code needs (the Lambert kernel, which
// shades most vertices, does not pay for Phong's pow() or the dielectric's Fresnel terms)
template <int LOBE, int TRAITS, bool HOT, int NL, bool FUSE>
__global__ void __launch_bounds__(SHADE_THREADS, shade_min_blocks(LOBE, HOT, NL)) k_shade(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters, int bounce)
{
    shade_queue<LOBE, TRAITS, HOT, NL, FUSE>(wp, w, counters, bounce);
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
        float4 p4 = v[V_POSITION], n4 = v[V_NORMAL], wo4 = v[V_WO], c4 = v[V_COLOR], rng4 = v[V_RNG], vb = v[V_BETA];
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
        smp.state = (unsigned long long)__float_as_uint(rng4.x) | ((unsigned long long)__float_as_uint(rng4.y) << 32);
        smp.skip(light_draw_offset(l, wp.rp.direct_sample));
        NeeRay qb, ql;
        light_sample_pair<TRAITS_ANY>(wp.rp.direct_sample, g, b, l, smp, &qb, &ql);
        store_nee_line(nee_line(w, wp.plane, l, slot), qb, ql, V3(vb.x, vb.y, vb.z));
    }
}

__global__ void __launch_bounds__(128) k_light_sample(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters)
{
    light_sample_queue<LOBE_LAMBERT>(wp, w, counters);
    light_sample_queue<LOBE_PHONG>(wp, w, counters);
}

// ---- k_nee: the light loop of the headline configuration with several lights, one thread per (vertex, light) -----------
// sample_all_light (ky.cpp:3834-3872) spends ~450 instructions per light on the two estimators' set-up (a BSDF sample, a
// cone sample of the sphere light, two BSDF evaluations, their pdfs) before any ray is traced; with Veach's five lights
// that loop was 80 % of a shade thread's work, serial, at 124 registers.  Here every (vertex, light) pair is a thread: the
// pairs of a vertex sit in adjacent lanes, so its 96-byte vertex record (written by shade) is one broadcast load, the set-up
// of all lights runs in parallel, and the two scene queries are traced on the spot through one copy of the closest-hit
// walk -- no light-sampling lines, no shadow stage.  The estimator value of the pair (0.5 Lb + 0.5 Ll, ky.cpp:4083) goes to
// results[slot * n_lights + light]; the path adds its vertex' values in light order at its next stage (add_pending).
#ifndef KYD_NEE_CULL
#define KYD_NEE_CULL 1
#endif
#ifndef KYD_NEE_PREFETCH
#define KYD_NEE_PREFETCH 0   // (L2 prefetch of the next vertex record: 703 vs 747 Msamples/s on C3 -- the records are already L1/L2 hits, the extra requests only compete)
#endif
#ifndef KYD_NEE_MIN_BLOCKS
#define KYD_NEE_MIN_BLOCKS 6   // the deferred form fits 80 registers without spills: 24 warps per SM (5 -> 829.5, 6 -> 839.2 Msamples/s on C3)
#endif
#ifndef KYD_NEE_SERIAL_MIN_BLOCKS
#define KYD_NEE_SERIAL_MIN_BLOCKS 5   // 94 registers, 20 warps per SM (A/B in profiles/r02_ab_variants.txt: 4 -> 705, 5 -> 747, 6 -> 716, 8 -> 696 Msamples/s on C3)
#endif
#ifndef KYD_NEE_DEFER
#define KYD_NEE_DEFER 1
#endif
#ifndef KYD_NEE_LIGHT_MAJOR
#define KYD_NEE_LIGHT_MAJOR 1
#endif
#ifndef KYD_NEE_LOCKSTEP
#define KYD_NEE_LOCKSTEP 0
#endif
#ifndef KYD_NEE_SUMMARY
#define KYD_NEE_SUMMARY 1   // per-vertex sums behind the results (wp.pair_kernel == 2): the next stage reads one 32-byte sector per vertex instead of 144 bytes
#endif
#if KYD_NEE_DEFER
// Deferred BSDF-sampled queries.  bsdf_query_certainly_misses settles all but a few per cent of a sphere light's
// BSDF-sampled queries, but a warp holds 32 pairs: with a survivor in most warps, nearly every warp walked the exact set-up
// (double-precision sincos, IEEE divisions and square roots, the sphere test, pdf_Li) and the query's traversal with ~5 of
// its 32 lanes busy -- a quarter of the kernel's issue slots.  Here a pair whose BSDF-sampled query survives the cull is
// parked in a per-warp shared-memory list together with its finished light-sampled half, and the warp runs the exact path
// on a FULL batch of 32 parked pairs whenever it has one (and once more for the rest when it runs out of new pairs).  Every
// iteration of the loop is, for the whole warp, either "32 new pairs" or "32 parked pairs": they share the record load, the
// frame set-up and -- through the NeeRay both produce -- one copy of the closest-hit walk.  The pair's value is the same
// expression of the same operands, 0.5 Lb + 0.5 Ll (ky.cpp:4083), whichever iteration writes it.
template <int LOBE, int TRAITS>
__global__ void __launch_bounds__(128, KYD_NEE_MIN_BLOCKS) k_nee(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters)
{
    stage_rects();
    __shared__ float4 s_parked[4][64];                   // {Ll.rgb, pair index}: < 32 left over + up to 32 new per iteration
#if KYD_NEE_LIGHT_MAJOR && KYD_NEE_SUMMARY
    __shared__ float4 s_run[128];                        // the running sum over the lights of this lane's vertex; w: a pair of it was parked
#endif
    const int n = (int)counters->queue[Q_NEE0 + (LOBE == LOBE_PHONG)];
    const int* __restrict__ queue = w.queue_nee[LOBE == LOBE_PHONG];
    const int n_lights = c_scene.n_lights;
    const int total = n * n_lights;                      // (<= 2^24 paths x 16 lights)
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    float4* parked = s_parked[threadIdx.x >> 5];
    int n_parked = 0;                                    // warp-uniform
    // A warp's new pairs are 32 VERTICES x ONE light, the vertex group's lights in consecutive iterations: the light's data
    // (constant memory, indexed by l) is then one address per warp instead of five -- a constant load with several
    // addresses in a warp is replayed once per address -- and the group's records are L1 hits after the first light.
    // (KYD_NEE_LIGHT_MAJOR=0: the lights of a vertex in adjacent lanes, one broadcast load of the record)
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    const int n_groups = KYD_NEE_LIGHT_MAJOR ? (n + 31) >> 5 : (total + 31) >> 5;
    int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int l_fresh = 0;
    unsigned rays = 0, traced = 0;
#if KYD_NEE_LOCKSTEP
    // The block's warps enter their new-pairs iterations together (as many of them as EVERY warp of the block has: the block's
    // last warp has the fewest groups), so that they fetch the same instruction lines at about the same time: the kernel is
    // instruction-fetch sensitive (profiles/r02_ab_variants.txt).  Parked-batch iterations run between barriers, unsynchronised.
    const int last_warp = ((blockIdx.x + 1) * blockDim.x - 1) >> 5;
    int lockstep_left = KYD_NEE_LIGHT_MAJOR && last_warp < n_groups ? ((n_groups - last_warp + warps_total - 1) / warps_total) * n_lights : 0;
#endif
    for (;;)
    {
        const bool fresh = n_parked < 32 && group < n_groups; // warp-uniform: new pairs while there are any and no full batch waits
        if (!fresh && n_parked == 0)
            break;
#if KYD_NEE_LOCKSTEP
        if (fresh && lockstep_left > 0)
        {
            --lockstep_left;
            __syncthreads();
        }
#endif
        int pair = -1;                                   // vertex | light << 24
        float3 Ll = KYD_BLACK;
        if (fresh)
        {
            if (KYD_NEE_LIGHT_MAJOR)
            {
                const int vertex = group * 32 + lane;
                if (vertex < n)
                    pair = vertex | (l_fresh << PAIR_LIGHT_SHIFT);
                if (++l_fresh == n_lights)
                {
                    l_fresh = 0;
                    group += warps_total;
                }
            }
            else
            {
                const int idx = group * 32 + lane;
                if (idx < total)
                    pair = (idx / n_lights) | ((idx % n_lights) << PAIR_LIGHT_SHIFT);
                group += warps_total;
            }
        }
        else
        {
            const int first = n_parked >= 32 ? n_parked - 32 : 0;
            if (first + lane < n_parked)
            {
                const float4 e = parked[first + lane];
                pair = __float_as_int(e.w);
                Ll = V3(e.x, e.y, e.z);
            }
            n_parked = first;
        }
        __syncwarp();
        bool park = false;
        if (pair >= 0)
        {
            const int vertex = pair & PAIR_SLOT_MASK;
            const int l = pair >> PAIR_LIGHT_SHIFT;
            const int slot = queue[vertex];
            const float4* v = vertex_line(w, slot);
            const float4 p4 = v[V_POSITION], n4 = v[V_NORMAL], wo4 = v[V_WO], c4 = v[V_COLOR], rng4 = v[V_RNG], vb = v[V_BETA];
            HitGeom g;
            g.position = V3(p4.x, p4.y, p4.z);
            g.normal = V3(n4.x, n4.y, n4.z);
            g.wo = V3(wo4.x, wo4.y, wo4.z);
            Bsdf b;
            b.f.s = V3(p4.w, wo4.w, c4.w);
            b.f.t = V3(rng4.z, rng4.w, vb.w);
            b.f.n = normalize(g.normal);                 // frame_t's z axis (ky.cpp:537-541)
            b.a = V3(c4.x, c4.y, c4.z);
            b.t = KYD_BLACK;
            b.eta_t = 1.f;
            b.exponent = n4.w;
            b.lobe = LOBE;

            Sampler smp;
            smp.debug = false;
            smp.state = (unsigned long long)__float_as_uint(rng4.x) | ((unsigned long long)__float_as_uint(rng4.y) << 32);
            smp.skip_lights(l);                          // both_mis: four draws per light before this one (ky.cpp:3866-3868)
            const float2 random_bsdf = smp.get_float2();
            NeeRay q;
            if (fresh)
            {
                // the BSDF-sampled half is settled here only if it certainly misses (the reference still traces that ray)
                park = !(KYD_NEE_CULL && bsdf_query_certainly_misses<TRAITS>(g, b, l, random_bsdf));
                rays += park ? 0u : 1u;
                const float2 random_light = smp.get_float2();
                q = nee_light_setup<TRAITS, KYD_NEE_CULL != 0>(g, b, l, random_light, true);              // ky.cpp:4035-4074
            }
            else
                q = nee_bsdf_setup<TRAITS, false>(g, b, l, random_bsdf, true);                            // ky.cpp:3968-4033
            rays += q.ref_query ? 1u : 0u;
            float3 value = KYD_BLACK;
            if (q.active)
            {
                traced++;
                float t;
                if (TRAITS == TRAITS_ANY && !fresh && q.light_surface < 0)
                {
                    // closest-hit form (a light carried by several surfaces, or the environment light; the host selects the
                    // sphere-light kernels only where every light's query is in its occlusion form)
                    const int s = wf_closest(q.ray, &t);
                    value = nee_bsdf_resolve(q, s, t);
                }
                else
                {
                    const int s = wf_closest_from(q.ray, q.light_surface, &t);
                    value = (fresh ? s < 0 : s == q.light_surface) ? q.value : KYD_BLACK;
                }
            }
            const float3 Lb = fresh ? KYD_BLACK : value;
            if (fresh)
                Ll = value;
            const float3 e = add(mul(Lb, 0.5f), mul(Ll, 0.5f));
            if (!park)
                w.nee[(size_t)slot * n_lights + l] = make_float4(e.x, e.y, e.z, 0.f);
#if KYD_NEE_LIGHT_MAJOR && KYD_NEE_SUMMARY
            if (fresh)
            {
                // sample_all_light's sum (ky.cpp:3864-3869) as the lane walks its vertex' lights in order, then L += beta * Ld
                // (ky.cpp:4575-4576) left as 16 bytes for the consumers (add_pending).  A parked pair's value arrives out of order:
                // its vertex is marked and summed by the reader from the lights' values instead.
                float4 run = l == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : s_run[threadIdx.x];
                if (park)
                    run.w = 1.f;
                else
                {
                    const float3 Ld = add(V3(run.x, run.y, run.z), e);
                    run.x = Ld.x; run.y = Ld.y; run.z = Ld.z;
                }
                if (l == n_lights - 1)
                {
                    // (the vertex' beta is read again here, an L2 hit, rather than kept in three registers across the walk: the
                    // kernel sits exactly at the 80 registers that six blocks per SM allow)
                    const float4 vb_again = __ldcg(&v[V_BETA]);
                    const float3 P = cmulc(V3(vb_again.x, vb_again.y, vb_again.z), V3(run.x, run.y, run.z));
                    nee_summary(w, wp.plane, n_lights)[slot] = make_float4(P.x, P.y, P.z, run.w == 0.f ? 1.f : 0.f);
                }
                else
                    s_run[threadIdx.x] = run;
            }
#endif
        }
        if (fresh)
        {
            const unsigned m = __ballot_sync(0xffffffffu, park);
            if (park)
                parked[n_parked + __popc(m & lt)] = make_float4(Ll.x, Ll.y, Ll.z, __int_as_float(pair));
            n_parked += __popc(m);
        }
        __syncwarp();
    }
    flush_counters(rays, traced, counters);
    if (LOBE == LOBE_LAMBERT && blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&counters->shade_lines, (unsigned long long)total);
}
#endif
// one pair per thread from start to end (lights the cull does not apply to: nothing would be settled early)
template <int LOBE, int TRAITS>
__global__ void __launch_bounds__(128, KYD_NEE_SERIAL_MIN_BLOCKS) k_nee_serial(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters)
{
    stage_rects();
    const int n = (int)counters->queue[Q_NEE0 + (LOBE == LOBE_PHONG)];
    const int* __restrict__ queue = w.queue_nee[LOBE == LOBE_PHONG];
    const int n_lights = c_scene.n_lights;
    const long long total = (long long)n * n_lights;
    const int stride = gridDim.x * blockDim.x;
    unsigned rays = 0, traced = 0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
    {
        const int vertex = (int)(idx / n_lights);        // vertex-major: the lights of a vertex are adjacent lanes
        const int l = (int)(idx - (long long)vertex * n_lights);
        const int slot = queue[vertex];
        const float4* v = vertex_line(w, slot);
#if KYD_NEE_PREFETCH
        if (idx + stride < total)
        {
            // the next iteration's vertex record on its way into L2 while this pair is worked on (no registers held)
            const float4* vn = vertex_line(w, queue[(int)((idx + stride) / n_lights)]);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vn));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vn + 4));
        }
#endif
        const float4 p4 = v[V_POSITION], n4 = v[V_NORMAL], wo4 = v[V_WO], c4 = v[V_COLOR], rng4 = v[V_RNG], vb = v[V_BETA];
        HitGeom g;
        g.position = V3(p4.x, p4.y, p4.z);
        g.normal = V3(n4.x, n4.y, n4.z);
        g.wo = V3(wo4.x, wo4.y, wo4.z);
        Bsdf b;
        b.f.s = V3(p4.w, wo4.w, c4.w);
        b.f.t = V3(rng4.z, rng4.w, vb.w);
        b.f.n = normalize(g.normal);                     // frame_t's z axis (ky.cpp:537-541)
        b.a = V3(c4.x, c4.y, c4.z);
        b.t = KYD_BLACK;
        b.eta_t = 1.f;
        b.exponent = n4.w;
        b.lobe = LOBE;

        Sampler smp;
        smp.debug = false;
        smp.state = (unsigned long long)__float_as_uint(rng4.x) | ((unsigned long long)__float_as_uint(rng4.y) << 32);
        smp.skip_lights(l);                              // both_mis: four draws per light before this one (ky.cpp:3866-3868)
        const float2 random_bsdf = smp.get_float2();
        const float2 random_light = smp.get_float2();
        float3 Lb = KYD_BLACK, Ll = KYD_BLACK;
#pragma unroll 1
        for (int trip = 0; trip < 2; ++trip)
        {
            NeeRay q;
            if (trip == 0) q = nee_bsdf_setup<TRAITS, KYD_NEE_CULL != 0>(g, b, l, random_bsdf, true);   // ky.cpp:3968-4033
            else q = nee_light_setup<TRAITS, KYD_NEE_CULL != 0>(g, b, l, random_light, true);           // ky.cpp:4035-4074
            rays += q.ref_query ? 1u : 0u;
            if (q.active)
            {
                traced++;
                float3 value;
                if (trip == 0 && q.light_surface < 0)
                {
                    // closest-hit form (a light carried by several surfaces, or the environment light)
                    float t;
                    const int s = wf_closest(q.ray, &t);
                    value = nee_bsdf_resolve(q, s, t);
                }
                else
                {
                    float t;
                    const int s = wf_closest_from(q.ray, q.light_surface, &t);
                    value = (trip == 0 ? s == q.light_surface : s < 0) ? q.value : KYD_BLACK;
                }
                if (trip == 0) Lb = value;
                else Ll = value;
            }
        }
        const float3 e = add(mul(Lb, 0.5f), mul(Ll, 0.5f));
        w.nee[(size_t)slot * n_lights + l] = make_float4(e.x, e.y, e.z, 0.f);
    }
    flush_counters(rays, traced, counters);
    if (LOBE == LOBE_LAMBERT && blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&counters->shade_lines, (unsigned long long)total);
}

// ---- shadow: the scene queries of the light loop and the estimators' second halves ------------------------
// closest-hit query for the BSDF-sampled direction, occlusion query for the light-sampled point; writes
// the estimator value of (vertex, light): Lb, Ll or 0.5 Lb + 0.5 Ll (ky.cpp:4083)
// PAIRS: the queue holds the (vertex, light) pairs that got a line (shade wrote them); !PAIRS (split light-sample stage):
// the queue holds vertices and every light has a line.
template <bool PAIRS>
__global__ void __launch_bounds__(256) k_shadow(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters)
{
    stage_rects();
    const int n_lights = c_scene.n_lights;
    const int stride = gridDim.x * blockDim.x;
    const int ds = wp.rp.direct_sample;
    unsigned rays = 0, traced = 0;
    for (int c = 0; c < 2; ++c)
    {
        const long long n = (long long)counters->queue[(PAIRS ? Q_PAIR0 : Q_NEE0) + c];
        const long long total = PAIRS ? n : n * n_lights;
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicAdd(&counters->shade_lines, (unsigned long long)total); // lines written for this stage
        const int* __restrict__ nee_queue = PAIRS ? w.queue_pair[c] : w.queue_nee[c];
        // block-uniform trip count and traversals outside divergent branches (idle lanes trace null rays): the
        // traversal loops stay in the uniform datapath, as in k_intersect
        for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += stride)
        {
            const long long idx = base + threadIdx.x;
            const bool valid = idx < total;
            int l = 0, slot = 0;
            if (valid)
            {
                if (PAIRS)
                {
                    const int entry = nee_queue[idx];
                    l = entry >> PAIR_LIGHT_SHIFT;
                    slot = entry & PAIR_SLOT_MASK;
                }
                else
                {
                    l = (int)(idx / n);
                    slot = nee_queue[idx - (long long)l * n];
                }
            }
            float4* line = nee_line(w, wp.plane, l, slot);
            float4 lo = make_float4(0.f, 0.f, 0.f, -1.f), ld = make_float4(0.f, 0.f, 0.f, 0.f), lv = ld, mx = ld;
            if (valid)
            {
                lo = line[N_LIGHT_O]; ld = line[N_LIGHT_D]; lv = line[N_LIGHT_VALUE]; mx = line[N_MIXED];
            }
            const int flags = __float_as_int(ld.w);
            rays += (unsigned)((flags & NEE_REF_BSDF) != 0) + (unsigned)((flags & NEE_REF_LIGHT) != 0);

            float3 Lb = KYD_BLACK, Ll = KYD_BLACK;
            const bool live = (flags & NEE_BSDF_LIVE) != 0;
            if (__any_sync(0xffffffffu, live))
            {
                NeeRay q;
                q.ray.o = q.ray.d = V3(0.f, 0.f, 0.f);
                q.ray.tmax = -1.f;
                q.value = KYD_BLACK;
                q.light = l;
                q.light_surface = -1;
                if (live)
                {
                    const float4 bo = line[N_BSDF_O], bd = line[N_BSDF_D];
                    q.ray.o = V3(bo.x, bo.y, bo.z);
                    q.ray.d = V3(bd.x, bd.y, bd.z);
                    q.ray.tmax = bo.w;
                    q.value = V3(mx.z, mx.w, bd.w);
                    q.light_surface = (flags >> NEE_LIGHT_SURFACE_SHIFT) - 1;
                }
                const bool occlusion_form = live && q.light_surface >= 0;
                if (__any_sync(0xffffffffu, occlusion_form))
                {
                    const bool blocked = wf_blocked_before(q.ray, occlusion_form ? q.light_surface : -1);
                    if (occlusion_form)
                        Lb = blocked ? KYD_BLACK : q.value;
                }
                if (__any_sync(0xffffffffu, live && !occlusion_form))
                {
                    Ray r = q.ray;
                    if (occlusion_form) r.tmax = -1.f;
                    float t;
                    const int s = wf_closest(r, &t);
                    if (live && !occlusion_form)
                        Lb = nee_bsdf_resolve(q, s, t);
                }
                if (live)
                    traced++;
            }
            {
                const bool light_live = (flags & NEE_LIGHT_LIVE) != 0;
                Ray r;
                r.o = V3(lo.x, lo.y, lo.z);
                r.d = V3(ld.x, ld.y, ld.z);
                r.tmax = light_live ? lo.w : -1.f;   // (no query: nothing can be hit)
                const bool occluded = wf_any_hit(r);
                if (light_live)
                {
                    Ll = occluded ? KYD_BLACK : V3(lv.x, lv.y, lv.z);
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
            if (valid)
            {
                line[N_RESULT] = make_float4(e.x, e.y, e.z, 0.f);
                line[N_VERTEX_BETA] = make_float4(lv.w, mx.x, mx.y, 0.f);
            }
        }
    }
    flush_counters(rays, traced, counters);
}

// ---- the three recursive integrators in wavefront form (ky.cpp:4191-4238, 4305-4402, 4409-4514) ---------------------------
// A recursion level returns  Lo_level + ((f * Li_deeper) * |cos|) / pdf  (ky.cpp:4233, 4400, 4512).  The forward pass is the
// iterative loop -- intersect, shade per lobe -- with the level's (Lo_level, f, |cos|, pdf) written to a record instead of a
// running throughput; k_unwind then applies the expression from the deepest level outwards, which is the recursion's FP32
// evaluation order.  Light queries are traced inside shade (no lines, no pending state); what differs between the three is
// spelled out at the reference lines cited below.
KYD_DEV float3 sample_all_light_inline(const HitGeom& g, const Bsdf& b, Sampler& smp, int ds, ShadeCounts* counts) // ky.cpp:3834-3872
{
    float3 Ld = KYD_BLACK;
    const int n = c_scene.n_lights;
    for (int l = 0; l < n; ++l)
    {
        NeeRay qb, ql;
        light_sample_pair<TRAITS_ANY>(ds, g, b, l, smp, &qb, &ql);
        counts->ref_rays += (qb.ref_query ? 1u : 0u) + (ql.ref_query ? 1u : 0u);
        Ld = add(Ld, nee_resolve_pair(ds, qb, ql, counts));
        smp.skip(4 + ((ds == KYD_DS_BSDF && !light_is_delta(c_scene.lights[l].kind)) ? 2 : 0));
    }
    return Ld;
}

// russian roulette shared by the three (ky.cpp:4219-4226, 4389-4397, 4501-4509)
KYD_DEV bool recursion_roulette_wf(Sampler& smp, int* depth, BsdfSample* bs)
{
    if (++*depth > 3)
    {
        const float m = max_component(bs->f);
        if (smp.get_float() < m)
            bs->f = mul(bs->f, 1 / m);
        else
            return false;
    }
    return true;
}

template <int LOBE, int INTEGRATOR>
__global__ void __launch_bounds__(SHADE_THREADS, 3) k_shade_rec(WaveParams wp, WaveBuffers w, DevCounters* __restrict__ counters, int bounce)
{
    constexpr bool SIMPLE = INTEGRATOR == KYD_INT_SIMPLE_PT_RECURSION, DEFERRED = INTEGRATOR == KYD_INT_PT_RECURSION_DEFERED;
    stage_rects();
    const int parity = bounce & 1;
    const int n = (int)counters->queue[Q_LOBE0 + 4 * parity + LOBE];
    const int* __restrict__ queue = w.queue_lobe[parity][LOBE];
    int* __restrict__ next_queue = parity ? w.queue_a : w.queue_b;
    unsigned long long* const tails[1] = { &counters->queue[Q_RAY0 + (parity ^ 1)] };
    int* const out_queues[1] = { next_queue };
    const int stride = gridDim.x * blockDim.x;
    const int ds = wp.rp.direct_sample, lighting = wp.rp.lighting, max_depth = wp.rp.max_depth;
    WarpPush<1> push;
    push.init();
    ShadeCounts counts = { 0u, 0u };
    long long ia = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += stride, ia += stride) // block-uniform trip count
    {
        bool alive = false;
        int slot = 0;
        if (ia < n)
        {
            slot = queue[ia];
            float4* p = path_line(w, slot);
            PathState st;
            unpack_path(st, p[P_ORIGIN], p[P_DIRECTION], p[P_BETA], p[P_TAIL]);
            Ray r;
            r.o = st.o; r.d = st.d; r.tmax = KYD_INF;
            const int surface = st.surface();
            const int depth = (st.flags & FLAG_DEPTH_MASK) >> FLAG_DEPTH_SHIFT;
            const bool prev_specular = (st.flags & FLAG_PREV_SPECULAR) != 0;
            const HitGeom g = shape_hit_geom(surface_shape(surface), r, st.t);
            const float3 emission = surface_emission(surface, g);
            Bsdf b;
            material_scattering(c_scene.materials[surface_material(surface)], g, &b);   // (the lobe is LOBE: classify_lobe made the same choice)
            Sampler smp;
            smp.debug = wp.rp.sampler == KYD_SAMPLER_DEBUG;
            smp.state = st.rng;

            float3 Lo = KYD_BLACK, f = KYD_BLACK;
            float a = 0.f, pdf = 1.f;
            Ray next = r;
            int next_depth = depth;
            bool next_specular = false;
            if (SIMPLE)
            {
                // ky.cpp:4201-4237: emission at every level; no light sampling; the next ray starts ON the surface (ky.cpp:4232)
                Lo = emission;
                if (depth < max_depth)
                {
                    BsdfSample bs = bsdf_sample(b, g.wo, smp.get_float2());
                    if (!(is_black(bs.f) || bs.pdf == 0.f) && recursion_roulette_wf(smp, &next_depth, &bs))
                    {
                        f = bs.f; a = abs_dot(bs.wi, g.normal); pdf = bs.pdf;
                        next.o = g.position; next.d = bs.wi;
                        alive = true;
                    }
                }
            }
            else
            {
                // ky.cpp:4321-4401 (path_tracing_recursion_t) and 4440-4513 (deferred form, with render_lighting_enum's filter)
                if (depth == 0 || (DEFERRED && prev_specular))
                {
                    const bool keep = !DEFERRED || (depth == 0 ? (lighting & KYD_LIGHTING_EMIT) : (lighting & KYD_LIGHTING_INDIRECT)) != 0;
                    Lo = add(Lo, keep ? emission : KYD_BLACK);
                }
                if (depth < max_depth)
                {
                    const bool delta = bsdf_is_delta(b.lobe);
                    if (!delta)
                    {
                        const float3 Ld = sample_all_light_inline(g, b, smp, ds, &counts);
                        const bool keep = !DEFERRED || (depth == 0 ? (lighting & KYD_LIGHTING_DIRECT) : (lighting & KYD_LIGHTING_INDIRECT)) != 0;
                        Lo = add(Lo, keep ? Ld : KYD_BLACK);
                    }
                    else if (!DEFERRED)
                    {
                        // a specular vertex looks for emitted light along its own sample, from ON the surface (ky.cpp:4340-4350)
                        const BsdfSample bs = bsdf_sample(b, g.wo, smp.get_float2());
                        Ray wi_ray;
                        wi_ray.o = g.position; wi_ray.d = bs.wi; wi_ray.tmax = KYD_INF;
                        float t;
                        const int s = wf_closest(wi_ray, &t);
                        counts.ref_rays++;
                        counts.traced++;
                        const float3 Le = s < 0 ? environment_lighting() : surface_emission(s, shape_hit_geom(surface_shape(s), wi_ray, t));
                        Lo = add(Lo, cdiv(mul(cmulc(bs.f, Le), abs_dot(bs.wi, g.normal)), bs.pdf));
                    }
                    BsdfSample bs = bsdf_sample(b, g.wo, smp.get_float2());
                    if (!(is_black(bs.f) || bs.pdf == 0.f) && recursion_roulette_wf(smp, &next_depth, &bs))
                    {
                        f = bs.f; a = abs_dot(bs.wi, g.normal); pdf = bs.pdf;
                        if (DEFERRED) { next.o = g.position; next.d = bs.wi; }   // no origin offset, ky.cpp:4511
                        else next = spawn_ray(g, bs.wi);                          // ky.cpp:4399
                        next_specular = delta;
                        alive = true;
                    }
                }
            }
            float4* lv = level_line(w, wp.plane, depth, slot);
            lv[0] = make_float4(Lo.x, Lo.y, Lo.z, a);
            lv[1] = make_float4(f.x, f.y, f.z, pdf);
            if (alive)
                store_path_ray(p, next.o, KYD_INF, next.d, (next_specular ? FLAG_PREV_SPECULAR : 0) | (next_depth << FLAG_DEPTH_SHIFT));
            else
                store_path_ray(p, r.o, KYD_INF, r.d, (depth + 1) << FLAG_LEVELS_SHIFT);
            store_path_tail(p, st.beta, st.Lo, smp.state);
        }
        push.commit(out_queues);
        push.reserve(alive ? 1u : 0u, slot, tails);
    }
    push.commit(out_queues);
    flush_counters(counts.ref_rays, counts.traced, counters);
}

// the recursion's return path: result = Lo_n + ((f_n * result) * |cos|_n) / pdf_n from the deepest level outwards; the value
// lands where k_accumulate expects a path's radiance
__global__ void __launch_bounds__(256) k_unwind(WaveParams wp, WaveBuffers w)
{
    const int stride = gridDim.x * blockDim.x;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < wp.nslots; slot += stride)
    {
        float4* p = path_line(w, slot);
        const int levels = (int)((unsigned)__float_as_int(p[P_DIRECTION].w) >> FLAG_LEVELS_SHIFT);
        float3 result = KYD_BLACK;
        for (int n = levels - 1; n >= 0; --n)
        {
            const float4* lv = level_line(w, wp.plane, n, slot);
            const float4 l0 = lv[0], l1 = lv[1];
            const float3 Lo = V3(l0.x, l0.y, l0.z);
            result = n == levels - 1 ? Lo : add(Lo, cdiv(mul(cmulc(V3(l1.x, l1.y, l1.z), result), l0.w), l1.w));
        }
        const float4 tail = p[P_TAIL];
        store_path_tail(p, V3(1.f, 1.f, 1.f), result, (unsigned long long)__float_as_uint(tail.z) | ((unsigned long long)__float_as_uint(tail.w) << 32));
    }
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
            float3 Li;
            if (wp.no_pending)
            {
                const float4 u2 = p[P_BETA], u3 = p[P_TAIL];   // Lo lives in sector 1
                Li = V3(u2.w, u3.x, u3.y);
            }
            else
            {
                PathState st;
                unpack_path(st, p[P_ORIGIN], p[P_DIRECTION], p[P_BETA], p[P_TAIL]);
                Li = add_pending(w, wp.plane, slot, st.pending(), st.Lo, wp.pair_kernel);
            }
            L = add(L, mul(Li, wp.rp.weight));
        }
        o[0] = film_value(L.x); o[1] = film_value(L.y); o[2] = film_value(L.z);
    }
}

__global__ void k_zero(float* __restrict__ p, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        p[i] = 0.f;
}

} // namespace KYD_KERNEL_NS
