// kyd_api.cu -- the C ABI of include/kyd.h: context, scene upload, render orchestration.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kyd_internal.h"

using namespace kyd;

struct kyd_ctx
{
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool events_recorded = false;   // ev_begin / ev_end hold a render's interval
    bool counters_in_flight = false; // an asynchronous render's counters have not been folded into stats yet
    std::string error;

    bool has_scene = false;
    DevScene scene{};
    unsigned long long scene_generation = 0;

    float* film_dev = nullptr;      // staging film for kyd_render (host destination)
    size_t film_capacity = 0;       // floats
    float* film_pinned = nullptr;   // pinned bounce buffer for the device->host copy
    size_t pinned_capacity = 0;
    uint8_t* body_dev = nullptr;    // encoded image body for kyd_film_encode (host destination)
    size_t body_capacity = 0;
    void* big_scene_dev = nullptr;  // one allocation: shapes | materials | lights | BVH nodes | BVH leaf order (large scenes)

    DevCounters* counters_dev = nullptr;
    DevCounters* counters_pinned = nullptr;
    kyd_stats stats{};

    WaveBuffers wave;
    int64_t wave_paths = 0;
    bool stage_timing = false; // KYD_STAGE_TIMING=1: per-stage CUDA-event times in kyd_stats::stage_ms
    StageTimer timer;

    // multi-GPU context (kyd_create_multi): this context is rank 0; the others render the rest of the sample range on
    // their own devices and their partial films are summed here in rank order (kyd_multi section below)
    std::vector<kyd_ctx*> peers;
    std::vector<bool> peer_direct;       // rank r's partial film is read in place (same device, or peer access enabled)
    std::vector<float*> peer_staging;    // staging buffers on this device for peers without peer access
    std::vector<size_t> peer_staging_capacity;
    std::vector<cudaEvent_t> peer_done;  // rank r's partial film is complete
    cudaEvent_t sum_done = nullptr;      // rank 0 has read the peers' partial films
};

namespace {

std::string g_create_error;

// c_scene is ONE __constant__ symbol per device (and per build of the kernels), shared by every context on that device.
// A render therefore owns the symbol for its whole launch sequence: `launch` is held while the sequence is enqueued, and
// a render that has to replace the symbol's contents, or that runs on another stream than the previous user, first
// makes its stream wait for `last_use` (recorded behind the previous user's last kernel).  Two contexts on one GPU
// driven from two host threads thus take turns per render instead of tracing each other's scene.
struct SceneSlot
{
    std::mutex launch;
    const kyd_ctx* owner = nullptr;
    unsigned long long generation = 0;
    cudaEvent_t last_use = nullptr;
    cudaStream_t last_stream = nullptr;
    bool used = false;
};
SceneSlot g_scene_slot[2][64];   // [small-scene build, large-scene build of the kernels][device]

int fail(kyd_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->error = msg;
    else g_create_error = msg;
    return code;
}

#define KYD_CUDA(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, KYD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)

float3 f3(const float* p) { return make_float3(p[0], p[1], p[2]); }

// The film comes back in chunks: chunk k is copied out of the pinned bounce buffer into the caller's (pageable) memory
// while chunks k+1.. are still crossing PCIe, and large chunks are split over a few host threads (a single-threaded
// memcpy of a 4K float film costs more than its PCIe transfer).
void copy_out(void* dst, const void* src, size_t bytes)
{
    const size_t min_per_thread = 4u << 20;
    size_t threads = bytes / min_per_thread;
    if (threads > 4) threads = 4;
    if (threads < 2) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> pool;
    const size_t part = (bytes / threads + 63) & ~(size_t)63;
    for (size_t t = 1; t < threads; ++t)
    {
        const size_t begin = t * part, end = (t + 1 == threads) ? bytes : (t + 1) * part;
        pool.emplace_back([=] { memcpy((char*)dst + begin, (const char*)src + begin, end - begin); });
    }
    memcpy(dst, src, part < bytes ? part : bytes);
    for (auto& th : pool) th.join();
}

int download_film(kyd_ctx* ctx, const float* dev, float* pinned, float* host, size_t bytes)
{
    const int chunks = bytes >= (32u << 20) ? 4 : 1;
    const size_t part = (bytes / chunks + 255) & ~(size_t)255;
    cudaEvent_t done[4] = {};
    for (int k = 0; k < chunks; ++k)
    {
        const size_t begin = (size_t)k * part, end = (k + 1 == chunks) ? bytes : (size_t)(k + 1) * part;
        KYD_CUDA(ctx, cudaMemcpyAsync((char*)pinned + begin, (const char*)dev + begin, end - begin, cudaMemcpyDeviceToHost, ctx->stream));
        KYD_CUDA(ctx, cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
        KYD_CUDA(ctx, cudaEventRecord(done[k], ctx->stream));
    }
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < chunks; ++k)
    {
        const size_t begin = (size_t)k * part, end = (k + 1 == chunks) ? bytes : (size_t)(k + 1) * part;
        if (e == cudaSuccess) e = cudaEventSynchronize(done[k]);
        if (e == cudaSuccess) copy_out((char*)host + begin, (const char*)pinned + begin, end - begin);
        cudaEventDestroy(done[k]);
    }
    KYD_CUDA(ctx, e);
    KYD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KYD_OK;
}

DevMaterial convert_material(const kyd_material& m)
{
    DevMaterial o{};
    o.kind = m.kind;
    o.diffuse = f3(m.diffuse); o.specular = f3(m.specular); o.transmission = f3(m.transmission);
    o.eta = m.eta; o.exponent = m.exponent;
    o.p_diffuse = m.diffuse_probability; o.p_specular = m.specular_probability;
    // color_t / float_t is three divisions (ky.cpp:231); IEEE on the host == IEEE on the device
    o.plastic_lambert = make_float3(m.diffuse[0] / m.diffuse_probability, m.diffuse[1] / m.diffuse_probability, m.diffuse[2] / m.diffuse_probability);
    o.plastic_phong = make_float3(m.specular[0] / m.specular_probability, m.specular[1] / m.specular_probability, m.specular[2] / m.specular_probability);
    return o;
}

DevShape convert_shape(const kyd_shape& s)
{
    DevShape d{};
    d.kind = s.kind;
    d.p0 = f3(s.p0); d.p1 = f3(s.p1); d.p2 = f3(s.p2); d.p3 = f3(s.p3); d.n = f3(s.normal);
    d.radius = s.radius; d.radius_sq = s.radius_sq; d.area = s.area;
    return d;
}

// The rectangle as the parallelogram p1 + u (p0 - p1) + v (p2 - p1), in double (kyd_device.cuh: rect_classify)
RectCull make_rect_cull(const DevShape& s)
{
    RectCull c{};
    const double b[3] = { s.p1.x, s.p1.y, s.p1.z };
    const double eu[3] = { s.p0.x - b[0], s.p0.y - b[1], s.p0.z - b[2] };
    const double ev[3] = { s.p2.x - b[0], s.p2.y - b[1], s.p2.z - b[2] };
    const double p3[3] = { s.p3.x, s.p3.y, s.p3.z };
    double n[3] = { eu[1] * ev[2] - eu[2] * ev[1], eu[2] * ev[0] - eu[0] * ev[2], eu[0] * ev[1] - eu[1] * ev[0] };
    const double area = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    double d1 = 0, d2 = 0, dev = 0;
    for (int a = 0; a < 3; ++a)
    {
        d1 += (eu[a] - ev[a]) * (eu[a] - ev[a]);              // |p0 - p2|^2
        d2 += (eu[a] + ev[a]) * (eu[a] + ev[a]);              // |p3 - p1|^2 of the ideal parallelogram
        const double ideal = b[a] + eu[a] + ev[a];
        dev += (p3[a] - ideal) * (p3[a] - ideal);
    }
    const double diag = std::sqrt(d1 > d2 ? d1 : d2);
    c.b = make_float3((float)b[0], (float)b[1], (float)b[2]);
    const bool degenerate = !(area > 1e-30) || !(area < 1e15) || !(diag < 1e15) || !(std::sqrt(dev) <= 0x1p-20 * diag);
    if (degenerate)
        return c;   // n = 0: every ray is a candidate
    for (int a = 0; a < 3; ++a) n[a] /= area;
    // dual basis in the plane: gu = (ev x n) / (eu . (ev x n)), gv = (n x eu) / (ev . (n x eu))
    double evxn[3] = { ev[1] * n[2] - ev[2] * n[1], ev[2] * n[0] - ev[0] * n[2], ev[0] * n[1] - ev[1] * n[0] };
    double nxeu[3] = { n[1] * eu[2] - n[2] * eu[1], n[2] * eu[0] - n[0] * eu[2], n[0] * eu[1] - n[1] * eu[0] };
    const double su = eu[0] * evxn[0] + eu[1] * evxn[1] + eu[2] * evxn[2];
    const double sv = ev[0] * nxeu[0] + ev[1] * nxeu[1] + ev[2] * nxeu[2];
    c.n = make_float3((float)n[0], (float)n[1], (float)n[2]);
    c.gu = make_float3((float)(evxn[0] / su), (float)(evxn[1] / su), (float)(evxn[2] / su));
    c.gv = make_float3((float)(nxeu[0] / sv), (float)(nxeu[1] / sv), (float)(nxeu[2] / sv));
    c.c_area = (float)(0x1p-15 / area);
    return c;
}

// axis-aligned rectangle: all four points share one coordinate and the edges p1->p0, p1->p2 each run along one axis;
// returns the normal axis or -1
int make_rect_aligned(const DevShape& s, RectAligned* out)
{
    const float p[4][3] = { { s.p0.x, s.p0.y, s.p0.z }, { s.p1.x, s.p1.y, s.p1.z }, { s.p2.x, s.p2.y, s.p2.z }, { s.p3.x, s.p3.y, s.p3.z } };
    for (int a = 0; a < 3; ++a)
    {
        if (!(p[0][a] == p[1][a] && p[1][a] == p[2][a] && p[2][a] == p[3][a]))
            continue;
        const int b = (a + 1) % 3, c = (a + 2) % 3;
        // edge p1->p0 along one in-plane axis, p1->p2 along the other, p3 the opposite corner (exactly: same floats)
        const bool u_b = p[0][c] == p[1][c] && p[2][b] == p[1][b] && p[3][b] == p[0][b] && p[3][c] == p[2][c];
        const bool u_c = p[0][b] == p[1][b] && p[2][c] == p[1][c] && p[3][c] == p[0][c] && p[3][b] == p[2][b];
        if (!u_b && !u_c)
            return -1;
        double lo_b = p[0][b], hi_b = p[0][b], lo_c = p[0][c], hi_c = p[0][c];
        for (int k = 1; k < 4; ++k)
        {
            lo_b = p[k][b] < lo_b ? p[k][b] : lo_b; hi_b = p[k][b] > hi_b ? p[k][b] : hi_b;
            lo_c = p[k][c] < lo_c ? p[k][c] : lo_c; hi_c = p[k][c] > hi_c ? p[k][c] : hi_c;
        }
        const double area = (hi_b - lo_b) * (hi_c - lo_c);
        if (!(area > 1e-30) || !(area < 1e15) || !std::isfinite(area))
            return -1;
        out->pa = p[0][a];
        out->lo_b = (float)lo_b;
        out->lo_c = (float)lo_c;
        out->gb = (float)(1.0 / (hi_b - lo_b));
        out->gc = (float)(1.0 / (hi_c - lo_c));
        out->c_area = (float)(0x1p-15 / area);
        return a;
    }
    return -1;
}

// classifiers of the first 32 rectangles of the sorted copy, regrouped aligned-x | aligned-y | aligned-z | general
void build_rect_classifiers(DevScene& d)
{
    const int n = d.kind_end[0] < 32 ? d.kind_end[0] : 32;
    RectAligned aligned[32];
    int axis[32];
    for (int k = 0; k < n; ++k)
        axis[k] = make_rect_aligned(d.sorted_shape[k], &aligned[k]);
    int pos = 0;
    for (int a = 0; a < 3; ++a)
    {
        for (int k = 0; k < n; ++k)
            if (axis[k] == a)
            {
                d.rect_aligned[pos] = aligned[k];
                d.rect_order[pos] = k;
                ++pos;
            }
        d.rect_aligned_end[a] = pos;
    }
    d.n_rect_general = 0;
    for (int k = 0; k < n; ++k)
        if (axis[k] < 0)
        {
            d.rect_general[d.n_rect_general++] = make_rect_cull(d.sorted_shape[k]);
            d.rect_order[pos++] = k;
        }
    d.n_rect_cull = pos;
    // L1 bound of the classified rectangles' vertices around the centre of their bounding box
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int k = 0; k < n; ++k)
    {
        const DevShape& s = d.sorted_shape[k];
        const float3 pts[4] = { s.p0, s.p1, s.p2, s.p3 };
        for (const float3& q : pts)
        {
            const double v[3] = { q.x, q.y, q.z };
            for (int a = 0; a < 3; ++a) { lo[a] = v[a] < lo[a] ? v[a] : lo[a]; hi[a] = v[a] > hi[a] ? v[a] : hi[a]; }
        }
    }
    if (n == 0) { for (int a = 0; a < 3; ++a) lo[a] = hi[a] = 0; }
    d.bound_center = make_float3((float)(0.5 * (lo[0] + hi[0])), (float)(0.5 * (lo[1] + hi[1])), (float)(0.5 * (lo[2] + hi[2])));
    d.bound_l1 = (float)((0.5 * ((hi[0] - lo[0]) + (hi[1] - lo[1]) + (hi[2] - lo[2]))) * (1.0 + 1e-6) + 1e-30);
}

int ensure_film(kyd_ctx* ctx, size_t floats)
{
    if (floats <= ctx->film_capacity)
        return KYD_OK;
    if (ctx->film_dev) cudaFree(ctx->film_dev);
    ctx->film_dev = nullptr;
    ctx->film_capacity = 0;
    KYD_CUDA(ctx, cudaMalloc(&ctx->film_dev, floats * sizeof(float)));
    ctx->film_capacity = floats;
    return KYD_OK;
}

int ensure_pinned(kyd_ctx* ctx, size_t floats)
{
    if (floats <= ctx->pinned_capacity)
        return KYD_OK;
    if (ctx->film_pinned) cudaFreeHost(ctx->film_pinned);
    ctx->film_pinned = nullptr;
    ctx->pinned_capacity = 0;
    KYD_CUDA(ctx, cudaMallocHost(&ctx->film_pinned, floats * sizeof(float)));
    ctx->pinned_capacity = floats;
    return KYD_OK;
}

int validate(kyd_ctx* ctx, const kyd_render_desc* d)
{
    if (!d) return fail(ctx, KYD_ERR_INVALID, "render desc is null");
    if (d->width <= 0 || d->height <= 0) return fail(ctx, KYD_ERR_INVALID, "film size must be positive");
    if (d->width >= 65536 || d->height >= 65536) return fail(ctx, KYD_ERR_INVALID, "film size must be below 65536 (sampler key)");
    if (d->spp <= 0) return fail(ctx, KYD_ERR_INVALID, "spp must be positive");
    if (d->sample_begin < 0 || d->sample_end < d->sample_begin || d->sample_end > (1 << 24))
        return fail(ctx, KYD_ERR_INVALID, "bad sample range");
    // every sample weighs 1/spp and the job has max(1, spp) of them (ky.cpp:3712-3723); indices beyond would push the
    // weights' sum above 1 (and the trapezoidal sampler's sub-pixel index out of its 2x2 grid)
    if (d->sample_end > (d->spp > 1 ? d->spp : 1))
        return fail(ctx, KYD_ERR_INVALID, "bad sample range: sample_end exceeds spp");
    switch (d->integrator)
    {
    case KYD_INT_POSITION: case KYD_INT_NORMAL: case KYD_INT_BASECOLOR: case KYD_INT_DIRECT_LIGHTING:
    case KYD_INT_PT_ITERATION:
        break;
    case KYD_INT_SIMPLE_PT_RECURSION: case KYD_INT_PT_RECURSION: case KYD_INT_PT_RECURSION_DEFERED:
        if (d->max_depth > KYD_MAX_RECURSION - 1)
            return fail(ctx, KYD_ERR_INVALID, "max_depth too large for a recursive integrator");
        break;
    default:
        return fail(ctx, KYD_ERR_INVALID, "unsupported integrator_enum_t value");
    }
    switch (d->direct_sample)
    {
    case KYD_DS_IDLE: case KYD_DS_BSDF: case KYD_DS_LIGHT: case KYD_DS_BSDF_MIS: case KYD_DS_LIGHT_MIS: case KYD_DS_BOTH_MIS:
        break;
    default:
        return fail(ctx, KYD_ERR_INVALID, "unsupported direct_sample_enum_t value");
    }
    if (d->sampler != KYD_SAMPLER_LCG48 && d->sampler != KYD_SAMPLER_DEBUG && d->sampler != KYD_SAMPLER_TRAPEZOIDAL)
        return fail(ctx, KYD_ERR_INVALID, "unsupported sampler");
    if (d->sampler == KYD_SAMPLER_TRAPEZOIDAL && (d->spp < 4 || d->spp % 4 != 0))
        return fail(ctx, KYD_ERR_INVALID, "the trapezoidal sampler needs spp to be a multiple of 4 (2x2 sub-pixels)");
    if (d->max_depth < 0) return fail(ctx, KYD_ERR_INVALID, "max_depth must be >= 0");
    return KYD_OK;
}

// makes the device's constant scene the one of this context; the caller holds slot.launch
int bind_scene(kyd_ctx* ctx, SceneSlot& slot, cudaStream_t stream)
{
    const bool big = ctx->scene.bvh_nodes != nullptr;
    if (!slot.last_use)
        KYD_CUDA(ctx, cudaEventCreateWithFlags(&slot.last_use, cudaEventDisableTiming));
    const bool replace = slot.owner != ctx || slot.generation != ctx->scene_generation;
    // kernels of the previous user may still read the symbol (replace), or ran on another stream whose order against
    // this one nothing else establishes (a later owner waits only for the LAST user's event, so users are chained)
    if (slot.used && (replace || slot.last_stream != stream))
        KYD_CUDA(ctx, cudaStreamWaitEvent(stream, slot.last_use, 0));
    if (replace)
    {
        if (big) kyd_big::upload_scene_constant(ctx->scene, stream);
        else upload_scene_constant(ctx->scene, stream);
        KYD_CUDA(ctx, cudaGetLastError());
        slot.owner = ctx;
        slot.generation = ctx->scene_generation;
    }
    return KYD_OK;
}

// bytes of wavefront state per path slot under `plan` (ensure_wave_buffers, kyd_kernels.cu)
size_t wave_bytes_per_path(const WavefrontPlan& plan, int n_lights)
{
    size_t b = 64 + 12 * 4;                                  // path record, 2 ray queues + 2 x 4 lobe queues + 2 vertex queues
    if (plan.nee && !plan.pair_kernel) b += (size_t)n_lights * (128 + 2 * 4);   // light-sampling lines + pair queues
    if (plan.pair_kernel) b += (size_t)(n_lights + 1) * 16;    // k_nee results + per-vertex summary
    if (plan.split || plan.pair_kernel) b += 96;               // vertex records
    return b;   // (+ 32 bytes per level for the recursive integrators, added by the caller)
}

int render_to_device(kyd_ctx* ctx, const kyd_render_desc* d, float* film_dev, cudaStream_t stream, bool timed)
{
    RenderParams rp{};
    rp.width = d->width; rp.height = d->height; rp.spp = d->spp;
    rp.sample_begin = d->sample_begin; rp.sample_end = d->sample_end;
    rp.integrator = d->integrator; rp.max_depth = d->max_depth; rp.direct_sample = d->direct_sample;
    rp.lighting = d->lighting; rp.sampler = d->sampler; rp.seed = d->seed; rp.flags = d->flags;
    rp.weight = (float)(1. / d->spp);

    const bool recursion = d->integrator == KYD_INT_SIMPLE_PT_RECURSION || d->integrator == KYD_INT_PT_RECURSION || d->integrator == KYD_INT_PT_RECURSION_DEFERED;
    // (the recursive integrators' wavefront form needs constant-memory scenes: large scenes keep the per-pixel kernel)
    const bool wavefront = !(d->flags & KYD_FLAG_FUSED) &&
        (d->integrator == KYD_INT_PT_ITERATION || d->integrator == KYD_INT_DIRECT_LIGHTING || (recursion && !ctx->scene.bvh_nodes));
    int64_t capacity = 0;
    if (wavefront)
    {
        // wave size: the caller's choice, else 16 Mi paths (or the whole job if it is smaller).  Measured on
        // B200 (profiles/r01_wave_sweep.txt): throughput grows with the wave up to the whole 4K film --
        // launch gaps and wave tails cost more than L2 residency of the path state would win.
        const int64_t job = (int64_t)d->width * d->height * (int64_t)(d->sample_end - d->sample_begin);
        capacity = ctx->wave_paths > 0 ? ctx->wave_paths : (int64_t)1 << 24;
        if (capacity > ((int64_t)1 << 24)) capacity = (int64_t)1 << 24;   // pair-queue entries carry the slot in 24 bits
        if (capacity > job) capacity = job;
        if (capacity < 1024) capacity = 1024;
        // light-sampling lines (128 B per light and path) and vertex records (96 B per path) only where the plan uses them
        const WavefrontPlan plan = wavefront_plan(rp, ctx->scene);
        const int nee_lights = plan.nee ? ctx->scene.n_lights + (plan.pair_kernel ? 1 : 0) : 0;   // (+1: k_nee's per-vertex summaries)
        // Results do not depend on the wave size, so memory decides it where it is short (a GPU shared with other work,
        // many lights: 136 B per light and path): the default is capped by what is free now, and an allocation that
        // still fails is retried with half the wave down to 64 Ki paths.
        const int nee_units = plan.pair_kernel ? 1 : 8;
        const bool vertex = plan.split || plan.pair_kernel;
        const int levels = plan.recursion ? d->max_depth + 1 : 0;
        if (levels > 0)
        {
            // level records: 32 bytes per level and path; keep them below ~6 GB by shrinking the wave
            const int64_t fit = (int64_t)6 << 30 >> 5;
            if (capacity > fit / levels) capacity = fit / levels < 65536 ? 65536 : fit / levels;
        }
        if (ctx->wave.max_levels < levels || ctx->wave.capacity < capacity || ctx->wave.max_lights < nee_lights || (nee_lights > 0 && ctx->wave.nee_units < nee_units) ||
            (vertex && !ctx->wave.has_vertex))
        {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
            {
                const size_t held = (size_t)ctx->wave.capacity * wave_bytes_per_path(plan, ctx->wave.max_lights);
                const int64_t fit = (int64_t)((free_b + held) / 10 * 9 / wave_bytes_per_path(plan, nee_lights > ctx->wave.max_lights ? nee_lights : ctx->wave.max_lights));
                if (capacity > fit && ctx->wave_paths == 0) capacity = fit < 65536 ? 65536 : fit;
            }
            for (;;)
            {
                const cudaError_t e = (cudaError_t)ensure_wave_buffers(ctx->wave, capacity, nee_lights, nee_units, vertex, levels);
                if (e == cudaSuccess) break;
                cudaGetLastError();   // clear the sticky-free allocation error
                if (e != cudaErrorMemoryAllocation || capacity <= 65536)
                    return fail(ctx, KYD_ERR_CUDA, std::string("wavefront buffers: ") + cudaGetErrorString(e));
                capacity /= 2;
            }
            if (capacity > ctx->wave.capacity) capacity = ctx->wave.capacity;
        }
    }

    SceneSlot& slot = g_scene_slot[ctx->scene.bvh_nodes != nullptr][ctx->device & 63];
    std::lock_guard<std::mutex> lock(slot.launch);   // the symbol is ours until the launch sequence is enqueued
    int rc = bind_scene(ctx, slot, stream);
    if (rc != KYD_OK) return rc;

    KYD_CUDA(ctx, cudaMemsetAsync(ctx->counters_dev, 0, sizeof(DevCounters), stream));
    if (timed) KYD_CUDA(ctx, cudaEventRecord(ctx->ev_begin, stream));

    uint64_t launches = 0;
    if (wavefront)
    {
        ctx->timer.stream = stream;
        if (ctx->scene.bvh_nodes)
            kyd_big::launch_render_wavefront(rp, ctx->scene, ctx->wave, capacity, film_dev, ctx->counters_dev, stream, ctx->sm_count, &launches,
                                             ctx->stage_timing ? &ctx->timer : nullptr);
        else
            launch_render_wavefront(rp, ctx->scene, ctx->wave, capacity, film_dev, ctx->counters_dev, stream, ctx->sm_count, &launches,
                                    ctx->stage_timing ? &ctx->timer : nullptr);
    }
    else
    {
        if (ctx->scene.bvh_nodes) kyd_big::launch_render_pixels(rp, film_dev, ctx->counters_dev, stream);
        else launch_render_pixels(rp, film_dev, ctx->counters_dev, stream);
        launches += 1;
    }
    KYD_CUDA(ctx, cudaGetLastError());

    if (timed)
    {
        KYD_CUDA(ctx, cudaEventRecord(ctx->ev_end, stream));
        ctx->events_recorded = true;
    }
    KYD_CUDA(ctx, cudaMemcpyAsync(ctx->counters_pinned, ctx->counters_dev, sizeof(DevCounters), cudaMemcpyDeviceToHost, stream));
    KYD_CUDA(ctx, cudaEventRecord(slot.last_use, stream));
    slot.last_stream = stream;
    slot.used = true;

    ctx->stats = kyd_stats{};
    ctx->stats.samples = (uint64_t)d->width * d->height * (uint64_t)(d->sample_end - d->sample_begin);
    ctx->stats.kernel_launches = launches;
    ctx->counters_in_flight = true;
    return KYD_OK;
}

// folds the counters of the last render into ctx->stats (the caller has synchronised with its stream)
void finish_stats(kyd_ctx* ctx)
{
    if (!ctx->counters_in_flight)
        return;
    ctx->counters_in_flight = false;
    ctx->stats.rays = ctx->counters_pinned->rays;
    ctx->stats.rays_traced = ctx->counters_pinned->rays_traced;
    ctx->stats.shade_vertices = ctx->counters_pinned->shade_vertices;
    ctx->stats.shade_light_lines = ctx->counters_pinned->shade_lines;
    ctx->stats.intersect_rays = ctx->counters_pinned->intersect_rays;
    if (ctx->stage_timing) ctx->timer.collect(ctx->stats.stage_ms);
    if (ctx->events_recorded)
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end) == cudaSuccess)
            ctx->stats.device_ms = ms;
        else
            cudaGetLastError();   // an interval that is not complete is not an error of the next call
    }
}

} // namespace

extern "C" {

int kyd_create(kyd_ctx** out_ctx, int device)
{
    if (!out_ctx) return fail(nullptr, KYD_ERR_INVALID, "out_ctx is null");
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, KYD_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (kyd has no CPU fallback)");
    if (device < 0 || device >= count)
        return fail(nullptr, KYD_ERR_INVALID, "device ordinal out of range");

    kyd_ctx* ctx = new kyd_ctx;
    ctx->device = device;
    auto cleanup = [&](const std::string& msg) { g_create_error = msg; kyd_destroy(ctx); return KYD_ERR_CUDA; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cleanup(cudaGetErrorString(e));
    if ((e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return cleanup(cudaGetErrorString(e));
    // a BLOCKING stream: ordered against the device's legacy default stream, so work a caller queued there (torch's
    // default stream: film.zero_(), an earlier reduce) is complete before a render touches the film, and vice versa
    if ((e = cudaStreamCreate(&ctx->stream)) != cudaSuccess) return cleanup(cudaGetErrorString(e));
    if ((e = cudaEventCreate(&ctx->ev_begin)) != cudaSuccess) return cleanup(cudaGetErrorString(e));
    if ((e = cudaEventCreate(&ctx->ev_end)) != cudaSuccess) return cleanup(cudaGetErrorString(e));
    if ((e = cudaMalloc(&ctx->counters_dev, sizeof(DevCounters))) != cudaSuccess) return cleanup(cudaGetErrorString(e));
    if ((e = cudaMallocHost(&ctx->counters_pinned, sizeof(DevCounters))) != cudaSuccess) return cleanup(cudaGetErrorString(e));
    memset(ctx->counters_pinned, 0, sizeof(DevCounters));
    { const char* e = getenv("KYD_STAGE_TIMING"); ctx->stage_timing = e && e[0] == '1'; }
    *out_ctx = ctx;
    return KYD_OK;
}

void kyd_destroy(kyd_ctx* ctx)
{
    if (!ctx) return;
    for (kyd_ctx* peer : ctx->peers) kyd_destroy(peer);
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();   // an asynchronous render on a caller's stream may still use the buffers freed below
    for (float* p : ctx->peer_staging) if (p) cudaFree(p);
    for (cudaEvent_t e : ctx->peer_done) if (e) cudaEventDestroy(e);
    if (ctx->sum_done) cudaEventDestroy(ctx->sum_done);
    for (int b = 0; b < 2; ++b)
    {
        SceneSlot& slot = g_scene_slot[b][ctx->device & 63];
        std::lock_guard<std::mutex> lock(slot.launch);
        if (slot.owner == ctx) slot.owner = nullptr;
    }
    free_wave_buffers(ctx->wave);
    if (ctx->film_dev) cudaFree(ctx->film_dev);
    if (ctx->film_pinned) cudaFreeHost(ctx->film_pinned);
    if (ctx->body_dev) cudaFree(ctx->body_dev);
    if (ctx->big_scene_dev) cudaFree(ctx->big_scene_dev);
    if (ctx->counters_dev) cudaFree(ctx->counters_dev);
    if (ctx->counters_pinned) cudaFreeHost(ctx->counters_pinned);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* kyd_last_error(const kyd_ctx* ctx)
{
    return ctx ? ctx->error.c_str() : g_create_error.c_str();
}

namespace {

// ---- bounding-volume hierarchy of a large scene (the reference's accel_t is an empty hook, ky.cpp:3097-3115) ----------
struct Box { float lo[3], hi[3]; };

Box shape_box(const DevShape& s)
{
    Box b;
    const float3 pts[4] = { s.p0, s.p1, s.p2, s.p3 };
    const int n = s.kind == KYD_SHAPE_RECTANGLE ? 4 : s.kind == KYD_SHAPE_TRIANGLE ? 3 : 1;
    const float r = (s.kind == KYD_SHAPE_SPHERE || s.kind == KYD_SHAPE_DISK) ? s.radius : 0.f;
    for (int a = 0; a < 3; ++a)
    {
        float lo = 3.4e38f, hi = -3.4e38f;
        for (int k = 0; k < n; ++k)
        {
            const float v = a == 0 ? pts[k].x : a == 1 ? pts[k].y : pts[k].z;
            lo = v - r < lo ? v - r : lo;
            hi = v + r > hi ? v + r : hi;
        }
        // padding: hits are accepted by the shapes' own arithmetic, whose points may sit a rounding error outside
        const float m = fabsf(lo) > fabsf(hi) ? fabsf(lo) : fabsf(hi);
        const float pad = 1e-3f + 1e-4f * m + 1e-3f * (hi - lo);
        b.lo[a] = lo - pad;
        b.hi[a] = hi + pad;
    }
    return b;
}

struct BvhBuilder
{
    const std::vector<Box>& boxes;
    std::vector<BvhNode> nodes;
    std::vector<int> prims;
    int max_depth = 0;   // deepest node (root = 0): the traversal stack holds at most max_depth + 1 entries

    explicit BvhBuilder(const std::vector<Box>& b) : boxes(b)
    {
        prims.resize(b.size());
        for (size_t i = 0; i < b.size(); ++i) prims[i] = (int)i;
        nodes.reserve(2 * b.size());
        nodes.push_back(BvhNode{});
        build(0, 0, (int)b.size(), 0);
    }

    void build(int node, int first, int count, int depth)
    {
        if (depth > max_depth) max_depth = depth;
        Box bound{ { 3.4e38f, 3.4e38f, 3.4e38f }, { -3.4e38f, -3.4e38f, -3.4e38f } };
        float clo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, chi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
        for (int k = first; k < first + count; ++k)
        {
            const Box& b = boxes[prims[k]];
            for (int a = 0; a < 3; ++a)
            {
                bound.lo[a] = b.lo[a] < bound.lo[a] ? b.lo[a] : bound.lo[a];
                bound.hi[a] = b.hi[a] > bound.hi[a] ? b.hi[a] : bound.hi[a];
                const float c = 0.5f * (b.lo[a] + b.hi[a]);
                clo[a] = c < clo[a] ? c : clo[a];
                chi[a] = c > chi[a] ? c : chi[a];
            }
        }
        for (int a = 0; a < 3; ++a) { nodes[node].bmin[a] = bound.lo[a]; nodes[node].bmax[a] = bound.hi[a]; }
        int axis = 0;
        for (int a = 1; a < 3; ++a)
            if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
        if (count <= 4 || !(chi[axis] > clo[axis]))
        {
            nodes[node].left = first;
            nodes[node].count = count;
            return;
        }
        // median split on the centroids along the widest axis
        const int mid = first + count / 2;
        std::nth_element(prims.begin() + first, prims.begin() + mid, prims.begin() + first + count, [&](int x, int y) {
            return boxes[x].lo[axis] + boxes[x].hi[axis] < boxes[y].lo[axis] + boxes[y].hi[axis];
        });
        const int left = (int)nodes.size();
        nodes.push_back(BvhNode{});
        nodes.push_back(BvhNode{});
        nodes[node].left = left;
        nodes[node].count = 0;
        build(left, first, mid - first, depth + 1);
        build(left + 1, mid, first + count - mid, depth + 1);
    }
};

// large scene: per-surface data and the hierarchy go to one device allocation; d gets the pointers
int upload_big_scene(kyd_ctx* ctx, const kyd_scene_desc* sc, DevScene& d)
{
    const int n = sc->surface_count;
    std::vector<DevShape> shapes(n);
    std::vector<int> materials(n), lights(n);
    std::vector<Box> boxes(n);
    for (int i = 0; i < n; ++i)
    {
        const kyd_surface& s = sc->surfaces[i];
        shapes[i] = convert_shape(sc->shapes[s.shape]);
        materials[i] = s.material;
        lights[i] = s.area_light;
        boxes[i] = shape_box(shapes[i]);
    }
    BvhBuilder bvh(boxes);
    // bvh_query (kyd_device.cuh) walks with a fixed stack of KYD_BVH_STACK = 48 entries: popping a node at depth k
    // leaves at most k siblings below it, so depth + 2 entries suffice.  The median split halves the count per level
    // (4000 surfaces: depth 10); anything deeper is refused here rather than silently skipped there.
    if (bvh.max_depth + 2 > 48)
        return fail(ctx, KYD_ERR_LIMIT, "bounding-volume hierarchy deeper than the traversal stack");
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_shapes = 0, o_mat = align(o_shapes + sizeof(DevShape) * n), o_light = align(o_mat + sizeof(int) * n),
                 o_nodes = align(o_light + sizeof(int) * n), o_prims = align(o_nodes + sizeof(BvhNode) * bvh.nodes.size()),
                 total = align(o_prims + sizeof(int) * n);
    std::vector<unsigned char> host(total, 0);
    memcpy(host.data() + o_shapes, shapes.data(), sizeof(DevShape) * n);
    memcpy(host.data() + o_mat, materials.data(), sizeof(int) * n);
    memcpy(host.data() + o_light, lights.data(), sizeof(int) * n);
    memcpy(host.data() + o_nodes, bvh.nodes.data(), sizeof(BvhNode) * bvh.nodes.size());
    memcpy(host.data() + o_prims, bvh.prims.data(), sizeof(int) * n);
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));
    void* fresh = nullptr;
    KYD_CUDA(ctx, cudaMalloc(&fresh, total));
    cudaError_t e = cudaMemcpy(fresh, host.data(), total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();   // kernels of an earlier render may still read the previous scene
    if (e != cudaSuccess) { cudaFree(fresh); KYD_CUDA(ctx, e); }
    if (ctx->big_scene_dev) cudaFree(ctx->big_scene_dev);
    ctx->big_scene_dev = fresh;
    const char* base = (const char*)ctx->big_scene_dev;
    d.big_shape = (const DevShape*)(base + o_shapes);
    d.big_material = (const int*)(base + o_mat);
    d.big_light = (const int*)(base + o_light);
    d.bvh_nodes = (const BvhNode*)(base + o_nodes);
    d.bvh_prims = (const int*)(base + o_prims);
    for (int l = 0; l < KYD_MAX_LIGHTS; ++l)
        d.light_surface[l] = -1;
    for (int i = 0; i < n; ++i)
        if (lights[i] >= 0)
            d.light_surface[lights[i]] = d.light_surface[lights[i]] == -1 ? i : -2;
    return KYD_OK;
}

} // namespace

namespace {
// builds the device form of the scene into a local copy and commits it to the context only when everything validated: a
// failed upload leaves the previous scene (host plan AND device symbol) in force
int upload_scene_one(kyd_ctx* ctx, const kyd_scene_desc* sc)
{
    if (!ctx) return KYD_ERR_INVALID;
    if (!sc) return fail(ctx, KYD_ERR_INVALID, "scene desc is null");
    if (sc->surface_count < 0 || sc->surface_count > KYD_MAX_SURFACES_BVH || sc->shape_count < 0 || sc->shape_count > KYD_MAX_SHAPES_BVH ||
        sc->material_count < 0 || sc->material_count > KYD_MAX_MATERIALS || sc->light_count < 0 || sc->light_count > KYD_MAX_LIGHTS)
        return fail(ctx, KYD_ERR_LIMIT, "scene exceeds KYD_MAX_* limits");
    const bool big = sc->surface_count > KYD_MAX_SURFACES;   // global memory + bounding-volume hierarchy instead of constant memory

    std::vector<DevScene> storage(1);
    DevScene& d = storage[0];
    memset(&d, 0, sizeof(d));
    d.camera.position = f3(sc->camera.position);
    d.camera.front = f3(sc->camera.front);
    d.camera.right = f3(sc->camera.right);
    d.camera.up = f3(sc->camera.up);
    d.camera.res_x = sc->camera.resolution[0];
    d.camera.res_y = sc->camera.resolution[1];
    d.camera.push = sc->camera.origin_push;
    d.n_surfaces = sc->surface_count;
    d.n_lights = sc->light_count;
    d.env_light = sc->environment_light;
    if (d.env_light >= sc->light_count) return fail(ctx, KYD_ERR_INVALID, "environment_light index out of range");

    for (int i = 0; i < sc->surface_count; ++i)
    {
        const kyd_surface& s = sc->surfaces[i];
        if (s.shape < 0 || s.shape >= sc->shape_count || s.material < 0 || s.material >= sc->material_count ||
            s.area_light < -1 || s.area_light >= sc->light_count)
            return fail(ctx, KYD_ERR_INVALID, "surface index out of range");
        if (sc->shapes[s.shape].kind < 0 || sc->shapes[s.shape].kind > KYD_SHAPE_DISK)
            return fail(ctx, KYD_ERR_INVALID, "unknown shape kind");
        if (big)
            continue;
        d.surf_shape[i] = convert_shape(sc->shapes[s.shape]);
        d.surf_material[i] = s.material;
        d.surf_light[i] = s.area_light;
    }
    if (!big)
    {
        // traversal copy grouped by kind; list order inside a group (ties are resolved by surface index, see scene_closest)
        const int group_kind[4] = { KYD_SHAPE_RECTANGLE, KYD_SHAPE_SPHERE, KYD_SHAPE_TRIANGLE, KYD_SHAPE_DISK };
        int k = 0;
        for (int g = 0; g < 4; ++g)
        {
            for (int i = 0; i < sc->surface_count; ++i)
                if (d.surf_shape[i].kind == group_kind[g])
                {
                    d.sorted_shape[k] = d.surf_shape[i];
                    d.sorted_surface[k] = i;
                    ++k;
                }
            d.kind_end[g] = k;
        }
        for (int l = 0; l < KYD_MAX_LIGHTS; ++l)
            d.light_surface[l] = -1;
        for (int i = 0; i < sc->surface_count; ++i)
            if (d.surf_light[i] >= 0)
                d.light_surface[d.surf_light[i]] = d.light_surface[d.surf_light[i]] == -1 ? i : -2;
        build_rect_classifiers(d);   // two-phase traversal (rectangles are group 0 of the sorted copy)
    }
    for (int i = 0; i < sc->material_count; ++i)
    {
        const kyd_material& m = sc->materials[i];
        if (m.kind < 0 || m.kind > KYD_MAT_PLASTIC) return fail(ctx, KYD_ERR_INVALID, "unknown material kind");
        d.materials[i] = convert_material(m);
    }
    d.n_nondelta_lights = 0;
    for (int i = 0; i < sc->light_count; ++i)
    {
        const kyd_light& l = sc->lights[i];
        if (l.kind < 0 || l.kind > KYD_LIGHT_ENVIRONMENT) return fail(ctx, KYD_ERR_INVALID, "unknown light kind");
        DevLight& o = d.lights[i];
        o.kind = l.kind;
        o.color = f3(l.color); o.position = f3(l.position); o.direction = f3(l.direction);
        o.world_radius = l.world_radius;
        o.shape = l.shape;
        if (l.kind == KYD_LIGHT_AREA)
        {
            if (l.shape < 0 || l.shape >= sc->shape_count) return fail(ctx, KYD_ERR_INVALID, "area light shape index out of range");
            d.light_shape[i] = convert_shape(sc->shapes[l.shape]);
        }
        if (l.kind == KYD_LIGHT_AREA || l.kind == KYD_LIGHT_ENVIRONMENT) d.n_nondelta_lights++;
    }
    // Where the occlusion form pays (A/B in profiles/r01_ab_variants.txt): sphere area lights, whose pdf_Li is positive
    // for every direction (ky.cpp:1509-1512) so that every BSDF-sampled query would otherwise be traced, in scenes with
    // several lights, where each traced query also costs a sector of a light-sampling line.  Elsewhere (one light:
    // queries are traced inside shade; rectangle lights: pdf_Li already tests the hit) the closest-hit form is as fast.
    // A scene's single area light (whatever its shape) uses it as well: shade then traces both of a vertex' queries with ONE
    // copy of the occlusion code, which matters more than the query's cost (the shade kernels are instruction-fetch bound).
    for (int l = 0; l < KYD_MAX_LIGHTS; ++l)
    {
        const bool area = l < sc->light_count && sc->lights[l].kind == KYD_LIGHT_AREA;
        const bool sphere_area = area && d.light_shape[l].kind == KYD_SHAPE_SPHERE;
        if (!((sphere_area && sc->light_count > 1) || (area && sc->light_count == 1 && !big)))
            d.light_surface[l] = -2;
    }
    if (big)
    {
        // last step that can fail; it replaces the device allocation only once the new one is filled
        int rc = upload_big_scene(ctx, sc, d);
        if (rc != KYD_OK) return rc;
        for (int l = 0; l < KYD_MAX_LIGHTS; ++l)
        {
            const bool sphere_area = l < sc->light_count && sc->lights[l].kind == KYD_LIGHT_AREA && d.light_shape[l].kind == KYD_SHAPE_SPHERE;
            if (!(sphere_area && sc->light_count > 1))
                d.light_surface[l] = -2;
        }
    }
    else if (ctx->big_scene_dev)
    {
        // back to a constant-memory scene: the large scene's allocation is no longer needed
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        cudaFree(ctx->big_scene_dev);
        ctx->big_scene_dev = nullptr;
    }
    ctx->scene = d;
    ctx->has_scene = true;
    ctx->scene_generation++;
    return KYD_OK;
}
} // namespace

int kyd_upload_scene(kyd_ctx* ctx, const kyd_scene_desc* sc)
{
    if (!ctx) return KYD_ERR_INVALID;
    int rc = upload_scene_one(ctx, sc);
    for (kyd_ctx* peer : ctx->peers)
    {
        if (rc != KYD_OK) break;
        rc = upload_scene_one(peer, sc);
        if (rc != KYD_OK) ctx->error = peer->error;
    }
    return rc;
}

namespace {

// ---- multi-GPU context (kyd_create_multi) ------------------------------------------------------------------------
// The path shards by SAMPLE INDEX: rank r of n renders a contiguous share of [sample_begin, sample_end)
// for every pixel into an unclamped partial film on its own GPU (each sample already weighs 1/spp), rank 0 then adds
// the partials to its own in rank order -- one kernel on rank 0's GPU that loads the peers' films through their NVLink
// peer mappings (or from a staging copy where peer access is not available) -- and clamps after the sum, because the
// reference clamps after the spp-sum (ky.cpp:3721-3726).  One host thread per GPU enqueues its rank's launches.
void sample_share(int begin, int end, int n, int r, int* b, int* e)
{
    const int total = end - begin, base = total / n, extra = total % n;
    *b = begin + r * base + (r < extra ? r : extra);
    *e = *b + base + (r < extra ? 1 : 0);
}

int multi_render(kyd_ctx* ctx, const kyd_render_desc* d, float* film_root, cudaStream_t stream)
{
    const int n = 1 + (int)ctx->peers.size();
    const size_t floats = (size_t)d->width * d->height * 3;
    std::vector<kyd_render_desc> share(n, *d);
    for (int r = 0; r < n; ++r)
    {
        sample_share(d->sample_begin, d->sample_end, n, r, &share[r].sample_begin, &share[r].sample_end);
        share[r].flags = d->flags & ~(uint32_t)KYD_FLAG_CLAMP;
        if (r > 0) share[r].flags &= ~(uint32_t)KYD_FLAG_ACCUMULATE;
    }
    std::vector<int> rcs(n, KYD_OK);
    std::vector<std::thread> pool;
    for (int r = 1; r < n; ++r)
    {
        if (share[r].sample_end == share[r].sample_begin) continue;   // nothing to add
        pool.emplace_back([&, r] {
            kyd_ctx* peer = ctx->peers[r - 1];
            cudaError_t e = cudaSetDevice(peer->device);
            if (e != cudaSuccess) { rcs[r] = fail(peer, KYD_ERR_CUDA, cudaGetErrorString(e)); return; }
            if ((rcs[r] = ensure_film(peer, floats)) != KYD_OK) return;
            if ((rcs[r] = render_to_device(peer, &share[r], peer->film_dev, peer->stream, true)) != KYD_OK) return;
            e = cudaEventRecord(ctx->peer_done[r - 1], peer->stream);
            if (e != cudaSuccess) rcs[r] = fail(peer, KYD_ERR_CUDA, cudaGetErrorString(e));
        });
    }
    rcs[0] = render_to_device(ctx, &share[0], film_root, stream, true);
    for (auto& t : pool) t.join();
    cudaSetDevice(ctx->device);
    for (int r = 0; r < n; ++r)
        if (rcs[r] != KYD_OK)
        {
            if (r > 0) ctx->error = "rank " + std::to_string(r) + ": " + ctx->peers[r - 1]->error;
            return rcs[r];
        }
    // the sum: rank order, peers' films read in place over their peer mappings
    const float* parts[KYD_MAX_MULTI] = {};
    int n_parts = 0;
    for (int r = 1; r < n; ++r)
    {
        if (share[r].sample_end == share[r].sample_begin) continue;
        kyd_ctx* peer = ctx->peers[r - 1];
        KYD_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->peer_done[r - 1], 0));
        const float* view = peer->film_dev;
        if (!ctx->peer_direct[r - 1])
        {
            if (ctx->peer_staging_capacity[r - 1] < floats)
            {
                if (ctx->peer_staging[r - 1]) cudaFree(ctx->peer_staging[r - 1]);
                ctx->peer_staging[r - 1] = nullptr;
                ctx->peer_staging_capacity[r - 1] = 0;
                KYD_CUDA(ctx, cudaMalloc(&ctx->peer_staging[r - 1], floats * sizeof(float)));
                ctx->peer_staging_capacity[r - 1] = floats;
            }
            KYD_CUDA(ctx, cudaMemcpyPeerAsync(ctx->peer_staging[r - 1], ctx->device, peer->film_dev, peer->device, floats * sizeof(float), stream));
            view = ctx->peer_staging[r - 1];
        }
        parts[n_parts++] = view;
    }
    launch_sum_partials(film_root, parts, n_parts, (int64_t)floats, (d->flags & KYD_FLAG_CLAMP) != 0, ctx->sm_count, stream);
    KYD_CUDA(ctx, cudaGetLastError());
    KYD_CUDA(ctx, cudaEventRecord(ctx->ev_end, stream));   // the job's interval on rank 0 ends behind the sum
    // the peers' film buffers are read by rank 0's stream: their next render must not overwrite them earlier
    KYD_CUDA(ctx, cudaEventRecord(ctx->sum_done, stream));
    for (kyd_ctx* peer : ctx->peers)
    {
        cudaSetDevice(peer->device);
        cudaStreamWaitEvent(peer->stream, ctx->sum_done, 0);
    }
    cudaSetDevice(ctx->device);
    ctx->stats.kernel_launches += 1;
    return KYD_OK;
}

int render_any(kyd_ctx* ctx, const kyd_render_desc* d, float* film_dev, cudaStream_t stream)
{
    if (!ctx->peers.empty())
        return multi_render(ctx, d, film_dev, stream);
    return render_to_device(ctx, d, film_dev, stream, true);
}

// folds the ranks' counters into rank 0's stats: sums, and the longest device interval
void finish_stats_all(kyd_ctx* ctx)
{
    const bool fresh = ctx->counters_in_flight;
    finish_stats(ctx);
    if (!fresh) return;
    for (kyd_ctx* peer : ctx->peers)
    {
        if (!peer->counters_in_flight) continue;
        cudaSetDevice(peer->device);
        cudaStreamSynchronize(peer->stream);
        finish_stats(peer);
        const kyd_stats& p = peer->stats;
        ctx->stats.rays += p.rays;
        ctx->stats.rays_traced += p.rays_traced;
        ctx->stats.kernel_launches += p.kernel_launches;
        ctx->stats.samples += p.samples;
        ctx->stats.shade_vertices += p.shade_vertices;
        ctx->stats.shade_light_lines += p.shade_light_lines;
        ctx->stats.intersect_rays += p.intersect_rays;
        for (int k = 0; k < 8; ++k)
            if (p.stage_ms[k] > ctx->stats.stage_ms[k]) ctx->stats.stage_ms[k] = p.stage_ms[k];
    }
    cudaSetDevice(ctx->device);
}

} // namespace

int kyd_create_multi(kyd_ctx** out_ctx, const int* devices, int n)
{
    if (!out_ctx) return fail(nullptr, KYD_ERR_INVALID, "out_ctx is null");
    *out_ctx = nullptr;
    if (!devices || n < 1 || n > KYD_MAX_MULTI) return fail(nullptr, KYD_ERR_INVALID, "device list must hold 1..KYD_MAX_MULTI ordinals");
    kyd_ctx* root = nullptr;
    int rc = kyd_create(&root, devices[0]);
    if (rc != KYD_OK) return rc;
    for (int r = 1; r < n; ++r)
    {
        kyd_ctx* peer = nullptr;
        rc = kyd_create(&peer, devices[r]);
        if (rc != KYD_OK) { kyd_destroy(root); return rc; }
        root->peers.push_back(peer);
        root->peer_staging.push_back(nullptr);
        root->peer_staging_capacity.push_back(0);
        // the partial film of a peer on another GPU is read in place when the GPUs are peers (NVLink / NVSwitch on a B200 box)
        bool direct = devices[r] == devices[0];
        if (!direct)
        {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[0], devices[r]);
            if (can)
            {
                cudaSetDevice(devices[0]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
                direct = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
                cudaGetLastError();
            }
        }
        root->peer_direct.push_back(direct);
        cudaEvent_t ev = nullptr;
        cudaSetDevice(devices[r]);
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
        {
            g_create_error = "cudaEventCreate failed";
            kyd_destroy(root);
            return KYD_ERR_CUDA;
        }
        root->peer_done.push_back(ev);
    }
    cudaSetDevice(devices[0]);
    if (n > 1 && cudaEventCreateWithFlags(&root->sum_done, cudaEventDisableTiming) != cudaSuccess)
    {
        g_create_error = "cudaEventCreate failed";
        kyd_destroy(root);
        return KYD_ERR_CUDA;
    }
    *out_ctx = root;
    return KYD_OK;
}

int kyd_device_count(const kyd_ctx* ctx) { return ctx ? 1 + (int)ctx->peers.size() : 0; }

int kyd_render(kyd_ctx* ctx, const kyd_render_desc* d, float* film_rgb)
{
    if (!ctx) return KYD_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, KYD_ERR_NO_SCENE, "kyd_render before kyd_upload_scene");
    if (!film_rgb) return fail(ctx, KYD_ERR_INVALID, "film pointer is null");
    int rc = validate(ctx, d);
    if (rc != KYD_OK) return rc;
    if (d->flags & KYD_FLAG_ACCUMULATE) return fail(ctx, KYD_ERR_INVALID, "KYD_FLAG_ACCUMULATE needs kyd_render_device");
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));

    const size_t floats = (size_t)d->width * d->height * 3;
    if ((rc = ensure_film(ctx, floats)) != KYD_OK) return rc;
    if ((rc = ensure_pinned(ctx, floats)) != KYD_OK) return rc;

    if ((rc = render_any(ctx, d, ctx->film_dev, ctx->stream)) != KYD_OK) return rc;
    if ((rc = download_film(ctx, ctx->film_dev, ctx->film_pinned, film_rgb, floats * sizeof(float))) != KYD_OK) return rc;
    finish_stats_all(ctx);
    return KYD_OK;
}

int kyd_render_device(kyd_ctx* ctx, const kyd_render_desc* d, float* film_rgb_device, void* cuda_stream)
{
    if (!ctx) return KYD_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, KYD_ERR_NO_SCENE, "kyd_render_device before kyd_upload_scene");
    if (!film_rgb_device) return fail(ctx, KYD_ERR_INVALID, "film pointer is null");
    int rc = validate(ctx, d);
    if (rc != KYD_OK) return rc;
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    if ((rc = render_any(ctx, d, film_rgb_device, stream)) != KYD_OK) return rc;
    if (!cuda_stream)
    {
        KYD_CUDA(ctx, cudaStreamSynchronize(stream));
        finish_stats_all(ctx);
    }
    return KYD_OK;
}

int kyd_clamp_device(kyd_ctx* ctx, float* film_rgb_device, int64_t n, void* cuda_stream)
{
    if (!ctx) return KYD_ERR_INVALID;
    if (!film_rgb_device || n < 0) return fail(ctx, KYD_ERR_INVALID, "bad film pointer / size");
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    if (n > 0) launch_clamp(film_rgb_device, n, stream);
    KYD_CUDA(ctx, cudaGetLastError());
    if (!cuda_stream) KYD_CUDA(ctx, cudaStreamSynchronize(stream));
    return KYD_OK;
}

int64_t kyd_film_body_bytes(int format, int width, int height)
{
    if (width <= 0 || height <= 0) return -1;
    if (format == KYD_FILM_GAMMA8 || format == KYD_FILM_BMP24) return (int64_t)width * height * 3;
    if (format == KYD_FILM_RGBE) return (int64_t)width * height * 4;
    return -1;
}

int kyd_film_header(int format, int width, int height, uint8_t* out, int cap)
{
    if (!out || kyd_film_body_bytes(format, width, height) < 0) return -1;
    char text[96];
    int n = 0;
    if (format == KYD_FILM_GAMMA8)
        n = snprintf(text, sizeof(text), "P3\n%d %d\n%d\n", width, height, 255);                       // ky.cpp:1673
    else if (format == KYD_FILM_RGBE)
        n = snprintf(text, sizeof(text), "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", height, width); // ky.cpp:1745-1748
    else
    {
        // ky.cpp:1692-1733: "BM" + 13 little-endian words; the file size counts padded lines although the body is unpadded
        const uint32_t padded = ((uint32_t)width * 3u + 3u) & ~3u;
        const uint32_t words[13] = { 14u + 40u + padded * (uint32_t)height, 0u, 54u, 40u, (uint32_t)width, (uint32_t)height,
                                     1u | (24u << 16), 0u, 0u, 0u, 0u, 0u, 0u };
        if (cap < 54) return -1;
        out[0] = 'B'; out[1] = 'M';
        memcpy(out + 2, words, sizeof(words));
        return 54;
    }
    if (n <= 0 || n > cap) return -1;
    memcpy(out, text, (size_t)n);
    return n;
}

int kyd_film_encode_device(kyd_ctx* ctx, const float* film_rgb_device, int width, int height, int format, uint8_t* out_body_device, void* cuda_stream)
{
    if (!ctx) return KYD_ERR_INVALID;
    if (kyd_film_body_bytes(format, width, height) < 0) return fail(ctx, KYD_ERR_INVALID, "unknown film format or non-positive film size");
    if (!film_rgb_device || !out_body_device) return fail(ctx, KYD_ERR_INVALID, "film / body pointer is null");
    if (((uintptr_t)film_rgb_device & 15u) || ((uintptr_t)out_body_device & 3u))
        return fail(ctx, KYD_ERR_INVALID, "film must be 16-byte aligned and the body 4-byte aligned");
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    KYD_CUDA(ctx, launch_film_encode(ctx->device, ctx->sm_count, film_rgb_device, width, height, format, out_body_device, stream));
    if (!cuda_stream) KYD_CUDA(ctx, cudaStreamSynchronize(stream));
    return KYD_OK;
}

int kyd_film_encode(kyd_ctx* ctx, const float* film_rgb, int width, int height, int format, uint8_t* out_body)
{
    if (!ctx) return KYD_ERR_INVALID;
    const int64_t bytes = kyd_film_body_bytes(format, width, height);
    if (bytes < 0) return fail(ctx, KYD_ERR_INVALID, "unknown film format or non-positive film size");
    if (!film_rgb || !out_body) return fail(ctx, KYD_ERR_INVALID, "film / body pointer is null");
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t floats = (size_t)width * height * 3;
    int rc;
    if ((rc = ensure_film(ctx, floats)) != KYD_OK) return rc;
    if ((rc = ensure_pinned(ctx, floats)) != KYD_OK) return rc;   // floats * 4 >= bytes: one bounce buffer serves both ways
    if ((size_t)bytes > ctx->body_capacity)
    {
        if (ctx->body_dev) cudaFree(ctx->body_dev);
        ctx->body_dev = nullptr;
        ctx->body_capacity = 0;
        KYD_CUDA(ctx, cudaMalloc(&ctx->body_dev, (size_t)bytes));
        ctx->body_capacity = (size_t)bytes;
    }
    memcpy(ctx->film_pinned, film_rgb, floats * sizeof(float));
    KYD_CUDA(ctx, cudaMemcpyAsync(ctx->film_dev, ctx->film_pinned, floats * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    KYD_CUDA(ctx, launch_film_encode(ctx->device, ctx->sm_count, ctx->film_dev, width, height, format, ctx->body_dev, ctx->stream));
    KYD_CUDA(ctx, cudaMemcpyAsync(ctx->film_pinned, ctx->body_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    KYD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out_body, ctx->film_pinned, (size_t)bytes);
    return KYD_OK;
}

int kyd_get_stats(kyd_ctx* ctx, kyd_stats* out)
{
    if (!ctx || !out) return KYD_ERR_INVALID;
    // an asynchronous kyd_render_device leaves the counters in flight: settle them
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    finish_stats_all(ctx);
    *out = ctx->stats;
    return KYD_OK;
}

int kyd_selftest(kyd_ctx* ctx, int which, uint64_t first, uint64_t count, uint64_t* out2)
{
    if (!ctx || !out2) return KYD_ERR_INVALID;
    if (which != KYD_SELFTEST_RSQRT && which != KYD_SELFTEST_POW && which != KYD_SELFTEST_TRAVERSAL)
        return fail(ctx, KYD_ERR_INVALID, "unknown self-test");
    if (which == KYD_SELFTEST_TRAVERSAL && (!ctx->has_scene || ctx->scene.bvh_nodes))
        return fail(ctx, KYD_ERR_NO_SCENE, "the traversal self-test needs an uploaded scene of at most KYD_MAX_SURFACES surfaces");
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long* dev = nullptr;
    KYD_CUDA(ctx, cudaMalloc(&dev, 2 * sizeof(unsigned long long)));
    cudaMemsetAsync(dev, 0, 2 * sizeof(unsigned long long), ctx->stream);
    if (which == KYD_SELFTEST_TRAVERSAL)
    {
        SceneSlot& slot = g_scene_slot[0][ctx->device & 63];
        std::lock_guard<std::mutex> lock(slot.launch);
        const int rc = bind_scene(ctx, slot, ctx->stream);
        if (rc != KYD_OK) { cudaFree(dev); return rc; }
        launch_selftest_traversal(first, count, dev, ctx->stream);
        cudaEventRecord(slot.last_use, ctx->stream);
        slot.last_stream = ctx->stream;
        slot.used = true;
    }
    else if (which == KYD_SELFTEST_RSQRT) launch_selftest_rsqrt(first, count, dev, ctx->stream);
    else launch_selftest_pow(first, count, dev, ctx->stream);
    unsigned long long host[2] = { 0, 0 };
    cudaError_t e = cudaMemcpyAsync(host, dev, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dev);
    KYD_CUDA(ctx, e);
    out2[0] = host[0];
    out2[1] = host[1];
    return KYD_OK;
}

int kyd_kat(kyd_ctx* ctx, int which, const void* object, int index, int traits, int n, const float* in, float* out)
{
    if (!ctx) return KYD_ERR_INVALID;
    static const int in_floats[7] = { 7, 8, 9, 14, 2, 11, 4 }, out_floats[7] = { 8, 7, 1, 13, 6, 11, 16 };
    if (which < 0 || which > KYD_KAT_SAMPLER) return fail(ctx, KYD_ERR_INVALID, "unknown known-answer kind");
    if (n < 0 || !in || !out) return fail(ctx, KYD_ERR_INVALID, "bad known-answer buffers");
    if (traits < 0 || traits > 2) return fail(ctx, KYD_ERR_INVALID, "traits must be 0, 1 or 2");
    const bool needs_object = which <= KYD_KAT_MATERIAL_BSDF, needs_scene = which == KYD_KAT_CAMERA_RAYS || which == KYD_KAT_LIGHT_SAMPLE;
    if (needs_object && !object) return fail(ctx, KYD_ERR_INVALID, "known-answer kind needs a shape / material");
    if (needs_scene && (!ctx->has_scene || ctx->scene.bvh_nodes)) return fail(ctx, KYD_ERR_NO_SCENE, "known-answer kind needs an uploaded scene of at most KYD_MAX_SURFACES surfaces");
    if (which == KYD_KAT_LIGHT_SAMPLE && (index < 0 || index >= ctx->scene.n_lights)) return fail(ctx, KYD_ERR_INVALID, "light index out of range");
    if (n == 0) return KYD_OK;
    DevShape shape{};
    DevMaterial material{};
    if (which <= KYD_KAT_SHAPE_PDF_DIRECTION)
    {
        const kyd_shape* s = (const kyd_shape*)object;
        if (s->kind < 0 || s->kind > KYD_SHAPE_DISK) return fail(ctx, KYD_ERR_INVALID, "unknown shape kind");
        shape = convert_shape(*s);
    }
    else if (which == KYD_KAT_MATERIAL_BSDF)
    {
        const kyd_material* m = (const kyd_material*)object;
        if (m->kind < 0 || m->kind > KYD_MAT_PLASTIC) return fail(ctx, KYD_ERR_INVALID, "unknown material kind");
        material = convert_material(*m);
    }
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));
    float *in_dev = nullptr, *out_dev = nullptr;
    const size_t in_bytes = sizeof(float) * in_floats[which] * (size_t)n, out_bytes = sizeof(float) * out_floats[which] * (size_t)n;
    KYD_CUDA(ctx, cudaMalloc(&in_dev, in_bytes));
    cudaError_t e = cudaMalloc(&out_dev, out_bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(in_dev, in, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(out_dev, 0, out_bytes, ctx->stream);
    int rc = KYD_OK;
    if (e == cudaSuccess)
    {
        SceneSlot& slot = g_scene_slot[0][ctx->device & 63];
        std::lock_guard<std::mutex> lock(slot.launch);
        if (needs_scene) rc = bind_scene(ctx, slot, ctx->stream);
        if (rc == KYD_OK)
        {
            launch_kat(which, shape, material, index, traits, n, in_dev, out_dev, ctx->stream);
            e = cudaGetLastError();
            if (needs_scene && slot.last_use)
            {
                cudaEventRecord(slot.last_use, ctx->stream);
                slot.last_stream = ctx->stream;
                slot.used = true;
            }
        }
    }
    if (e == cudaSuccess && rc == KYD_OK) e = cudaMemcpyAsync(out, out_dev, out_bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(in_dev);
    cudaFree(out_dev);
    if (rc != KYD_OK) return rc;
    KYD_CUDA(ctx, e);
    return KYD_OK;
}

int kyd_render_smallpt_f64(kyd_ctx* ctx, int width, int height, int samples_per_pixel, double* film_rgb)
{
    if (!ctx) return KYD_ERR_INVALID;
    if (width <= 0 || height <= 0 || height > 65535 || samples_per_pixel <= 0) return fail(ctx, KYD_ERR_INVALID, "bad film size / sample count");
    if (!film_rgb) return fail(ctx, KYD_ERR_INVALID, "film pointer is null");
    KYD_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * 3 * (size_t)width * height;
    double* dev = nullptr;
    KYD_CUDA(ctx, cudaMalloc(&dev, bytes));
    cudaError_t e = cudaEventRecord(ctx->ev_begin, ctx->stream);
    if (e == cudaSuccess) e = launch_smallpt_f64(width, height, samples_per_pixel, dev, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_end, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(film_rgb, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dev);
    KYD_CUDA(ctx, e);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end);
    ctx->stats = kyd_stats{};
    ctx->stats.samples = (uint64_t)width * height * samples_per_pixel;
    ctx->stats.kernel_launches = 1;
    ctx->stats.device_ms = ms;
    return KYD_OK;
}

int kyd_set_wave_paths(kyd_ctx* ctx, int64_t paths)
{
    if (!ctx) return KYD_ERR_INVALID;
    if (paths < 0) return fail(ctx, KYD_ERR_INVALID, "wave size must be >= 0");
    ctx->wave_paths = paths;
    for (kyd_ctx* peer : ctx->peers) peer->wave_paths = paths;
    return KYD_OK;
}

} // extern "C"
