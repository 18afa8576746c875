// kyd_internal.h -- host-side declarations shared by kyd_api.cu and kyd_kernels.cu
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <vector>

#include "kyd_scene.h"

namespace kyd {

// device-side counters, one block per context
struct DevCounters
{
    unsigned long long rays;        // scene queries the reference issues for the same samples
    unsigned long long rays_traced; // scene queries actually traversed on the device
    unsigned long long queue[16];   // wavefront queue tails (layout: kyd_wavefront.cuh)
    unsigned long long shade_vertices, shade_lines; // wavefront shade: vertices shaded, light-sampling lines written
    unsigned long long intersect_rays;              // closest-hit queries traversed by k_intersect (its share of rays_traced)
};

// wavefront buffers (device memory owned by the context), capacity = paths per wave
struct WaveBuffers
{
    int64_t capacity = 0;
    int max_lights = 0;           // lights the light-sampling buffer was sized for (0: not allocated)
    int nee_units = 0;            // float4 units per (light, path slot) of that buffer: 8 = a 128-byte line, 1 = a k_nee result
    bool has_vertex = false;
    // layouts: kyd_wavefront.cuh
    float4* path = nullptr;       // one 64-byte record (4 float4) per path slot
    float4* vertex = nullptr;     // 6 float4 per path slot: vertex record of the split light-sample stage
    float4* nee = nullptr;        // one 128-byte line per (light, path slot): the two NEE queries, vertex beta, result
    float4* levels = nullptr;     // recursive integrators: 2 float4 per (level, path slot), level-major: {Lo.rgb, |cos|}, {f.rgb, pdf}
    int max_levels = 0;           // levels that buffer was sized for (0: not allocated)
    // queues of path slots
    int* queue_a = nullptr;        // ray queues (ping-pong by bounce parity)
    int* queue_b = nullptr;
    int* queue_lobe[2][4] = {};    // hit paths sorted by BSDF lobe: Lambert, mirror, glass, Phong (LOBE_* order); [bounce parity]:
                                   // with the intersect stage fused into shade, shade reads one set while it fills the other
    int* queue_nee[2] = {};        // vertices for the stand-alone light-sample stage (split mode): Lambert, Phong
    int* queue_pair[2] = {};       // (vertex, light) pairs with a light-sampling line: slot | light << 24; capacity * lights entries
};

// optional per-stage device timing (KYD_STAGE_TIMING=1): CUDA events around every kernel of the wavefront
struct StageTimer
{
    enum { RAYGEN = 0, INTERSECT = 1, SHADE = 2, LIGHT_SAMPLE = 3, SHADOW = 4, ACCUMULATE = 6, PIXEL = 7 };
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> events; // pairs
    std::vector<int> stages;
    size_t used = 0;
    void begin(int stage);
    void end();
    void collect(double* stage_ms); // synchronises
    ~StageTimer();
};

void upload_scene_constant(const DevScene& scene, cudaStream_t stream);

// megakernel path: one thread per pixel, samples in order (used for every integrator; the only path
// for the debug, direct-lighting and recursive integrators)
void launch_render_pixels(const RenderParams& rp, float* film_dev, DevCounters* counters, cudaStream_t stream);

void launch_clamp(float* film_dev, int64_t n, cudaStream_t stream);

// film[i] = clamp?(((film[i] + parts[0][i]) + parts[1][i]) + ...): the end of a multi-GPU job on rank 0's device; the
// parts are the other ranks' partial films, addressable from this device (peer mappings or staging copies)
void launch_sum_partials(float* film_dev, const float* const* parts, int n_parts, int64_t n, bool clamp, int sm_count, cudaStream_t stream);

void launch_selftest_pow(unsigned long long first, unsigned long long count, unsigned long long* out_dev, cudaStream_t stream);
void launch_kat(int which, const DevShape& shape, const DevMaterial& material, int index, int traits, int n, const float* in_dev, float* out_dev, cudaStream_t stream);
void launch_selftest_traversal(unsigned long long first, unsigned long long count, unsigned long long* out_dev, cudaStream_t stream);
void launch_selftest_rsqrt(unsigned long long first, unsigned long long count, unsigned long long* out_dev, cudaStream_t stream);

// film output stage (kyd_film.cu): float film -> body bytes of `format` (kyd_film_format); film_dev 16-byte and
// out_dev 4-byte aligned
cudaError_t launch_film_encode(int device, int sm_count, const float* film_dev, int width, int height, int format, uint8_t* out_dev, cudaStream_t stream);

// FP64 smallpt validation mode (kyd_smallpt.cu)
cudaError_t launch_smallpt_f64(int width, int height, int samples_per_pixel, double* film_dev, cudaStream_t stream);

// which wavefront kernels a render uses (decides the buffers it needs, too)
struct WavefrontPlan
{
    bool hot;             // headline configuration: path_tracing_iteration_t, both_mis, no debug sampler, fused light-sample
    bool inline_queries;  // hot and one light: shade traces its own light queries (no light-sampling lines, no shadow stage)
    bool nee;             // the shadow stage runs (light-sampling lines are written)
    bool split;           // KYD_FLAG_SPLIT_LIGHT_SAMPLE: vertex records for the stand-alone light-sample kernel
    bool pair_kernel;     // hot with several lights: the light loop runs in k_nee over (vertex, light) pairs (vertex records + 16-byte results)
    bool recursion;       // one of the three recursive integrators: forward pass with a per-level record, then k_unwind
    bool fused;           // headline configurations unless KYD_FUSE_INTERSECT=0: shade also traces the path's next ray (closest hit, lobe
                          // classification), so only the camera rays go through the intersect kernel
};
inline WavefrontPlan wavefront_plan(const RenderParams& rp, const DevScene& scene)
{
    WavefrontPlan p;
    const bool direct_only = rp.integrator == KYD_INT_DIRECT_LIGHTING;
    p.split = (rp.flags & KYD_FLAG_SPLIT_LIGHT_SAMPLE) != 0;
    // (large scenes run the general kernels only: the specialised ones are not built for them)
    p.recursion = rp.integrator == KYD_INT_SIMPLE_PT_RECURSION || rp.integrator == KYD_INT_PT_RECURSION || rp.integrator == KYD_INT_PT_RECURSION_DEFERED;
    p.hot = !p.recursion && !direct_only && rp.direct_sample == KYD_DS_BOTH_MIS && rp.sampler != KYD_SAMPLER_DEBUG && !p.split && scene.bvh_nodes == nullptr;
    p.inline_queries = p.hot && scene.n_lights == 1;
    p.nee = rp.direct_sample != KYD_DS_IDLE && scene.n_lights > 0 && !p.inline_queries && !p.recursion;   // (recursion: light queries inside shade)
    p.pair_kernel = p.hot && p.nee;
    static const bool fuse_enabled = !(getenv("KYD_FUSE_INTERSECT") && getenv("KYD_FUSE_INTERSECT")[0] == '0');
    static const bool fuse_many_enabled = !(getenv("KYD_FUSE_INTERSECT_MANY") && getenv("KYD_FUSE_INTERSECT_MANY")[0] == '0');
    // Several lights (k_nee): a vertex' light values are pending when its continuation is traced, but a ray that leaves the scene
    // adds the environment's light only after a SPECULAR vertex (ky.cpp:4548-4563), and those sample no lights -- nothing is
    // pending then; a Lambert / Phong vertex' path that leaves the scene just ends, and k_accumulate adds its pending values.
    p.fused = (p.inline_queries && fuse_enabled) || (p.pair_kernel && fuse_enabled && fuse_many_enabled);
    return p;
}

void free_wave_buffers(WaveBuffers& w);
// (re)allocates the wavefront buffers for `capacity` path slots; light-sampling lines for `nee_lights` lights (0: none
// needed) and vertex records only if `vertex`; returns a cudaError_t
int ensure_wave_buffers(WaveBuffers& w, int64_t capacity, int nee_lights, int nee_units, bool vertex, int levels = 0);

// wavefront path: path_tracing_iteration_t and direct_lighting_t.  Adds to `launches` the kernels it launched.
void launch_render_wavefront(const RenderParams& rp, const DevScene& scene, WaveBuffers& w, int64_t wave_paths, float* film_dev, DevCounters* counters,
                             cudaStream_t stream, int sm_count, uint64_t* launches, StageTimer* timer);

} // namespace kyd

// the large-scene build of the kernels (kyd_kernels_big.cu): same entry points over a scene in global memory + BVH
namespace kyd_big {
void upload_scene_constant(const kyd::DevScene& scene, cudaStream_t stream);
void launch_render_pixels(const kyd::RenderParams& rp, float* film_dev, kyd::DevCounters* counters, cudaStream_t stream);
void launch_render_wavefront(const kyd::RenderParams& rp, const kyd::DevScene& scene, kyd::WaveBuffers& w, int64_t wave_paths, float* film_dev,
                             kyd::DevCounters* counters, cudaStream_t stream, int sm_count, uint64_t* launches, kyd::StageTimer* timer);
} // namespace kyd_big
