// kyd_internal.h -- host-side declarations shared by kyd_api.cu and kyd_kernels.cu
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "kyd_scene.h"

namespace kyd {

// device-side counters, one block per context
struct DevCounters
{
    unsigned long long rays;        // scene queries the reference issues for the same samples
    unsigned long long rays_traced; // scene queries actually traversed on the device
    unsigned long long queue[8];    // wavefront queue tails (reset per use)
};

// wavefront buffers (device memory owned by the context), capacity = paths per wave
struct WaveBuffers
{
    int64_t capacity = 0;
    int max_lights = 0;
    // path state, one entry per path slot
    float4* ray_o = nullptr;      // origin.xyz, tmax
    float4* ray_d = nullptr;      // direction.xyz, -
    float2* hit = nullptr;        // distance, surface index (int bits; -1 = miss)
    float4* beta = nullptr;       // throughput.rgb, flags (bit 0 previous vertex specular, bits 8.. bounce)
    float4* radiance = nullptr;   // Lo.rgb, -
    uint2* rng = nullptr;         // 48-bit LCG state
    // vertex record written by shade for the light-sample stage
    float4* vx_position = nullptr; // position.xyz, lobe
    float4* vx_normal = nullptr;   // isect normal.xyz, exponent
    float4* vx_wo = nullptr;       // wo.xyz, eta_t
    float4* vx_color = nullptr;    // lobe colour a.rgb, -
    float4* vx_beta = nullptr;     // throughput at the vertex (for the deferred Lo += beta * Ld), pending flag
    uint2* vx_rng = nullptr;       // sampler state before the vertex' light loop
    // NEE queries: two per (vertex, light)
    float4* nee_o = nullptr;       // origin.xyz, tmax
    float4* nee_d = nullptr;       // direction.xyz, bit 0: the reference issues this query
    float4* nee_value = nullptr;   // contribution if the query succeeds
    float4* nee_result = nullptr;  // per (light, slot): Ld of that light
    // queues of path slots
    int* queue_a = nullptr;
    int* queue_b = nullptr;
    int* queue_nee = nullptr;
};

void upload_scene_constant(const DevScene& scene, cudaStream_t stream);

// megakernel path: one thread per pixel, samples in order (used for every integrator; the only path
// for the debug, direct-lighting and recursive integrators)
void launch_render_pixels(const RenderParams& rp, float* film_dev, DevCounters* counters, cudaStream_t stream);

void launch_clamp(float* film_dev, int64_t n, cudaStream_t stream);

void free_wave_buffers(WaveBuffers& w);
// (re)allocates the wavefront buffers for `capacity` path slots and `lights` lights; returns a cudaError_t
int ensure_wave_buffers(WaveBuffers& w, int64_t capacity, int lights);

// wavefront path: path_tracing_iteration_t and direct_lighting_t.  Adds to `launches` the kernels it launched.
void launch_render_wavefront(const RenderParams& rp, const DevScene& scene, WaveBuffers& w, int64_t wave_paths, float* film_dev, DevCounters* counters,
                             cudaStream_t stream, int sm_count, uint64_t* launches);

} // namespace kyd
