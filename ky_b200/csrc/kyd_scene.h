// kyd_scene.h -- PODs shared by the host side (kyd_api.cu) and the device side (kyd_kernels.cu):
// the flattened scene as it sits in constant memory, and the per-launch render parameters.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "kyd.h"

namespace kyd {

// deepest max_depth the stack-based recursive integrators accept
#define KYD_MAX_RECURSION 32

// ---- flattened scene in constant memory ---------------------------------------------------------
struct DevShape
{
    float3 p0, p1, p2, p3, n; // meaning per kind as in kyd_shape
    float radius, radius_sq, area;
    int kind;
};

struct DevMaterial
{
    int kind;
    float3 diffuse, specular, transmission;
    float eta, exponent, p_diffuse, p_specular;
    float3 plastic_lambert; // diffuse / p_diffuse  (ky.cpp:2670), three IEEE divisions done at upload
    float3 plastic_phong;   // specular / p_specular (ky.cpp:2666)
};

struct DevLight
{
    int kind;
    float3 color, position, direction;
    float world_radius;
    int shape;
};

struct DevCamera
{
    float3 position, front, right, up;
    float res_x, res_y, push;
};

// node of the bounding-volume hierarchy of a large scene: an inner node's children are nodes[left] and nodes[left + 1];
// a leaf (count > 0) covers bvh_prims[left .. left + count)
struct BvhNode
{
    float bmin[3];
    int left;
    float bmax[3];
    int count;
};

// Conservative classifiers of the rectangles (kyd_device.cuh, "two-phase traversal"), precomputed in double at upload.
// General form: the rectangle as the parallelogram b + u eu + v ev in its plane.  A rectangle whose four points are not a
// planar parallelogram to 2^-20 of its diagonal gets n = 0: its classification is always "candidate", i.e. every ray
// takes the reference's own test.
struct RectCull
{
    float3 b; float c_area;    // corner p1; 2^-15 / area
    float3 n; int pad0;        // unit normal (0: exact test only)
    float3 gu; int pad1;       // dual basis: u = gu . (p - b), v = gv . (p - b)
    float3 gv; int pad2;
};
// Axis-aligned form (every Cornell wall): normal along axis A, the rectangle spans [lo_b, lo_b + 1/gb] x [lo_c, lo_c + 1/gc]
// along axes B = (A + 1) % 3 and C = (A + 2) % 3 in the plane x_A = pa
struct RectAligned
{
    float pa, lo_b, lo_c, gb;
    float gc, c_area; int pad0, pad1;
};

struct DevScene
{
    DevCamera camera;
    int n_surfaces, n_lights, env_light;
    int n_nondelta_lights;
    DevShape surf_shape[KYD_MAX_SURFACES];  // geometry of surface i, copied from its shape: traversal order
    int surf_material[KYD_MAX_SURFACES];
    int surf_light[KYD_MAX_SURFACES];
    DevMaterial materials[KYD_MAX_MATERIALS];
    DevLight lights[KYD_MAX_LIGHTS];
    DevShape light_shape[KYD_MAX_LIGHTS];   // area_light_t::shape_, independent of the surface list
    // traversal copy of the surfaces' geometry grouped by shape kind (rectangle, sphere, triangle, disk), list order
    // kept inside a group: one branch-free loop per kind whose loads are warp-uniform (kyd_device.cuh, scene_closest)
    int light_surface[KYD_MAX_LIGHTS];      // occlusion form of the BSDF-sampled query (nee_bsdf_setup): the one surface whose
                                            // area_light is light l; -1: none carries it; -2: form not used for this light
    DevShape sorted_shape[KYD_MAX_SURFACES];
    int sorted_surface[KYD_MAX_SURFACES];   // surface index of sorted_shape[k]
    int kind_end[4];                        // sorted_shape[kind_end[g-1] .. kind_end[g]) is group g
    // two-phase traversal: the first n_rect_cull = min(kind_end[0], 32) rectangles of the sorted copy, regrouped as
    // aligned-x | aligned-y | aligned-z | general; bit k of the traversal's masks is position k of that order
    RectAligned rect_aligned[32];           // positions [0, rect_aligned_end[2])
    RectCull rect_general[32];              // positions rect_aligned_end[2] + j, j < n_rect_general
    int rect_aligned_end[3];                // aligned rectangles with normal axis x: [0, end[0]); y: [end[0], end[1]); z: [end[1], end[2])
    int n_rect_general;
    int n_rect_cull;                        // = rect_aligned_end[2] + n_rect_general
    int rect_order[32];                     // sorted_shape index of position k
    float3 bound_center; float bound_l1;    // every vertex v of those rectangles: |v - bound_center|_1 <= bound_l1
    // large scenes (more than KYD_MAX_SURFACES surfaces): per-surface data and the hierarchy in global memory, the arrays
    // above unused; null pointers otherwise
    const DevShape* big_shape;
    const int* big_material;
    const int* big_light;
    const BvhNode* bvh_nodes;
    const int* bvh_prims;                   // surface indices in leaf order
};

struct RenderParams
{
    int width, height;
    int spp;
    int sample_begin, sample_end;
    int integrator, max_depth, direct_sample, lighting, sampler;
    unsigned long long seed;
    unsigned flags;
    float weight; // (float)(1.0 / spp), ky.cpp:3717
};

} // namespace kyd
