/*
 * kyd.h -- C ABI of libkyd.so, the B200 (sm_100a) rendering core for ky's per-pixel Monte Carlo
 * integrator loop.
 *
 * Drop-in boundary: everything below replaces the BODY of
 *     void integrator_t::render(scene_t*, sampler_t*, film_t*)        reference ky.cpp:3689-3729
 * (pixel loop + spp loop + virtual Li() + scene_t::intersect + BSDF / light sampling), and nothing
 * else.  The host side keeps ky's class surface (include/ky.hpp); its integrator_t::render()
 * flattens scene / camera / integrator parameters into the PODs declared here, calls
 * kyd_render(), and feeds the returned pixels to film_t::add_color() (ky.cpp:1586, 3726).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; the caller owns every pointer it passes in; the
 *     library copies what it needs during the call.
 *   - every function returns 0 on success, a KYD_ERR_* code otherwise; kyd_last_error()
 *     returns the message (the reference throws from LOG_ERROR, ky.cpp:75-82; the C++ host
 *     wrapper rethrows non-zero codes as std::runtime_error).
 *   - there is no CPU fallback: without a CUDA device every entry point fails with
 *     KYD_ERR_CUDA.
 *   - one host thread per context at a time (the reference's render() is not re-entrant
 *     either: it owns the film rows it writes, ky.cpp:3696-3728).
 */
#ifndef KYD_H
#define KYD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KYD_VERSION 1

enum kyd_error
{
    KYD_OK = 0,
    KYD_ERR_INVALID = 1,   /* bad argument / unsupported enum value */
    KYD_ERR_CUDA = 2,      /* CUDA runtime error, or no device */
    KYD_ERR_NO_SCENE = 3,  /* kyd_render before kyd_upload_scene */
    KYD_ERR_LIMIT = 4      /* scene larger than KYD_MAX_* */
};

enum { KYD_MAX_SURFACES = 64, KYD_MAX_SHAPES = 64, KYD_MAX_MATERIALS = 32, KYD_MAX_LIGHTS = 16 };
/* Scenes with more than KYD_MAX_SURFACES surfaces (up to KYD_MAX_SURFACES_BVH, e.g. small triangle meshes) leave constant
   memory: their surfaces live in global memory under a bounding-volume hierarchy built at upload -- the accelerator the
   reference leaves as an empty hook (accel_t, ky.cpp:3097-3115).  Results are those of the reference's linear walk
   (closest hit, lowest surface index among equal distances); smaller scenes keep the linear walk itself. */
enum { KYD_MAX_SURFACES_BVH = 4000, KYD_MAX_SHAPES_BVH = 4000 };

/* ---- flattened scene: one POD per reference object ------------------------------------------- */

/* shape_t subclasses, ky.cpp:1100-1519 */
enum kyd_shape_kind { KYD_SHAPE_SPHERE = 0, KYD_SHAPE_RECTANGLE = 1, KYD_SHAPE_TRIANGLE = 2, KYD_SHAPE_DISK = 3 };

typedef struct kyd_shape
{
    int32_t kind;
    float p0[3];      /* sphere: center_; disk: position_; triangle/rectangle: p0_ */
    float p1[3];      /* triangle/rectangle: p1_ */
    float p2[3];      /* triangle/rectangle: p2_ */
    float p3[3];      /* rectangle: p3_ */
    float normal[3];  /* rectangle/triangle/disk: stored normal_ (after flip_normal), ky.cpp:1105,1174,1256 */
    float radius;     /* sphere/disk radius_ */
    float radius_sq;  /* sphere radius_sq_ as the constructor rounds it, ky.cpp:1332 */
    float area;       /* shape_t::area(), ky.cpp:1141,1222,1304,1401 */
} kyd_shape;

/* material_t subclasses, ky.cpp:2579-2682 */
enum kyd_material_kind { KYD_MAT_MATTE = 0, KYD_MAT_MIRROR = 1, KYD_MAT_GLASS = 2, KYD_MAT_PLASTIC = 3 };

typedef struct kyd_material
{
    int32_t kind;
    float diffuse[3];       /* matte diffuse_color_; plastic diffuse_color_ */
    float specular[3];      /* mirror specular_color_; plastic specular_color_; glass reflection_color_ */
    float transmission[3];  /* glass transmission_color_ */
    float eta;              /* glass eta_ (eta_i is 1, ky.cpp:2630) */
    float exponent;         /* plastic exponent_ */
    float diffuse_probability;   /* plastic diffuse_probility_, ky.cpp:2657 */
    float specular_probability;  /* plastic specular_probility_, ky.cpp:2658 */
} kyd_material;

/* light_t subclasses, ky.cpp:2810-3062 */
enum kyd_light_kind { KYD_LIGHT_POINT = 0, KYD_LIGHT_DIRECTION = 1, KYD_LIGHT_AREA = 2, KYD_LIGHT_ENVIRONMENT = 3 };

typedef struct kyd_light
{
    int32_t kind;
    float color[3];      /* point intensity_; direction irradiance_; area/environment radiance_ */
    float position[3];   /* point light world_position_ */
    float direction[3];  /* direction light world_direction_ (normalised by its ctor, ky.cpp:2874) */
    float world_radius;  /* direction/environment world_radius_ from preprocess(), ky.cpp:3555-3574 */
    int32_t shape;       /* area light shape_ as an index into shapes[] (NOT derived from surfaces) */
} kyd_light;

/* surface_t, ky.cpp:3071-3089: three independent links */
typedef struct kyd_surface
{
    int32_t shape;
    int32_t material;
    int32_t area_light;  /* index into lights[] or -1 */
} kyd_surface;

/* camera_t after its constructor ran, ky.cpp:1864-1880 */
typedef struct kyd_camera
{
    float position[3];
    float front[3];       /* normalised */
    float right[3];       /* scaled by tan(fov/2) * aspect */
    float up[3];          /* scaled by tan(fov/2) */
    float resolution[2];  /* film_t::get_resolution(): the SUB-film size for film_grid_t, ky.cpp:1815 */
    float origin_push;    /* 0 for ky; 140 for the smallpt camera (smallpt_rewrite.cpp:676) */
} kyd_camera;

typedef struct kyd_scene_desc
{
    kyd_camera camera;
    int32_t shape_count;     const kyd_shape* shapes;
    int32_t material_count;  const kyd_material* materials;
    int32_t light_count;     const kyd_light* lights;       /* scene_t::light_list() order */
    int32_t surface_count;   const kyd_surface* surfaces;   /* scene_t surface_list_ order = traversal order */
    int32_t environment_light;                              /* scene_t::environment_light_ as light index or -1 */
} kyd_scene_desc;

/* ---- render request --------------------------------------------------------------------------- */

/* integrator_enum_t, ky.cpp:3625-3654 (same numeric values) */
enum kyd_integrator
{
    KYD_INT_POSITION = 0, KYD_INT_NORMAL = 1, KYD_INT_BASECOLOR = 2,   /* debug_integrator_t, ky.cpp:4094 */
    KYD_INT_DIRECT_LIGHTING = 6,                                        /* direct_lighting_t, ky.cpp:4125 */
    KYD_INT_SIMPLE_PT_RECURSION = 8,                                    /* ky.cpp:4191 */
    KYD_INT_PT_RECURSION = 9,                                           /* ky.cpp:4305 */
    KYD_INT_PT_RECURSION_DEFERED = 10,                                  /* ky.cpp:4409 */
    KYD_INT_PT_ITERATION = 11                                           /* ky.cpp:4523 */
};

/* direct_sample_enum_t, ky.cpp:3608-3623 (same numeric values) */
enum kyd_direct_sample
{
    KYD_DS_IDLE = 0, KYD_DS_BSDF = 4, KYD_DS_LIGHT = 8, KYD_DS_BSDF_MIS = 16, KYD_DS_LIGHT_MIS = 32, KYD_DS_BOTH_MIS = 48
};

/* lighting_enum_t, ky.cpp:3591-3603 (lighting filter of render_lighting_enum, ky.cpp:4907-4935) */
enum kyd_lighting
{
    KYD_LIGHTING_EMIT = 1, KYD_LIGHTING_DIRECT = 2, KYD_LIGHTING_INDIRECT = 4, KYD_LIGHTING_ALL = 31
};

enum kyd_sampler
{
    KYD_SAMPLER_LCG48 = 0,  /* counter-seeded 48-bit LCG, the contract of DESIGN.md "Sampling" */
    KYD_SAMPLER_DEBUG = 1,  /* debug_sampler_t: every draw is 0.5, ky.cpp:922-947 */
    KYD_SAMPLER_TRAPEZOIDAL = 2 /* LCG48 with tent-filtered 2x2 sub-pixel camera samples (smallpt's filter,
                               smallpt2pbrt/smallpt_rewrite.cpp:397-475, in float): spp counts all samples of a pixel and
                               must be a multiple of 4; sample s belongs to sub-pixel s / (spp / 4) */
};

enum kyd_render_flags
{
    KYD_FLAG_CLAMP = 1u,       /* film = clamp01(sum), what render() hands to add_color (ky.cpp:3726);
                                  without it the raw partial sum is returned (multi-GPU partials) */
    KYD_FLAG_FUSED = 2u,       /* one thread per pixel running the whole integrator instead of the wavefront stages */
    KYD_FLAG_ACCUMULATE = 4u,  /* device film: add to what the buffer holds instead of overwriting */
    KYD_FLAG_SPLIT_LIGHT_SAMPLE = 8u /* wavefront: light-sample (NEE + MIS set-up) as its own kernel over (vertex, light)
                                  pairs instead of inside shade; same results, kept for measurement */
};

typedef struct kyd_render_desc
{
    int32_t width, height;       /* film (or sub-film) resolution in pixels */
    int32_t spp;                 /* samples per pixel of the WHOLE job: every sample weighs 1/spp (ky.cpp:3717) */
    int32_t sample_begin;        /* this call renders sample indices [sample_begin, sample_end) of each pixel; */
    int32_t sample_end;          /* (0, spp) is the whole job; sub-ranges are what a multi-GPU split hands out */
    int32_t integrator;          /* enum kyd_integrator */
    int32_t max_depth;           /* path_integrator_t::max_path_depth_, ky.cpp:4182 */
    int32_t direct_sample;       /* enum kyd_direct_sample */
    int32_t lighting;            /* enum kyd_lighting; KYD_LIGHTING_ALL unless rendering render_lighting_enum panels */
    int32_t sampler;             /* enum kyd_sampler */
    uint64_t seed;               /* rng_t's seed, 1234 in the reference (ky.cpp:833) */
    uint32_t flags;              /* enum kyd_render_flags */
    uint32_t reserved;
} kyd_render_desc;

typedef struct kyd_stats
{
    uint64_t samples;           /* camera paths started by the last kyd_render call */
    uint64_t rays;              /* scene queries the reference issues for these samples (scene_t::intersect + occluded calls) */
    uint64_t rays_traced;       /* scene queries the device actually traversed (it skips those that cannot change the result) */
    uint64_t kernel_launches;   /* kernels launched by the last call */
    double device_ms;           /* CUDA-event time of the last call's kernels (excludes host copies) */
    double stage_ms[8];         /* raygen, intersect, shade, light_sample, shadow, -, accumulate, per-pixel kernel;
                                   filled when the context was created with KYD_STAGE_TIMING=1 in the environment */
    uint64_t shade_vertices;      /* wavefront: path vertices shaded */
    uint64_t shade_light_lines;   /* wavefront: light-sampling lines written by shade (64 B, 96 B with a live BSDF query) */
    uint64_t intersect_rays;      /* wavefront: closest-hit queries traversed by the intersect stage (part of rays_traced) */
} kyd_stats;

typedef struct kyd_ctx kyd_ctx;

/* creates a context on CUDA device `device` (cudaSetDevice ordinal) with its own stream.  The stream is a blocking one:
   work queued on the device's legacy default stream before a call is complete before the call's kernels run. */
int kyd_create(kyd_ctx** out_ctx, int device);

/* Multi-GPU context behind the same calls: `devices` lists n CUDA ordinals (1..KYD_MAX_MULTI; an
   ordinal may repeat).  kyd_upload_scene copies the scene to every device; kyd_render / kyd_render_device split the
   sample range [sample_begin, sample_end) evenly over the devices (every device renders its sample indices of EVERY
   pixel into an unclamped partial film, one host thread per device), the partial films are added on devices[0] in rank
   order by one kernel that reads the peers' films over their NVLink peer mappings, and KYD_FLAG_CLAMP is applied after
   that sum -- the reference clamps after the spp-sum (ky.cpp:3721-3726).  The film pointer of kyd_render_device lives
   on devices[0].  The re-associated per-pixel sum differs from the single-device film by rounding only (<= ~1e-6
   relative) and is the same from run to run.  kyd_get_stats reports sums over the devices (device_ms: devices[0]'s
   interval, which ends behind the final sum).  Everything else (film stage, self-tests) runs on devices[0]. */
enum { KYD_MAX_MULTI = 16 };
int kyd_create_multi(kyd_ctx** out_ctx, const int* devices, int n);
int kyd_device_count(const kyd_ctx* ctx);   /* devices behind the context (1 for kyd_create) */
void kyd_destroy(kyd_ctx* ctx);
const char* kyd_last_error(const kyd_ctx* ctx); /* ctx may be NULL: last creation error */

/* copies the flattened scene to the device (replaces the previous one) */
int kyd_upload_scene(kyd_ctx* ctx, const kyd_scene_desc* scene);

/* renders into HOST memory: film_rgb[height*width*3], row-major, y down (film_t::pixels_, ky.cpp:1574).
   The device->host copy is part of the call. */
int kyd_render(kyd_ctx* ctx, const kyd_render_desc* desc, float* film_rgb);

/* same, but film_rgb is a DEVICE pointer on the context's device (what a multi-GPU driver reduces
   with NCCL before the final clamp).  Runs asynchronously on `cuda_stream` (a cudaStream_t, or NULL
   for the context's own stream, in which case the call synchronises before returning). */
int kyd_render_device(kyd_ctx* ctx, const kyd_render_desc* desc, float* film_rgb_device, void* cuda_stream);

/* film[i] = clamp01(film[i]) on the device, n floats: the final step of a multi-GPU job after the reduce */
int kyd_clamp_device(kyd_ctx* ctx, float* film_rgb_device, int64_t n, void* cuda_stream);

int kyd_get_stats(kyd_ctx* ctx, kyd_stats* out);

/* device self-tests of exactness-critical fast paths.  KYD_SELFTEST_RSQRT: compares the FP32 fast path of
   vec3 normalize's reciprocal square root with its definition (float)(1.0 / sqrt((double)s)), ky.cpp:314, for
   the `count` float bit patterns starting at `first`; out2[0] = mismatches, out2[1] = inputs that took the slow path.
   KYD_SELFTEST_POW: compares powf's fast path (exp2(y log2 x) + rounding-interval check) with its definition
   (float)pow((double)x, (double)y) on `count` argument pairs derived from the indices first.. (Phong's exponents
   30 / 90 / 5000 and their 1/(n+1), random exponents; bases dense next to 1; negative bases with integer exponents)
   KYD_SELFTEST_TRAVERSAL: compares the two-phase traversal of the wavefront kernels (conservative classification of a ray
   against each rectangle, the reference's own tests only for candidates; kyd_device.cuh) with the reference's list walk
   (ky.cpp:3172-3206) on `count` adversarial rays through the UPLOADED scene (at most KYD_MAX_SURFACES surfaces) -- closest
   hit, occlusion and occlusion-before-a-surface queries; out2[0] = queries that disagree, out2[1] = rays that hit a surface. */
enum kyd_selftest_kind { KYD_SELFTEST_RSQRT = 0, KYD_SELFTEST_POW = 1, KYD_SELFTEST_TRAVERSAL = 2 };
int kyd_selftest(kyd_ctx* ctx, int which, uint64_t first, uint64_t count, uint64_t* out2);

/* Device known-answer harness: runs n inputs through ONE device function of the path and returns what it computed, so that
   fixtures generated from the reference (tests/golden/golden_kat.npz) pin the device functions one by one -- including branches
   no film reaches.  Layouts (floats per item) are those of the oracle's test entry points (oracle/kyo.c):
     KYD_KAT_SHAPE_INTERSECT        object = kyd_shape*     in 7 {o, d, tmax}            out 8 {hit, tmax, position, normal}   ky.cpp:1111-1393
     KYD_KAT_SHAPE_SAMPLE_DIRECTION object = kyd_shape*     in 8 {p, n, u0, u1}          out 7 {lp, ln, pdf}                   ky.cpp:1028-1051, 1419-1501
     KYD_KAT_SHAPE_PDF_DIRECTION    object = kyd_shape*     in 9 {p, n, wi}              out 1 pdf                             ky.cpp:1055-1090, 1503-1513
     KYD_KAT_MATERIAL_BSDF          object = kyd_material*  in 14 {p, n, wo, wi, u0, u1} out 13 {sample f, wi, pdf, type; eval; pdf; is_delta}  ky.cpp:2147-2682
     KYD_KAT_CAMERA_RAYS            uploaded scene          in 2 {px, py}                out 6 {o, d}                          ky.cpp:1884-1892
     KYD_KAT_LIGHT_SAMPLE           uploaded scene, light `index`  in 11 {p, n, u0, u1, wi}  out 11 {position, wi, pdf, Li, pdf_Li(wi)}  ky.cpp:2810-3062
     KYD_KAT_SAMPLER                seed `index`            in 4 {x, y, sample, -}       out 16 draws of the counter-seeded sampler
   `traits` selects the instantiation the specialised shade kernels use (0: general; 1: rectangle area light; 2: sphere area
   lights), for the functions that have one.  in / out are HOST pointers. */
enum kyd_kat_kind { KYD_KAT_SHAPE_INTERSECT = 0, KYD_KAT_SHAPE_SAMPLE_DIRECTION = 1, KYD_KAT_SHAPE_PDF_DIRECTION = 2, KYD_KAT_MATERIAL_BSDF = 3,
                    KYD_KAT_CAMERA_RAYS = 4, KYD_KAT_LIGHT_SAMPLE = 5, KYD_KAT_SAMPLER = 6 };
int kyd_kat(kyd_ctx* ctx, int which, const void* object, int index, int traits, int n, const float* in, float* out);

/* size in paths of one wavefront (0 = library default, 2^24; larger values are clamped to 2^24); tuning knob, results do
   not depend on it */
int kyd_set_wave_paths(kyd_ctx* ctx, int64_t paths);

/* ---- film output stage: the step right after the path (replaces the per-pixel loops of
   film_t::store_ppm_impl / store_bmp_impl / store_hdr_impl, ky.cpp:1661-1782).  The device turns the float
   film into the BODY bytes of the image file; kyd_film_header() gives the bytes the reference writes in front.
     KYD_FILM_GAMMA8 : 3 bytes/pixel, R G B, rows top-down; byte = gamma_encoding(x) =
                       (uint8_t)(pow((double)clamp01(x), 1/2.2) * 255 + .5)  (ky.cpp:1548).  The numbers the
                       reference prints into a P3 ppm (ky.cpp:1669-1681); header = "P3\n<w> <h>\n255\n".
     KYD_FILM_BMP24  : the same bytes as B G R, rows bottom-up, lines NOT padded to 4 bytes although the
                       header's file size is computed from padded lines (ky.cpp:1719-1733, reference quirk kept).
     KYD_FILM_RGBE   : 4 bytes/pixel Radiance RGBE, flat (no RLE), rows top-down (ky.cpp:1739-1782).
   NaN / out-of-range float->uint8 conversions follow the reference binary on x86-64 (cvttss2si, low byte). */
enum kyd_film_format { KYD_FILM_GAMMA8 = 0, KYD_FILM_BMP24 = 1, KYD_FILM_RGBE = 2 };

/* body size in bytes (3*w*h or 4*w*h); -1 for an unknown format or non-positive size */
int64_t kyd_film_body_bytes(int format, int width, int height);
/* writes the file header into `out` (capacity `cap`), returns its length, or -1.  Pure host function. */
int kyd_film_header(int format, int width, int height, uint8_t* out, int cap);
/* host film (width*height*3 floats, row-major top-down) -> host body bytes; copies in, encodes on the device, copies out */
int kyd_film_encode(kyd_ctx* ctx, const float* film_rgb, int width, int height, int format, uint8_t* out_body);
/* same with DEVICE pointers on the context's device: a film left resident by kyd_render_device is encoded without
   ever crossing PCIe as floats.  Stream semantics as kyd_render_device. */
int kyd_film_encode_device(kyd_ctx* ctx, const float* film_rgb_device, int width, int height, int format,
                           uint8_t* out_body_device, void* cuda_stream);

/* ---- FP64 validation mode for BASELINE config 1 ---------------------------------------
   The reference's double-precision smallpt, smallpt2pbrt/smallpt_kernel.cpp (Device::Render :403-438, Radiance :184-296:
   recursive, no light sampling, the nine-sphere scene compiled into that file, one 32-bit LCG per sample seeded with
   y * width + x * spp + s) on the device.  film_rgb: width * height * 3 doubles, clamp01'ed, in the reference's own
   layout (rows bottom-up).  Independent of kyd_upload_scene.  Agrees with the reference to ~1e-12 per pixel except where
   a last-bit difference between CUDA's and glibc's double sin/cos flips a branch (none in the test cases). */
int kyd_render_smallpt_f64(kyd_ctx* ctx, int width, int height, int samples_per_pixel, double* film_rgb);

#ifdef __cplusplus
}
#endif

#endif /* KYD_H */
