// oracle/ref/ref_addon.cpp -- TEST INFRASTRUCTURE (never part of the product).
//
// Wraps the REFERENCE's own classes (the patched copy of /root/reference/ky.cpp that
// build_ref.sh writes to oracle/_ref/ky_ref.cpp and that is #included below) behind a
// small C API, so tests/ and bench.py can
//   * render any BASELINE config with the reference's integrators into a raw float film,
//   * call single reference functions (shape intersect, BSDF sample/eval/pdf, light
//     sampling, sampler draws) on arbitrary inputs to make known-answer vectors.
// Nothing here re-implements rendering arithmetic: it only constructs reference objects
// through their public constructors and calls their public / virtual interfaces.
//
// Two libraries are built from this file:
//   libky_ref_verbatim.so : reference + compile-only patches; glibc float libm;
//                           random_sampler_t (per-row mt19937_64) is the default sampler
//   libky_ref_det.so      : -DKY_ORACLE_DETERMINISTIC (stateless plastic lobe choice) and
//                           linked with crlibm_shim.c; THE parity oracle when driven with
//                           sampler = 1 (lcg48_sampler_t below)

#include "ky_ref.cpp" // generated: patched copy of the reference, see build_ref.sh

#include <chrono>
#include <omp.h>

thread_local unsigned long long kyref_ray_count = 0;

// ------------------------------------------------------------------------------------------
// counter-seeded sampler: a plug-in for the reference's own sampler_t interface
// (ky.cpp:877-920).  Precedent inside the reference: per-sample seeding in
// smallpt2pbrt/smallpt_kernel.cpp:334,412 and the 48-bit LCG of smallpt2pbrt/erand48.h.
// ------------------------------------------------------------------------------------------
class lcg48_sampler_t : public sampler_t
{
public:
    lcg48_sampler_t(int samples_per_pixel, uint64_t seed, int sample_offset) :
        sampler_t(samples_per_pixel), seed_{ seed }, sample_offset_{ sample_offset }
    {
    }

    std::unique_ptr<sampler_t> clone() override
    {
        return std::make_unique<lcg48_sampler_t>(samples_per_pixel_, seed_, sample_offset_);
    }

    float_t get_float() override
    {
        state_ = (state_ * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull;
        return (float_t)(state_ >> 24) * 0x1p-24f;
    }

    vec2_t get_float2() override
    {
        float_t x = get_float();
        float_t y = get_float();
        return vec2_t(x, y);
    }

    // first call of every sample (ky.cpp:3714): reseed from (seed, x, y, sample index)
    camera_sample_t get_camera_sample(point2_t p_film) override
    {
        uint64_t x = (uint64_t)(int)p_film.x, y = (uint64_t)(int)p_film.y;
        uint64_t s = (uint64_t)(current_sample_index_ + sample_offset_);
        uint64_t key = s | (x << 24) | (y << 40);
        state_ = kyref_mix64(seed_ * 0x9E3779B97F4A7C15ull + key) >> 16;
        return { p_film + get_float2() };
    }

protected:
    uint64_t seed_{};
    int sample_offset_{};
    uint64_t state_{};
};

// Tent-filtered 2x2 sub-pixel camera samples (smallpt's filter; smallpt2pbrt/smallpt_rewrite.cpp:397-475
// TrapezoidalSampler) on top of the counter-seeded LCG, in ky's float_t: samples_per_pixel counts ALL samples of a pixel
// (a multiple of 4); sample s belongs to sub-pixel s / (spp / 4) -- sub-pixel after sub-pixel like the original -- and
// lands at p_film + ((sub_x + dx + 0.5) / 2, (sub_y + dy + 0.5) / 2) with dx, dy tent-distributed in [-1, 1).
class lcg48_trapezoidal_sampler_t : public lcg48_sampler_t
{
public:
    using lcg48_sampler_t::lcg48_sampler_t;

    std::unique_ptr<sampler_t> clone() override
    {
        return std::make_unique<lcg48_trapezoidal_sampler_t>(samples_per_pixel_, seed_, sample_offset_);
    }

    camera_sample_t get_camera_sample(point2_t p_film) override
    {
        uint64_t x = (uint64_t)(int)p_film.x, y = (uint64_t)(int)p_film.y;
        int sample = current_sample_index_ + sample_offset_;
        uint64_t key = (uint64_t)sample | (x << 24) | (y << 40);
        state_ = kyref_mix64(seed_ * 0x9E3779B97F4A7C15ull + key) >> 16;

        int sub_pixel = sample / (samples_per_pixel_ / 4);
        int sub_x = sub_pixel % 2, sub_y = sub_pixel / 2;
        float_t random1 = 2 * get_float();
        float_t random2 = 2 * get_float();
        float_t delta_x = random1 < 1 ? std::sqrt(random1) - 1 : 1 - std::sqrt(2 - random1);
        float_t delta_y = random2 < 1 ? std::sqrt(random2) - 1 : 1 - std::sqrt(2 - random2);
        vec2_t sample_point{ ((float_t)sub_x + delta_x + 0.5f) / 2, ((float_t)sub_y + delta_y + 0.5f) / 2 };
        return { p_film + sample_point };
    }
};

// camera with smallpt's ray-origin push (smallpt2pbrt/smallpt_rewrite.cpp:676), needed for
// BASELINE config 1.  ky's camera_t keeps its basis private, so the subclass rebuilds it with
// the same expressions as ky.cpp:1864-1880 from the same constructor arguments.
class pushed_camera_t : public camera_t
{
public:
    pushed_camera_t(vec3_t position, vec3_t front, vec3_t up, degree_t fov, vec2_t resolution, float_t push) :
        camera_t(position, front, up, fov, resolution),
        position_{ position }, front_{ front.normalize() }, up_{ up.normalize() }, resolution_{ resolution }, push_{ push }
    {
        float_t tan_fov = std::tan(radians(fov) / 2);
        right_ = up_.cross(front_).normalize() * tan_fov * (resolution_.x / resolution_.y);
        up_ = front_.cross(right_).normalize() * tan_fov;
    }

    ray_t generate_ray(const camera_sample_t& sample) const override
    {
        vec3_t direction =
            front_ +
            right_ * (sample.p_film.x / resolution_.x - 0.5) +
               up_ * (0.5 - sample.p_film.y / resolution_.y);
        return ray_t{ position_ + direction * push_, direction.normalize() };
    }

private:
    vec3_t position_, front_, right_, up_;
    vec2_t resolution_;
    float_t push_;
};

// ------------------------------------------------------------------------------------------
// extra scenes, built only from reference classes through scene_t's public constructor
// ------------------------------------------------------------------------------------------

// BASELINE config 1: the smallpt scene (geometry/materials: smallpt2pbrt/smallpt_rewrite.cpp:
// 1199-1244, camera :1391-1392) expressed with ky's FP32 classes.
static scene_t create_smallpt_scene(point2_t res)
{
    const_camera_sptr_t camera = std::make_shared<pushed_camera_t>(
        point3_t{ 50, 52, -295.6 }, vec3_t{ 0, -0.042612, 1 }, vec3_t{ 0, 1, 0 }, 53, res, 140);

    shape_sptr_t left   = std::make_shared<sphere_t>(vec3_t(1e5 + 1, 40.8, -81.6), 1e5);
    shape_sptr_t right  = std::make_shared<sphere_t>(vec3_t(-1e5 + 99, 40.8, -81.6), 1e5);
    shape_sptr_t back   = std::make_shared<sphere_t>(vec3_t(50, 40.8, -1e5), 1e5);
    shape_sptr_t front  = std::make_shared<sphere_t>(vec3_t(50, 40.8, 1e5 - 170), 1e5);
    shape_sptr_t bottom = std::make_shared<sphere_t>(vec3_t(50, 1e5, -81.6), 1e5);
    shape_sptr_t top    = std::make_shared<sphere_t>(vec3_t(50, -1e5 + 81.6, -81.6), 1e5);
    shape_sptr_t mirror = std::make_shared<sphere_t>(vec3_t(27, 16.5, -47), 16.5);
    shape_sptr_t glass  = std::make_shared<sphere_t>(vec3_t(73, 16.5, -78), 16.5);
    shape_sptr_t light  = std::make_shared<sphere_t>(vec3_t(50, 681.6 - .27, -81.6), 600);
    shape_list_t shape_list{ left, right, back, front, bottom, top, mirror, glass, light };

    material_sptr_t red   = std::make_shared<matte_material_t>(color_t(.75, .25, .25));
    material_sptr_t blue  = std::make_shared<matte_material_t>(color_t(.25, .25, .75));
    material_sptr_t gray  = std::make_shared<matte_material_t>(color_t(.75, .75, .75));
    material_sptr_t black = std::make_shared<matte_material_t>(color_t());
    material_sptr_t mirror_mat = std::make_shared<mirror_material_t>(color_t(.999, .999, .999));
    material_sptr_t glass_mat  = std::make_shared<glass_material_t>(1.5, color_t(.999, .999, .999), color_t(.999, .999, .999));
    material_list_t material_list{ red, blue, gray, black, mirror_mat, glass_mat };

    auto area = std::make_shared<area_light_t>(point3_t(), 1, color_t(12, 12, 12), light.get());
    light_list_t light_list{ area };

    surface_list_t surface_list
    {
        {   left.get(),   red.get(), nullptr },
        {  right.get(),  blue.get(), nullptr },
        {   back.get(),  gray.get(), nullptr },
        {  front.get(), black.get(), nullptr },
        { bottom.get(),  gray.get(), nullptr },
        {    top.get(),  gray.get(), nullptr },
        { mirror.get(), mirror_mat.get(), nullptr },
        {  glass.get(),  glass_mat.get(), nullptr },
        {  light.get(), black.get(), area.get() },
    };

    return scene_t{ camera, shape_list, material_list, light_list, surface_list };
}

// coverage scene for the shapes no shipped scene instantiates (disk_t, triangle_t) and for
// rectangle / triangle / disk area lights, point + direction + environment lights at once
static scene_t create_shapes_scene(point2_t res)
{
    const_camera_sptr_t camera = std::make_shared<camera_t>(
        point3_t{ 0.1f, 3.6f, 0.4f }, vec3_t{ -0.02f, -1.f, -0.08f }, vec3_t{ 0, 0, 1 }, 70, res);

    material_sptr_t black  = std::make_shared<matte_material_t>(color_t());
    material_sptr_t white  = std::make_shared<matte_material_t>(color_t(.7, .7, .7));
    material_sptr_t orange = std::make_shared<matte_material_t>(color_t(.8, .45, .15));
    material_sptr_t glossy = std::make_shared<plastic_material_t>(color_t(.2, .25, .3), color_t(.5, .5, .5), 30.);
    material_sptr_t mirror_mat = std::make_shared<mirror_material_t>(color_t(.9, .9, .9));
    material_sptr_t glass_mat  = std::make_shared<glass_material_t>(1.45);
    material_list_t material_list{ black, white, orange, glossy, mirror_mat, glass_mat };

    shape_sptr_t floor  = std::make_shared<rectangle_t>(point3_t(-2, -2, -1), point3_t(2, -2, -1), point3_t(2, 2, -1), point3_t(-2, 2, -1));
    shape_sptr_t wall   = std::make_shared<rectangle_t>(point3_t(-2, -2, -1), point3_t(-2, -2, 2), point3_t(2, -2, 2), point3_t(2, -2, -1));
    shape_sptr_t tri0   = std::make_shared<triangle_t>(point3_t(-1.6f, -1.2f, -1), point3_t(-0.4f, -1.5f, -1), point3_t(-1.1f, -1.4f, 0.7f));
    shape_sptr_t tri1   = std::make_shared<triangle_t>(point3_t(1.7f, -0.9f, -0.99f), point3_t(0.6f, -1.3f, -0.99f), point3_t(1.2f, -1.6f, 0.9f), true);
    shape_sptr_t disk0  = std::make_shared<disk_t>(point3_t(0.2f, -0.3f, -0.6f), vec3_t(0.1f, 0.4f, 1.f), 0.55f);
    shape_sptr_t ball   = std::make_shared<sphere_t>(vec3_t(-0.9f, 0.4f, -0.6f), 0.4f);
    shape_sptr_t gball  = std::make_shared<sphere_t>(vec3_t(0.9f, 0.6f, -0.65f), 0.35f);
    shape_sptr_t ltri   = std::make_shared<triangle_t>(point3_t(-0.5f, -0.6f, 1.6f), point3_t(0.5f, -0.6f, 1.6f), point3_t(0.f, 0.4f, 1.7f), true);
    shape_sptr_t ldisk  = std::make_shared<disk_t>(point3_t(1.5f, 0.2f, 1.2f), vec3_t(-1.f, 0.f, -0.6f), 0.3f);
    shape_sptr_t lrect  = std::make_shared<rectangle_t>(point3_t(-1.9f, 0.5f, 0.2f), point3_t(-1.9f, 1.1f, 0.2f), point3_t(-1.9f, 1.1f, 0.8f), point3_t(-1.9f, 0.5f, 0.8f));
    shape_sptr_t lball  = std::make_shared<sphere_t>(vec3_t(0.f, 1.2f, 0.2f), 0.12f);
    shape_list_t shape_list{ floor, wall, tri0, tri1, disk0, ball, gball, ltri, ldisk, lrect, lball };

    auto l_tri  = std::make_shared<area_light_t>(point3_t(), 1, color_t(18, 17, 15), ltri.get());
    auto l_disk = std::make_shared<area_light_t>(point3_t(), 1, color_t(9, 14, 20), ldisk.get());
    auto l_rect = std::make_shared<area_light_t>(point3_t(), 1, color_t(6, 9, 5), lrect.get());
    auto l_ball = std::make_shared<area_light_t>(point3_t(), 1, color_t(30, 22, 12), lball.get());
    auto l_pnt  = std::make_shared<point_light_t>(point3_t(-1.2f, 1.5f, 1.4f), 1, color_t(1.5f, 1.2f, 2.f));
    auto l_dir  = std::make_shared<direction_light_t>(point3_t(), 1, color_t(.6f, .5f, .3f), vec3_t(.4f, -1.f, -.7f));
    auto l_env  = std::make_shared<environment_light_t>(point3_t(), 1, color_t(.12f, .16f, .22f));
    light_list_t light_list{ l_tri, l_disk, l_rect, l_ball, l_pnt, l_dir, l_env };

    surface_list_t surface_list
    {
        { floor.get(), glossy.get(), nullptr },
        {  wall.get(),  white.get(), nullptr },
        {  tri0.get(), orange.get(), nullptr },
        {  tri1.get(), glossy.get(), nullptr },
        { disk0.get(),  white.get(), nullptr },
        {  ball.get(), mirror_mat.get(), nullptr },
        { gball.get(), glass_mat.get(), nullptr },
        {  ltri.get(),  black.get(), l_tri.get() },
        { ldisk.get(),  black.get(), l_disk.get() },
        { lrect.get(),  black.get(), l_rect.get() },
        { lball.get(),  black.get(), l_ball.get() },
    };

    return scene_t{ camera, shape_list, material_list, light_list, surface_list, l_env.get() };
}

static scene_t make_scene(int scene, int scene_flags, point2_t res)
{
    switch (scene)
    {
    case 0: return scene_t::create_cornell_box_scene((cornell_box_enum_t)scene_flags, res);
    case 1: return scene_t::create_mis_scene(res);
    case 2: return create_smallpt_scene(res);
    case 3: return create_shapes_scene(res);
    }
    throw std::runtime_error("kyref: unknown scene");
}

static std::unique_ptr<integrator_t> make_integrator(int integrator, int depth, int direct_sample)
{
    auto e = (integrator_enum_t)integrator;
    switch (e)
    {
    case integrator_enum_t::position:
    case integrator_enum_t::normal:
    case integrator_enum_t::basecolor:
        return std::make_unique<debug_integrator_t>(e);
    default:
        break;
    }
    auto r = create_integrator(e, depth, (direct_sample_enum_t)direct_sample);
    if (!r) throw std::runtime_error("kyref: unknown integrator");
    return r;
}

extern "C" {

struct kyref_render_desc
{
    int scene;          // 0 cornell, 1 veach mis, 2 smallpt (ky classes), 3 shapes coverage scene
    int scene_flags;    // cornell_box_enum_t bits (scene 0 only)
    int width, height;
    int spp;
    int sample_offset;  // lcg48 sampler: global index of this render's first sample
    int integrator;     // integrator_enum_t
    int max_depth;
    int direct_sample;  // direct_sample_enum_t
    int sampler;        // 0 random_sampler_t (reference), 1 lcg48_sampler_t, 2 debug_sampler_t
    int threads;        // OpenMP threads, 0 = runtime default
    int clamp;          // 1: film = what the reference stores (clamp01 per pixel); 0 is not available
    unsigned long long seed;
};

int kyref_is_deterministic_build()
{
#ifdef KY_ORACLE_DETERMINISTIC
    return 1;
#else
    return 0;
#endif
}

// film_rgb: width*height*3 floats, row-major, y down, exactly film_t's pixels_
int kyref_render(const kyref_render_desc* d, float* film_rgb, double* seconds, unsigned long long* rays)
{
    try
    {
        film_t film(d->width, d->height);
        scene_t scene = make_scene(d->scene, d->scene_flags, film.get_resolution());

        std::unique_ptr<sampler_t> sampler;
        if (d->sampler == 0) sampler = std::make_unique<random_sampler_t>(d->spp);
        else if (d->sampler == 1) sampler = std::make_unique<lcg48_sampler_t>(d->spp, d->seed, d->sample_offset);
        else if (d->sampler == 3) sampler = std::make_unique<lcg48_trapezoidal_sampler_t>(d->spp, d->seed, d->sample_offset);
        else sampler = std::make_unique<debug_sampler_t>(d->spp);

        auto integrator = make_integrator(d->integrator, d->max_depth, d->direct_sample);

        if (d->threads > 0) omp_set_num_threads(d->threads);

        unsigned long long total_rays = 0;
        #pragma omp parallel
        { kyref_ray_count = 0; }

        auto t0 = std::chrono::steady_clock::now();
        integrator->render(&scene, sampler.get(), &film);
        auto t1 = std::chrono::steady_clock::now();

        #pragma omp parallel reduction(+ : total_rays)
        { total_rays += kyref_ray_count; }

        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        if (rays) *rays = total_rays;

        for (int y = 0; y < d->height; ++y)
            for (int x = 0; x < d->width; ++x)
            {
                color_t c = film(x, y);
                float* o = film_rgb + 3 * ((size_t)y * d->width + x);
                o[0] = c.r; o[1] = c.g; o[2] = c.b;
            }
        return 0;
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "kyref_render: %s\n", e.what());
        return 1;
    }
}

// ---- known-answer helpers: single reference functions on caller-supplied inputs -------------

// sampler draws: camera sample (2 floats, pixel-relative) then n-2 get_float()
void kyref_sampler_floats(int kind, unsigned long long seed, int x, int y, int sample_index, int n, float* out)
{
    std::unique_ptr<sampler_t> s;
    if (kind == 0) s = std::make_unique<random_sampler_t>(1);
    else s = std::make_unique<lcg48_sampler_t>(1, seed, sample_index);
    s->start_pixel();
    camera_sample_t cs = s->get_camera_sample({ (float_t)x, (float_t)y });
    out[0] = cs.p_film.x - (float_t)x;
    out[1] = cs.p_film.y - (float_t)y;
    for (int i = 2; i < n; ++i) out[i] = s->get_float();
}

float kyref_plastic_random_of(const float* position, const float* wo)
{
    return kyref_plastic_random(position, wo);
}

static std::unique_ptr<shape_t> make_shape(int kind, const float* p)
{
    switch (kind)
    {
    case 0: return std::make_unique<sphere_t>(vec3_t(p[0], p[1], p[2]), p[3]);
    case 1: return std::make_unique<rectangle_t>(point3_t(p[0], p[1], p[2]), point3_t(p[3], p[4], p[5]),
                point3_t(p[6], p[7], p[8]), point3_t(p[9], p[10], p[11]), p[12] != 0);
    case 2: return std::make_unique<triangle_t>(point3_t(p[0], p[1], p[2]), point3_t(p[3], p[4], p[5]),
                point3_t(p[6], p[7], p[8]), p[9] != 0);
    case 3: return std::make_unique<disk_t>(point3_t(p[0], p[1], p[2]), vec3_t(p[3], p[4], p[5]), p[6]);
    }
    return nullptr;
}

// rays: n x {o.xyz, d.xyz, tmax}; out: n x {hit, t, p.xyz, n.xyz}
void kyref_shape_intersect(int kind, const float* params, int n, const float* rays, float* out)
{
    auto shape = make_shape(kind, params);
    for (int i = 0; i < n; ++i)
    {
        const float* r = rays + 7 * i;
        ray_t ray(point3_t(r[0], r[1], r[2]), vec3_t(r[3], r[4], r[5]), r[6]);
        isect_t isect;
        bool hit = shape->intersect(ray, &isect);
        float* o = out + 8 * i;
        o[0] = hit ? 1.f : 0.f; o[1] = ray.distance();
        o[2] = isect.position.x; o[3] = isect.position.y; o[4] = isect.position.z;
        o[5] = isect.normal.x; o[6] = isect.normal.y; o[7] = isect.normal.z;
    }
}

float kyref_shape_area(int kind, const float* params)
{
    return make_shape(kind, params)->area();
}

// in: n x {p.xyz, n.xyz, u0, u1}; out: n x {lp.xyz, ln.xyz, pdf}
void kyref_shape_sample_direction(int kind, const float* params, int n, const float* in, float* out)
{
    auto shape = make_shape(kind, params);
    for (int i = 0; i < n; ++i)
    {
        const float* a = in + 8 * i;
        isect_t isect(point3_t(a[0], a[1], a[2]), vec3_t(a[3], a[4], a[5]), vec3_t(0, 0, 1));
        float_t pdf{};
        isect_t li = shape->sample_direction(isect, float2_t(a[6], a[7]), pdf);
        float* o = out + 7 * i;
        o[0] = li.position.x; o[1] = li.position.y; o[2] = li.position.z;
        o[3] = li.normal.x; o[4] = li.normal.y; o[5] = li.normal.z; o[6] = pdf;
    }
}

// in: n x {p.xyz, n.xyz, wi.xyz}; out: n pdfs
void kyref_shape_pdf_direction(int kind, const float* params, int n, const float* in, float* out)
{
    auto shape = make_shape(kind, params);
    for (int i = 0; i < n; ++i)
    {
        const float* a = in + 9 * i;
        isect_t isect(point3_t(a[0], a[1], a[2]), vec3_t(a[3], a[4], a[5]), vec3_t(0, 0, 1));
        out[i] = shape->pdf_direction(isect, vec3_t(a[6], a[7], a[8]));
    }
}

static std::unique_ptr<material_t> make_material(int kind, const float* p)
{
    switch (kind)
    {
    case 0: return std::make_unique<matte_material_t>(color_t(p[0], p[1], p[2]));
    case 1: return std::make_unique<mirror_material_t>(color_t(p[0], p[1], p[2]));
    case 2: return std::make_unique<glass_material_t>(p[6], color_t(p[0], p[1], p[2]), color_t(p[3], p[4], p[5]));
    case 3: return std::make_unique<plastic_material_t>(color_t(p[0], p[1], p[2]), color_t(p[3], p[4], p[5]), p[6]);
    }
    return nullptr;
}

// in: n x {p.xyz, n.xyz, wo.xyz, wi.xyz, u0, u1}
// out: n x {sample.f rgb, sample.wi xyz, sample.pdf, sample.type, eval rgb, pdf, is_delta}
void kyref_material_bsdf(int kind, const float* params, int n, const float* in, float* out)
{
    auto material = make_material(kind, params);
    for (int i = 0; i < n; ++i)
    {
        const float* a = in + 14 * i;
        isect_t isect(point3_t(a[0], a[1], a[2]), vec3_t(a[3], a[4], a[5]), vec3_t(a[6], a[7], a[8]));
        bsdf_uptr_t bsdf = material->scattering(isect);
        bsdf_sample_t bs = bsdf->sample(isect.wo, float2_t(a[12], a[13]));
        vec3_t wi(a[9], a[10], a[11]);
        color_t f = bsdf->eval(isect.wo, wi);
        float_t pdf = bsdf->pdf(isect.wo, wi);
        float* o = out + 13 * i;
        o[0] = bs.f.r; o[1] = bs.f.g; o[2] = bs.f.b;
        o[3] = bs.wi.x; o[4] = bs.wi.y; o[5] = bs.wi.z;
        o[6] = bs.pdf; o[7] = (float)(int)bs.bsdf_type;
        o[8] = f.r; o[9] = f.g; o[10] = f.b; o[11] = pdf; o[12] = bsdf->is_delta() ? 1.f : 0.f;
    }
}

// camera rays: in n x {px, py} (film coordinates incl. jitter); out n x {o.xyz, d.xyz}
void kyref_camera_rays(int scene, int scene_flags, int width, int height, int n, const float* in, float* out)
{
    scene_t sc = make_scene(scene, scene_flags, { (float_t)width, (float_t)height });
    for (int i = 0; i < n; ++i)
    {
        ray_t ray = sc.get_camera()->generate_ray({ point2_t(in[2 * i], in[2 * i + 1]) });
        float* o = out + 6 * i;
        o[0] = ray.origin().x; o[1] = ray.origin().y; o[2] = ray.origin().z;
        o[3] = ray.direction().x; o[4] = ray.direction().y; o[5] = ray.direction().z;
    }
}

// light sampling through the scene's own light list:
// in: n x {p.xyz, n.xyz, u0, u1, wi.xyz}; out: n x {ls.position xyz, ls.wi xyz, ls.pdf, ls.Li rgb, pdf_Li(wi)}
int kyref_light_sample(int scene, int scene_flags, int light_index, int n, const float* in, float* out)
{
    scene_t sc = make_scene(scene, scene_flags, { 64.f, 64.f });
    if (light_index < 0 || light_index >= sc.light_count()) return 1;
    const light_t& light = *sc.light_list()[light_index];
    for (int i = 0; i < n; ++i)
    {
        const float* a = in + 11 * i;
        isect_t isect(point3_t(a[0], a[1], a[2]), vec3_t(a[3], a[4], a[5]), vec3_t(0, 0, 1));
        light_sample_t ls = light.sample_Li(isect, float2_t(a[6], a[7]));
        float* o = out + 11 * i;
        o[0] = ls.position.x; o[1] = ls.position.y; o[2] = ls.position.z;
        o[3] = ls.wi.x; o[4] = ls.wi.y; o[5] = ls.wi.z;
        o[6] = ls.pdf; o[7] = ls.Li.r; o[8] = ls.Li.g; o[9] = ls.Li.b;
        o[10] = light.pdf_Li(isect, vec3_t(a[8], a[9], a[10]));
    }
    return 0;
}

// film output stage (ky.cpp:1548, 1661-1782): the reference's own static writers on caller-provided floats;
// format 0 = ppm (text), 1 = bmp, 2 = hdr.  gamma: n floats -> n bytes through gamma_encoding().
int kyref_store_film(int format, const char* path, int width, int height, const float* floats)
{
    switch (format)
    {
    case 0: return film_t::store_ppm_impl(path, width, height, 3, floats) ? 0 : 1;
    case 1: return film_t::store_bmp_impl(path, width, height, 3, floats) ? 0 : 1;
    case 2: return film_t::store_hdr_impl(path, width, height, 3, floats) ? 0 : 1;
    }
    return 2;
}

void kyref_gamma_encoding(long long n, const float* in, unsigned char* out)
{
    for (long long i = 0; i < n; ++i)
        out[i] = gamma_encoding(in[i]);
}

} // extern "C"
