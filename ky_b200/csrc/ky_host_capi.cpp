// ky_host_capi.cpp -- thin C entry points over the C++ host surface (include/ky.hpp) so that
// non-C++ callers (the Python tests and bench.py) can obtain flattened scenes and run the
// reference-named entry points.  Built into libky_host.so, which links libkyd.so.
#include "ky.hpp"
#include "ky_entry.hpp"

using namespace ky;

namespace {
thread_local std::string g_error;
}

extern "C" {

const char* ky_host_last_error() { return g_error.c_str(); }

// scene: 0 cornell box (flags = cornell_box_enum_t bits), 1 Veach MIS, 2 smallpt (config 1), 3 shapes coverage
void* ky_host_scene_create(int scene, int flags, int width, int height)
{
    try
    {
        point2_t res{ (float)width, (float)height };
        scene_t s;
        switch (scene)
        {
        case 0: s = scene_t::create_cornell_box_scene((cornell_box_enum_t)flags, res); break;
        case 1: s = scene_t::create_mis_scene(res); break;
        case 2: s = scene_t::create_smallpt_scene(res); break;
        case 3: s = scene_t::create_shapes_scene(res); break;
        default: throw std::runtime_error("unknown scene id");
        }
        return s.flatten().release();
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return nullptr;
    }
}

const kyd_scene_desc* ky_host_scene_desc(void* handle) { return &static_cast<flat_scene_t*>(handle)->desc; }

void ky_host_scene_destroy(void* handle) { delete static_cast<flat_scene_t*>(handle); }

// describes one shape / material built through the host classes' constructors (tests: constructor
// arithmetic -- stored normals, areas, plastic lobe probabilities -- against the reference's)
// shape params: sphere {c.xyz, r}; rectangle {p0,p1,p2,p3, flip}; triangle {p0,p1,p2, flip}; disk {p, n, r}
int ky_host_shape_describe(int kind, const float* p, kyd_shape* out)
{
    try
    {
        std::unique_ptr<shape_t> s;
        switch (kind)
        {
        case KYD_SHAPE_SPHERE: s = std::make_unique<sphere_t>(vec3_t(p[0], p[1], p[2]), p[3]); break;
        case KYD_SHAPE_RECTANGLE: s = std::make_unique<rectangle_t>(point3_t(p[0], p[1], p[2]), point3_t(p[3], p[4], p[5]), point3_t(p[6], p[7], p[8]), point3_t(p[9], p[10], p[11]), p[12] != 0); break;
        case KYD_SHAPE_TRIANGLE: s = std::make_unique<triangle_t>(point3_t(p[0], p[1], p[2]), point3_t(p[3], p[4], p[5]), point3_t(p[6], p[7], p[8]), p[9] != 0); break;
        case KYD_SHAPE_DISK: s = std::make_unique<disk_t>(point3_t(p[0], p[1], p[2]), vec3_t(p[3], p[4], p[5]), p[6]); break;
        default: throw std::runtime_error("unknown shape kind");
        }
        *out = s->describe();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// material params: {diffuse/specular rgb, second colour rgb, eta or exponent} as in oracle/ref/ref_addon.cpp
int ky_host_material_describe(int kind, const float* p, kyd_material* out)
{
    try
    {
        std::unique_ptr<material_t> m;
        switch (kind)
        {
        case KYD_MAT_MATTE: m = std::make_unique<matte_material_t>(color_t(p[0], p[1], p[2])); break;
        case KYD_MAT_MIRROR: m = std::make_unique<mirror_material_t>(color_t(p[0], p[1], p[2])); break;
        case KYD_MAT_GLASS: m = std::make_unique<glass_material_t>(p[6], color_t(p[0], p[1], p[2]), color_t(p[3], p[4], p[5])); break;
        case KYD_MAT_PLASTIC: m = std::make_unique<plastic_material_t>(color_t(p[0], p[1], p[2]), color_t(p[3], p[4], p[5]), p[6]); break;
        default: throw std::runtime_error("unknown material kind");
        }
        *out = m->describe();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// the host-side LCG48 sampler: camera sample (2 floats) then n-2 get_float()
void ky_host_sampler_floats(unsigned long long seed, int x, int y, int sample_index, int n, float* out)
{
    lcg48_sampler_t s(sample_index + 1, seed);
    s.start_pixel();
    for (int i = 0; i < sample_index; ++i) s.next_sample();
    camera_sample_t cs = s.get_camera_sample({ (float)x, (float)y });
    out[0] = cs.p_film.x - (float)x;
    out[1] = cs.p_film.y - (float)y;
    for (int i = 2; i < n; ++i) out[i] = s.get_float();
}

// runs one of the reference's entry points (ky.cpp:4675-4935) and copies the resulting film out.
// name: "render_single_scene", "render_debug", "render_multiple_integrator", "render_direct_sample_enum",
//       "render_multiple_scene", "render_mis_scene", "render_lighting_enum"
// out_rgb may be NULL to query the film size only.
int ky_host_render_entry(const char* name, const ky_entry_params* params, float* out_rgb, int* out_width, int* out_height)
{
    try
    {
        std::unique_ptr<film_t> film = run_entry(name, params ? *params : ky_entry_params{}, out_rgb != nullptr);
        if (out_width) *out_width = film->get_width();
        if (out_height) *out_height = film->get_height();
        if (out_rgb)
            std::memcpy(out_rgb, film->data(), sizeof(float) * 3 * (size_t)film->get_pixel_num());
        return 0;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// film_t writers on caller-provided pixels: kind 0 = store_image (device film stage: <path>.bmp, or .hdr when the
// library is built with KY_OUTPUT_HDR), 1/2/3 = device-encoded ppm / bmp / hdr written to `path` as given,
// 11/12/13 = the host-side writers (reference arithmetic on the CPU) for ppm / bmp / hdr.
int ky_host_film_store(int kind, const char* path, int width, int height, const float* rgb)
{
    try
    {
        film_t film(width, height);
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x)
            {
                const float* p = rgb + 3 * ((size_t)y * width + x);
                film.set_color(x, y, color_t(p[0], p[1], p[2]));
            }
        bool ok = false;
        switch (kind)
        {
        case 0: ok = film.store_image(path); break;
        case 1: ok = film.store_device(path, KYD_FILM_GAMMA8); break;
        case 2: ok = film.store_device(path, KYD_FILM_BMP24); break;
        case 3: ok = film.store_device(path, KYD_FILM_RGBE); break;
        case 11: ok = film.store_ppm(path); break;
        case 12: ok = film.store_bmp(path); break;
        case 13: ok = film.store_hdr(path); break;
        default: g_error = "unknown film writer"; return 2;
        }
        if (!ok) { g_error = std::string("cannot write ") + path; return 1; }
        return 0;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

} // extern "C"
