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

} // extern "C"
