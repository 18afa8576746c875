// ky_entry.hpp -- the reference's entry points (reference ky.cpp:4675-4935) on top of ky.hpp.
//
// Each entry point exists twice: with the reference's own signature (hard-coded resolution / spp,
// writes <name>.bmp), and parameterised (ky_entry_params; zero fields mean "the reference's value"),
// returning the film so that callers and tests can look at the pixels.
#pragma once

#include "ky.hpp"

extern "C" {
typedef struct ky_entry_params
{
    int sub_width, sub_height; // film (or panel) size; 0 = the reference's hard-coded size
    int spp;                   // 0 = the reference's value (per scene where it differs)
    int depth;                 // max path depth; 0 = the reference's value
} ky_entry_params;
}

namespace ky {

namespace detail {
inline int pick(int requested, int reference_value) { return requested > 0 ? requested : reference_value; }

struct lit_scene_t { cornell_box_enum_t light; int spp; };
} // namespace detail

// reference ky.cpp:4675-4713: Cornell box with the environment light, path_tracing_iteration_t(5, both_mis)
inline std::unique_ptr<film_t> render_single_scene(const ky_entry_params& p, bool render = true)
{
    auto film = std::make_unique<film_t>(detail::pick(p.sub_width, 1024), detail::pick(p.sub_height, 1024));
    if (!render) return film;
    scene_t scene = scene_t::create_cornell_box_scene(
        cornell_box_enum_t::both_small_spheres | cornell_box_enum_t::light_environment, film->get_resolution());
    random_sampler_t sampler(detail::pick(p.spp, 16));
    auto integrator = create_integrator(integrator_enum_t::path_tracing_iteration, detail::pick(p.depth, 5), direct_sample_enum_t::both_mis);
    integrator->render(&scene, &sampler, film.get());
    return film;
}

// reference ky.cpp:4715-4738: position / normal / basecolor panels of the Veach scene
inline std::unique_ptr<film_t> render_debug(const ky_entry_params& p, bool render = true)
{
    auto film = std::make_unique<film_grid_t>(1, 3, detail::pick(p.sub_width, 512), detail::pick(p.sub_height, 308));
    if (!render) return film;
    random_sampler_t sampler(detail::pick(p.spp, 10));
    scene_t scene = scene_t::create_mis_scene(film->get_resolution());
    for (auto e : { integrator_enum_t::position, integrator_enum_t::normal, integrator_enum_t::basecolor })
    {
        debug_integrator_t integrator(e);
        integrator.render(&scene, &sampler, film.get());
        film->next_subfilm();
    }
    return film;
}

// reference ky.cpp:4740-4777: 4 light variants (rows) x 5 integrators (columns)
inline std::unique_ptr<film_t> render_multiple_integrator(const ky_entry_params& p, bool render = true)
{
    auto film = std::make_unique<film_grid_t>(4, 5, detail::pick(p.sub_width, 256), detail::pick(p.sub_height, 256));
    if (!render) return film;
    const detail::lit_scene_t rows[] = { { cornell_box_enum_t::light_point, 1 }, { cornell_box_enum_t::light_direction, 10 },
        { cornell_box_enum_t::light_area, 1 }, { cornell_box_enum_t::light_environment, 10 } };
    for (const auto& row : rows)
    {
        scene_t scene = scene_t::create_cornell_box_scene(cornell_box_enum_t::both_small_spheres | row.light, film->get_resolution());
        random_sampler_t sampler(detail::pick(p.spp, row.spp));
        for (auto e : { integrator_enum_t::direct_lighting, integrator_enum_t::simple_path_tracing_recursion,
                 integrator_enum_t::path_tracing_recursion, integrator_enum_t::path_tracing_recursion_defered,
                 integrator_enum_t::path_tracing_iteration })
        {
            auto integrator = create_integrator(e, detail::pick(p.depth, 5), direct_sample_enum_t::both_mis);
            integrator->render(&scene, &sampler, film.get());
            film->next_subfilm();
        }
    }
    return film;
}

// reference ky.cpp:4779-4817: 4 light variants (rows) x 5 direct-sampling strategies (columns)
inline std::unique_ptr<film_t> render_direct_sample_enum(const ky_entry_params& p, bool render = true)
{
    auto film = std::make_unique<film_grid_t>(4, 5, detail::pick(p.sub_width, 256), detail::pick(p.sub_height, 256));
    if (!render) return film;
    const detail::lit_scene_t rows[] = { { cornell_box_enum_t::light_point, 1 }, { cornell_box_enum_t::light_direction, 10 },
        { cornell_box_enum_t::light_area, 1 }, { cornell_box_enum_t::light_environment, 10 } };
    for (const auto& row : rows)
    {
        random_sampler_t sampler(detail::pick(p.spp, row.spp));
        scene_t scene = scene_t::create_cornell_box_scene(cornell_box_enum_t::both_small_spheres | row.light, film->get_resolution());
        for (auto ds : { direct_sample_enum_t::bsdf, direct_sample_enum_t::light, direct_sample_enum_t::bsdf_mis,
                 direct_sample_enum_t::light_mis, direct_sample_enum_t::both_mis })
        {
            path_tracing_iteration_t integrator(detail::pick(p.depth, 5), ds);
            integrator.render(&scene, &sampler, film.get());
            film->next_subfilm();
        }
    }
    return film;
}

// reference ky.cpp:4819-4876 ("multi_scene_mis"): 3 strategies (rows) x 4 light variants (columns)
inline std::unique_ptr<film_t> render_multiple_scene(const ky_entry_params& p, bool render = true)
{
    auto film = std::make_unique<film_grid_t>(3, 4, detail::pick(p.sub_width, 256), detail::pick(p.sub_height, 256));
    if (!render) return film;
    const detail::lit_scene_t columns[] = { { cornell_box_enum_t::light_point, 10 }, { cornell_box_enum_t::light_direction, 40 },
        { cornell_box_enum_t::light_area, 40 }, { cornell_box_enum_t::light_environment, 10 } };
    for (auto ds : { direct_sample_enum_t::bsdf, direct_sample_enum_t::light, direct_sample_enum_t::both_mis })
    {
        path_tracing_iteration_t integrator(detail::pick(p.depth, 5), ds);
        for (const auto& column : columns)
        {
            random_sampler_t sampler(detail::pick(p.spp, column.spp));
            scene_t scene = scene_t::create_cornell_box_scene(cornell_box_enum_t::both_small_spheres | column.light, film->get_resolution());
            integrator.render(&scene, &sampler, film.get());
            film->next_subfilm();
        }
    }
    return film;
}

// reference ky.cpp:4878-4905 ("veach_mis"): six direct-sampling strategies on the Veach scene
inline std::unique_ptr<film_t> render_mis_scene(const ky_entry_params& p, bool render = true)
{
    auto film = std::make_unique<film_grid_t>(2, 3, detail::pick(p.sub_width, 512), detail::pick(p.sub_height, 308));
    if (!render) return film;
    random_sampler_t sampler(detail::pick(p.spp, 10));
    scene_t scene = scene_t::create_mis_scene(film->get_resolution());
    for (auto ds : { direct_sample_enum_t::bsdf, direct_sample_enum_t::light, direct_sample_enum_t::idle,
             direct_sample_enum_t::bsdf_mis, direct_sample_enum_t::light_mis, direct_sample_enum_t::both_mis })
    {
        path_tracing_iteration_t integrator(detail::pick(p.depth, 5), ds);
        integrator.render(&scene, &sampler, film.get());
        film->next_subfilm();
    }
    return film;
}

// reference ky.cpp:4907-4935 (commented out there; its film.next_cell() does not exist and its
// lighting_enum_ member is never read): emit / direct / indirect / all panels of the Cornell box with
// path_tracing_recursion_defered_t(10, both_mis, lighting).  The filter is defined in DESIGN.md.
inline std::unique_ptr<film_t> render_lighting_enum(const ky_entry_params& p, bool render = true)
{
    auto film = std::make_unique<film_grid_t>(1, 4, detail::pick(p.sub_width, 256), detail::pick(p.sub_height, 256));
    if (!render) return film;
    random_sampler_t sampler(detail::pick(p.spp, 10));
    scene_t scene = scene_t::create_cornell_box_scene(cornell_box_enum_t::both_small_spheres | cornell_box_enum_t::light_area, film->get_resolution());
    for (auto le : { lighting_enum_t::emit, lighting_enum_t::direct, lighting_enum_t::indirect, lighting_enum_t::all })
    {
        path_tracing_recursion_defered_t integrator(detail::pick(p.depth, 10), direct_sample_enum_t::both_mis, le);
        integrator.render(&scene, &sampler, film.get());
        film->next_subfilm();
    }
    return film;
}

inline std::unique_ptr<film_t> run_entry(const std::string& name, const ky_entry_params& p, bool render = true)
{
    if (name == "render_single_scene") return render_single_scene(p, render);
    if (name == "render_debug") return render_debug(p, render);
    if (name == "render_multiple_integrator") return render_multiple_integrator(p, render);
    if (name == "render_direct_sample_enum") return render_direct_sample_enum(p, render);
    if (name == "render_multiple_scene") return render_multiple_scene(p, render);
    if (name == "render_mis_scene") return render_mis_scene(p, render);
    if (name == "render_lighting_enum") return render_lighting_enum(p, render);
    throw std::runtime_error("unknown entry point: " + name);
}

// ---- the reference's own signatures ------------------------------------------------------------------
inline void render_single_scene(int argc, char* argv[])
{
    ky_entry_params p{};
    p.spp = argc == 2 ? std::atoi(argv[1]) / 4 : 16; // ky.cpp:4690
    render_single_scene(p)->store_image("single");
}
inline void render_debug(int, char*[]) { render_debug(ky_entry_params{})->store_image("render_debug"); }
inline void render_multiple_integrator() { render_multiple_integrator(ky_entry_params{})->store_image("direct_sample"); }
inline void render_direct_sample_enum(int, char*[]) { render_direct_sample_enum(ky_entry_params{})->store_image("direct_sample"); }
inline void render_multiple_scene(int, char*[]) { render_multiple_scene(ky_entry_params{})->store_image("light_mis"); }
inline void render_mis_scene(int, char*[]) { render_mis_scene(ky_entry_params{})->store_image("veach_mis"); }
inline void render_lighting_enum() { render_lighting_enum(ky_entry_params{})->store_image("lighting"); }

} // namespace ky
