// ky_main.cpp -- command-line front end with the reference's entry points (reference ky.cpp:4937-4949 picks one by
// commenting lines in main(); here the first argument picks it).
//
//   ky <entry> [spp] [sub_width sub_height] [depth]
//   entry: render_single_scene | render_debug | render_multiple_integrator | render_direct_sample_enum |
//          render_multiple_scene | render_mis_scene | render_lighting_enum
// Writes <entry>.bmp like the reference's store_image (24-bit BGR, bottom-up, gamma 1/2.2).
#include <chrono>
#include <cstdio>

#include "ky.hpp"
#include "ky_entry.hpp"

int main(int argc, char* argv[])
{
    const std::string entry = argc > 1 ? argv[1] : "render_single_scene";
    ky_entry_params p{};
    if (argc > 2) p.spp = std::atoi(argv[2]);
    if (argc > 4) { p.sub_width = std::atoi(argv[3]); p.sub_height = std::atoi(argv[4]); }
    if (argc > 5) p.depth = std::atoi(argv[5]);
    try
    {
        auto t0 = std::chrono::steady_clock::now();
        std::unique_ptr<ky::film_t> film = ky::run_entry(entry, p);
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        film->store_image(entry);
        std::printf("%s: %dx%d in %.3f s -> %s.bmp\n", entry.c_str(), film->get_width(), film->get_height(), s, entry.c_str());
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "ky: %s\n", e.what());
        return 1;
    }
    return 0;
}
