// oracle/ref/ref_entries_main.cpp -- TEST INFRASTRUCTURE: proof of the drop-in.
//
// The reference's own entry points and main() (/root/reference/ky.cpp from `void render_single_scene(` to the end of the file:
// render_single_scene, render_debug, render_multiple_integrator, render_direct_sample_enum, render_multiple_scene,
// render_mis_scene, main), extracted VERBATIM by build_ref.sh into oracle/_ref/ref_entries.inc, compiled against
// include/ky.hpp instead of the reference's class definitions and linked with libkyd.so.  Nothing of the reference's code
// is edited: that its scene factories, samplers, integrators, film grids and image writers are used through the same names
// and call shapes is what "keeps ky's class surface" means.
//
// The reference's main() renders render_single_scene only (the others are commented out there, ky.cpp:4941-4946), so it is
// renamed by the preprocessor and a dispatcher picks the entry point by name: ky_ref_entries <entry> [4 x spp].
#include "ky.hpp"

using namespace ky;

#define main ky_reference_main
#include "ref_entries.inc"
#undef main

#include <cstring>

int main(int argc, char* argv[])
{
    const char* entry = argc > 1 ? argv[1] : "main";
    try
    {
        if (!std::strcmp(entry, "main")) return ky_reference_main(argc - 1, argv + 1);
        if (!std::strcmp(entry, "render_single_scene")) render_single_scene(argc - 1, argv + 1);
        else if (!std::strcmp(entry, "render_debug")) render_debug(argc - 1, argv + 1);
        else if (!std::strcmp(entry, "render_multiple_integrator")) render_multiple_integrator();
        else if (!std::strcmp(entry, "render_direct_sample_enum")) render_direct_sample_enum(argc - 1, argv + 1);
        else if (!std::strcmp(entry, "render_multiple_scene")) render_multiple_scene(argc - 1, argv + 1);
        else if (!std::strcmp(entry, "render_mis_scene")) render_mis_scene(argc - 1, argv + 1);
        else { std::fprintf(stderr, "unknown entry point %s\n", entry); return 2; }
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "ky_ref_entries: %s\n", e.what());
        return 1;
    }
    return 0;
}
