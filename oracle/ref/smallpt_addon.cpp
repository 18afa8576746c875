// oracle/ref/smallpt_addon.cpp -- TEST INFRASTRUCTURE.  C entry into the reference's FP64 smallpt
// (/root/reference/smallpt2pbrt/smallpt_kernel.cpp, CPU_RENDER configuration), compiled by build_ref.sh with the
// patched copy of that file included below (its main() cut off, its progress print removed; nothing else touched).
#include "smallpt_kernel_ref.cpp"

#include <cstring>

extern "C" {

// film: width * height * 3 doubles in the reference's own layout (rows bottom-up, smallpt_kernel.cpp:431-432)
int smallpt_ref_render(int width, int height, int samples_per_pixel, double* film_rgb)
{
    Device device;
    Color* film = device.Render(width, height, samples_per_pixel);
    std::memcpy(film_rgb, film, sizeof(double) * 3 * (size_t)width * height);
    return 0;
}

} // extern "C"
