// oracle/ref/smallpt_cuda_addon.cu -- TEST / MEASUREMENT INFRASTRUCTURE.  C entry into the reference's own CUDA path
// (/root/reference/smallpt2pbrt/smallpt_kernel.cu = "#define USE_CUDA" + smallpt_kernel.cpp: one thread per pixel, recursive
// FP64 Radiance, managed film), compiled by build_ref.sh for sm_100a from the patched copy included below (main() cut off;
// nothing else touched).  Used to time "the reference's GPU path" beside kyd_render_smallpt_f64; never linked into the product.
#define USE_CUDA
#include "smallpt_kernel_cuda_ref.cu"

#include <chrono>
#include <cstring>

extern "C" {

// Runs the reference's Device::Render (stack-limit set-up, managed allocation, Kernel<<<>>>, synchronize) and returns its
// wall-clock seconds in *seconds; film: width * height * 3 doubles, rows bottom-up.  The Device object is leaked on purpose:
// its destructor frees an uninitialised member (Render's local `film` shadows it) and exits the process.
// base_stack_bytes > 0: the per-thread stack limit is set to this value BEFORE the reference's Render reads it and triples
// it (smallpt_kernel.cpp:352-355).  As written -- the default 1024 bytes tripled -- the recursive FP64 Radiance overflows its
// stack on sm_100a and the kernel faults ("an illegal memory access was encountered", observed on B200).
int smallpt_ref_cuda_render(int width, int height, int samples_per_pixel, double* film_rgb, double* seconds, int base_stack_bytes)
{
    if (base_stack_bytes > 0)
        cudaDeviceSetLimit(cudaLimitStackSize, (size_t)base_stack_bytes);
    Device* device = new Device;
    const auto t0 = std::chrono::steady_clock::now();
    Color* film = device->Render(width, height, samples_per_pixel);
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (film_rgb)
        std::memcpy(film_rgb, film, sizeof(double) * 3 * (size_t)width * height);
    cudaFree(film);
    return 0;
}

} // extern "C"
