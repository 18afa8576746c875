#!/usr/bin/env bash
# oracle/ref/build_ref.sh -- TEST INFRASTRUCTURE.
#
# Builds the reference (/root/reference/ky.cpp) where it lies, into oracle/_ref/ only
# (git-ignored; travels to the GPU box with the snapshot like any built .so):
#
#   oracle/_ref/libky_ref_verbatim.so   reference + compile-only patches P1-P4
#   oracle/_ref/libky_ref_det.so        + P5 (stateless plastic lobe draw) + crlibm_shim.c
#   oracle/_ref/libky_ref_glibc.so      + P5, glibc's own float libm (no crlibm_shim): what the libm contract costs,
#                                       measured by tests/test_libm_contract.py
#   oracle/_ref/ky_ref_entries          the reference's own render_* entry points and main(), verbatim, compiled against
#                                       include/ky.hpp and linked with libkyd.so (the drop-in proof; needs libkyd.so built)
#   oracle/_ref/libsmallpt_kernel_ref.so       smallpt2pbrt/smallpt_kernel.cpp, CPU_RENDER configuration
#   oracle/_ref/libsmallpt_kernel_cuda_ref.so  the same file as smallpt_kernel.cu compiles it (USE_CUDA): the reference's
#                                       own CUDA kernel, built for sm_100a (timed beside kyd_render_smallpt_f64)
#
# No reference source is copied into the repository: the patched copy lives in oracle/_ref/.
# The reference's own build system (CMake, MSVC presets) is not used.
#
# Patches (each must hit exactly one line, else the build fails):
#   P1  ky.cpp:81    throw std::exception(msg.c_str())  -> std::runtime_error  (MSVC-only ctor)
#   P2  ky.cpp:4937  int main(...)                      -> int ky_main(...)    (library build)
#   P3  ky.cpp:3703  per-row progress printf            -> removed             (I/O only)
#   P4  ky.cpp:3174  scene_t::intersect                 -> + ray counter hook  (no arithmetic)
#   P5  ky.cpp:2663  plastic_material_t rng_ draw       -> KY_PLASTIC_RANDOM   (verbatim build:
#                    expands to the original rng_.uniform_float(); det build: hash of the hit)
# plus oracle/ref/shim_print as <print> (libstdc++ 13 lacks it; ky.cpp:20).
set -euo pipefail

here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${KY_REFERENCE_DIR:-/root/reference}"
out="$here/../_ref"

if [ ! -f "$ref/ky.cpp" ]; then
    echo "build_ref: $ref/ky.cpp not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
    exit 0
fi

mkdir -p "$out/shim"
cp "$here/shim_print" "$out/shim/print"

src="$out/ky_ref.cpp"
cp "$ref/ky.cpp" "$src"

patch1() { # description, sed-script, grep-pattern-that-must-match-once-before
    local n
    n=$(grep -c -- "$3" "$src" || true)
    if [ "$n" != "1" ]; then echo "build_ref: patch '$1' anchor matched $n lines" >&2; exit 1; fi
    sed -i "$2" "$src"
}
patch1 P1 's/throw std::exception(msg.c_str());/throw std::runtime_error(msg);/' 'throw std::exception(msg.c_str());'
patch1 P2 's/^int main(int argc, char\* argv\[\])/int ky_main(int argc, char* argv[])/' '^int main(int argc, char\* argv\[\])'
patch1 P3 '/std::printf("%s", std::format("rendering\.\.\. {} spp/d' 'std::printf("%s", std::format("rendering\.\.\. {} spp'
patch1 P4 's/bool is_hit = false;/bool is_hit = false; KYREF_COUNT_RAY();/' 'bool is_hit = false;'
patch1 P5 's/float_t random = rng_\.uniform_float();/float_t random = KY_PLASTIC_RANDOM(isect, rng_);/' 'float_t random = rng_\.uniform_float();'

# pinned flags: no -march, no fast-math, no FMA contraction
CXXFLAGS="-std=c++23 -O2 -ffp-contract=off -fopenmp -fPIC -w -I$out/shim -I$out -include $here/ref_prelude.h"

g++ $CXXFLAGS -shared "$here/ref_addon.cpp" -o "$out/libky_ref_verbatim.so"

gcc -O2 -fPIC -fno-builtin -fvisibility=hidden -c "$here/crlibm_shim.c" -o "$out/crlibm_shim.o"
g++ $CXXFLAGS -DKY_ORACLE_DETERMINISTIC -shared "$here/ref_addon.cpp" "$out/crlibm_shim.o" \
    -Wl,-Bsymbolic -o "$out/libky_ref_det.so"

# the deterministic configuration on glibc's own float functions (sinf / cosf / sincosf / powf / acosf): same sampler, same
# plastic lobe draw, stock libm -- the build a maintainer gets by dropping lcg48_sampler_t and P5 into the reference
g++ $CXXFLAGS -DKY_ORACLE_DETERMINISTIC -shared "$here/ref_addon.cpp" -Wl,-Bsymbolic -o "$out/libky_ref_glibc.so"

# ---- the FP64 smallpt of the teaching ladder: smallpt2pbrt/smallpt_kernel.cpp, CPU_RENDER ----
#   S1  smallpt_kernel.cpp:440-  main() (MSVC-only fopen_s / errno_t, writes a ppm)  -> cut off
#   S2  smallpt_kernel.cpp:417   per-row progress fprintf                            -> removed (I/O only)
src="$out/smallpt_kernel_ref.cpp"
cp "$ref/smallpt2pbrt/smallpt_kernel.cpp" "$src"
patch1 S1 '/^int main(int argc, char\* argv\[\])/,$d' '^int main(int argc, char\* argv\[\])'
patch1 S2 '/fprintf(stderr, "\\rRendering (%d spp) %5.2f%%"/d' 'fprintf(stderr, "\\rRendering (%d spp) %5.2f%%"'
g++ -std=c++20 -O2 -ffp-contract=off -fopenmp -fPIC -w -I"$out" -shared "$here/smallpt_addon.cpp" -o "$out/libsmallpt_kernel_ref.so"

# the reference's CUDA kernel (smallpt_kernel.cu = "#define USE_CUDA" + the same file), for sm_100a; the reference ships no
# arch flags (smallpt2pbrt/CMakeLists.txt:22-25).  Same S1 cut; the GPU configuration has no progress print.
if command -v nvcc >/dev/null 2>&1; then
    src="$out/smallpt_kernel_cuda_ref.cu"
    cp "$ref/smallpt2pbrt/smallpt_kernel.cpp" "$src"
    patch1 S1 '/^int main(int argc, char\* argv\[\])/,$d' '^int main(int argc, char\* argv\[\])'
    nvcc -std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -w -I"$out" -shared "$here/smallpt_cuda_addon.cu" \
        -o "$out/libsmallpt_kernel_cuda_ref.so"
fi

# ---- the drop-in proof: the reference's own entry points + main(), verbatim, against include/ky.hpp + libkyd.so ----
kyd_lib="$here/../../ky_b200/lib"
if [ -f "$kyd_lib/libkyd.so" ]; then
    n=$(grep -c '^void render_single_scene(int argc, char\* argv\[\])' "$ref/ky.cpp" || true)
    if [ "$n" != "1" ]; then echo "build_ref: entry-point anchor matched $n lines" >&2; exit 1; fi
    sed -n '/^void render_single_scene(int argc, char\* argv\[\])/,$p' "$ref/ky.cpp" > "$out/ref_entries.inc"
    g++ -std=c++20 -O1 -ffp-contract=off -w -I"$here/../../include" -I"$out" "$here/ref_entries_main.cpp" -o "$out/ky_ref_entries" \
        -L"$kyd_lib" -lkyd -Wl,-rpath,'$ORIGIN/../../ky_b200/lib'
fi

echo "build_ref: built $(ls "$out"/*.so | tr '\n' ' ')"
