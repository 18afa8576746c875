// oracle/ref/ref_prelude.h -- TEST INFRASTRUCTURE (never part of the product).
//
// Force-included (-include) in front of the patched copy of /root/reference/ky.cpp when
// oracle/ref/build_ref.sh builds oracle/_ref/.  It only supplies the names the sed
// patches refer to; see build_ref.sh for the patches themselves.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>

// P4: every scene_t::intersect() call (ky.cpp:3172) is one "ray" (primary, bsdf, shadow)
extern thread_local unsigned long long kyref_ray_count;
#define KYREF_COUNT_RAY() (++kyref_ray_count)

// ---- the deterministic-sampling contract (DESIGN.md "Sampling contract") -------------------
inline uint64_t kyref_mix64(uint64_t z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

inline uint32_t kyref_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// stateless replacement of plastic_material_t's private, shared, racy rng_t draw
// (ky.cpp:2663, 2681): a hash of the hit record, so scattering() is a pure function
inline float kyref_plastic_random(const float* p, const float* wo)
{
    uint64_t h = kyref_mix64((uint64_t)kyref_bits(wo[1]) | ((uint64_t)kyref_bits(wo[2]) << 32));
    h = kyref_mix64(((uint64_t)kyref_bits(p[2]) | ((uint64_t)kyref_bits(wo[0]) << 32)) ^ h);
    h = kyref_mix64(((uint64_t)kyref_bits(p[0]) | ((uint64_t)kyref_bits(p[1]) << 32)) ^ h);
    return (float)(h >> 40) * 0x1p-24f;
}

#ifdef KY_ORACLE_DETERMINISTIC
    #define KY_PLASTIC_RANDOM(isect, rng) kyref_plastic_random(&(isect).position.x, &(isect).wo.x)
#else
    #define KY_PLASTIC_RANDOM(isect, rng) (rng).uniform_float()
#endif
