// ky.hpp -- host-side C++20 class surface of ky-b200.
//
// Keeps the names, constructor signatures and call shapes of the reference's scene / shape /
// material / light / camera / sampler / film / integrator classes (reference ky.cpp, regions
// "geometry" .. "main"), so code written against ky's classes -- in particular its scene
// factories and its render_* entry points -- compiles against this header unchanged.  What is
// different is where the work happens: the objects here only DESCRIBE the scene.
// integrator_t::render() (reference ky.cpp:3689-3729) flattens them into the PODs of kyd.h and
// hands the whole pixel / sample / bounce loop to the sm_100a kernels behind libkyd.so; there is
// no host implementation of intersection, BSDF or light sampling in this file and no CPU
// fallback.
//
// Host-side arithmetic that feeds the device (camera basis, stored normals, areas, plastic lobe
// probabilities, light preprocessing) follows the reference's expressions operation by
// operation, including its use of double sqrt inside vec3_t::normalize (ky.cpp:314).
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <concepts>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <format>
#include <fstream>
#include <limits>
#include <memory>
#include <numbers>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "kyd.h"

// ---- host utilities the reference's entry points use (ky.cpp:34-165) -------------------------
// Build configuration: the reference defines KY_DEBUG under MSVC's _DEBUG and KY_RELEASE otherwise (ky.cpp:36-40), and its
// render_single_scene selects the real render under KY_RELEASE (ky.cpp:4688-4709).
#if defined(_DEBUG)
    #if !defined(KY_DEBUG)
        #define KY_DEBUG
    #endif
#elif !defined(KY_RELEASE)
    #define KY_RELEASE
#endif

namespace ky {
// LOG(fmt, args...): std::format to stdout; LOG_ERROR additionally throws (the reference's error convention, ky.cpp:75-82)
template <class... Ts>
inline void log_message(std::format_string<Ts...> fmt, Ts&&... args)
{
    std::fputs(std::format(fmt, std::forward<Ts>(args)...).c_str(), stdout);
}
template <class... Ts>
[[noreturn]] inline void log_error(std::format_string<Ts...> fmt, Ts&&... args)
{
    const std::string msg = std::format(fmt, std::forward<Ts>(args)...);
    std::fputs(msg.c_str(), stdout);
    throw std::runtime_error(msg);
}
// seconds a callable took.  Wall clock: the reference's clock() adds up the CPU time of all OpenMP threads (ky.cpp:156-163),
// which says nothing about a render that runs on GPUs.
inline float timing_seconds(std::invocable auto function)
{
    const auto start = std::chrono::steady_clock::now();
    function();
    return std::chrono::duration<float>(std::chrono::steady_clock::now() - start).count();
}
} // namespace ky
#ifndef LOG
    #define LOG(...) ::ky::log_message(__VA_ARGS__)
#endif
#ifndef LOG_ERROR
    #define LOG_ERROR(...) ::ky::log_error(__VA_ARGS__)
#endif

namespace ky {

using float01_t = float;
using radian_t = float;
using degree_t = float;

// ---- constants (ky.cpp:180-191) --------------------------------------------------------------
inline constexpr float k_infinity = std::numeric_limits<float>::infinity();
inline constexpr float k_pi = std::numbers::pi;
inline constexpr float k_inv_pi = std::numbers::inv_pi;
inline constexpr float k_inv_4pi = k_inv_pi / 4.f;
constexpr radian_t radians(degree_t degree) { return (k_pi / 180.f) * degree; }

// float transcendental used on the host = the correctly rounded value (DESIGN.md "libm contract")
inline float tan_cr(float x) { return (float)std::tan((double)x); }

// ---- color / vectors (ky.cpp:226-388) ----------------------------------------------------------
struct color_t
{
    float r{}, g{}, b{};

    constexpr color_t() = default;
    constexpr color_t(float r_, float g_, float b_) : r{ r_ }, g{ g_ }, b{ b_ } {}
    constexpr color_t(double r_, double g_, double b_) : r{ (float)r_ }, g{ (float)g_ }, b{ (float)b_ } {}
    constexpr color_t(int r_, int g_, int b_) : r{ (float)r_ }, g{ (float)g_ }, b{ (float)b_ } {}

    color_t operator*(float s) const { return { r * s, g * s, b * s }; }
    color_t operator/(float s) const { return { r / s, g / s, b / s }; }
    color_t operator+(color_t c) const { return { r + c.r, g + c.g, b + c.b }; }
    color_t operator*(color_t c) const { return { r * c.r, g * c.g, b * c.b }; }
    color_t& operator+=(color_t c) { r += c.r; g += c.g; b += c.b; return *this; }
    friend color_t operator*(float s, color_t c) { return { s * c.r, s * c.g, s * c.b }; }

    float max_component_value() const { return std::max({ r, g, b }); }
    float luminance() const { return 0.212671f * r + 0.715160f * g + 0.072169f * b; }
    bool is_black() const { return (r <= 0) && (g <= 0) && (b <= 0); }
};

struct vec2_t
{
    float x{}, y{};
    constexpr vec2_t() = default;
    constexpr vec2_t(float x_, float y_) : x{ x_ }, y{ y_ } {}
    float operator[](int i) const { return i == 0 ? x : y; }
    vec2_t operator+(vec2_t v) const { return { x + v.x, y + v.y }; }
    vec2_t operator-(vec2_t v) const { return { x - v.x, y - v.y }; }
};
using point2_t = vec2_t;
using float2_t = vec2_t;

struct vec3_t
{
    float x{}, y{}, z{};

    constexpr vec3_t() = default;
    constexpr vec3_t(float x_, float y_, float z_) : x{ x_ }, y{ y_ }, z{ z_ } {}
    // the reference's scene data mixes int / float / double literals; each is narrowed to float once
    template <class A, class B, class C>
    constexpr vec3_t(A a, B b, C c) : x{ (float)a }, y{ (float)b }, z{ (float)c } {}

    float operator[](int i) const { return (&x)[i]; }
    vec3_t operator-() const { return { -x, -y, -z }; }
    vec3_t operator+(vec3_t v) const { return { x + v.x, y + v.y, z + v.z }; }
    vec3_t operator-(vec3_t v) const { return { x - v.x, y - v.y, z - v.z }; }
    vec3_t operator*(float s) const { return { x * s, y * s, z * s }; }
    vec3_t operator/(float s) const { return { x / s, y / s, z / s }; }
    friend vec3_t operator*(float s, vec3_t v) { return { v.x * s, v.y * s, v.z * s }; }

    float magnitude_squared() const { return x * x + y * y + z * z; }
    // ky.cpp:310: unqualified sqrt() on a float is ::sqrt(double), rounded back to float
    float magnitude() const { return (float)std::sqrt((double)magnitude_squared()); }
    // ky.cpp:314: v * (float)(1.0 / sqrt((double)|v|^2))
    vec3_t normalize() const { return *this * (float)(1 / std::sqrt((double)(x * x + y * y + z * z))); }

    float dot(vec3_t v) const { return x * v.x + y * v.y + z * v.z; }
    vec3_t cross(vec3_t v) const { return { y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x }; }
};
using point3_t = vec3_t;
using normal_t = vec3_t;
using unit_vec3_t = vec3_t;

inline float dot(vec3_t u, vec3_t v) { return u.dot(v); }
inline vec3_t cross(vec3_t u, vec3_t v) { return u.cross(v); }
inline vec3_t normalize(vec3_t v) { return v.normalize(); }
inline float distance(point3_t a, point3_t b) { return (a - b).magnitude(); }
inline vec3_t lerp(vec3_t u, vec3_t v, float t) { return u + t * (v - u); }
inline vec3_t min(vec3_t a, vec3_t b) { return { std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z) }; }
inline vec3_t max(vec3_t a, vec3_t b) { return { std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z) }; }

inline void store3(float* dst, vec3_t v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
inline void store3(float* dst, color_t c) { dst[0] = c.r; dst[1] = c.g; dst[2] = c.b; }

// ---- bounds / frame: host-only helpers of light preprocessing (ky.cpp:461-578) ----------------
class bounds3_t
{
public:
    bounds3_t()
    {
        constexpr float lo = std::numeric_limits<float>::lowest();
        constexpr float hi = std::numeric_limits<float>::max();
        min_ = point3_t(hi, hi, hi);
        max_ = point3_t(lo, lo, lo);
    }
    bounds3_t(point3_t a, point3_t b) : min_{ ky::min(a, b) }, max_{ ky::max(a, b) } {}

    bounds3_t join(point3_t p) const { return bounds3_t(ky::min(min_, p), ky::max(max_, p)); }
    bounds3_t join(const bounds3_t& b) const { return bounds3_t(ky::min(min_, b.min_), ky::max(max_, b.max_)); }

    bool contain(point3_t p) const
    {
        return p.x >= min_.x && p.x <= max_.x && p.y >= min_.y && p.y <= max_.y && p.z >= min_.z && p.z <= max_.z;
    }

    void bounding_sphere(point3_t* center, float* radius) const
    {
        *center = lerp(min_, max_, 0.5f);
        *radius = contain(*center) ? distance(*center, max_) : 0;
    }

private:
    point3_t min_, max_;
};

class frame_t
{
public:
    frame_t() = default;
    frame_t(normal_t n) : n_{ n.normalize() }
    {
        vec3_t helper = (std::abs(n_.x) > 0.99f) ? vec3_t(0, 1, 0) : vec3_t(1, 0, 0);
        t_ = normalize(cross(n_, helper));
        s_ = normalize(cross(t_, n_));
    }
    vec3_t binormal() const { return s_; }
    vec3_t tangent() const { return t_; }
    vec3_t normal() const { return n_; }

private:
    vec3_t s_{ 1, 0, 0 }, t_{ 0, 1, 0 }, n_{ 0, 0, 1 };
};

// ---- samplers (ky.cpp:877-975) -----------------------------------------------------------------
struct camera_sample_t { point2_t p_film{}; };

inline uint64_t mix64(uint64_t z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

class sampler_t
{
public:
    virtual ~sampler_t() = default;
    sampler_t(int samples_per_pixel) : samples_per_pixel_{ samples_per_pixel } {}

    virtual std::unique_ptr<sampler_t> clone() = 0;

    virtual int ge_samples_per_pixel() { return samples_per_pixel_; } // (sic) the reference's spelling
    virtual void set_samples_per_pixel(int spp) { samples_per_pixel_ = spp; }

    virtual void start_pixel() { current_sample_index_ = 0; }
    virtual bool next_sample() { return ++current_sample_index_ < samples_per_pixel_; }

    virtual float get_float() = 0;
    virtual vec2_t get_float2() = 0;
    virtual camera_sample_t get_camera_sample(point2_t p_film) = 0;

    // what the device needs to know about this sampler
    virtual int device_sampler() const = 0;   // enum kyd_sampler
    virtual uint64_t device_seed() const { return 1234; }

protected:
    int samples_per_pixel_{};
    int current_sample_index_{};
};

// every draw is 0.5 (ky.cpp:922-947)
class debug_sampler_t : public sampler_t
{
public:
    using sampler_t::sampler_t;
    std::unique_ptr<sampler_t> clone() override { return std::make_unique<debug_sampler_t>(samples_per_pixel_); }
    float get_float() override { return 0.5f; }
    vec2_t get_float2() override { return { 0.5f, 0.5f }; }
    camera_sample_t get_camera_sample(point2_t p) override { return { p + vec2_t{ 0.5f, 0.5f } }; }
    int device_sampler() const override { return KYD_SAMPLER_DEBUG; }
};

// The sampler the device reproduces bit for bit: a 48-bit LCG (erand48's multiplier and increment)
// re-seeded at the first draw of every sample from (seed, pixel x, pixel y, sample index).
// Usable on the host too, which is how tests pin the device stream.
class lcg48_sampler_t : public sampler_t
{
public:
    lcg48_sampler_t(int samples_per_pixel, uint64_t seed = 1234) : sampler_t(samples_per_pixel), seed_{ seed } {}

    std::unique_ptr<sampler_t> clone() override { return std::make_unique<lcg48_sampler_t>(samples_per_pixel_, seed_); }

    float get_float() override
    {
        state_ = (state_ * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull;
        return (float)(state_ >> 24) * 0x1p-24f;
    }
    vec2_t get_float2() override
    {
        float x = get_float();
        float y = get_float();
        return { x, y };
    }
    camera_sample_t get_camera_sample(point2_t p_film) override
    {
        uint64_t key = (uint64_t)current_sample_index_ | ((uint64_t)(int)p_film.x << 24) | ((uint64_t)(int)p_film.y << 40);
        state_ = mix64(seed_ * 0x9E3779B97F4A7C15ull + key) >> 16;
        return { p_film + get_float2() };
    }

    int device_sampler() const override { return KYD_SAMPLER_LCG48; }
    uint64_t device_seed() const override { return seed_; }

protected:
    uint64_t seed_{};
    uint64_t state_{};
};

// smallpt's tent filter on a 2x2 sub-pixel grid (smallpt2pbrt/smallpt_rewrite.cpp:397-475 TrapezoidalSampler) over the
// counter-seeded LCG, in float.  samples_per_pixel counts every sample of the pixel and must be a multiple of 4;
// sample s belongs to sub-pixel s / (spp / 4), sub-pixel after sub-pixel like the original.
class trapezoidal_sampler_t : public lcg48_sampler_t
{
public:
    using lcg48_sampler_t::lcg48_sampler_t;

    std::unique_ptr<sampler_t> clone() override { return std::make_unique<trapezoidal_sampler_t>(samples_per_pixel_, seed_); }

    camera_sample_t get_camera_sample(point2_t p_film) override
    {
        uint64_t key = (uint64_t)current_sample_index_ | ((uint64_t)(int)p_film.x << 24) | ((uint64_t)(int)p_film.y << 40);
        state_ = mix64(seed_ * 0x9E3779B97F4A7C15ull + key) >> 16;
        int sub_pixel = current_sample_index_ / (samples_per_pixel_ / 4);
        int sub_x = sub_pixel % 2, sub_y = sub_pixel / 2;
        float random1 = 2 * get_float();
        float random2 = 2 * get_float();
        float delta_x = random1 < 1 ? std::sqrt(random1) - 1 : 1 - std::sqrt(2 - random1);
        float delta_y = random2 < 1 ? std::sqrt(random2) - 1 : 1 - std::sqrt(2 - random2);
        return { p_film + vec2_t{ ((float)sub_x + delta_x + 0.5f) / 2, ((float)sub_y + delta_y + 0.5f) / 2 } };
    }

    int device_sampler() const override { return KYD_SAMPLER_TRAPEZOIDAL; }
};

// The reference's random_sampler_t restarts mt19937_64(1234) on every image ROW and consumes a
// data-dependent number of draws per pixel (ky.cpp:3701, 949-975, 833): a serial dependence across a
// row that no pixel-parallel device can follow.  It is kept as a name; its stream is the
// counter-seeded LCG with the reference's seed 1234 (DESIGN.md "Sampling contract").
class random_sampler_t : public lcg48_sampler_t
{
public:
    random_sampler_t(int samples_per_pixel) : lcg48_sampler_t(samples_per_pixel, 1234) {}
    std::unique_ptr<sampler_t> clone() override { return std::make_unique<random_sampler_t>(samples_per_pixel_); }
};

// ---- shapes (ky.cpp:1009-1519): description only ---------------------------------------------
class shape_t
{
public:
    virtual ~shape_t() = default;
    virtual bounds3_t world_bound() const = 0;
    virtual float area() const = 0;
    virtual kyd_shape describe() const = 0;
};
using shape_sptr_t = std::shared_ptr<shape_t>;
using shape_list_t = std::vector<shape_sptr_t>;

class disk_t : public shape_t
{
public:
    disk_t(point3_t position, normal_t normal, float radius) :
        position_{ position }, normal_{ normalize(normal) }, radius_{ radius }
    {
    }
    bounds3_t world_bound() const override
    {
        frame_t frame{ normal_ };
        vec3_t offset = frame.binormal() * radius_ + frame.tangent() * radius_;
        return bounds3_t(position_ - offset, position_ + offset);
    }
    float area() const override { return k_pi * radius_ * radius_; }
    kyd_shape describe() const override
    {
        kyd_shape s{};
        s.kind = KYD_SHAPE_DISK;
        store3(s.p0, position_);
        store3(s.normal, normal_);
        s.radius = radius_;
        s.area = area();
        return s;
    }

    point3_t position_;
    normal_t normal_;
    float radius_;
};

class triangle_t : public shape_t
{
public:
    triangle_t(point3_t p0, point3_t p1, point3_t p2, bool flip_normal = false) : p0_{ p0 }, p1_{ p1 }, p2_{ p2 }
    {
        normal_ = normalize(cross(p1_ - p0_, p2_ - p0_));
        if (flip_normal)
            normal_ = -normal_;
    }
    bounds3_t world_bound() const override { return bounds3_t(p0_, p1_).join(p2_); }
    // ky.cpp:1222: 0.5 (double) * magnitude (float) is evaluated in double, then narrowed
    float area() const override { return (float)(0.5 * (double)cross(p1_ - p0_, p2_ - p0_).magnitude()); }
    kyd_shape describe() const override
    {
        kyd_shape s{};
        s.kind = KYD_SHAPE_TRIANGLE;
        store3(s.p0, p0_); store3(s.p1, p1_); store3(s.p2, p2_);
        store3(s.normal, normal_);
        s.area = area();
        return s;
    }

    point3_t p0_, p1_, p2_;
    normal_t normal_;
};

class rectangle_t : public shape_t
{
public:
    rectangle_t(point3_t p0, point3_t p1, point3_t p2, point3_t p3, bool flip_normal = false) :
        p0_{ p0 }, p1_{ p1 }, p2_{ p2 }, p3_{ p3 }
    {
        normal_ = normalize(cross(p1_ - p0_, p2_ - p0_));
        if (flip_normal)
            normal_ = -normal_;
    }
    bounds3_t world_bound() const override { return bounds3_t(p0_, p1_).join(p2_).join(p3_); }
    float area() const override { return cross(p0_ - p1_, p2_ - p1_).magnitude(); }
    kyd_shape describe() const override
    {
        kyd_shape s{};
        s.kind = KYD_SHAPE_RECTANGLE;
        store3(s.p0, p0_); store3(s.p1, p1_); store3(s.p2, p2_); store3(s.p3, p3_);
        store3(s.normal, normal_);
        s.area = area();
        return s;
    }

    point3_t p0_, p1_, p2_, p3_;
    normal_t normal_;
};

class sphere_t : public shape_t
{
public:
    sphere_t(vec3_t center, float radius) : center_{ center }, radius_{ radius }, radius_sq_{ radius * radius } {}
    bounds3_t world_bound() const override
    {
        vec3_t half(radius_, radius_, radius_);
        return bounds3_t(center_ + half, center_ - half);
    }
    float area() const override { return 4 * k_pi * radius_sq_; }
    kyd_shape describe() const override
    {
        kyd_shape s{};
        s.kind = KYD_SHAPE_SPHERE;
        store3(s.p0, center_);
        s.radius = radius_;
        s.radius_sq = radius_sq_;
        s.area = area();
        return s;
    }

private:
    vec3_t center_;
    float radius_;
    float radius_sq_;
};

// ---- film (ky.cpp:1545-1836) ---------------------------------------------------------------------
// one device context per process and CUDA device, created on first use
class device_t
{
public:
    // The process-wide device context behind integrator_t::render.  Which GPUs it spans is the environment's choice, so
    // that unchanged ky code scales: KY_CUDA_DEVICES = "all" or a comma-separated list of CUDA ordinals ("0,1,2,3")
    // makes it a multi-GPU context (kyd_create_multi: the samples of every render are split over the devices and the
    // partial films summed on the first); otherwise it is the single device KY_CUDA_DEVICE (default 0).
    static device_t& instance(int cuda_device = -1)
    {
        static device_t dev(cuda_device < 0 ? env_devices() : std::vector<int>{ cuda_device });
        return dev;
    }
    kyd_ctx* ctx() const { return ctx_; }
    int device_count() const { return kyd_device_count(ctx_); }
    void check(int rc) const
    {
        if (rc != KYD_OK)
            throw std::runtime_error(std::string("kyd: ") + kyd_last_error(ctx_));
    }
    ~device_t() { kyd_destroy(ctx_); }

private:
    explicit device_t(const std::vector<int>& devices)
    {
        const int rc = devices.size() == 1 ? kyd_create(&ctx_, devices[0]) : kyd_create_multi(&ctx_, devices.data(), (int)devices.size());
        if (rc != KYD_OK)
            throw std::runtime_error(std::string("kyd_create: ") + kyd_last_error(nullptr));
    }
    static std::vector<int> env_devices()
    {
        std::vector<int> out;
        if (const char* list = std::getenv("KY_CUDA_DEVICES"))
        {
            const std::string s(list);
            if (s == "all")
            {
                // probe by creating contexts until an ordinal is refused
                for (int d = 0; d < KYD_MAX_MULTI; ++d)
                {
                    kyd_ctx* probe = nullptr;
                    if (kyd_create(&probe, d) != KYD_OK) break;
                    kyd_destroy(probe);
                    out.push_back(d);
                }
            }
            else
            {
                size_t pos = 0;
                while (pos < s.size())
                {
                    size_t comma = s.find(',', pos);
                    if (comma == std::string::npos) comma = s.size();
                    if (comma > pos) out.push_back(std::atoi(s.substr(pos, comma - pos).c_str()));
                    pos = comma + 1;
                }
            }
        }
        if (out.empty())
        {
            const char* e = std::getenv("KY_CUDA_DEVICE");
            out.push_back(e ? std::atoi(e) : 0);
        }
        return out;
    }
    kyd_ctx* ctx_{};
};

constexpr float clamp01(float x) { return std::clamp(x, 0.f, 1.f); }
inline color_t clamp01(color_t c) { return color_t(clamp01(c.r), clamp01(c.g), clamp01(c.b)); }
inline uint8_t gamma_encoding(float x) { return (uint8_t)(std::pow((double)clamp01(x), 1 / 2.2) * 255 + .5); }

class film_t
{
public:
    film_t(int width, int height) : width_{ width }, height_{ height }, pixels_((size_t)width * height) {}
    virtual ~film_t() = default;
    film_t(const film_t&) = delete;
    film_t& operator=(const film_t&) = delete;

    int get_width() const { return width_; }
    int get_height() const { return height_; }
    int get_pixel_num() const { return width_ * height_; }
    int get_channels() const { return 3; }

    virtual vec2_t get_resolution() const { return { (float)width_, (float)height_ }; }
    virtual color_t& operator()(int x, int y) { return pixels_[(size_t)width_ * y + x]; }

    void set_color(int x, int y, color_t c) { operator()(x, y) = c; }
    void clear_color(int x, int y) { set_color(x, y, color_t{}); }
    void add_color(int x, int y, color_t delta)
    {
        color_t& c = operator()(x, y);
        c = c + delta;
    }
    void clear(color_t c) { std::fill(pixels_.begin(), pixels_.end(), c); }

    const float* data() const { return &pixels_[0].r; }

    // writes <filename>.bmp (24-bit BGR, bottom-up, gamma 1/2.2), or <filename>.hdr with KY_OUTPUT_HDR, like
    // ky.cpp:1623-1659; the per-pixel work runs on the device (film output stage); no viewer is launched
    virtual bool store_image(std::string filename) const
    {
#ifdef KY_OUTPUT_HDR
        return store_device(filename + ".hdr", KYD_FILM_RGBE);
#else
        return store_device(filename + ".bmp", KYD_FILM_BMP24);
#endif
    }

    // header from kyd_film_header, body from the device (kyd_film_encode); KYD_FILM_GAMMA8 is printed as a P3 ppm
    bool store_device(const std::string& path, int format) const
    {
        const int64_t bytes = kyd_film_body_bytes(format, width_, height_);
        uint8_t header[128];
        const int header_bytes = kyd_film_header(format, width_, height_, header, (int)sizeof(header));
        if (bytes < 0 || header_bytes < 0) return false;
        std::vector<uint8_t> body((size_t)bytes);
        device_t& dev = device_t::instance();
        dev.check(kyd_film_encode(dev.ctx(), data(), width_, height_, format, body.data()));
        std::ofstream out(path, std::ios::binary);
        if (!out) return false;
        out.write((const char*)header, header_bytes);
        if (format == KYD_FILM_GAMMA8)
        {
            std::string text;
            text.reserve(body.size() * 4);
            for (uint8_t v : body) { text += std::to_string((int)v); text += ' '; }
            out.write(text.data(), (std::streamsize)text.size());
        }
        else
            out.write((const char*)body.data(), (std::streamsize)body.size());
        return (bool)out;
    }

    // host-side writers with the reference's arithmetic (store_bmp_impl / store_ppm_impl / store_hdr_impl)

    bool store_bmp(const std::string& path) const
    {
        std::ofstream out(path, std::ios::binary);
        if (!out) return false;
        const uint32_t line = (uint32_t)width_ * 3, padded = (line + 3u) & ~3u;
        const uint32_t header[13] = { 14 + 40 + padded * (uint32_t)height_, 0, 54, 40,
            (uint32_t)width_, (uint32_t)height_, 1u | (24u << 16), 0, 0, 0, 0, 0, 0 };
        out.write("BM", 2);
        out.write((const char*)header, sizeof(header));
        std::vector<uint8_t> row(line);
        for (int y = height_ - 1; y >= 0; --y) // the reference writes unpadded lines, ky.cpp:1730-1733
        {
            for (int x = 0; x < width_; ++x)
            {
                const color_t& c = pixels_[(size_t)width_ * y + x];
                row[3 * x + 0] = gamma_encoding(c.b);
                row[3 * x + 1] = gamma_encoding(c.g);
                row[3 * x + 2] = gamma_encoding(c.r);
            }
            out.write((const char*)row.data(), line);
        }
        return true;
    }

    bool store_ppm(const std::string& path) const
    {
        std::ofstream out(path, std::ios::binary);
        if (!out) return false;
        out << "P3\n" << width_ << ' ' << height_ << "\n255\n";
        const float* f = data();
        for (size_t i = 0; i < pixels_.size() * 3; ++i)
            out << (int)gamma_encoding(f[i]) << ' ';
        return true;
    }

    // Radiance RGBE, flat (ky.cpp:1739-1782)
    bool store_hdr(const std::string& path) const
    {
        std::ofstream out(path, std::ios::binary);
        if (!out) return false;
        out << "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y " << height_ << " +X " << width_ << "\n";
        for (const color_t& c : pixels_)
        {
            uint8_t rgbe[4]{};
            float v = std::max({ c.r, c.g, c.b });
            if (v >= 1e-32f)
            {
                int e;
                float m = float(std::frexp(v, &e) * 256.f / v);
                rgbe[0] = uint8_t(c.r * m); rgbe[1] = uint8_t(c.g * m); rgbe[2] = uint8_t(c.b * m);
                rgbe[3] = uint8_t(e + 128);
            }
            out.write((const char*)rgbe, 4);
        }
        return true;
    }

private:
    int32_t width_{}, height_{};
    std::vector<color_t> pixels_;
};

// several sub-films in one image; get_resolution() is the SUB-film size (ky.cpp:1802-1836)
class film_grid_t : public film_t
{
public:
    film_grid_t(int row, int column, int sub_width, int sub_height) :
        film_t(column * sub_width, row * sub_height), row_{ row }, column_{ column }, sub_width_{ sub_width }, sub_height_{ sub_height }
    {
    }
    vec2_t get_resolution() const override { return { (float)sub_width_, (float)sub_height_ }; }
    color_t& operator()(int x, int y) override
    {
        int col = subfilm_index % column_, row = subfilm_index / column_;
        return film_t::operator()(x + col * sub_width_, y + row * sub_height_);
    }
    void next_subfilm() { ++subfilm_index; }

private:
    int row_{}, column_{};
    int subfilm_index{};
    int sub_width_{}, sub_height_{};
};

// ---- camera (ky.cpp:1859-1906) -------------------------------------------------------------------
class camera_t
{
public:
    virtual ~camera_t() = default;
    // ray_origin_push: 0 in ky; smallpt starts rays 140 units along the unnormalised direction
    // (smallpt2pbrt/smallpt_rewrite.cpp:676), needed for BASELINE config 1
    camera_t(vec3_t position, vec3_t front, vec3_t up, degree_t fov, vec2_t resolution, float ray_origin_push = 0) :
        position_{ position }, front_{ front.normalize() }, up_{ up.normalize() }, resolution_{ resolution }, push_{ ray_origin_push }
    {
        float tan_fov = tan_cr(radians(fov) / 2);
        right_ = up_.cross(front_).normalize() * tan_fov * (resolution_.x / resolution_.y);
        up_ = front_.cross(right_).normalize() * tan_fov;
    }

    kyd_camera describe() const
    {
        kyd_camera c{};
        store3(c.position, position_); store3(c.front, front_); store3(c.right, right_); store3(c.up, up_);
        c.resolution[0] = resolution_.x; c.resolution[1] = resolution_.y;
        c.origin_push = push_;
        return c;
    }

private:
    vec3_t position_, front_, right_, up_;
    vec2_t resolution_;
    float push_{};
};
using const_camera_sptr_t = std::shared_ptr<const camera_t>;

// ---- materials (ky.cpp:2568-2682) ------------------------------------------------------------------
class material_t
{
public:
    virtual ~material_t() = default;
    virtual kyd_material describe() const = 0;
};
using material_sptr_t = std::shared_ptr<material_t>;
using material_list_t = std::vector<material_sptr_t>;

class matte_material_t : public material_t
{
public:
    matte_material_t(color_t diffuse_color) : diffuse_color_{ diffuse_color } {}
    kyd_material describe() const override
    {
        kyd_material m{};
        m.kind = KYD_MAT_MATTE;
        store3(m.diffuse, diffuse_color_);
        return m;
    }
private:
    color_t diffuse_color_{};
};

class mirror_material_t : public material_t
{
public:
    mirror_material_t(color_t specular_color) : specular_color_{ specular_color } {}
    kyd_material describe() const override
    {
        kyd_material m{};
        m.kind = KYD_MAT_MIRROR;
        store3(m.specular, specular_color_);
        return m;
    }
private:
    color_t specular_color_{};
};

class glass_material_t : public material_t
{
public:
    glass_material_t(float eta, color_t reflection_color = color_t{ 1, 1, 1 }, color_t transmission_color = color_t{ 1, 1, 1 }) :
        eta_{ eta }, reflection_color_{ reflection_color }, transmission_color_{ transmission_color }
    {
    }
    kyd_material describe() const override
    {
        kyd_material m{};
        m.kind = KYD_MAT_GLASS;
        store3(m.specular, reflection_color_);
        store3(m.transmission, transmission_color_);
        m.eta = eta_;
        return m;
    }
private:
    float eta_{};
    color_t reflection_color_{}, transmission_color_{};
};

class plastic_material_t : public material_t
{
public:
    plastic_material_t(color_t diffuse_color, color_t specular_color, float shininess) :
        diffuse_color_{ diffuse_color }, specular_color_{ specular_color }, exponent_{ shininess }
    {
        float diffuse = diffuse_color.luminance();
        float specular = specular_color.luminance();
        float luminance = diffuse + specular;
        diffuse_probility_ = diffuse / luminance;
        specular_probility_ = specular / luminance;
    }
    kyd_material describe() const override
    {
        kyd_material m{};
        m.kind = KYD_MAT_PLASTIC;
        store3(m.diffuse, diffuse_color_);
        store3(m.specular, specular_color_);
        m.exponent = exponent_;
        m.diffuse_probability = diffuse_probility_;
        m.specular_probability = specular_probility_;
        return m;
    }
private:
    color_t diffuse_color_{}, specular_color_{};
    float exponent_{};
    float diffuse_probility_{}, specular_probility_{};
};

// ---- lights (ky.cpp:2764-3062) ---------------------------------------------------------------------
class scene_t;

class light_t
{
public:
    virtual ~light_t() = default;
    light_t(point3_t world_position, int samples_num = 1) : world_position_{ world_position }, samples_num_{ samples_num } {}

    virtual bool is_delta() const = 0;
    virtual bool is_finite() const = 0;
    virtual void preprocess(const scene_t&) {}
    virtual color_t power() const = 0;
    virtual kyd_light describe(const std::vector<const shape_t*>& shape_index) const = 0;

protected:
    point3_t world_position_;
    int samples_num_;
};
using light_sptr_t = std::shared_ptr<light_t>;
using light_list_t = std::vector<light_sptr_t>;

class point_light_t : public light_t
{
public:
    point_light_t(point3_t world_position, int samples_num, color_t intensity) :
        light_t(world_position, samples_num), intensity_{ intensity }
    {
    }
    bool is_delta() const override { return true; }
    bool is_finite() const override { return true; }
    color_t power() const override { return 4 * k_pi * intensity_; }
    kyd_light describe(const std::vector<const shape_t*>&) const override
    {
        kyd_light l{};
        l.kind = KYD_LIGHT_POINT;
        store3(l.color, intensity_);
        store3(l.position, world_position_);
        l.shape = -1;
        return l;
    }
private:
    color_t intensity_{};
};

class direction_light_t : public light_t
{
public:
    direction_light_t(point3_t world_position, int samples_num, color_t irradiance, const vec3_t world_direction) :
        light_t(world_position, samples_num), irradiance_{ irradiance }, world_direction_{ normalize(world_direction) }
    {
    }
    bool is_delta() const override { return true; }
    bool is_finite() const override { return false; }
    void preprocess(const scene_t& scene) override;
    color_t power() const override { return power_; }
    kyd_light describe(const std::vector<const shape_t*>&) const override
    {
        kyd_light l{};
        l.kind = KYD_LIGHT_DIRECTION;
        store3(l.color, irradiance_);
        store3(l.direction, world_direction_);
        l.world_radius = world_radius_;
        l.shape = -1;
        return l;
    }
private:
    color_t irradiance_{};
    point3_t world_center_{};
    float world_radius_{};
    color_t power_{};
    unit_vec3_t world_direction_{};
};

class area_light_t : public light_t
{
public:
    area_light_t(point3_t world_position, int samples_num, color_t radiance, const shape_t* shape) :
        light_t(world_position, samples_num), radiance_{ radiance }, shape_{ shape }, power_{ radiance_ * shape->area() * k_pi }
    {
    }
    bool is_delta() const override { return false; }
    bool is_finite() const override { return true; }
    color_t power() const override { return power_; }
    kyd_light describe(const std::vector<const shape_t*>& shape_index) const override
    {
        kyd_light l{};
        l.kind = KYD_LIGHT_AREA;
        store3(l.color, radiance_);
        auto it = std::find(shape_index.begin(), shape_index.end(), shape_);
        if (it == shape_index.end())
            throw std::runtime_error("area_light_t: shape is not in the scene's shape list");
        l.shape = (int32_t)(it - shape_index.begin());
        return l;
    }
private:
    color_t radiance_{};
    const shape_t* shape_{};
    color_t power_{};
};

class environment_light_t : public light_t
{
public:
    environment_light_t(point3_t world_position, int samples_num, color_t radiance) :
        light_t(world_position, samples_num), radiance_{ radiance }
    {
    }
    bool is_delta() const override { return false; }
    bool is_finite() const override { return false; }
    void preprocess(const scene_t& scene) override;
    color_t power() const override { return power_; }
    kyd_light describe(const std::vector<const shape_t*>&) const override
    {
        kyd_light l{};
        l.kind = KYD_LIGHT_ENVIRONMENT;
        store3(l.color, radiance_);
        l.world_radius = world_radius_;
        l.shape = -1;
        return l;
    }
private:
    color_t radiance_{};
    point3_t world_center_{};
    float world_radius_{};
    color_t power_{};
};

// ---- surface + scene (ky.cpp:3071-3237) ------------------------------------------------------------
struct surface_t
{
    const shape_t* shape{};
    const material_t* material{};
    const area_light_t* area_light{};
};
using surface_list_t = std::vector<surface_t>;

enum class cornell_box_enum_t
{
    none,
    light_area = 1, light_direction = 2, light_point = 4, light_environment = 8,
    large_mirror_sphere = 16, large_glass_sphere = 32, small_mirror_sphere = 64, small_glass_sphere = 128,
    glossy_floor = 256,
    both_small_spheres = small_mirror_sphere | small_glass_sphere,
    both_large_spheres = large_mirror_sphere | large_glass_sphere,
    default_scene = both_small_spheres | light_area,
};
constexpr cornell_box_enum_t operator|(cornell_box_enum_t a, cornell_box_enum_t b) { return (cornell_box_enum_t)((int)a | (int)b); }
constexpr cornell_box_enum_t operator&(cornell_box_enum_t a, cornell_box_enum_t b) { return (cornell_box_enum_t)((int)a & (int)b); }
constexpr bool enum_have(cornell_box_enum_t group, cornell_box_enum_t value) { return (group & value) != (cornell_box_enum_t)0; }

// flattened scene: owns the arrays a kyd_scene_desc points into
struct flat_scene_t
{
    std::vector<kyd_shape> shapes;
    std::vector<kyd_material> materials;
    std::vector<kyd_light> lights;
    std::vector<kyd_surface> surfaces;
    kyd_scene_desc desc{};
};

class scene_t
{
public:
    scene_t() = default;
    scene_t(const_camera_sptr_t camera, shape_list_t shape_list, material_list_t material_list, light_list_t light_list,
        surface_list_t surface_list, environment_light_t* env_light = nullptr) :
        camera_{ std::move(camera) }, shape_list_{ std::move(shape_list) }, material_list_{ std::move(material_list) },
        light_list_{ std::move(light_list) }, environment_light_{ env_light }, surface_list_{ std::move(surface_list) }
    {
        for (light_sptr_t& light : light_list_)
            light->preprocess(*this);
    }
    scene_t(scene_t&&) = default;
    scene_t& operator=(scene_t&&) = default;
    scene_t(const scene_t&) = delete;
    scene_t& operator=(const scene_t&) = delete;

    bounds3_t world_bound() const
    {
        bounds3_t b;
        for (const surface_t& s : surface_list_)
            b = b.join(s.shape->world_bound());
        return b;
    }

    const camera_t* get_camera() const { return camera_.get(); }
    int light_count() const { return (int)light_list_.size(); }
    const light_list_t& light_list() const { return light_list_; }
    const environment_light_t* environment_light() const { return environment_light_; }

    // the SoA-ready description handed to kyd_upload_scene(); pointers become indices
    std::unique_ptr<flat_scene_t> flatten() const
    {
        auto flat = std::make_unique<flat_scene_t>();

        // shapes: the shape list first (keeps its order), then any shape only a surface or light knows
        std::vector<const shape_t*> shape_index;
        for (const shape_sptr_t& s : shape_list_) shape_index.push_back(s.get());
        for (const surface_t& s : surface_list_)
            if (std::find(shape_index.begin(), shape_index.end(), s.shape) == shape_index.end())
                shape_index.push_back(s.shape);
        for (const shape_t* s : shape_index) flat->shapes.push_back(s->describe());

        std::vector<const material_t*> material_index;
        for (const material_sptr_t& m : material_list_) material_index.push_back(m.get());
        for (const surface_t& s : surface_list_)
            if (std::find(material_index.begin(), material_index.end(), s.material) == material_index.end())
                material_index.push_back(s.material);
        for (const material_t* m : material_index) flat->materials.push_back(m->describe());

        std::vector<const light_t*> light_index;
        for (const light_sptr_t& l : light_list_)
        {
            light_index.push_back(l.get());
            flat->lights.push_back(l->describe(shape_index));
        }

        for (const surface_t& s : surface_list_)
        {
            kyd_surface f{};
            f.shape = (int32_t)(std::find(shape_index.begin(), shape_index.end(), s.shape) - shape_index.begin());
            f.material = (int32_t)(std::find(material_index.begin(), material_index.end(), s.material) - material_index.begin());
            f.area_light = -1;
            if (s.area_light)
            {
                auto it = std::find(light_index.begin(), light_index.end(), (const light_t*)s.area_light);
                if (it == light_index.end())
                    throw std::runtime_error("scene_t: a surface's area light is not in the light list");
                f.area_light = (int32_t)(it - light_index.begin());
            }
            flat->surfaces.push_back(f);
        }

        kyd_scene_desc& d = flat->desc;
        d.camera = camera_->describe();
        d.shape_count = (int32_t)flat->shapes.size();       d.shapes = flat->shapes.data();
        d.material_count = (int32_t)flat->materials.size(); d.materials = flat->materials.data();
        d.light_count = (int32_t)flat->lights.size();       d.lights = flat->lights.data();
        d.surface_count = (int32_t)flat->surfaces.size();   d.surfaces = flat->surfaces.data();
        d.environment_light = -1;
        if (environment_light_)
            d.environment_light = (int32_t)(std::find(light_index.begin(), light_index.end(), (const light_t*)environment_light_) - light_index.begin());
        return flat;
    }

public:
    static scene_t create_cornell_box_scene(cornell_box_enum_t scene_enum, point2_t film_resolution);
    static scene_t create_mis_scene(point2_t film_resolution);
    // BASELINE config 1: smallpt's nine spheres expressed with these FP32 classes
    static scene_t create_smallpt_scene(point2_t film_resolution);
    // coverage scene: disk / triangle shapes, every light kind (tests)
    static scene_t create_shapes_scene(point2_t film_resolution);

private:
    const_camera_sptr_t camera_;
    shape_list_t shape_list_;
    material_list_t material_list_;
    light_list_t light_list_;
    environment_light_t* environment_light_{};
    surface_list_t surface_list_;
};

inline void direction_light_t::preprocess(const scene_t& scene)
{
    scene.world_bound().bounding_sphere(&world_center_, &world_radius_);
    power_ = irradiance_ * (k_pi * world_radius_ * world_radius_);
}

inline void environment_light_t::preprocess(const scene_t& scene)
{
    scene.world_bound().bounding_sphere(&world_center_, &world_radius_);
    power_ = radiance_ * (k_pi * world_radius_ * world_radius_);
}

// ---- scene data (ky.cpp:3240-3533; Appendix B of SURVEY.md) ------------------------------------------

inline scene_t scene_t::create_cornell_box_scene(cornell_box_enum_t scene_enum, point2_t film_resolution)
{
    using enum cornell_box_enum_t;
    if (enum_have(scene_enum, large_mirror_sphere) && enum_have(scene_enum, large_glass_sphere))
        throw std::runtime_error("create_cornell_box_scene: cannot set both large balls");

    auto camera = std::make_shared<camera_t>(
        point3_t{ -0.0439815f, 4.12529f, 0.222539f }, vec3_t{ 0.00688625f, -0.998505f, -0.0542161f },
        vec3_t{ 3.73896e-4f, -0.0542148f, 0.998529f }, 80, film_resolution);

    auto matte = [](color_t c) { return std::make_shared<matte_material_t>(c); };
    material_sptr_t black = matte(color_t()), white = matte(color_t(.8, .8, .8));
    material_sptr_t red = matte(color_t(0.803922f, 0.152941f, 0.152941f));
    material_sptr_t green = matte(color_t(0.156863f, 0.803922f, 0.172549f));
    material_sptr_t blue = matte(color_t(0.156863f, 0.172549f, 0.803922f));
    material_sptr_t glossy = std::make_shared<plastic_material_t>(color_t(.1, .1, .1), color_t(.7, .7, .7), 90.f);
    material_sptr_t mirror_mat = std::make_shared<mirror_material_t>(color_t(1, 1, 1));
    material_sptr_t glass_mat = std::make_shared<glass_material_t>((float)1.6);
    material_list_t material_list{ black, white, red, green, blue, glossy, mirror_mat, glass_mat };

    // box corners: x in {x0,x1}, y in {y0,y1}, z in {-zc,zc}; index = 4*(y==y1) + corner-in-quad
    const float x0 = -1.27029f, x1 = 1.28975f, y0 = -1.30455f, y1 = 1.25549f, zc = 1.28002f;
    const vec3_t cb[8] = {
        { x0, y0, -zc }, { x1, y0, -zc }, { x1, y0, zc }, { x0, y0, zc },
        { x0, y1, -zc }, { x1, y1, -zc }, { x1, y1, zc }, { x0, y1, zc } };
    auto quad = [](const vec3_t* v, int a, int b, int c, int d) { return std::make_shared<rectangle_t>(v[a], v[b], v[c], v[d]); };
    shape_sptr_t left = quad(cb, 3, 0, 4, 7), right = quad(cb, 1, 2, 6, 5), back = quad(cb, 0, 3, 2, 1);
    shape_sptr_t bottom = quad(cb, 0, 1, 5, 4), top = quad(cb, 2, 3, 7, 6);

    const float large_radius = 0.8f, small_radius = 0.5f;
    vec3_t large_center = (cb[0] + cb[4] + cb[5] + cb[1]) * (1.f / 4.f) + vec3_t(0, 0, large_radius);
    vec3_t left_wall_center = (cb[0] + cb[4]) * (1.f / 2.f) + vec3_t(0, 0, small_radius);
    vec3_t right_wall_center = (cb[1] + cb[5]) * (1.f / 2.f) + vec3_t(0, 0, small_radius);
    float length_x = right_wall_center.x - left_wall_center.x;
    vec3_t left_center = left_wall_center + vec3_t(2.f * length_x / 7.f, 0.f, 0.f);
    vec3_t right_center = right_wall_center - vec3_t(2.f * length_x / 7.f, 0.f, 0.f);
    shape_sptr_t large_ball = std::make_shared<sphere_t>(large_center, large_radius);
    shape_sptr_t left_ball = std::make_shared<sphere_t>(left_center, small_radius);
    shape_sptr_t right_ball = std::make_shared<sphere_t>(right_center, small_radius);

    // light box under the ceiling
    const float h = 0.25f, z_lo = 1.26002f, z_hi = 1.28002f;
    const vec3_t lb[8] = {
        { -h, -h, z_lo }, { h, -h, z_lo }, { h, -h, z_hi }, { -h, -h, z_hi },
        { -h, h, z_lo }, { h, h, z_lo }, { h, h, z_hi }, { -h, h, z_hi } };
    shape_sptr_t left2 = quad(lb, 3, 7, 4, 0), right2 = quad(lb, 1, 5, 6, 2), front2 = quad(lb, 4, 7, 6, 5);
    shape_sptr_t back2 = quad(lb, 0, 1, 2, 3), bottom2 = quad(lb, 0, 4, 5, 1);

    shape_list_t shape_list{ left, right, back, bottom, top, large_ball, left_ball, right_ball, left2, right2, front2, back2, bottom2 };

    light_list_t light_list{};
    if (enum_have(scene_enum, light_area))
        light_list.push_back(std::make_shared<area_light_t>(point3_t(), 1, color_t(25, 25, 25), bottom2.get()));
    if (enum_have(scene_enum, light_direction))
        light_list.push_back(std::make_shared<direction_light_t>(point3_t(), 1, color_t(10, 4, 0), vec3_t(-1, -1.5, -1)));
    if (enum_have(scene_enum, light_point))
    {
        float I = 70 * k_inv_4pi;
        light_list.push_back(std::make_shared<point_light_t>(point3_t(0.0, 0.5, 1.0), 1, color_t(I, I, I)));
    }
    environment_light_t* environment_light{};
    if (enum_have(scene_enum, light_environment))
    {
        auto light = std::make_shared<environment_light_t>(point3_t(), 1, color_t(135. / 255, 206. / 255, 250. / 255));
        light_list.push_back(light);
        environment_light = light.get();
    }

    surface_list_t surface_list{
        { left.get(), green.get(), nullptr }, { right.get(), red.get(), nullptr }, { top.get(), white.get(), nullptr },
        { bottom.get(), glossy.get(), nullptr }, { back.get(), blue.get(), nullptr } };
    if (enum_have(scene_enum, large_mirror_sphere))
        surface_list.push_back({ large_ball.get(), mirror_mat.get(), nullptr });
    else if (enum_have(scene_enum, large_glass_sphere))
        surface_list.push_back({ large_ball.get(), glass_mat.get(), nullptr });
    if (enum_have(scene_enum, small_mirror_sphere))
        surface_list.push_back({ left_ball.get(), mirror_mat.get(), nullptr });
    if (enum_have(scene_enum, small_glass_sphere))
        surface_list.push_back({ right_ball.get(), glass_mat.get(), nullptr });
    if (enum_have(scene_enum, light_area))
    {
        for (const shape_sptr_t& side : { left2, right2, front2, back2 })
            surface_list.push_back({ side.get(), white.get(), nullptr });
        surface_list.push_back({ bottom2.get(), black.get(), (area_light_t*)light_list[0].get() });
    }

    return scene_t{ camera, shape_list, material_list, light_list, surface_list, environment_light };
}

inline scene_t scene_t::create_mis_scene(point2_t film_resolution)
{
    auto camera = std::make_shared<camera_t>(point3_t{ 0, 2, -15 }, vec3_t{ 0, -4, 12.5 }, vec3_t{ 0, 1, 0 }, 50, film_resolution);

    material_sptr_t black = std::make_shared<matte_material_t>(color_t());
    material_sptr_t gray = std::make_shared<matte_material_t>(color_t(.4, .4, .4));
    material_sptr_t silver = std::make_shared<plastic_material_t>(color_t(0.07, 0.09, 0.13), color_t(1, 1, 1), 5000.f);
    material_list_t material_list{ black, gray, silver };

    auto flipped = [](point3_t a, point3_t b, point3_t c, point3_t d) { return std::make_shared<rectangle_t>(a, b, c, d, true); };
    shape_sptr_t bottom = flipped({ -10, -4.14615, 10 }, { -10, -4.14615, -10 }, { 10, -4.14615, -10 }, { 10, -4.14615, 10 });
    shape_sptr_t back = flipped({ -10, -10, 2 }, { -10, 10, 2 }, { 10, 10, 2 }, { 10, -10, 2 });
    // four planks: (y,z) of the near edge, (y,z) of the far edge, x from 4 to -4
    auto plank = [&](double ya, double za, double yb, double zb) { return flipped({ 4, ya, za }, { 4, yb, zb }, { -4, yb, zb }, { -4, ya, za }); };
    shape_sptr_t plank0 = plank(-2.70651, -0.25609, -2.08375, 0.526323);
    shape_sptr_t plank1 = plank(-3.28825, -1.36972, -2.83856, -0.476536);
    shape_sptr_t plank2 = plank(-3.73096, -2.70046, -3.43378, -1.74564);
    shape_sptr_t plank3 = plank(-3.99615, -4.0667, -3.82069, -3.08221);

    shape_sptr_t ball0 = std::make_shared<sphere_t>(point3_t(10, 10, -4), 0.5f);
    shape_sptr_t ball1 = std::make_shared<sphere_t>(point3_t(-3.75, 0, 0), (float)0.03333);
    shape_sptr_t ball2 = std::make_shared<sphere_t>(point3_t(-1.25, 0, 0), (float)0.1);
    shape_sptr_t ball3 = std::make_shared<sphere_t>(point3_t(1.25, 0, 0), (float)0.3);
    shape_sptr_t ball4 = std::make_shared<sphere_t>(point3_t(3.75, 0, 0), (float)0.9);
    shape_list_t shape_list{ bottom, back, plank0, plank1, plank2, plank3, ball0, ball1, ball2, ball3, ball4 };

    // NOTE the reference wires light1 to ball2's shape and light2 to ball1's (ky.cpp:3498-3499) while the
    // surfaces below attach light1 to ball1 and light2 to ball2 (ky.cpp:3525-3526); kept as is
    auto emitter = [](double radiance, const shape_sptr_t& shape) { return std::make_shared<area_light_t>(point3_t(), 1, color_t(radiance, radiance, radiance), shape.get()); };
    auto light0 = emitter(800, ball0), light1 = emitter(901.803, ball2), light2 = emitter(100, ball1);
    auto light3 = emitter(11.1111, ball3), light4 = emitter(1.23457, ball4);
    light_list_t light_list{ light0, light1, light2, light3, light4 };

    surface_list_t surface_list{
        { bottom.get(), gray.get(), nullptr }, { back.get(), gray.get(), nullptr },
        { plank0.get(), silver.get(), nullptr }, { plank1.get(), silver.get(), nullptr },
        { plank2.get(), silver.get(), nullptr }, { plank3.get(), silver.get(), nullptr },
        { ball0.get(), black.get(), light0.get() }, { ball1.get(), black.get(), light1.get() },
        { ball2.get(), black.get(), light2.get() }, { ball3.get(), black.get(), light3.get() },
        { ball4.get(), black.get(), light4.get() } };

    return scene_t{ camera, shape_list, material_list, light_list, surface_list };
}

inline scene_t scene_t::create_smallpt_scene(point2_t film_resolution)
{
    // smallpt2pbrt/smallpt_rewrite.cpp:1199-1244 (scene), :1391-1392 (camera), :676 (origin push)
    auto camera = std::make_shared<camera_t>(point3_t{ 50, 52, -295.6 }, vec3_t{ 0, -0.042612, 1 }, vec3_t{ 0, 1, 0 }, 53, film_resolution, 140.f);

    struct ball_t { double x, y, z, r; };
    const ball_t balls[9] = {
        { 1e5 + 1, 40.8, -81.6, 1e5 }, { -1e5 + 99, 40.8, -81.6, 1e5 }, { 50, 40.8, -1e5, 1e5 }, { 50, 40.8, 1e5 - 170, 1e5 },
        { 50, 1e5, -81.6, 1e5 }, { 50, -1e5 + 81.6, -81.6, 1e5 },
        { 27, 16.5, -47, 16.5 }, { 73, 16.5, -78, 16.5 }, { 50, 681.6 - .27, -81.6, 600 } };
    shape_list_t shape_list;
    for (const ball_t& b : balls)
        shape_list.push_back(std::make_shared<sphere_t>(vec3_t(b.x, b.y, b.z), (float)b.r));

    material_sptr_t red = std::make_shared<matte_material_t>(color_t(.75, .25, .25));
    material_sptr_t blue = std::make_shared<matte_material_t>(color_t(.25, .25, .75));
    material_sptr_t gray = std::make_shared<matte_material_t>(color_t(.75, .75, .75));
    material_sptr_t black = std::make_shared<matte_material_t>(color_t());
    material_sptr_t mirror_mat = std::make_shared<mirror_material_t>(color_t(.999, .999, .999));
    material_sptr_t glass_mat = std::make_shared<glass_material_t>((float)1.5, color_t(.999, .999, .999), color_t(.999, .999, .999));
    material_list_t material_list{ red, blue, gray, black, mirror_mat, glass_mat };

    auto area = std::make_shared<area_light_t>(point3_t(), 1, color_t(12, 12, 12), shape_list[8].get());
    light_list_t light_list{ area };

    const material_sptr_t order[9] = { red, blue, gray, black, gray, gray, mirror_mat, glass_mat, black };
    surface_list_t surface_list;
    for (int i = 0; i < 9; ++i)
        surface_list.push_back({ shape_list[i].get(), order[i].get(), i == 8 ? area.get() : nullptr });

    return scene_t{ camera, shape_list, material_list, light_list, surface_list };
}

inline scene_t scene_t::create_shapes_scene(point2_t film_resolution)
{
    auto camera = std::make_shared<camera_t>(point3_t{ 0.1f, 3.6f, 0.4f }, vec3_t{ -0.02f, -1.f, -0.08f }, vec3_t{ 0, 0, 1 }, 70, film_resolution);

    material_sptr_t black = std::make_shared<matte_material_t>(color_t());
    material_sptr_t white = std::make_shared<matte_material_t>(color_t(.7, .7, .7));
    material_sptr_t orange = std::make_shared<matte_material_t>(color_t(.8, .45, .15));
    material_sptr_t glossy = std::make_shared<plastic_material_t>(color_t(.2, .25, .3), color_t(.5, .5, .5), 30.f);
    material_sptr_t mirror_mat = std::make_shared<mirror_material_t>(color_t(.9, .9, .9));
    material_sptr_t glass_mat = std::make_shared<glass_material_t>((float)1.45);
    material_list_t material_list{ black, white, orange, glossy, mirror_mat, glass_mat };

    shape_sptr_t floor = std::make_shared<rectangle_t>(point3_t(-2, -2, -1), point3_t(2, -2, -1), point3_t(2, 2, -1), point3_t(-2, 2, -1));
    shape_sptr_t wall = std::make_shared<rectangle_t>(point3_t(-2, -2, -1), point3_t(-2, -2, 2), point3_t(2, -2, 2), point3_t(2, -2, -1));
    shape_sptr_t tri0 = std::make_shared<triangle_t>(point3_t(-1.6f, -1.2f, -1.f), point3_t(-0.4f, -1.5f, -1.f), point3_t(-1.1f, -1.4f, 0.7f));
    shape_sptr_t tri1 = std::make_shared<triangle_t>(point3_t(1.7f, -0.9f, -0.99f), point3_t(0.6f, -1.3f, -0.99f), point3_t(1.2f, -1.6f, 0.9f), true);
    shape_sptr_t disk0 = std::make_shared<disk_t>(point3_t(0.2f, -0.3f, -0.6f), vec3_t(0.1f, 0.4f, 1.f), 0.55f);
    shape_sptr_t ball = std::make_shared<sphere_t>(vec3_t(-0.9f, 0.4f, -0.6f), 0.4f);
    shape_sptr_t gball = std::make_shared<sphere_t>(vec3_t(0.9f, 0.6f, -0.65f), 0.35f);
    shape_sptr_t ltri = std::make_shared<triangle_t>(point3_t(-0.5f, -0.6f, 1.6f), point3_t(0.5f, -0.6f, 1.6f), point3_t(0.f, 0.4f, 1.7f), true);
    shape_sptr_t ldisk = std::make_shared<disk_t>(point3_t(1.5f, 0.2f, 1.2f), vec3_t(-1.f, 0.f, -0.6f), 0.3f);
    shape_sptr_t lrect = std::make_shared<rectangle_t>(point3_t(-1.9f, 0.5f, 0.2f), point3_t(-1.9f, 1.1f, 0.2f), point3_t(-1.9f, 1.1f, 0.8f), point3_t(-1.9f, 0.5f, 0.8f));
    shape_sptr_t lball = std::make_shared<sphere_t>(vec3_t(0.f, 1.2f, 0.2f), 0.12f);
    shape_list_t shape_list{ floor, wall, tri0, tri1, disk0, ball, gball, ltri, ldisk, lrect, lball };

    auto l_tri = std::make_shared<area_light_t>(point3_t(), 1, color_t(18, 17, 15), ltri.get());
    auto l_disk = std::make_shared<area_light_t>(point3_t(), 1, color_t(9, 14, 20), ldisk.get());
    auto l_rect = std::make_shared<area_light_t>(point3_t(), 1, color_t(6, 9, 5), lrect.get());
    auto l_ball = std::make_shared<area_light_t>(point3_t(), 1, color_t(30, 22, 12), lball.get());
    auto l_pnt = std::make_shared<point_light_t>(point3_t(-1.2f, 1.5f, 1.4f), 1, color_t(1.5f, 1.2f, 2.f));
    auto l_dir = std::make_shared<direction_light_t>(point3_t(), 1, color_t(.6f, .5f, .3f), vec3_t(.4f, -1.f, -.7f));
    auto l_env = std::make_shared<environment_light_t>(point3_t(), 1, color_t(.12f, .16f, .22f));
    light_list_t light_list{ l_tri, l_disk, l_rect, l_ball, l_pnt, l_dir, l_env };

    surface_list_t surface_list{
        { floor.get(), glossy.get(), nullptr }, { wall.get(), white.get(), nullptr }, { tri0.get(), orange.get(), nullptr },
        { tri1.get(), glossy.get(), nullptr }, { disk0.get(), white.get(), nullptr }, { ball.get(), mirror_mat.get(), nullptr },
        { gball.get(), glass_mat.get(), nullptr }, { ltri.get(), black.get(), l_tri.get() }, { ldisk.get(), black.get(), l_disk.get() },
        { lrect.get(), black.get(), l_rect.get() }, { lball.get(), black.get(), l_ball.get() } };

    return scene_t{ camera, shape_list, material_list, light_list, surface_list, l_env.get() };
}

// ---- integrators (ky.cpp:3591-4639) -----------------------------------------------------------------
enum class lighting_enum_t
{
    emit = 1, direct = 2, indirect = 4, all_lighting = emit | direct | indirect,
    diffuse = 8, specular = 16, all_scattering = diffuse | specular,
    all = all_lighting | all_scattering
};

enum class direct_sample_enum_t
{
    idle,
    sample_single_light = 1, sample_all_light = 2,
    bsdf = 4, light = 8,
    bsdf_mis = 16, light_mis = 32, both_mis = bsdf_mis | light_mis,
    default_stragtgy = sample_all_light | both_mis
};

enum class integrator_enum_t
{
    position, normal, basecolor,
    delta_bsdf, delta_light, direct_lighting_point,
    direct_lighting,
    stochastic_raytracing,
    simple_path_tracing_recursion,
    path_tracing_recursion, path_tracing_recursion_defered, path_tracing_iteration,
};

class integrator_t
{
public:
    virtual ~integrator_t() = default;

    // reference ky.cpp:3689-3729.  Same contract: adds clamp01(mean radiance) of every pixel to the
    // film (so film_grid_t panels and repeated calls accumulate exactly as in the reference).
    void render(scene_t* scene, sampler_t* original_sampler, film_t* film)
    {
        vec2_t resolution = film->get_resolution();
        kyd_render_desc desc = describe();
        desc.width = (int)resolution.x;
        desc.height = (int)resolution.y;
        // do { } while (next_sample()) renders at least one sample (ky.cpp:3712-3723); the weight stays 1/spp
        int spp = original_sampler->ge_samples_per_pixel();
        desc.spp = spp;
        desc.sample_begin = 0;
        desc.sample_end = std::max(1, spp);
        desc.sampler = original_sampler->device_sampler();
        desc.seed = original_sampler->device_seed();
        desc.flags = KYD_FLAG_CLAMP;

        device_t& dev = device_t::instance();
        auto flat = scene->flatten();
        dev.check(kyd_upload_scene(dev.ctx(), &flat->desc));

        std::vector<float> pixels((size_t)desc.width * desc.height * 3);
        dev.check(kyd_render(dev.ctx(), &desc, pixels.data()));

        for (int y = 0; y < desc.height; ++y)
            for (int x = 0; x < desc.width; ++x)
            {
                const float* p = &pixels[3 * ((size_t)y * desc.width + x)];
                film->add_color(x, y, color_t(p[0], p[1], p[2]));
            }
    }

protected:
    virtual kyd_render_desc describe() const = 0;

    static kyd_render_desc base_desc(integrator_enum_t e, int depth, direct_sample_enum_t ds, lighting_enum_t le = lighting_enum_t::all)
    {
        kyd_render_desc d{};
        d.integrator = (int32_t)e;
        d.max_depth = depth;
        d.direct_sample = (int32_t)ds;
        d.lighting = (int32_t)le;
        return d;
    }
};

class debug_integrator_t : public integrator_t
{
public:
    debug_integrator_t(integrator_enum_t integrator_enum) : integrator_enum_{ integrator_enum } {}
protected:
    kyd_render_desc describe() const override { return base_desc(integrator_enum_, 0, direct_sample_enum_t::idle); }
private:
    integrator_enum_t integrator_enum_;
};

class direct_lighting_t : public integrator_t
{
public:
    direct_lighting_t(direct_sample_enum_t direct_sample_enum) : direct_sample_enum_{ direct_sample_enum } {}
protected:
    kyd_render_desc describe() const override { return base_desc(integrator_enum_t::direct_lighting, 0, direct_sample_enum_); }
private:
    direct_sample_enum_t direct_sample_enum_;
};

class path_integrator_t : public integrator_t
{
public:
    path_integrator_t(int max_path_depth, direct_sample_enum_t direct_sample_enum) :
        max_path_depth_{ max_path_depth }, direct_sample_enum_{ direct_sample_enum }
    {
    }
protected:
    int max_path_depth_;
    direct_sample_enum_t direct_sample_enum_;
};

class simple_path_tracing_recursion_t : public path_integrator_t
{
public:
    using path_integrator_t::path_integrator_t;
protected:
    kyd_render_desc describe() const override { return base_desc(integrator_enum_t::simple_path_tracing_recursion, max_path_depth_, direct_sample_enum_); }
};

class path_tracing_recursion_t : public path_integrator_t
{
public:
    using path_integrator_t::path_integrator_t;
protected:
    kyd_render_desc describe() const override { return base_desc(integrator_enum_t::path_tracing_recursion, max_path_depth_, direct_sample_enum_); }
};

class path_tracing_recursion_defered_t : public path_integrator_t
{
public:
    path_tracing_recursion_defered_t(int max_path_depth, direct_sample_enum_t direct_sample_enum, lighting_enum_t lighting_enum) :
        path_integrator_t(max_path_depth, direct_sample_enum), lighting_enum_{ lighting_enum }
    {
    }
protected:
    kyd_render_desc describe() const override
    {
        return base_desc(integrator_enum_t::path_tracing_recursion_defered, max_path_depth_, direct_sample_enum_, lighting_enum_);
    }
private:
    lighting_enum_t lighting_enum_;
};

class path_tracing_iteration_t : public path_integrator_t
{
public:
    using path_integrator_t::path_integrator_t;
protected:
    kyd_render_desc describe() const override { return base_desc(integrator_enum_t::path_tracing_iteration, max_path_depth_, direct_sample_enum_); }
};

inline std::unique_ptr<integrator_t> create_integrator(integrator_enum_t integrator_enum, int depth, direct_sample_enum_t direct_sample_enum)
{
    switch (integrator_enum)
    {
    case integrator_enum_t::direct_lighting:
        return std::make_unique<direct_lighting_t>(direct_sample_enum);
    case integrator_enum_t::simple_path_tracing_recursion:
        return std::make_unique<simple_path_tracing_recursion_t>(depth, direct_sample_enum);
    case integrator_enum_t::path_tracing_recursion:
        return std::make_unique<path_tracing_recursion_t>(depth, direct_sample_enum);
    case integrator_enum_t::path_tracing_recursion_defered:
        return std::make_unique<path_tracing_recursion_defered_t>(depth, direct_sample_enum, lighting_enum_t::all);
    case integrator_enum_t::path_tracing_iteration:
        return std::make_unique<path_tracing_iteration_t>(depth, direct_sample_enum);
    default:
        return nullptr;
    }
}

} // namespace ky
