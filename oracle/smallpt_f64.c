/*
 * oracle/smallpt_f64.c -- TEST INFRASTRUCTURE.  C restatement of the reference's double-precision smallpt
 * (/root/reference/smallpt2pbrt/smallpt_kernel.cpp, CPU_RENDER configuration; every function cites its lines):
 * FP64, recursive, no next-event estimation, nine spheres, one 32-bit LCG per sample.  The checker of the FP64
 * validation mode kyd_render_smallpt_f64.  Pinned: tests/test_smallpt_f64.py compares it
 * bit for bit with oracle/_ref/libsmallpt_kernel_ref.so (the reference file itself, compiled by oracle/ref/build_ref.sh)
 * and with the fixture tests/golden/golden_smallpt_f64.npz generated from that build.
 *
 * Evaluation orders that C++ leaves unspecified and g++ 13 -O2 fixed in the reference binary (found by comparing with it):
 *   - camera sample (smallpt_kernel.cpp:423-425): the rng() of the cy term is drawn BEFORE the one of the cx term;
 *   - glass, both branches (smallpt_kernel.cpp:289): the refracted Radiance() is evaluated BEFORE the reflected one;
 *   - camera ray (smallpt_kernel.cpp:427): Ray(camera.origin + direction * 140, direction.Normalize()) -- Normalize() works
 *     in place and the second argument is evaluated first, so the origin is pushed along the NORMALISED direction.
 *
 * Build: part of oracle/_build/libkyo.so (gcc -O2 -ffp-contract=off -fno-fast-math).
 */
#include <math.h>
#include <stdint.h>

typedef struct { double x, y, z; } d3;

static d3 D3(double x, double y, double z) { d3 v = { x, y, z }; return v; }
static d3 dadd(d3 a, d3 b) { return D3(a.x + b.x, a.y + b.y, a.z + b.z); }
static d3 dsub(d3 a, d3 b) { return D3(a.x - b.x, a.y - b.y, a.z - b.z); }
static d3 dmul(d3 a, double b) { return D3(a.x * b, a.y * b, a.z * b); }
static d3 ddiv(d3 a, double b) { return D3(a.x / b, a.y / b, a.z / b); }
static d3 dcmul(d3 a, d3 b) { return D3(a.x * b.x, a.y * b.y, a.z * b.z); }
static double ddot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static d3 dcross(d3 a, d3 b) { return D3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* smallpt_kernel.cpp:73: *this * (1 / sqrt(...)) */
static d3 dnormalize(d3 a) { return dmul(a, 1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); }

static const double k_pi = 3.14159265358979323846; /* std::numbers::pi, smallpt_kernel.cpp:44 */

/* smallpt_kernel.cpp:47-53 */
typedef struct { uint32_t seed; } lcg32;
static double lcg_next(lcg32* r) { r->seed = 214013u * r->seed + 2531011u; return r->seed * (1.0 / 4294967296); }

enum { SP_DIFFUSE = 0, SP_SPECULAR = 1, SP_REFRACT = 2 };
typedef struct { double radius; d3 center, emission, color; int material; } sphere_t;

/* smallpt_kernel.cpp:144-157 (scene data) */
static const sphere_t k_scene[9] = {
    { 1e5, { 1e5 + 1, 40.8, 81.6 }, { 0, 0, 0 }, { .75, .25, .25 }, SP_DIFFUSE },
    { 1e5, { -1e5 + 99, 40.8, 81.6 }, { 0, 0, 0 }, { .25, .25, .75 }, SP_DIFFUSE },
    { 1e5, { 50, 40.8, 1e5 }, { 0, 0, 0 }, { .75, .75, .75 }, SP_DIFFUSE },
    { 1e5, { 50, 40.8, -1e5 + 170 }, { 0, 0, 0 }, { 0, 0, 0 }, SP_DIFFUSE },
    { 1e5, { 50, 1e5, 81.6 }, { 0, 0, 0 }, { .75, .75, .75 }, SP_DIFFUSE },
    { 1e5, { 50, -1e5 + 81.6, 81.6 }, { 0, 0, 0 }, { .75, .75, .75 }, SP_DIFFUSE },
    { 16.5, { 27, 16.5, 47 }, { 0, 0, 0 }, { 1, 1, 1 }, SP_SPECULAR },
    { 16.5, { 73, 16.5, 78 }, { 0, 0, 0 }, { 1, 1, 1 }, SP_REFRACT },
    { 600, { 50, 681.6 - .27, 81.6 }, { 12, 12, 12 }, { 0, 0, 0 }, SP_DIFFUSE },
};

/* smallpt_kernel.cpp:113-139 */
static double sphere_intersect(const sphere_t* s, d3 o, d3 d)
{
    d3 oc = dsub(s->center, o);
    double neg_b = ddot(oc, d);
    double det = neg_b * neg_b - ddot(oc, oc) + s->radius * s->radius;
    if (det < 0)
        return 0;
    det = sqrt(det);
    double epsilon = 1e-4, t;
    if ((t = neg_b - det) > epsilon)
        return t;
    if ((t = neg_b + det) > epsilon)
        return t;
    return 0;
}

/* smallpt_kernel.cpp:163-182: spheres visited from the last to the first, strict < */
static int scene_intersect(d3 o, d3 d, double* min_distance, int* id)
{
    double infinity = 1e20, distance;
    *min_distance = infinity;
    for (int i = 9; i--;)
        if ((distance = sphere_intersect(&k_scene[i], o, d)) != 0 && distance < *min_distance)
        {
            *min_distance = distance;
            *id = i;
        }
    return *min_distance < infinity;
}

/* smallpt_kernel.cpp:184-296 */
static d3 radiance(d3 o, d3 dir, int depth, lcg32* rng)
{
    double distance;
    int id = 0;
    if (!scene_intersect(o, dir, &distance, &id))
        return D3(0, 0, 0);
    const sphere_t* obj = &k_scene[id];
    if (depth > 5)
        return obj->emission;

    d3 position = dadd(o, dmul(dir, distance));
    d3 normal = dnormalize(dsub(position, obj->center));
    d3 shading_normal = ddot(normal, dir) < 0 ? normal : dmul(normal, -1);

    d3 f = obj->color;
    double max_component = (f.x > f.y && f.x > f.z) ? f.x : (f.y > f.z ? f.y : f.z);
    if (++depth > 3)
    {
        if (lcg_next(rng) < max_component)
            f = dmul(f, 1 / max_component);
        else
            return obj->emission;
    }

    if (obj->material == SP_DIFFUSE)
    {
        double random1 = 2 * k_pi * lcg_next(rng);
        double random2 = lcg_next(rng);
        double random2_sqrt = sqrt(random2);
        d3 w = shading_normal;
        d3 u = dnormalize(dcross(fabs(w.x) > .1 ? D3(0, 1, 0) : D3(1, 0, 0), w));
        d3 v = dcross(w, u);
        d3 direction = dnormalize(dadd(dadd(dmul(dmul(u, cos(random1)), random2_sqrt), dmul(dmul(v, sin(random1)), random2_sqrt)),
                                       dmul(w, sqrt(1 - random2))));
        f = ddiv(f, k_pi);
        double abs_cos_theta = fabs(ddot(shading_normal, direction));
        double pdf = abs_cos_theta / k_pi;
        d3 li = radiance(position, direction, depth, rng);
        return dadd(obj->emission, ddiv(dmul(dcmul(f, li), abs_cos_theta), pdf));
    }
    else if (obj->material == SP_SPECULAR)
    {
        d3 direction = dsub(dir, dmul(dmul(normal, 2), ddot(normal, dir)));
        return dadd(obj->emission, dcmul(f, radiance(position, direction, depth, rng)));
    }
    else
    {
        int into = ddot(normal, shading_normal) > 0;
        double eta_i = 1, eta_t = 1.5;
        double eta = into ? eta_i / eta_t : eta_t / eta_i;
        d3 reflect_dir = dsub(dir, dmul(dmul(normal, 2), ddot(normal, dir)));
        double cos_theta_i = ddot(dir, shading_normal);
        double cos_theta_t2 = 1 - eta * eta * (1 - cos_theta_i * cos_theta_i);
        if (cos_theta_t2 < 0)
            return dadd(obj->emission, dcmul(f, radiance(position, reflect_dir, depth, rng)));
        double cos_theta_t = sqrt(cos_theta_t2);
        d3 refract_dir = dnormalize(dsub(dmul(dir, eta), dmul(normal, (into ? 1 : -1) * (cos_theta_i * eta + cos_theta_t))));

        double a = eta_t - eta_i, b = eta_t + eta_i;
        double r0 = a * a / (b * b);
        double c = 1 - (into ? -cos_theta_i : ddot(refract_dir, normal));
        double re = r0 + (1 - r0) * c * c * c * c * c;
        double tr = 1 - re;
        double p = .25 + .5 * re;
        double rp = re / p, tp = tr / (1 - p);

        d3 li;
        if (depth > 2)
        {
            if (lcg_next(rng) < p)
                li = dmul(radiance(position, reflect_dir, depth, rng), rp);
            else
                li = dmul(radiance(position, refract_dir, depth, rng), tp);
        }
        else
        {
#ifdef SMALLPT_REFLECT_FIRST
            d3 lr = radiance(position, reflect_dir, depth, rng);
            d3 lt = radiance(position, refract_dir, depth, rng);
#else
            d3 lt = radiance(position, refract_dir, depth, rng);
            d3 lr = radiance(position, reflect_dir, depth, rng);
#endif
            li = dadd(dmul(lr, re), dmul(lt, tr));
        }
        return dadd(obj->emission, dcmul(f, li));
    }
}

static double clamp01(double x) { return x < 0 ? 0 : x > 1 ? 1 : x; }

/* smallpt_kernel.cpp:403-438 (CPU_RENDER Device::Render): film rows bottom-up, width * height * 3 doubles */
int kyo_smallpt_f64(int width, int height, int samples_per_pixel, double* film)
{
    d3 cam_o = D3(50, 52, 295.6), cam_d = dnormalize(D3(0, -0.042612, -1));
    d3 cx = D3(width * .5135 / height, 0, 0);
    d3 cy = dmul(dnormalize(dcross(cx, cam_d)), .5135);
    for (long i = 0; i < (long)width * height * 3; ++i)
        film[i] = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
        {
            d3 li = D3(0, 0, 0);
            for (int s = 0; s < samples_per_pixel; s++)
            {
                lcg32 rng = { (uint32_t)(y * width + x * samples_per_pixel + s) };
#ifdef SMALLPT_CX_FIRST
                double rx = lcg_next(&rng), ry = lcg_next(&rng);
#else
                double ry = lcg_next(&rng), rx = lcg_next(&rng);
#endif
                d3 direction = dadd(dadd(dmul(cx, (rx + x) / width - .5), dmul(cy, (ry + y) / height - .5)), cam_d);
#ifdef SMALLPT_PUSH_UNNORMALISED
                d3 l = radiance(dadd(cam_o, dmul(direction, 140)), dnormalize(direction), 0, &rng);
#else
                /* Ray(camera.origin + direction * 140, direction.Normalize()): Normalize() works in place and g++ evaluates
                   the second argument first, so the push uses the NORMALISED direction */
                direction = dnormalize(direction);
                d3 l = radiance(dadd(cam_o, dmul(direction, 140)), direction, 0, &rng);
#endif
                li = dadd(li, dmul(l, 1. / samples_per_pixel));
            }
            double* o = film + 3 * ((long)(height - y - 1) * width + x);
            o[0] += clamp01(li.x); o[1] += clamp01(li.y); o[2] += clamp01(li.z);
        }
    return 0;
}
