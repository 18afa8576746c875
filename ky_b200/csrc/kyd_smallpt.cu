// FP64 validation mode: the reference's double-precision smallpt
// (/root/reference/smallpt2pbrt/smallpt_kernel.cpp, CPU_RENDER Device::Render :403-438 and Radiance :184-296) on the
// device.  Recursive in the reference; here a depth-first walk with an explicit stack that visits the calls in the order
// the reference binary does (oracle/smallpt_f64.c lists the three evaluation orders g++ fixed): one 32-bit LCG per
// sample is shared by the whole call tree, so the order is part of the result.  One thread per pixel, samples in order.
// Double arithmetic is IEEE and uncontracted (-fmad=false); sin / cos are CUDA's double functions (<= 2 ulp) where the
// reference calls glibc's (< 1 ulp): results agree to ~1e-12 except where such a last-bit difference flips a branch.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kyd_internal.h"

namespace {

struct d3 { double x, y, z; };

__device__ __forceinline__ d3 D3(double x, double y, double z) { d3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return D3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return D3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator*(d3 a, double b) { return D3(a.x * b, a.y * b, a.z * b); }
__device__ __forceinline__ d3 operator/(d3 a, double b) { return D3(a.x / b, a.y / b, a.z / b); }
__device__ __forceinline__ d3 cmul(d3 a, d3 b) { return D3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ d3 cross(d3 a, d3 b) { return D3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ d3 normalize(d3 a) { return a * (1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); } // :73

#define SP_PI 3.14159265358979323846 /* std::numbers::pi, :44 */

enum { SP_DIFFUSE = 0, SP_SPECULAR = 1, SP_REFRACT = 2 };
struct Sphere { double radius; double center[3]; double emission[3]; double color[3]; int material; };

// scene data, smallpt_kernel.cpp:144-157
__constant__ Sphere c_spheres[9] = {
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

__device__ __forceinline__ d3 ld3(const double* p) { return D3(p[0], p[1], p[2]); }

__device__ __forceinline__ double lcg_next(unsigned& seed) // :47-53
{
    seed = 214013u * seed + 2531011u;
    return seed * (1.0 / 4294967296);
}

__device__ __forceinline__ double sphere_intersect(const Sphere& s, d3 o, d3 d) // :113-139
{
    d3 oc = ld3(s.center) - o;
    double neg_b = dot(oc, d);
    double det = neg_b * neg_b - dot(oc, oc) + s.radius * s.radius;
    if (det < 0)
        return 0;
    det = sqrt(det);
    const double epsilon = 1e-4;
    double t = neg_b - det;
    if (t > epsilon)
        return t;
    t = neg_b + det;
    return t > epsilon ? t : 0;
}

__device__ __forceinline__ bool scene_intersect(d3 o, d3 d, double& min_distance, int& id) // :163-182, last sphere first
{
    const double infinity = 1e20;
    min_distance = infinity;
    for (int i = 9; i--;)
    {
        const double distance = sphere_intersect(c_spheres[i], o, d);
        if (distance != 0 && distance < min_distance)
        {
            min_distance = distance;
            id = i;
        }
    }
    return min_distance < infinity;
}

// one pending call of Radiance(): what to do with the value its child returns
enum { F_DIFFUSE, F_SPECULAR, F_GLASS_ONE, F_GLASS_REFRACTED, F_GLASS_REFLECTED };
struct Frame
{
    int kind, depth;      // depth = the value the children were called with
    d3 emission, f;
    double a, b;          // diffuse: |cos|, pdf; glass one branch: scale; glass both: Re, Tr
    d3 position, reflect_dir, refracted;
};

#define SP_MAX_FRAMES 8

// Radiance(ray, 0, rng), smallpt_kernel.cpp:184-296
__device__ d3 radiance(d3 o, d3 dir, unsigned& rng)
{
    Frame stack[SP_MAX_FRAMES];
    int top = 0;
    int depth = 0;
    d3 value = D3(0, 0, 0);
    for (;;)
    {
        // ---- descend: one call of Radiance(o, dir, depth) up to its first recursive call ----
        bool leaf = true;
        double distance;
        int id = 0;
        if (!scene_intersect(o, dir, distance, id))
            value = D3(0, 0, 0);
        else
        {
            const Sphere& obj = c_spheres[id];
            const d3 emission = ld3(obj.emission);
            value = emission;                       // what every early return below returns
            if (!(depth > 5))
            {
                const d3 position = o + dir * distance;
                const d3 normal = normalize(position - ld3(obj.center));
                const d3 shading_normal = dot(normal, dir) < 0 ? normal : normal * -1;
                d3 f = ld3(obj.color);
                const double max_component = (f.x > f.y && f.x > f.z) ? f.x : (f.y > f.z ? f.y : f.z);
                bool alive = true;
                if (++depth > 3)
                {
                    if (lcg_next(rng) < max_component)
                        f = f * (1 / max_component);
                    else
                        alive = false;
                }
                if (alive)
                {
                    leaf = false;
                    Frame& fr = stack[top++];
                    fr.depth = depth;
                    fr.emission = emission;
                    if (obj.material == SP_DIFFUSE)
                    {
                        const double random1 = 2 * SP_PI * lcg_next(rng);
                        const double random2 = lcg_next(rng);
                        const double random2_sqrt = sqrt(random2);
                        const d3 w = shading_normal;
                        const d3 u = normalize(cross(fabs(w.x) > .1 ? D3(0, 1, 0) : D3(1, 0, 0), w));
                        const d3 v = cross(w, u);
                        const d3 direction = normalize(u * cos(random1) * random2_sqrt + v * sin(random1) * random2_sqrt + w * sqrt(1 - random2));
                        fr.kind = F_DIFFUSE;
                        fr.f = f / SP_PI;
                        fr.a = fabs(dot(shading_normal, direction));
                        fr.b = fr.a / SP_PI;
                        o = position;
                        dir = direction;
                    }
                    else if (obj.material == SP_SPECULAR)
                    {
                        fr.kind = F_SPECULAR;
                        fr.f = f;
                        o = position;
                        dir = dir - normal * 2 * dot(normal, dir);
                    }
                    else
                    {
                        const bool into = dot(normal, shading_normal) > 0;
                        const double eta_i = 1, eta_t = 1.5;
                        const double eta = into ? eta_i / eta_t : eta_t / eta_i;
                        const d3 reflect_dir = dir - normal * 2 * dot(normal, dir);
                        const double cos_theta_i = dot(dir, shading_normal);
                        const double cos_theta_t2 = 1 - eta * eta * (1 - cos_theta_i * cos_theta_i);
                        fr.f = f;
                        o = position;
                        if (cos_theta_t2 < 0)
                        {
                            fr.kind = F_SPECULAR;   // total internal reflection: emission + f * Radiance(reflected)
                            dir = reflect_dir;
                        }
                        else
                        {
                            const double cos_theta_t = sqrt(cos_theta_t2);
                            const d3 refract_dir = normalize(dir * eta - normal * ((into ? 1 : -1) * (cos_theta_i * eta + cos_theta_t)));
                            const double a = eta_t - eta_i, b = eta_t + eta_i;
                            const double r0 = a * a / (b * b);
                            const double c = 1 - (into ? -cos_theta_i : dot(refract_dir, normal));
                            const double re = r0 + (1 - r0) * c * c * c * c * c;
                            const double tr = 1 - re;
                            const double p = .25 + .5 * re;
                            const double rp = re / p, tp = tr / (1 - p);
                            if (depth > 2)
                            {
                                fr.kind = F_GLASS_ONE;
                                if (lcg_next(rng) < p) { fr.a = rp; dir = reflect_dir; }
                                else { fr.a = tp; dir = refract_dir; }
                            }
                            else
                            {
                                // both branches: the reference binary evaluates the refracted one first
                                fr.kind = F_GLASS_REFRACTED;
                                fr.a = re;
                                fr.b = tr;
                                fr.position = position;
                                fr.reflect_dir = reflect_dir;
                                dir = refract_dir;
                            }
                        }
                    }
                }
            }
        }
        if (!leaf)
            continue;
        // ---- return: hand `value` to the pending calls ----
        bool descend = false;
        while (top > 0 && !descend)
        {
            Frame& fr = stack[top - 1];
            switch (fr.kind)
            {
            case F_DIFFUSE: value = fr.emission + (cmul(fr.f, value) * fr.a) / fr.b; --top; break;
            case F_SPECULAR: value = fr.emission + cmul(fr.f, value); --top; break;
            case F_GLASS_ONE: value = fr.emission + cmul(fr.f, value * fr.a); --top; break;
            case F_GLASS_REFRACTED:
                fr.refracted = value;
                fr.kind = F_GLASS_REFLECTED;
                o = fr.position;
                dir = fr.reflect_dir;
                depth = fr.depth;
                descend = true;
                break;
            default: // F_GLASS_REFLECTED
                value = fr.emission + cmul(fr.f, value * fr.a + fr.refracted * fr.b);
                --top;
                break;
            }
        }
        if (!descend)
            return value;
    }
}

__device__ __forceinline__ double clamp01(double x) { return x < 0 ? 0 : x > 1 ? 1 : x; }

// Device::Render, smallpt_kernel.cpp:403-438
__global__ void __launch_bounds__(64) k_smallpt_f64(int width, int height, int samples_per_pixel, double* __restrict__ film)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= width || y >= height)
        return;
    const d3 cam_o = D3(50, 52, 295.6), cam_d = normalize(D3(0, -0.042612, -1));
    const d3 cx = D3(width * .5135 / height, 0, 0);
    const d3 cy = normalize(cross(cx, cam_d)) * .5135;
    d3 li = D3(0, 0, 0);
    for (int s = 0; s < samples_per_pixel; ++s)
    {
        unsigned rng = (unsigned)(y * width + x * samples_per_pixel + s);
        const double ry = lcg_next(rng), rx = lcg_next(rng);   // the cy term's draw comes first in the reference binary
        d3 direction = cx * ((rx + x) / width - .5) + cy * ((ry + y) / height - .5) + cam_d;
        direction = normalize(direction);                      // ... and Normalize() runs before the origin push (in place)
        const d3 l = radiance(cam_o + direction * 140, direction, rng);
        li = li + l * (1. / samples_per_pixel);
    }
    double* o = film + 3 * ((size_t)(height - y - 1) * width + x);
    o[0] = clamp01(li.x); o[1] = clamp01(li.y); o[2] = clamp01(li.z);
}

} // namespace

namespace kyd {

cudaError_t launch_smallpt_f64(int width, int height, int samples_per_pixel, double* film_dev, cudaStream_t stream)
{
    dim3 block(64, 1), grid((width + 63) / 64, height);
    k_smallpt_f64<<<grid, block, 0, stream>>>(width, height, samples_per_pixel, film_dev);
    return cudaGetLastError();
}

} // namespace kyd
