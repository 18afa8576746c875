/*
 * oracle/ref/crlibm_shim.c -- TEST INFRASTRUCTURE (never linked into the product).
 *
 * Link-time replacement of the float transcendentals the reference calls
 * (ky.cpp:724-732, 753-755, 766-768, 1482-1490, 1875, 2499, 2535-2549, 3032-3047)
 * by "evaluate in double, round once" versions.
 *
 * Why: glibc 2.39's sinf/cosf/sincosf/powf/acosf are < 1 ulp but NOT correctly
 * rounded (measured here against the double functions: 1.3 % of sinf, 1.3 % of
 * cosf, 0.1-0.2 % of powf and 7.7 % of acosf results differ by one ulp), and glibc
 * dispatches FMA / non-FMA variants by CPU (ifunc), so the reference's float
 * results are a property of the host it runs on.  A correctly rounded libm is a
 * conforming libm; with it the deterministic reference build has ONE answer that a
 * device can reproduce: CUDA's double sin/cos/pow/acos (<= 2 ulp in double) rounded
 * once to float give the same float unless the exact value lies within ~2^-51
 * (relative) of a float rounding boundary, i.e. about 1e-8 of the calls.
 *
 * The verbatim reference build (timing baseline, statistical anchor) does NOT link
 * this file and keeps glibc's own float functions.
 */
#include <math.h>

float sinf(float x) { return (float)sin((double)x); }
float cosf(float x) { return (float)cos((double)x); }
float tanf(float x) { return (float)tan((double)x); }
float acosf(float x) { return (float)acos((double)x); }
float powf(float x, float y) { return (float)pow((double)x, (double)y); }
void sincosf(float x, float* s, float* c)
{
    *s = (float)sin((double)x);
    *c = (float)cos((double)x);
}
