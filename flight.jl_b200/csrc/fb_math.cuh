// fb_math.cuh — FP64 vector/quaternion helpers for the batched C172 kernels (sm_100a).
// Reference semantics: lib/FlightPhysics/src/quaternions.jl:109-115, attitude.jl:98-118.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define FB_DEV __device__ __forceinline__
#define FB_PI 3.14159265358979323846

// min / max / clamp as compare + select (1 DSETP + 2 FSEL each). fmin()/fmax() carry IEEE NaN handling that costs
// ~9 instructions per call on sm_100a, and this path is NaN-free by construction.
__host__ __device__ inline double fb_max(double a, double b) { return (a < b) ? b : a; }
__host__ __device__ inline double fb_min(double a, double b) { return (a > b) ? b : a; }
// Both compares read the ORIGINAL x, so they issue back to back instead of the second waiting for the first select (lo <= hi:
// at most one of them holds); same result as max(lo, min(x, hi)) for every non-NaN x.
__host__ __device__ inline double clampd(double x, double lo, double hi) { const bool below = x < lo, above = x > hi; const double r = below ? lo : x; return above ? hi : r; }

struct d3 { double x, y, z; };
struct q4 { double w, x, y, z; };

// ---- fast FP64 primitives (normal-range operands only; no denormal / inf / nan slow paths) -------------
// MUFU seed (~20 bits) + two Newton iterations: relative error <= ~1 ulp, a third of the instructions
// (and a tenth of the code bytes) of the IEEE-exact sequences nvcc emits for `/`, sqrt() and rsqrt().
FB_DEV double fb_rcp(double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    return fma(r, e, r);
}
FB_DEV double fb_div(double a, double b) {
    double r = fb_rcp(b);
    double q = a * r;
    return fma(fma(-b, q, a), r, q);
}
FB_DEV double fb_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}
FB_DEV double fb_sqrt(double a) {   // a > 0
    double y = fb_rsqrt(a);
    double s = a * y;
    return fma(fma(-s, s, a), 0.5 * y, s);
}

// atan2 for finite, not-both-zero arguments. One division: with t = min/max in [0,1] and c the nearest
// multiple of 1/8, atan(t) = atan(c) + atan(z), z = (min - c max) / (max + c min), |z| <= 1/16, and the odd
// series to z^13 leaves a truncation error below 1e-18. Absolute error ~1 ulp of the result's quadrant offset.
static __constant__ double c_atan_tab[9] = {0.0, 0.12435499454676144, 0.24497866312686414, 0.35877067027057225, 0.46364760900080609,
                                     0.55859931534356244, 0.64350110879328437, 0.71882999962162453, 0.78539816339744831};
FB_DEV double fb_atan2(double y, double x) {   // branch-free; atan2(+-0, +-0) gives +-0 like an isolated zero vector should
    double ax = fabs(x), ay = fabs(y);
    double mx = fb_max(ax, ay), mn = fb_min(ax, ay);
    mx = (mx > 0.0) ? mx : 1.0;
    float r = __fdividef((float)mn, (float)mx);
    int idx = min(max(__float2int_rn(r * 8.0f), 0), 8);
    double c = (double)idx * 0.125;
    double z = fb_div(fma(-c, mx, mn), fma(c, mn, mx));
    double z2 = z * z;
    double p = fma(z2, 1.0 / 13.0, -1.0 / 11.0);
    p = fma(p, z2, 1.0 / 9.0);
    p = fma(p, z2, -1.0 / 7.0);
    p = fma(p, z2, 1.0 / 5.0);
    p = fma(p, z2, -1.0 / 3.0);
    double a = fma(p * z2, z, z) + c_atan_tab[idx];
    a = (ay > ax) ? (0.5 * FB_PI - a) : a;
    a = (x < 0.0) ? (FB_PI - a) : a;
    return copysign(a, y);
}

FB_DEV d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
FB_DEV d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
FB_DEV d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
FB_DEV d3 operator-(d3 a) { return mk3(-a.x, -a.y, -a.z); }
FB_DEV d3 operator*(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
FB_DEV double dot(d3 a, d3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
FB_DEV d3 cross(d3 a, d3 b) {
    return mk3(fma(a.y, b.z, -a.z * b.y), fma(a.z, b.x, -a.x * b.z), fma(a.x, b.y, -a.y * b.x));
}
FB_DEV q4 mkq(double w, double x, double y, double z) { q4 r; r.w = w; r.x = x; r.y = y; r.z = z; return r; }
FB_DEV q4 qconj(q4 a) { return mkq(a.w, -a.x, -a.y, -a.z); }
// Hamilton product
FB_DEV q4 qmul(q4 a, q4 b) {
    return mkq(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
               a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
               a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
               a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
// v_a = q_ab * v_b  (v + 2 q_im x (q_re v + q_im x v))
FB_DEV d3 qrot(q4 q, d3 v) {
    d3 qi = mk3(q.x, q.y, q.z);
    d3 t = cross(qi, v);
    t = mk3(fma(q.w, v.x, t.x), fma(q.w, v.y, t.y), fma(q.w, v.z, t.z));
    d3 c = cross(qi, t);
    return mk3(fma(2.0, c.x, v.x), fma(2.0, c.y, v.y), fma(2.0, c.z, v.z));
}
FB_DEV d3 qrot_inv(q4 q, d3 v) { return qrot(qconj(q), v); }
// 0.5 * q * (0, w)
FB_DEV q4 qdt(q4 q, d3 w) {
    return mkq(-0.5 * (q.x * w.x + q.y * w.y + q.z * w.z),
               0.5 * (q.w * w.x + q.y * w.z - q.z * w.y),
               0.5 * (q.w * w.y - q.x * w.z + q.z * w.x),
               0.5 * (q.w * w.z + q.x * w.y - q.y * w.x));
}
// cos(a/2), sin(a/2) of a = atan2(y, x) without evaluating the angle (|(x,y)| > 0)
FB_DEV void half_angle_cs(double c, double s, double& ch, double& sh) {   // (c, s) = (cos a, sin a)
    double t = 0.5 * (1.0 + fabs(c));
    double it = fb_rsqrt(t);
    double big = t * it;                 // sqrt((1+|c|)/2)
    double small_ = 0.5 * s * it;        // s / (2 sqrt((1+|c|)/2))
    const bool pos = (c >= 0.0);
    double sbig = (s >= 0.0) ? big : -big;
    ch = pos ? big : fabs(small_);
    sh = pos ? small_ : sbig;
}
FB_DEV void half_angle(double y, double x, double& ch, double& sh) {
    double ir = fb_rsqrt(fma(x, x, y * y));
    half_angle_cs(x * ir, y * ir, ch, sh);
}
FB_DEV int isign(double a) { return (a > 0.0) - (a < 0.0); }
