// pb_math.cuh -- f32 device math for the rs_pbrt PathIntegrator hot path on sm_100a.
//
// Arithmetic contract (DESIGN.md "Numerics"): every expression is evaluated in the same
// operation order as the reference's Rust (file:line cited per function), with IEEE
// round-to-nearest +,-,*,/,sqrt and NO fused multiply-add (the translation unit is compiled with
// -fmad=false; rustc/LLVM never contracts).  Transcendentals (sin, cos) go through f64 and are
// rounded once to f32, which reproduces glibc's (almost always correctly rounded) sinf/cosf that
// the reference calls through Rust's std.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PB_HD __host__ __device__ __forceinline__
#define PB_D __device__ __forceinline__

namespace pb {

// src/core/pbrt.rs:16-23, src/core/rng.rs:13
#define PB_MACHINE_EPSILON 5.9604644775390625e-8f
#define PB_SHADOW_EPSILON 0.0001f
#define PB_PI 3.14159265358979323846f
#define PB_INV_PI 0.31830988618379067154f
#define PB_INV_2_PI 0.15915494309189533577f
#define PB_PI_OVER_2 1.57079632679489661923f
#define PB_PI_OVER_4 0.78539816339744830961f
#define PB_TAU 6.28318530717958647692f
#define PB_ONE_MINUS_EPSILON 0.99999994f

// gamma(n) = n*eps / (1 - n*eps)   src/core/pbrt.rs:94-96   (folded at compile time: IEEE either way)
PB_HD float gamma_n(int n) { return ((float)n * PB_MACHINE_EPSILON) / (1.0f - (float)n * PB_MACHINE_EPSILON); }

struct V3 {
    float x, y, z;
};
PB_HD V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
PB_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PB_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PB_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
PB_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
// vector / f32 multiplies by the reciprocal (geometry.rs:1262-1291)
PB_HD V3 vdiv(V3 a, float s) { float inv = 1.0f / s; return mk3(a.x * inv, a.y * inv, a.z * inv); }
PB_HD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // geometry.rs:630
PB_HD float absdot3(V3 a, V3 b) { return fabsf(dot3(a, b)); }
PB_HD float len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
PB_HD float len3(V3 a) { return sqrtf(len2(a)); }
PB_HD V3 norm3(V3 a) { return vdiv(a, len3(a)); }  // geometry.rs:412
PB_HD V3 abs3(V3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
PB_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// cross product in f64, rounded once (geometry.rs:680-692)
PB_HD V3 cross3(V3 a, V3 b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return mk3((float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)), (float)((ax * by) - (ay * bx)));
}
PB_HD float maxcomp(V3 v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }  // geometry.rs:711
PB_HD int maxdim(V3 v) {                                           // geometry.rs:721-734
    if (v.x > v.y) return (v.x > v.z) ? 0 : 2;
    return (v.y > v.z) ? 1 : 2;
}
PB_HD V3 faceforward3(V3 n, V3 v) { return (dot3(n, v) < 0.0f) ? -n : n; }  // geometry.rs:1842-1858
PB_HD void coordinate_system(V3 v1, V3& v2, V3& v3) {                       // geometry.rs:779-794
    if (fabsf(v1.x) > fabsf(v1.y)) v2 = vdiv(mk3(-v1.z, 0.0f, v1.x), sqrtf(v1.x * v1.x + v1.z * v1.z));
    else v2 = vdiv(mk3(0.0f, v1.z, -v1.y), sqrtf(v1.y * v1.y + v1.z * v1.z));
    v3 = cross3(v1, v2);
}

PB_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
PB_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
// src/core/pbrt.rs:61-91
PB_HD float next_float_up(float v) {
    if (isinf(v) && v > 0.0f) return v;
    if (v == -0.0f) v = 0.0f;
    uint32_t ui = f2u(v);
    if (v >= 0.0f) ui += 1; else ui -= 1;
    return u2f(ui);
}
PB_HD float next_float_down(float v) {
    if (isinf(v) && v < 0.0f) return v;
    if (v == 0.0f) v = -0.0f;
    uint32_t ui = f2u(v);
    if (v > 0.0f) ui -= 1; else ui += 1;
    return u2f(ui);
}
// pnt3_offset_ray_origin  src/core/geometry.rs:1535-1556
PB_HD V3 offset_ray_origin(V3 p, V3 p_error, V3 n, V3 w) {
    float d = dot3(abs3(n), p_error);
    V3 off = n * d;
    if (dot3(w, n) < 0.0f) off = -off;
    V3 po = p + off;
    if (off.x > 0.0f) po.x = next_float_up(po.x); else if (off.x < 0.0f) po.x = next_float_down(po.x);
    if (off.y > 0.0f) po.y = next_float_up(po.y); else if (off.y < 0.0f) po.y = next_float_down(po.y);
    if (off.z > 0.0f) po.z = next_float_up(po.z); else if (off.z < 0.0f) po.z = next_float_down(po.z);
    return po;
}
// Rust `x as i32`: saturating, NaN -> 0.  cvt.rzi.s32.f32 has exactly these semantics on the
// device; the host branch spells them out.
PB_HD int f2i_sat(float x) {
#ifdef __CUDA_ARCH__
    return __float2int_rz(x);
#else
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
#endif
}
PB_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // pbrt.rs:108-121
PB_HD float lerpf(float t, float a, float b) { return a * (1.0f - t) + b * t; }              // pbrt.rs:235-245

// f32 sin/cos as glibc computes them (f64 evaluation, one rounding)
PB_D float sin_rn(float x) { return (float)sin((double)x); }
PB_D float cos_rn(float x) { return (float)cos((double)x); }
PB_D void sincos_rn(float x, float& s, float& c) {  // one f64 range reduction for both
    double ds, dc;
    sincos((double)x, &ds, &dc);
    s = (float)ds;
    c = (float)dc;
}

// RGBSpectrum (src/core/spectrum.rs:1530-1780)
struct Sp {
    float r, g, b;
};
PB_HD Sp mksp(float r, float g, float b) { Sp s; s.r = r; s.g = g; s.b = b; return s; }
PB_HD Sp sp1(float v) { return mksp(v, v, v); }
PB_HD Sp operator+(Sp a, Sp b) { return mksp(a.r + b.r, a.g + b.g, a.b + b.b); }
PB_HD Sp operator-(Sp a, Sp b) { return mksp(a.r - b.r, a.g - b.g, a.b - b.b); }
PB_HD Sp operator*(Sp a, Sp b) { return mksp(a.r * b.r, a.g * b.g, a.b * b.b); }
PB_HD Sp operator/(Sp a, Sp b) { return mksp(a.r / b.r, a.g / b.g, a.b / b.b); }
PB_HD Sp operator*(Sp a, float s) { return mksp(a.r * s, a.g * s, a.b * s); }
PB_HD Sp operator/(Sp a, float s) { return mksp(a.r / s, a.g / s, a.b / s); }  // true division (spectrum.rs:1752)
PB_HD bool is_black(Sp a) { return a.r == 0.0f && a.g == 0.0f && a.b == 0.0f; }
PB_HD bool has_nans(Sp a) { return a.r != a.r || a.g != a.g || a.b != a.b; }
PB_HD float lum(Sp a) { return 0.212671f * a.r + 0.715160f * a.g + 0.072169f * a.b; }  // spectrum.rs:1581
PB_HD float maxsp(Sp a) { return fmaxf(fmaxf(a.r, a.g), a.b); }
PB_HD Sp sqrtsp(Sp a) { return mksp(sqrtf(a.r), sqrtf(a.g), sqrtf(a.b)); }

}  // namespace pb
