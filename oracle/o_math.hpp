// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the rs_pbrt PathIntegrator
// hot path, used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs as the CHECKER.  Nothing under rs_pbrt_b200/ may include,
// link or call this.  PARITY UNPINNED: the reference ships no golden vectors or
// asserting tests for this path and cannot be built here (no
// Rust toolchain); what is pinned independently is listed in oracle/README.md.
//
// o_math.hpp: Float = f32 math of src/core/pbrt.rs, geometry.rs, spectrum.rs.
// Build with -ffp-contract=off: Rust/LLVM never contracts a*b+c into an FMA.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

typedef float Float;

// src/core/pbrt.rs:16-23
static const Float MACHINE_EPSILON = 5.9604644775390625e-8f;  // f32::EPSILON * 0.5 = 2^-24
static const Float SHADOW_EPSILON = 0.0001f;
static const Float PI = 3.14159265358979323846f;       // std::f32::consts::PI
static const Float INV_PI = 0.31830988618379067154f;
static const Float INV_2_PI = 0.15915494309189533577f;  // pbrt.rs:19
static const Float PI_OVER_2 = 1.57079632679489661923f;
static const Float PI_OVER_4 = 0.78539816339744830961f;
static const Float TAU = 6.28318530717958647692f;      // std::f32::consts::TAU
static const Float FLOAT_ONE_MINUS_EPSILON = 0.99999994f;  // 0x1.fffffep-1  src/core/rng.rs:13
static const Float INF = std::numeric_limits<float>::infinity();

inline uint32_t float_to_bits(Float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline Float bits_to_float(uint32_t u) { Float f; std::memcpy(&f, &u, 4); return f; }

// src/core/pbrt.rs:61-91
inline Float next_float_up(Float v) {
    if (std::isinf(v) && v > 0.0f) return v;
    if (v == -0.0f) v = 0.0f;
    uint32_t ui = float_to_bits(v);
    if (v >= 0.0f) ui += 1; else ui -= 1;
    return bits_to_float(ui);
}
inline Float next_float_down(Float v) {
    if (std::isinf(v) && v < 0.0f) return v;
    if (v == 0.0f) v = -0.0f;
    uint32_t ui = float_to_bits(v);
    if (v > 0.0f) ui -= 1; else ui += 1;
    return bits_to_float(ui);
}
// src/core/pbrt.rs:94-96
inline Float gamma(int n) { return ((Float)n * MACHINE_EPSILON) / (1.0f - (Float)n * MACHINE_EPSILON); }

// src/core/pbrt.rs:108-121
template <typename T> inline T clamp_t(T val, T low, T high) {
    if (val < low) return low;
    if (val > high) return high;
    return val;
}
// Rust f32::max / f32::min: a NaN operand is ignored (== fmaxf / fminf)
inline Float fmax_(Float a, Float b) { return std::fmax(a, b); }
inline Float fmin_(Float a, Float b) { return std::fmin(a, b); }
// Rust `x as i32` / `as u8`: saturating, NaN -> 0
inline int32_t f2i(Float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int32_t)x;
}
inline Float lerp(Float t, Float a, Float b) { return a * (1.0f - t) + b * t; }  // pbrt.rs:235-245
inline Float radians(Float deg) { return (PI / 180.0f) * deg; }                 // pbrt.rs:144

struct Vec3 {
    Float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
// Points, vectors and normals share one storage type here; the reference keeps
// three types whose arithmetic is identical component-wise f32.
typedef Vec3 Point3;
typedef Vec3 Normal3;

inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(const Vec3& a, Float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
// geometry.rs:1262-1291: `/ Float` multiplies by the reciprocal (quirk Q10)
inline Vec3 operator/(const Vec3& a, Float s) { Float inv = 1.0f / s; return Vec3(a.x * inv, a.y * inv, a.z * inv); }
inline Float dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // geometry.rs:630
inline Float abs_dot(const Vec3& a, const Vec3& b) { return std::fabs(dot(a, b)); }
inline Float length_squared(const Vec3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline Float length(const Vec3& a) { return std::sqrt(length_squared(a)); }
inline Vec3 normalize(const Vec3& a) { return a / length(a); }   // geometry.rs:412
inline Vec3 vabs(const Vec3& a) { return Vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
// geometry.rs:680-692: cross products are evaluated in f64 and rounded once to f32
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return Vec3((Float)((ay * bz) - (az * by)), (Float)((az * bx) - (ax * bz)), (Float)((ax * by) - (ay * bx)));
}
inline Float max_component(const Vec3& v) { return fmax_(v.x, fmax_(v.y, v.z)); }  // geometry.rs:711
// geometry.rs:721-734
inline int max_dimension(const Vec3& v) {
    if (v.x > v.y) return (v.x > v.z) ? 0 : 2;
    return (v.y > v.z) ? 1 : 2;
}
inline Vec3 permute(const Vec3& v, int x, int y, int z) { return Vec3(v[x], v[y], v[z]); }
// geometry.rs:779-794
inline void coordinate_system(const Vec3& v1, Vec3& v2, Vec3& v3) {
    if (std::fabs(v1.x) > std::fabs(v1.y)) v2 = Vec3(-v1.z, 0.0f, v1.x) / std::sqrt(v1.x * v1.x + v1.z * v1.z);
    else v2 = Vec3(0.0f, v1.z, -v1.y) / std::sqrt(v1.y * v1.y + v1.z * v1.z);
    v3 = cross(v1, v2);
}
inline Vec3 faceforward(const Vec3& n, const Vec3& v) { return (dot(n, v) < 0.0f) ? -n : n; }  // geometry.rs:1842-1858

// geometry.rs:1535-1556
inline Point3 offset_ray_origin(const Point3& p, const Vec3& p_error, const Normal3& n, const Vec3& w) {
    Float d = dot(vabs(n), p_error);
    Vec3 offset = n * d;
    if (dot(w, n) < 0.0f) offset = -offset;
    Point3 po = p + offset;
    for (int i = 0; i < 3; ++i) {
        if (offset[i] > 0.0f) po.at(i) = next_float_up(po[i]);
        else if (offset[i] < 0.0f) po.at(i) = next_float_down(po[i]);
    }
    return po;
}

struct Vec2 { Float x, y; Vec2() : x(0), y(0) {} Vec2(Float a, Float b) : x(a), y(b) {} };

// RGBSpectrum, src/core/spectrum.rs:1530-1780
struct Spectrum {
    Float c[3];
    Spectrum() { c[0] = c[1] = c[2] = 0.0f; }
    explicit Spectrum(Float v) { c[0] = c[1] = c[2] = v; }
    Spectrum(Float r, Float g, Float b) { c[0] = r; c[1] = g; c[2] = b; }
    bool is_black() const { return c[0] == 0.0f && c[1] == 0.0f && c[2] == 0.0f; }
    bool has_nans() const { return c[0] != c[0] || c[1] != c[1] || c[2] != c[2]; }
    Float y() const { return 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]; }  // spectrum.rs:1581
    Float max_component_value() const { return fmax_(fmax_(c[0], c[1]), c[2]); }
};
inline Spectrum operator+(const Spectrum& a, const Spectrum& b) { return Spectrum(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]); }
inline Spectrum operator-(const Spectrum& a, const Spectrum& b) { return Spectrum(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]); }
inline Spectrum operator*(const Spectrum& a, const Spectrum& b) { return Spectrum(a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2]); }
inline Spectrum operator/(const Spectrum& a, const Spectrum& b) { return Spectrum(a.c[0] / b.c[0], a.c[1] / b.c[1], a.c[2] / b.c[2]); }
inline Spectrum operator*(const Spectrum& a, Float s) { return Spectrum(a.c[0] * s, a.c[1] * s, a.c[2] * s); }
inline Spectrum operator*(Float s, const Spectrum& a) { return Spectrum(s * a.c[0], s * a.c[1], s * a.c[2]); }
// spectrum.rs:1752-1762: `/ Float` on a spectrum is a true division (quirk Q10)
inline Spectrum operator/(const Spectrum& a, Float s) { return Spectrum(a.c[0] / s, a.c[1] / s, a.c[2] / s); }
inline Spectrum& operator+=(Spectrum& a, const Spectrum& b) { a.c[0] += b.c[0]; a.c[1] += b.c[1]; a.c[2] += b.c[2]; return a; }
inline Spectrum& operator*=(Spectrum& a, const Spectrum& b) { a.c[0] *= b.c[0]; a.c[1] *= b.c[1]; a.c[2] *= b.c[2]; return a; }
inline Spectrum sqrt(const Spectrum& a) { return Spectrum(std::sqrt(a.c[0]), std::sqrt(a.c[1]), std::sqrt(a.c[2])); }
inline Spectrum clamp_spectrum(const Spectrum& a, Float lo, Float hi) {
    return Spectrum(clamp_t(a.c[0], lo, hi), clamp_t(a.c[1], lo, hi), clamp_t(a.c[2], lo, hi));
}

}  // namespace orc
