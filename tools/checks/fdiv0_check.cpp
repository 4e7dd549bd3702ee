// fdiv0 / spdiv0 (rs_pbrt_b200/csrc/pb_math.cuh) against the hardware IEEE division, on the host: every special value against
// every special value, and 2e8 random bit patterns with a zero or random numerator.  Bit patterns must agree (NaNs compare equal
// as a class).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
using std::isinf;
// host stand-ins for the device intrinsics that the device-only parts of the header use
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline int __double2int_rz(double d) { return (int)d; }
#include "../../rs_pbrt_b200/csrc/pb_math.cuh"
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float fl(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static bool same(float a, float b) { return (a != a && b != b) || bits(a) == bits(b); }
int main() {
    const float sp[] = {0.0f, -0.0f, 1.0f, -1.0f, 1e-45f, -1e-45f, 1.1754944e-38f, 3.4028235e38f, -3.4028235e38f, INFINITY, -INFINITY, NAN, 0.5f, 3.0f};
    uint64_t bad = 0, n = 0;
    for (float a : sp) for (float b : sp) { volatile float q = a / b; if (!same(pb::fdiv0(a, b), q)) { bad++; printf("bad %a / %a\n", a, b); } n++; }
    uint64_t st = 88172645463325252ull;
    for (uint64_t i = 0; i < 200000000ull; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        float b = fl((uint32_t)st), a = (i & 1) ? fl((uint32_t)(st >> 32)) : ((i & 2) ? 0.0f : -0.0f);
        volatile float q = a / b;
        if (!same(pb::fdiv0(a, b), q)) { if (bad < 5) printf("bad %a / %a\n", a, b); bad++; }
        n++;
    }
    printf("fdiv0: %llu mismatches of %llu\n", (unsigned long long)bad, (unsigned long long)n);
    // while the device header is compiled for the host anyway: its libm restatements against libm on a sample
    uint64_t bad2 = 0;
    for (uint64_t i = 0; i < 20000000ull; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        float x = (float)((double)(st & 0xffffffffu) / 4294967296.0 * 16.0 - 8.0), y = (float)((double)(st >> 32) / 4294967296.0 * 2.0 - 1.0);
        float s, c;
        pb::sincos_rn(x, s, c);
        if (!same(s, sinf(x)) || !same(c, cosf(x)) || !same(pb::acos_rn(y), acosf(y)) || !same(pb::atan2_rn(y, x), atan2f(y, x))) bad2++;
    }
    printf("device-header sin/cos/acos/atan2 vs libm: %llu mismatches of 20000000\n", (unsigned long long)bad2);
    bad += bad2;
    return bad ? 1 : 0;
}
