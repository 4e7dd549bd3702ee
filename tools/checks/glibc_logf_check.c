// Validate a restatement of glibc 2.39's logf (sysdeps/ieee754/flt-32/e_logf.c: 16-entry table of 1/c and ln(c), degree-3
// polynomial in double; the -mfma ifunc variant, fused operations read off the disassembly of libm.so.6) against the libm on this
// box, over every non-negative float.  The device code in rs_pbrt_b200/csrc/pb_math.cuh (log_rn) is the same text.
//   gcc -O2 -ffp-contract=off -o /tmp/logf_check tools/checks/glibc_logf_check.c -lm -lpthread && /tmp/logf_check
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline uint32_t fu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float uf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static const double T[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2,
    0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3, 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5,
    0x1p+0, 0x0p+0, 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3, 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,
    0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};
static float my_logf(float x) {
    uint32_t ix = fu(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2u == 0u) return -INFINITY;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return (x - x) / (x - x);
        ix = fu(x * 0x1p23f);
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    uint32_t i = (tmp >> 19) & 15u;
    int32_t k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = T[2 * i], logc = T[2 * i + 1];
    double z = (double)uf(iz);
    double r = fma(z, invc, -1.0);
    double y0 = fma((double)k, 0x1.62e42fefa39efp-1, logc);
    double r2 = r * r;
    double y = fma(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = fma(-0x1.00ea348b88334p-2, r2, y);
    y = fma(y, r2, y0 + r);
    return (float)y;
}
static uint64_t bad[64];
static void* run(void* arg) {
    long t = (long)arg;
    uint64_t lo = (uint64_t)t << 25, hi = lo + (1u << 25), b = 0;
    for (uint64_t u = lo; u < hi; ++u) {
        float x = uf((uint32_t)u);
        float a = logf(x), m = my_logf(x);
        if (fu(a) != fu(m) && !(a != a && m != m)) { if (b < 3) printf("x=%a libm=%a mine=%a\n", x, a, m); ++b; }
    }
    bad[t] = b;
    return 0;
}
int main(void) {
    pthread_t th[64];
    for (long t = 0; t < 64; ++t) pthread_create(&th[t], 0, run, (void*)t);
    uint64_t b = 0;
    for (int t = 0; t < 64; ++t) { pthread_join(th[t], 0); b += bad[t]; }
    printf("logf: 2^31 non-negative patterns, mismatches = %llu\n", (unsigned long long)b);
    return b != 0;
}
