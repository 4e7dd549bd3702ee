// Validate a restatement of glibc 2.39's log2f (sysdeps/ieee754/flt-32/e_log2f.c, the ARM optimized-routines algorithm: 16-entry
// table of 1/c and log2(c), degree-4 polynomial in double; on x86-64 CPUs with FMA the ifunc picks the -mfma build, in which every
// a*b+c of the source is one fused operation -- read off the disassembly of libm.so.6) against the libm on this box, over every
// non-negative float.  The device code in rs_pbrt_b200/csrc/pb_math.cuh (log2_rn) is the same text with CUDA intrinsics.
//   gcc -O2 -ffp-contract=off -o /tmp/log2f_check tools/checks/glibc_log2f_check.c -lm -lpthread && /tmp/log2f_check
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline uint32_t fu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float uf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static const double T[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2}, {0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2},
    {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2}, {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4}, {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5},
    {0x1p+0, 0x0p+0}, {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4}, {0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3}, {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2}, {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},
    {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};
static const double A[4] = {-0x1.712b6f70a7e4dp-2, 0x1.ecabf496832ep-2, -0x1.715479ffae3dep-1, 0x1.715475f35c8b8p0};
static float my_log2f(float x) {
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
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int32_t k = (int32_t)tmp >> 23;
    double invc = T[i][0], logc = T[i][1];
    double z = (double)uf(iz);
    double r = fma(z, invc, -1.0);
    double y0 = logc + (double)k;
    double r2 = r * r;
    double y = fma(A[1], r, A[2]);
    y = fma(A[0], r2, y);
    double p = fma(A[3], r, y0);
    y = fma(y, r2, p);
    return (float)y;
}
static uint64_t bad[64];
static void* run(void* arg) {
    long t = (long)arg;
    uint64_t lo = (uint64_t)t << 25, hi = lo + (1u << 25), b = 0;  // 64 threads x 2^25 = every pattern with the sign bit clear
    for (uint64_t u = lo; u < hi; ++u) {
        float x = uf((uint32_t)u);
        float a = log2f(x), m = my_log2f(x);
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
    float neg[] = {-0.0f, -1.0f, -INFINITY};
    for (int i = 0; i < 3; ++i) { float a = log2f(neg[i]), m = my_log2f(neg[i]); if (fu(a) != fu(m) && !(a != a && m != m)) { printf("x=%a libm=%a mine=%a\n", neg[i], a, m); ++b; } }
    printf("log2f: 2^31 non-negative patterns (+3 negative), mismatches = %llu\n", (unsigned long long)b);
    return b != 0;
}
