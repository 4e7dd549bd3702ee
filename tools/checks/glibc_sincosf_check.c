// Validate a restatement of glibc 2.28+ sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h algorithm) against the libm on this box.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <pthread.h>
#ifndef USE_FMA
#define USE_FMA 0
#endif
static inline double madd(double a, double b, double c) {  // a + b*c
#if USE_FMA
    return fma(b, c, a);
#else
    return a + b * c;
#endif
}
typedef struct { double sign[4]; double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3; } sincos_t;
static const sincos_t T[2] = {
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10,
     0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10,
     -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }
static inline float sinf_poly(double x, double x2, const sincos_t* p, int n) {
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = madd(p->s2, x2, p->s3);
        double x7 = x3 * x2;
        double s = madd(x, x3, p->s1);
        return (float)madd(s, x7, s1);
    } else {
        double x4 = x2 * x2;
        double c2 = madd(p->c3, x2, p->c4);
        double c1 = madd(p->c0, x2, p->c1);
        double x6 = x4 * x2;
        double c = madd(c1, x4, p->c2);
        return (float)madd(c, x6, c2);
    }
}
static inline double reduce_fast(double x, const sincos_t* p, int* np) {
    double r = x * p->hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
#if USE_FMA
    return fma(-(double)n, p->hpi, x);
#else
    return x - n * p->hpi;
#endif
}
static float my_sinf(float y) {
    double x = y, s; int n; const sincos_t* p = &T[0];
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        s = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sinf_poly(x, s, p, 0);
    }
    x = reduce_fast(x, p, &n);
    s = p->sign[n & 3];
    if (n & 2) p = &T[1];
    return sinf_poly(x * s, x * x, p, n);
}
static float my_cosf(float y) {
    double x = y, s; int n; const sincos_t* p = &T[0];
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        s = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sinf_poly(x, s, p, 1);
    }
    x = reduce_fast(x, p, &n);
    s = p->sign[n & 3];
    if (n & 2) p = &T[1];
    return sinf_poly(x * s, x * x, p, n ^ 1);
}
typedef struct { uint32_t lo, hi; uint64_t bad_s, bad_c, n; uint32_t first_bad; } job;
static void* run(void* a) {
    job* j = (job*)a;
    for (uint32_t u = j->lo; u < j->hi; ++u) {
        float x; memcpy(&x, &u, 4);
        for (int sgn = 0; sgn < 2; ++sgn) {
            float v = sgn ? -x : x;
            float a1 = sinf(v), a2 = my_sinf(v), b1 = cosf(v), b2 = my_cosf(v);
            if (asuint(a1) != asuint(a2)) { if (!j->bad_s) j->first_bad = asuint(v); j->bad_s++; }
            if (asuint(b1) != asuint(b2)) j->bad_c++;
            j->n++;
        }
    }
    return 0;
}
int main() {
    // all floats with |x| < 120 (top of reduce_fast's range)
    float lim = 120.0f; uint32_t top = asuint(lim);
    enum { NT = 8 };
    pthread_t th[NT]; job jobs[NT];
    for (int t = 0; t < NT; ++t) { jobs[t] = (job){(uint32_t)((uint64_t)top * t / NT), (uint32_t)((uint64_t)top * (t + 1) / NT), 0, 0, 0, 0}; pthread_create(&th[t], 0, run, &jobs[t]); }
    uint64_t bs = 0, bc = 0, n = 0; uint32_t fb = 0;
    for (int t = 0; t < NT; ++t) { pthread_join(th[t], 0); bs += jobs[t].bad_s; bc += jobs[t].bad_c; n += jobs[t].n; if (!fb) fb = jobs[t].first_bad; }
    printf("USE_FMA=%d tested %llu values: sin mismatches %llu, cos mismatches %llu, first bad bits %08x\n", USE_FMA, (unsigned long long)n, (unsigned long long)bs, (unsigned long long)bc, fb);
    return 0;
}
