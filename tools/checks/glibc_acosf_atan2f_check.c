// Validate a restatement of glibc 2.39's acosf / atanf / atan2f (the fdlibm-derived single-precision routines in
// sysdeps/ieee754/flt-32/{e_acosf,s_atanf,e_atan2f}.c: plain float arithmetic, no FMA) against the libm on this box.
// The device code in rs_pbrt_b200/csrc/pb_math.cuh is the same text with CUDA intrinsics.
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint32_t fu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float uf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define F(u) uf(u)
static float my_acosf(float x) {
    const float one = 1.0f, pi = F(0x40490fda), pio2_hi = F(0x3fc90fda), pio2_lo = F(0x33a22168);
    const float pS0 = F(0x3e2aaaab), pS1 = -F(0x3ea6b090), pS2 = F(0x3e4e0aa8), pS3 = -F(0x3d241146), pS4 = F(0x3a4f7f04), pS5 = F(0x3811ef08);
    const float qS1 = -F(0x4019d139), qS2 = F(0x4001572d), qS3 = -F(0x3f303361), qS4 = F(0x3d9dc62e);
    int32_t hx = (int32_t)fu(x), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
    if (ix > 0x3f800000) return (x - x) / (x - x);
    float z, p, q, r, w, s, c, df;
    if (ix < 0x3f000000) {
        if (ix <= 0x32800000) return pio2_hi + pio2_lo;
        z = x * x;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (hx < 0) {
        z = (one + x) * 0.5f;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        s = sqrtf(z);
        r = p / q;
        w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    } else {
        z = (one - x) * 0.5f;
        s = sqrtf(z);
        df = uf(fu(s) & 0xfffff000u);
        c = (z - df * df) / (s + df);
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        w = r * s + c;
        return 2.0f * (df + w);
    }
}
static float my_atanf(float x) {
    static const uint32_t hi[4] = {0x3eed6338, 0x3f490fda, 0x3f7b985e, 0x3fc90fda}, lo[4] = {0x31ac3769, 0x33222168, 0x33140fb4, 0x33a22168};
    const float aT0 = F(0x3eaaaaab), aT1 = -F(0x3e4ccccd), aT2 = F(0x3e124925), aT3 = -F(0x3de38e38), aT4 = F(0x3dba2e6e), aT5 = -F(0x3d9d8795),
                aT6 = F(0x3d886b35), aT7 = -F(0x3d6ef16b), aT8 = F(0x3d4bda59), aT9 = F(0xbd15a221), aT10 = F(0x3c8569d7);
    int32_t hx = (int32_t)fu(x), ix = hx & 0x7fffffff, id;
    if (ix >= 0x4c000000) {
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? F(hi[3]) + F(lo[3]) : -F(hi[3]) - F(lo[3]);
    }
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    float z = x * x, w = z * z;
    float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    z = F(hi[id]) - ((x * (s1 + s2) - F(lo[id])) - x);
    return hx < 0 ? -z : z;
}
static float my_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = F(0x3f490fdb), pi_o_2 = F(0x3fc90fdb), pi = F(0x40490fdb), pi_lo = -F(0x33bbbd2e);
    int32_t hx = (int32_t)fu(x), ix = hx & 0x7fffffff, hy = (int32_t)fu(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return my_atanf(y);
    int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) { case 0: case 1: return y; case 2: return pi + tiny; default: return -pi - tiny; }
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) { case 0: return pi_o_4 + tiny; case 1: return -pi_o_4 - tiny; case 2: return 3.0f * pi_o_4 + tiny; default: return -3.0f * pi_o_4 - tiny; }
        } else {
            switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return pi + tiny; default: return -pi - tiny; }
        }
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = my_atanf(fabsf(y / x));
    switch (m) {
        case 0: return z;
        case 1: return uf(fu(z) ^ 0x80000000u);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}
typedef struct { int t; uint64_t bad_acos, bad_atan, bad_atan2, n; uint32_t first[3]; } job;
enum { NT = 8 };
static void* run(void* a) {
    job* j = (job*)a;
    // acosf: every float in [-1, 1]; atanf: every finite float (strided over threads)
    for (uint64_t u = (uint64_t)j->t; u <= 0x3f800000u; u += NT) {
        for (int sg = 0; sg < 2; ++sg) {
            float x = uf((uint32_t)u | (sg ? 0x80000000u : 0u));
            if (fu(acosf(x)) != fu(my_acosf(x))) { if (!j->bad_acos) j->first[0] = fu(x); j->bad_acos++; }
            j->n++;
        }
    }
    for (uint64_t u = (uint64_t)j->t; u < 0x7f800000u; u += NT) {
        for (int sg = 0; sg < 2; ++sg) {
            float x = uf((uint32_t)u | (sg ? 0x80000000u : 0u));
            if (fu(atanf(x)) != fu(my_atanf(x))) { if (!j->bad_atan) j->first[1] = fu(x); j->bad_atan++; }
        }
    }
    // atan2f: random pairs -- unit-vector-like components, full-range floats, and special values
    uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)(j->t + 1);
    const float special[] = {0.0f, -0.0f, 1.0f, -1.0f, 1e-30f, -1e-30f, 1e30f, 3.0e-39f, 0.5f, -0.5f, INFINITY, -INFINITY};
    for (uint64_t i = 0; i < 120000000ull; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        float y, x;
        int mode = (int)(i % 4);
        if (mode == 0) { y = (float)((double)(st & 0xffffffff) / 4294967296.0 * 2.0 - 1.0); x = (float)((double)(st >> 32) / 4294967296.0 * 2.0 - 1.0); }
        else if (mode == 1) { y = uf((uint32_t)st); x = uf((uint32_t)(st >> 32)); }
        else if (mode == 2) { y = (float)((double)(st & 0xffffffff) / 4294967296.0 * 2.0 - 1.0); x = special[(st >> 40) % 12]; }
        else { x = (float)((double)(st & 0xffffffff) / 4294967296.0 * 2.0 - 1.0); y = special[(st >> 40) % 12]; }
        float a = atan2f(y, x), b = my_atan2f(y, x);
        if (fu(a) != fu(b) && !(a != a && b != b)) { if (!j->bad_atan2) { j->first[2] = fu(y); } j->bad_atan2++; }
    }
    return 0;
}
int main(void) {
    pthread_t th[NT]; job jobs[NT];
    for (int t = 0; t < NT; ++t) { memset(&jobs[t], 0, sizeof(job)); jobs[t].t = t; pthread_create(&th[t], 0, run, &jobs[t]); }
    uint64_t a = 0, b = 0, c = 0, n = 0;
    for (int t = 0; t < NT; ++t) { pthread_join(th[t], 0); a += jobs[t].bad_acos; b += jobs[t].bad_atan; c += jobs[t].bad_atan2; n += jobs[t].n; }
    printf("acosf: %llu mismatches of %llu; atanf: %llu mismatches of all finite floats; atan2f: %llu mismatches of %llu pairs\n", (unsigned long long)a,
           (unsigned long long)n, (unsigned long long)b, (unsigned long long)c, (unsigned long long)(120000000ull * NT));
    for (int t = 0; t < NT; ++t) if (jobs[t].bad_acos || jobs[t].bad_atan || jobs[t].bad_atan2) { printf("first bad (thread %d): acos %08x atan %08x atan2.y %08x\n", t, jobs[t].first[0], jobs[t].first[1], jobs[t].first[2]); break; }
    return (a || b || c) ? 1 : 0;
}
