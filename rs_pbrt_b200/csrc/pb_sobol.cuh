// pb_sobol.cuh -- SobolSampler (global sampler) on the device.
//   sobol_interval_to_index   src/core/lowdiscrepancy.rs:1014-1043
//   sobol_sample_float        src/core/lowdiscrepancy.rs:1053-1076
//   SobolSampler::sample_dimension / get_1d / get_2d   src/samplers/sobol.rs:125-200
//   radical_inverse           src/core/lowdiscrepancy.rs:1080-1145
// The generator matrices live in data/sobol_tables.bin (see tools/extract_sobol_tables.py).
#pragma once
#include "pb_math.cuh"

namespace pb {

#define PB_SOBOL_MATRIX_SIZE 52
#define PB_SOBOL_DIMS 1024

// vdc / vdci: the rows for resolution exponent m (52 u64 each)
PB_D uint64_t sobol_interval_to_index(const uint64_t* __restrict__ vdc, const uint64_t* __restrict__ vdci, uint32_t m, uint64_t frame,
                                      int px, int py) {
    if (m == 0) return 0;
    const uint32_t m2 = m << 1;
    uint64_t index = frame << m2;
    uint64_t delta = 0;
    for (int c = 0; frame > 0; frame >>= 1, ++c)
        if (frame & 1) delta ^= vdc[c];
    uint64_t b = ((uint64_t)(((uint32_t)px) << m) | (uint64_t)(int64_t)py) ^ delta;
    for (int c = 0; b > 0; b >>= 1, ++c)
        if (b & 1) index ^= vdci[c];
    return index;
}

// m32: SOBOL_MATRICES_32 (row `dim`, 52 columns)
PB_D float sobol_sample_float(const uint32_t* __restrict__ m32, uint64_t a, uint32_t dim) {
    uint32_t v = 0;
    const uint32_t* row = m32 + dim * PB_SOBOL_MATRIX_SIZE;
    for (int i = 0; a != 0; a >>= 1, ++i)
        if (a & 1) v ^= row[i];
    // u32 -> f32 rounds to nearest even, like Rust's `as f32`
    return fminf(__uint2float_rn(v) * 2.3283064365386963e-10f, PB_ONE_MINUS_EPSILON);
}

struct SobolCtx {
    const uint32_t* m32;   // generator matrices (shared or global memory)
    uint64_t index;        // interval_sample_index
    uint32_t dim;          // next dimension
    bool overflow;         // the reference panics past 1024 dimensions (sobol.rs:119-124); we flag
};
// dims 0/1 are only drawn by the camera sample (raygen); every later draw is a plain dimension
PB_D float sobol_get_1d(SobolCtx& s) {
    if (s.dim >= PB_SOBOL_DIMS) { s.overflow = true; return 0.0f; }
    float r = sobol_sample_float(s.m32, s.index, s.dim);
    s.dim += 1;
    return r;
}
PB_D float2 sobol_get_2d(SobolCtx& s) {
    if (s.dim + 1 >= PB_SOBOL_DIMS) { s.overflow = true; return make_float2(0.0f, 0.0f); }
    float y = sobol_sample_float(s.m32, s.index, s.dim + 1);
    float x = sobol_sample_float(s.m32, s.index, s.dim);
    s.dim += 2;
    return make_float2(x, y);
}

PB_D uint32_t reverse_bits_32(uint32_t n) { return __brev(n); }
PB_D uint64_t reverse_bits_64(uint64_t n) { return __brevll(n); }
PB_D float radical_inverse_specialized(uint32_t base, uint64_t a) {
    const float inv_base = 1.0f / (float)base;
    uint64_t reversed = 0;
    float inv_base_n = 1.0f;
    while (a != 0) {
        uint64_t next = a / base;
        uint64_t digit = a - next * base;
        reversed = reversed * base + digit;
        inv_base_n *= inv_base;
        a = next;
    }
    return fminf(__ull2float_rn(reversed) * inv_base_n, PB_ONE_MINUS_EPSILON);
}
PB_D float radical_inverse(int base_index, uint64_t a) {
    switch (base_index) {
        case 0: return __ull2float_rn(reverse_bits_64(a)) * 5.421010862427522e-20f;  // 0x1p-64
        case 1: return radical_inverse_specialized(3, a);
        case 2: return radical_inverse_specialized(5, a);
        case 3: return radical_inverse_specialized(7, a);
        default: return radical_inverse_specialized(11, a);
    }
}

}  // namespace pb
