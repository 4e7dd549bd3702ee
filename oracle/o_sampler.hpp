// ORACLE -- TEST INFRASTRUCTURE ONLY (see o_math.hpp).
// o_sampler.hpp: SobolSampler (global sampler), Sobol' index/sample functions, radical inverse.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "o_math.hpp"

namespace orc {

// data/sobol_tables.bin, produced by tools/extract_sobol_tables.py from
// src/core/sobolmatrices.rs (SOBOL_MATRICES_32 :7, VD_C_SOBOL_MATRICES :53463, VD_C_SOBOL_MATRICES_INV :54155)
struct SobolTables {
    std::vector<uint32_t> m32;   // 1024 * 52
    std::vector<uint64_t> vdc;   // 25 * 52 (row m-1)
    std::vector<uint64_t> vdci;  // 26 * 52
    bool loaded = false;
    void load(const char* path) {
        FILE* f = std::fopen(path, "rb");
        if (!f) throw std::runtime_error(std::string("cannot open ") + path);
        uint32_t hdr[8];
        if (std::fread(hdr, 4, 8, f) != 8 || hdr[0] != 0x4C424F53u || hdr[1] != 1024 || hdr[2] != 52) {
            std::fclose(f);
            throw std::runtime_error("bad sobol table header");
        }
        m32.resize(1024 * 52);
        vdc.resize(25 * 52);
        vdci.resize(26 * 52);
        bool ok = std::fread(m32.data(), 4, m32.size(), f) == m32.size() && std::fread(vdc.data(), 8, vdc.size(), f) == vdc.size() &&
                  std::fread(vdci.data(), 8, vdci.size(), f) == vdci.size();
        std::fclose(f);
        if (!ok) throw std::runtime_error("short sobol table file");
        loaded = true;
    }
};
inline SobolTables& sobol_tables() { static SobolTables t; return t; }

static const int NUM_SOBOL_DIMENSIONS = 1024;
static const int SOBOL_MATRIX_SIZE = 52;

// src/core/lowdiscrepancy.rs:1014-1043
inline uint64_t sobol_interval_to_index(uint32_t m, uint64_t frame, int32_t px, int32_t py) {
    if (m == 0) return 0;
    const SobolTables& T = sobol_tables();
    const uint32_t m2 = m << 1;
    uint64_t index = frame << m2;
    uint64_t delta = 0;
    for (int c = 0; frame > 0; frame >>= 1, ++c)
        if (frame & 1) delta ^= T.vdc[(m - 1) * 52 + c];
    uint64_t b = ((uint64_t)(((uint32_t)px) << m) | (uint64_t)(int64_t)py) ^ delta;
    for (int c = 0; b > 0; b >>= 1, ++c)
        if (b & 1) index ^= T.vdci[(m - 1) * 52 + c];
    return index;
}

// src/core/lowdiscrepancy.rs:1053-1076
inline Float sobol_sample_float(int64_t a, int dimension, uint32_t scramble) {
    if (dimension >= NUM_SOBOL_DIMENSIONS) throw std::runtime_error("Integrator has consumed too many Sobol' dimensions");
    const SobolTables& T = sobol_tables();
    uint32_t v = scramble;
    for (size_t i = (size_t)dimension * SOBOL_MATRIX_SIZE; a != 0; a >>= 1, ++i)
        if (a & 1) v ^= T.m32[i];
    return fmin_((Float)v * 2.3283064365386963e-10f /* 0x1.0p-32 */, FLOAT_ONE_MINUS_EPSILON);
}

inline uint32_t reverse_bits_32(uint32_t n) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
    n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
    n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
    n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
    return n;
}
inline uint64_t reverse_bits_64(uint64_t n) {
    uint64_t n0 = reverse_bits_32((uint32_t)n), n1 = reverse_bits_32((uint32_t)(n >> 32));
    return (n0 << 32) | n1;
}
// src/core/lowdiscrepancy.rs:1080-1145 (bases 2,3,5,7,11 are all the light grid needs)
inline Float radical_inverse_specialized(uint64_t base, uint64_t a) {
    const Float inv_base = 1.0f / (Float)base;
    uint64_t reversed = 0;
    Float inv_base_n = 1.0f;
    while (a != 0) {
        uint64_t next = a / base;
        uint64_t digit = a - next * base;
        reversed = reversed * base + digit;
        inv_base_n *= inv_base;
        a = next;
    }
    return fmin_((Float)reversed * inv_base_n, FLOAT_ONE_MINUS_EPSILON);
}
inline Float radical_inverse(int base_index, uint64_t a) {
    static const uint64_t primes[5] = {2, 3, 5, 7, 11};
    if (base_index == 0) return (Float)reverse_bits_64(a) * 5.421010862427522e-20f /* 0x1.0p-64 */;
    return radical_inverse_specialized(primes[base_index], a);
}

inline bool is_power_of_2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }
inline int32_t round_up_pow2_32(int32_t v) {
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v + 1;
}
inline int log2_int(uint32_t v) { int r = 0; while (v > 1) { v >>= 1; ++r; } return r; }

// src/samplers/sobol.rs:15-272 (no sample arrays are requested by the path integrator,
// so array_start_dim == array_end_dim == 5)
struct SobolSampler {
    int64_t samples_per_pixel;
    int32_t sb_min[2], sb_max[2];
    int32_t resolution, log2_resolution;
    int64_t dimension = 0;
    uint64_t interval_sample_index = 0;
    int32_t current_pixel[2] = {0, 0};
    int64_t current_pixel_sample_index = 0;
    static const int64_t array_start_dim = 5, array_end_dim = 5;

    SobolSampler(int64_t spp, const int32_t sample_bounds[4]) {
        samples_per_pixel = spp;  // already rounded up by the caller (sobol.rs:39-45)
        sb_min[0] = sample_bounds[0]; sb_min[1] = sample_bounds[1];
        sb_max[0] = sample_bounds[2]; sb_max[1] = sample_bounds[3];
        int32_t dx = sb_max[0] - sb_min[0], dy = sb_max[1] - sb_min[1];
        resolution = round_up_pow2_32(std::max(dx, dy));
        log2_resolution = log2_int((uint32_t)resolution);
    }
    uint64_t get_index_for_sample(uint64_t sample_num) const {
        return sobol_interval_to_index((uint32_t)log2_resolution, sample_num, current_pixel[0] - sb_min[0], current_pixel[1] - sb_min[1]);
    }
    Float sample_dimension(uint64_t index, int64_t dim) const {
        Float s = sobol_sample_float((int64_t)index, (int)dim, 0);
        if (dim == 0 || dim == 1) {
            s = s * (Float)resolution + (Float)sb_min[dim];
            s = clamp_t(s - (Float)current_pixel[dim], 0.0f, FLOAT_ONE_MINUS_EPSILON);
        }
        return s;
    }
    void start_pixel(int32_t x, int32_t y) {
        current_pixel[0] = x; current_pixel[1] = y;
        current_pixel_sample_index = 0;
        dimension = 0;
        interval_sample_index = get_index_for_sample(0);
    }
    Float get_1d() {
        if (dimension >= array_start_dim && dimension < array_end_dim) dimension = array_end_dim;
        Float r = sample_dimension(interval_sample_index, dimension);
        dimension += 1;
        return r;
    }
    Vec2 get_2d() {
        if (dimension + 1 >= array_start_dim && dimension < array_end_dim) dimension = array_end_dim;
        Float y = sample_dimension(interval_sample_index, dimension + 1);
        Float x = sample_dimension(interval_sample_index, dimension);
        dimension += 2;
        return Vec2(x, y);
    }
    bool start_next_sample() {
        dimension = 0;
        interval_sample_index = get_index_for_sample((uint64_t)current_pixel_sample_index + 1);
        current_pixel_sample_index += 1;
        return current_pixel_sample_index < samples_per_pixel;
    }
    bool set_sample_number(int64_t n) {
        dimension = 0;
        interval_sample_index = get_index_for_sample((uint64_t)n);
        current_pixel_sample_index = n;
        return n < samples_per_pixel;
    }
};

}  // namespace orc
