// ORACLE -- TEST INFRASTRUCTURE ONLY (see o_math.hpp).
// o_sampler.hpp: SobolSampler and HaltonSampler (global samplers), Sobol' index/sample functions, (scrambled) radical inverse, PCG32.
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
// lowdiscrepancy.rs:18-147: PRIMES / PRIME_SUMS are the first 1000 primes and their exclusive prefix sums
static const int PRIME_TABLE_SIZE = 1000;
struct PrimeTables {
    std::vector<uint32_t> primes, sums;
    PrimeTables() {
        for (uint32_t c = 2; primes.size() < (size_t)PRIME_TABLE_SIZE; ++c) {
            bool is_prime = true;
            for (uint32_t q : primes) { if (q * q > c) break; if (c % q == 0) { is_prime = false; break; } }
            if (is_prime) primes.push_back(c);
        }
        uint32_t acc = 0;
        for (uint32_t q : primes) { sums.push_back(acc); acc += q; }
    }
};
inline const PrimeTables& prime_tables() { static PrimeTables t; return t; }
inline Float radical_inverse(int base_index, uint64_t a) {
    if (base_index == 0) return (Float)reverse_bits_64(a) * 5.421010862427522e-20f /* 0x1.0p-64 */;
    return radical_inverse_specialized(prime_tables().primes[base_index], a);
}

// src/core/rng.rs:13-83 (PCG32)
struct Rng {
    uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
    uint32_t uniform_uint32() {
        uint64_t oldstate = state;
        state = oldstate * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((oldstate >> 18) ^ oldstate) >> 27);
        uint32_t rot = (uint32_t)(oldstate >> 59);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    uint32_t uniform_uint32_bounded(uint32_t b) {
        uint32_t threshold = (~b + 1u) & b;  // rng.rs:61: `&`, where pbrt-v3 has `%` -- restated as written
        for (;;) {
            uint32_t r = uniform_uint32();
            if (r >= threshold) return r % b;
        }
    }
};
// sampling.rs:202-212 with n_dimensions = 1; lowdiscrepancy.rs:2165-2187
inline std::vector<uint16_t> compute_radical_inverse_permutations(Rng& rng) {
    const PrimeTables& T = prime_tables();
    std::vector<uint16_t> perms((size_t)T.sums.back() + T.primes.back());
    size_t p = 0;
    for (int i = 0; i < PRIME_TABLE_SIZE; ++i) {
        const int32_t count = (int32_t)T.primes[i];
        for (int32_t j = 0; j < count; ++j) perms[p + j] = (uint16_t)j;
        for (int32_t k = 0; k < count; ++k) {
            int32_t other = k + (int32_t)rng.uniform_uint32_bounded((uint32_t)(count - k));
            std::swap(perms[p + k], perms[p + other]);
        }
        p += (size_t)count;
    }
    return perms;
}
inline const std::vector<uint16_t>& radical_inverse_permutations() {  // halton.rs:18-26: one table per process, Rng::new()
    static std::vector<uint16_t> perms = [] { Rng rng; return compute_radical_inverse_permutations(rng); }();
    return perms;
}
// lowdiscrepancy.rs:1101-1122
inline Float scrambled_radical_inverse(int base_index, uint64_t a, const uint16_t* perm) {
    const uint64_t base = prime_tables().primes[base_index];
    const Float inv_base = 1.0f / (Float)base;
    uint64_t reversed_digits = 0;
    Float inv_base_n = 1.0f;
    while (a != 0) {
        uint64_t next = a / base;
        uint64_t digit = a - next * base;
        reversed_digits = reversed_digits * base + perm[digit];
        inv_base_n *= inv_base;
        a = next;
    }
    return fmin_(inv_base_n * ((Float)reversed_digits + inv_base * (Float)perm[0] / (1.0f - inv_base)), FLOAT_ONE_MINUS_EPSILON);
}
// lowdiscrepancy.rs:788-797
inline uint64_t inverse_radical_inverse(uint64_t base, uint64_t inverse, uint64_t n_digits) {
    uint64_t index = 0;
    for (uint64_t i = 0; i < n_digits; ++i) {
        uint64_t digit = inverse % base;
        inverse /= base;
        index = index * base + digit;
    }
    return index;
}

inline bool is_power_of_2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }
inline int32_t round_up_pow2_32(int32_t v) {
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v + 1;
}
inline int log2_int(uint32_t v) { int r = 0; while (v > 1) { v >>= 1; ++r; } return r; }

// src/samplers/sobol.rs:15-272 (no sample arrays are requested by the path integrator,
// so array_start_dim == array_end_dim == 5)
// Sampler enum (src/core/sampler.rs): the two global samplers in scope share GlobalSampler's get_1d / get_2d bookkeeping
struct Sampler {
    int64_t samples_per_pixel = 0;
    int64_t dimension = 0;
    uint64_t interval_sample_index = 0;
    int32_t current_pixel[2] = {0, 0};
    int64_t current_pixel_sample_index = 0;
    static const int64_t array_start_dim = 5;
    int64_t array_end_dim = 5;
    // 2D sample arrays (sobol.rs / halton.rs request_2d_array, start_pixel, get_2d_array); no integrator in scope asks for 1D ones
    std::vector<int32_t> samples_2d_array_sizes;
    std::vector<std::vector<Vec2>> sample_array_2d;
    size_t array_2d_offset = 0;
    virtual ~Sampler() {}
    virtual uint64_t get_index_for_sample(uint64_t sample_num) = 0;
    virtual Float sample_dimension(uint64_t index, int64_t dim) const = 0;
    void request_2d_array(int32_t n) {  // round_count(n) == n for both global samplers
        samples_2d_array_sizes.push_back(n);
        sample_array_2d.emplace_back((size_t)n * (size_t)samples_per_pixel);
    }
    void start_pixel(int32_t x, int32_t y) {
        current_pixel[0] = x; current_pixel[1] = y;
        current_pixel_sample_index = 0;
        array_2d_offset = 0;
        dimension = 0;
        interval_sample_index = get_index_for_sample(0);
        array_end_dim = array_start_dim + 2 * (int64_t)sample_array_2d.size();
        int64_t dim = array_start_dim;
        for (size_t i = 0; i < samples_2d_array_sizes.size(); ++i) {
            const size_t n_samples = (size_t)samples_2d_array_sizes[i] * (size_t)samples_per_pixel;
            for (size_t j = 0; j < n_samples; ++j) {
                const uint64_t idx = get_index_for_sample((uint64_t)j);
                const Float ax = sample_dimension(idx, dim);
                const Float ay = sample_dimension(idx, dim + 1);
                sample_array_2d[i][j] = Vec2(ax, ay);
            }
            dim += 2;
        }
    }
    const Vec2* get_2d_array(int32_t n) {
        if (array_2d_offset == sample_array_2d.size()) return nullptr;
        const size_t start = (size_t)current_pixel_sample_index * (size_t)n;
        array_2d_offset += 1;
        return sample_array_2d[array_2d_offset - 1].data() + start;
    }
    // get_2d_array_idxs (sobol.rs:227-237): false = every requested array has been handed out for this sample
    bool get_2d_array_idxs(int32_t n, size_t& idx, size_t& start) {
        if (array_2d_offset == sample_array_2d.size()) return false;
        start = (size_t)current_pixel_sample_index * (size_t)n;
        idx = array_2d_offset;
        array_2d_offset += 1;
        return true;
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
        array_2d_offset = 0;
        interval_sample_index = get_index_for_sample((uint64_t)current_pixel_sample_index + 1);
        current_pixel_sample_index += 1;
        return current_pixel_sample_index < samples_per_pixel;
    }
    bool set_sample_number(int64_t n) {
        dimension = 0;
        array_2d_offset = 0;
        interval_sample_index = get_index_for_sample((uint64_t)n);
        current_pixel_sample_index = n;
        return n < samples_per_pixel;
    }
};

struct SobolSampler : Sampler {
    int32_t sb_min[2], sb_max[2];
    int32_t resolution, log2_resolution;

    SobolSampler(int64_t spp, const int32_t sample_bounds[4]) {
        samples_per_pixel = spp;  // already rounded up by the caller (sobol.rs:39-45)
        sb_min[0] = sample_bounds[0]; sb_min[1] = sample_bounds[1];
        sb_max[0] = sample_bounds[2]; sb_max[1] = sample_bounds[3];
        int32_t dx = sb_max[0] - sb_min[0], dy = sb_max[1] - sb_min[1];
        resolution = round_up_pow2_32(std::max(dx, dy));
        log2_resolution = log2_int((uint32_t)resolution);
    }
    uint64_t get_index_for_sample(uint64_t sample_num) override {
        return sobol_interval_to_index((uint32_t)log2_resolution, sample_num, current_pixel[0] - sb_min[0], current_pixel[1] - sb_min[1]);
    }
    Float sample_dimension(uint64_t index, int64_t dim) const override {
        Float s = sobol_sample_float((int64_t)index, (int)dim, 0);
        if (dim == 0 || dim == 1) {
            s = s * (Float)resolution + (Float)sb_min[dim];
            s = clamp_t(s - (Float)current_pixel[dim], 0.0f, FLOAT_ONE_MINUS_EPSILON);
        }
        return s;
    }
};

// src/samplers/halton.rs:28-260
struct HaltonSampler : Sampler {
    static const int32_t K_MAX_RESOLUTION = 128;
    int32_t base_scales[2], base_exponents[2];
    uint64_t sample_stride;
    int64_t mult_inverse[2];
    int32_t pixel_for_offset[2] = {0, 0};
    uint64_t offset_for_current_pixel = 0;
    bool sample_at_pixel_center;

    static void extended_gcd(uint64_t a, uint64_t b, int64_t& x, int64_t& y) {
        if (b == 0) { x = 1; y = 0; return; }
        int64_t d = (int64_t)a / (int64_t)b, xp = 0, yp = 0;
        extended_gcd(b, a % b, xp, yp);
        x = yp;
        y = xp - (d * yp);
    }
    static uint64_t multiplicative_inverse(int64_t a, int64_t n) {
        int64_t x = 0, y = 0;
        extended_gcd((uint64_t)a, (uint64_t)n, x, y);
        int64_t r = x - (x / n) * n;  // mod_t pbrt.rs:127-140
        if (r < 0) r += n;
        return (uint64_t)r;
    }
    HaltonSampler(int64_t spp, const int32_t sample_bounds[4], bool at_center) {
        samples_per_pixel = spp;
        sample_at_pixel_center = at_center;
        const int32_t res[2] = {sample_bounds[2] - sample_bounds[0], sample_bounds[3] - sample_bounds[1]};
        for (int i = 0; i < 2; ++i) {
            const int32_t base = i == 0 ? 2 : 3;
            int32_t scale = 1, exp = 0;
            while (scale < std::min(res[i], K_MAX_RESOLUTION)) { scale *= base; exp += 1; }
            base_scales[i] = scale;
            base_exponents[i] = exp;
        }
        sample_stride = (uint64_t)base_scales[0] * (uint64_t)base_scales[1];
        mult_inverse[0] = (int64_t)multiplicative_inverse(base_scales[1], base_scales[0]);
        mult_inverse[1] = (int64_t)multiplicative_inverse(base_scales[0], base_scales[1]);
    }
    uint64_t get_index_for_sample(uint64_t sample_num) override {
        if (current_pixel[0] != pixel_for_offset[0] || current_pixel[1] != pixel_for_offset[1]) {
            offset_for_current_pixel = 0;
            if (sample_stride > 1) {
                for (int i = 0; i < 2; ++i) {
                    int32_t pm = current_pixel[i] - (current_pixel[i] / K_MAX_RESOLUTION) * K_MAX_RESOLUTION;  // mod_t
                    if (pm < 0) pm += K_MAX_RESOLUTION;
                    uint64_t dim_offset = inverse_radical_inverse(i == 0 ? 2 : 3, (uint64_t)pm, (uint64_t)base_exponents[i]);
                    offset_for_current_pixel += dim_offset * (sample_stride / (uint64_t)base_scales[i]) * (uint64_t)mult_inverse[i];
                }
                offset_for_current_pixel %= sample_stride;
            }
            pixel_for_offset[0] = current_pixel[0];
            pixel_for_offset[1] = current_pixel[1];
        }
        return offset_for_current_pixel + sample_num * sample_stride;
    }
    Float sample_dimension(uint64_t index, int64_t dim) const override {
        if (sample_at_pixel_center && (dim == 0 || dim == 1)) return 0.5f;
        if (dim == 0) return radical_inverse(0, index >> (uint64_t)base_exponents[0]);
        if (dim == 1) return radical_inverse(1, index / (uint64_t)base_scales[1]);
        if (dim >= PRIME_TABLE_SIZE) throw std::runtime_error("HaltonSampler can only sample 1000 dimensions");
        return scrambled_radical_inverse((int)dim, index, radical_inverse_permutations().data() + prime_tables().sums[dim]);
    }
};

}  // namespace orc
