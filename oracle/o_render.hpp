// ORACLE -- TEST INFRASTRUCTURE ONLY (see o_math.hpp).
// o_render.hpp: Scene, SurfaceInteraction, DiffuseAreaLight, light distributions,
// uniform_sample_one_light / estimate_direct, PathIntegrator::li, camera, film, tile render loop.
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "o_geom.hpp"
#include "o_reflection.hpp"
#include "o_sampler.hpp"

namespace orc {

struct Counters {
    uint64_t camera_rays = 0, closest_rays = 0, shadow_rays = 0, nodes_visited = 0, tris_tested = 0, light_tri_tests = 0;
    void add(const Counters& o) {
        camera_rays += o.camera_rays; closest_rays += o.closest_rays; shadow_rays += o.shadow_rays;
        nodes_visited += o.nodes_visited; tris_tested += o.tris_tested; light_tri_tests += o.light_tri_tests;
    }
};

// sampling.rs:17-147
struct Distribution1D {
    std::vector<Float> func, cdf;
    Float func_int = 0.0f;
    Distribution1D() {}
    explicit Distribution1D(const std::vector<Float>& f) : func(f) {
        size_t n = f.size();
        cdf.resize(n + 1);
        cdf[0] = 0.0f;
        for (size_t i = 1; i <= n; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / (Float)n;
        func_int = cdf[n];
        if (func_int == 0.0f) for (size_t i = 1; i <= n; ++i) cdf[i] = (Float)i / (Float)n;
        else for (size_t i = 1; i <= n; ++i) cdf[i] /= func_int;
    }
    size_t count() const { return func.size(); }
    Float sample_continuous(Float u, Float* pdf, size_t* off) const {  // sampling.rs:53-102
        size_t first = 0, len = cdf.size();
        while (len > 0) {
            size_t half = len >> 1, middle = first + half;
            if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; }
            else len = half;
        }
        size_t offset = (size_t)clamp_t((long)first - 1, 0L, (long)cdf.size() - 2);
        if (off) *off = offset;
        Float du = u - cdf[offset];
        if ((cdf[offset + 1] - cdf[offset]) > 0.0f) du /= cdf[offset + 1] - cdf[offset];
        if (pdf) *pdf = (func_int > 0.0f) ? func[offset] / func_int : 0.0f;
        return ((Float)offset + du) / (Float)count();
    }
    size_t sample_discrete(Float u, Float& pdf) const {
        size_t first = 0, len = cdf.size();
        while (len > 0) {
            size_t half = len >> 1, middle = first + half;
            if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; }
            else len = half;
        }
        long off = clamp_t((long)first - 1, 0L, (long)cdf.size() - 2);
        pdf = (func_int > 0.0f) ? func[off] / (func_int * (Float)func.size()) : 0.0f;
        return (size_t)off;
    }
};

// sampling.rs:150-198
struct Distribution2D {
    std::vector<Distribution1D> p_conditional_v;
    Distribution1D p_marginal;
    Distribution2D() {}
    Distribution2D(const std::vector<Float>& func, int nu, int nv) {
        std::vector<Float> marginal;
        for (int v = 0; v < nv; ++v) {
            p_conditional_v.emplace_back(std::vector<Float>(func.begin() + (size_t)v * nu, func.begin() + (size_t)(v + 1) * nu));
            marginal.push_back(p_conditional_v.back().func_int);
        }
        p_marginal = Distribution1D(marginal);
    }
    Vec2 sample_continuous(const Vec2& u, Float& pdf) const {
        Float pdfs[2] = {0.0f, 0.0f};
        size_t v = 0;
        Float d1 = p_marginal.sample_continuous(u.y, &pdfs[1], &v);
        Float d0 = p_conditional_v[v].sample_continuous(u.x, &pdfs[0], nullptr);
        pdf = pdfs[0] * pdfs[1];
        return Vec2(d0, d1);
    }
    Float pdf(const Vec2& p) const {
        size_t nu = p_conditional_v[0].count(), nv = p_marginal.count();
        size_t iu = (size_t)clamp_t((long)f2i(p.x * (Float)nu), 0L, (long)nu - 1);
        size_t iv = (size_t)clamp_t((long)f2i(p.y * (Float)nv), 0L, (long)nv - 1);
        return p_conditional_v[iv].func[iu] / p_marginal.func_int;
    }
};

// MipMap<Spectrum> (mipmap.rs:36-396): InfiniteAreaLight's map (ImageWrap::Repeat, isotropic lookups) and ImageTexture<Spectrum>
// (any wrap mode, trilinear or EWA lookups).
struct MipMapRGB {
    struct Level { int us, vs; std::vector<Spectrum> t; };
    std::vector<Level> pyramid;
    uint32_t wrap = PBRT_WRAP_REPEAT;
    bool do_trilinear = false;
    Float max_anisotropy = 8.0f;
    Float weight_lut[128];
    int width() const { return pyramid[0].us; }
    int height() const { return pyramid[0].vs; }
    size_t levels() const { return pyramid.size(); }
    MipMapRGB() {}
    // texture.rs:426-439
    static Float lanczos(Float x, Float tau) {
        x = std::fabs(x);
        if (x < 1e-5f) return 1.0f;
        if (x > 1.0f) return 0.0f;
        x *= PI;
        Float s = std::sin(x * tau) / (x * tau);
        Float l = std::sin(x) / x;
        return s * l;
    }
    struct ResampleWeight { int32_t first_texel; Float weight[4]; };
    static std::vector<ResampleWeight> resample_weights(int32_t old_res, int32_t new_res) {  // mipmap.rs:298-322
        std::vector<ResampleWeight> wt;
        const Float filterwidth = 2.0f;
        for (int32_t i = 0; i < new_res; ++i) {
            Float center = ((Float)i + 0.5f) * (Float)old_res / (Float)new_res;
            ResampleWeight rw;
            rw.first_texel = f2i(std::floor((center - filterwidth) + 0.5f));
            for (int j = 0; j < 4; ++j) {
                Float pos = (Float)rw.first_texel + (Float)j + 0.5f;
                rw.weight[j] = lanczos((pos - center) / filterwidth, 2.0f);
            }
            Float inv_sum_wts = 1.0f / (rw.weight[0] + rw.weight[1] + rw.weight[2] + rw.weight[3]);
            for (int j = 0; j < 4; ++j) rw.weight[j] *= inv_sum_wts;
            wt.push_back(rw);
        }
        return wt;
    }
    static int32_t mod_i(int32_t a, int32_t b) { int32_t r = a - (a / b) * b; return r < 0 ? r + b : r; }  // pbrt.rs mod_t
    static int32_t wrap_index(uint32_t wrap, int32_t i, int32_t n) {  // the match in the resampling loops, mipmap.rs:88-92
        if (wrap == PBRT_WRAP_REPEAT) return mod_i(i, n);
        if (wrap == PBRT_WRAP_CLAMP) return clamp_t(i, 0, n - 1);
        return i;
    }
    MipMapRGB(int w, int h, const float* rgb, uint32_t wrap_mode = PBRT_WRAP_REPEAT, bool trilinear = false, Float max_aniso = 8.0f)
        : wrap(wrap_mode), do_trilinear(trilinear), max_anisotropy(max_aniso) {
        for (int i = 0; i < 128; ++i) {  // EWA filter weights, mipmap.rs:188-195
            Float alpha = 2.0f;
            Float r2 = (Float)i / (Float)(128 - 1);
            weight_lut[i] = std::exp(-alpha * r2) - std::exp(-alpha);
        }
        std::vector<Spectrum> img((size_t)w * h);
        for (size_t i = 0; i < (size_t)w * h; ++i) img[i] = Spectrum(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        if ((w & (w - 1)) || (h & (h - 1))) {  // resample to power-of-two resolution, mipmap.rs:65-149
            const int pw = round_up_pow2_32(w), ph = round_up_pow2_32(h);
            std::vector<Spectrum> res((size_t)pw * ph);
            std::vector<ResampleWeight> sw = resample_weights(w, pw);
            for (int t = 0; t < h; ++t)
                for (int s = 0; s < pw; ++s) {
                    Spectrum acc;
                    for (int j = 0; j < 4; ++j) {
                        int32_t orig_s = wrap_index(wrap, sw[s].first_texel + j, w);
                        if (orig_s >= 0 && orig_s < w) acc += img[(size_t)t * w + orig_s] * sw[s].weight[j];
                    }
                    res[(size_t)t * pw + s] = acc;
                }
            std::vector<ResampleWeight> tw = resample_weights(h, ph);
            std::vector<Spectrum> work(ph);
            for (int s = 0; s < pw; ++s) {
                for (int t = 0; t < ph; ++t) {
                    work[t] = Spectrum();
                    for (int j = 0; j < 4; ++j) {
                        int32_t offset = wrap_index(wrap, tw[t].first_texel + j, h);
                        if (offset >= 0 && offset < h) work[t] += res[(size_t)offset * pw + s] * tw[t].weight[j];
                    }
                }
                for (int t = 0; t < ph; ++t)
                    res[(size_t)t * pw + s] = Spectrum(clamp_t(work[t].c[0], 0.0f, INF), clamp_t(work[t].c[1], 0.0f, INF), clamp_t(work[t].c[2], 0.0f, INF));
            }
            img.swap(res);
            w = pw; h = ph;
        }
        Level l0{w, h, std::move(img)};
        pyramid.push_back(std::move(l0));
        size_t n_levels = 1 + (size_t)f2i(std::log2((Float)std::max(w, h)));
        for (size_t i = 1; i < n_levels; ++i) {
            int s_res = std::max(1, pyramid[i - 1].us / 2), t_res = std::max(1, pyramid[i - 1].vs / 2);
            Level l{s_res, t_res, std::vector<Spectrum>((size_t)s_res * t_res)};
            for (int t = 0; t < t_res; ++t)
                for (int s = 0; s < s_res; ++s)
                    l.t[(size_t)t * s_res + s] =
                        (texel(i - 1, 2 * s, 2 * t) + texel(i - 1, 2 * s + 1, 2 * t) + texel(i - 1, 2 * s, 2 * t + 1) + texel(i - 1, 2 * s + 1, 2 * t + 1)) * 0.25f;
            pyramid.push_back(std::move(l));
        }
    }
    // mipmap.rs:208-232.  Repeat: (s as usize) mod size, sizes are powers of two.  Black answers the clamped texel too (the
    // reference's "TMP" branch), so it differs from Clamp only in the resampling above.
    const Spectrum& texel(size_t level, long s, long t) const {
        const Level& l = pyramid[level];
        size_t ss, tt;
        if (wrap == PBRT_WRAP_REPEAT) { ss = (size_t)s % (size_t)l.us; tt = (size_t)t % (size_t)l.vs; }
        else { ss = (size_t)clamp_t(s, 0L, (long)l.us - 1); tt = (size_t)clamp_t(t, 0L, (long)l.vs - 1); }
        return l.t[tt * l.us + ss];
    }
    Spectrum triangle(size_t level, const Vec2& st) const {  // mipmap.rs:323-336
        level = std::min(level, levels() - 1);
        Float s = st.x * (Float)pyramid[level].us - 0.5f, t = st.y * (Float)pyramid[level].vs - 0.5f;
        long s0 = (long)std::floor(s), t0 = (long)std::floor(t);
        Float ds = s - (Float)s0, dt = t - (Float)t0;
        Spectrum tmp1 = texel(level, s0 + 1, t0 + 1) * (ds * dt);
        Spectrum tmp2 = texel(level, s0 + 1, t0) * (ds * (1.0f - dt));
        Spectrum tmp3 = texel(level, s0, t0 + 1) * ((1.0f - ds) * dt);
        Spectrum tmp4 = texel(level, s0, t0) * ((1.0f - ds) * (1.0f - dt));
        return tmp4 + tmp3 + tmp2 + tmp1;
    }
    Spectrum lookup(const Vec2& st, Float width) const {  // lookup_pnt_flt mipmap.rs:233-252
        Float level = (Float)levels() - 1.0f + std::log2(std::max(width, 1e-8f));
        if (level < 0.0f) return triangle(0, st);
        if (level >= (Float)levels() - 1.0f) return texel(levels() - 1, 0, 0);
        size_t i_level = (size_t)std::floor(level);
        Float delta = level - (Float)i_level;
        return triangle(i_level, st) * (1.0f - delta) + triangle(i_level + 1, st) * delta;
    }
    // lookup_pnt_vec_vec, mipmap.rs:253-296
    Spectrum lookup(const Vec2& st, Vec2 dst0, Vec2 dst1) const {
        if (do_trilinear) {
            Float width = fmax_(fmax_(std::fabs(dst0.x), std::fabs(dst0.y)), fmax_(std::fabs(dst1.x), std::fabs(dst1.y)));
            return lookup(st, width);
        }
        if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) std::swap(dst0, dst1);
        Float major_length = std::sqrt(dst0.x * dst0.x + dst0.y * dst0.y);
        Float minor_length = std::sqrt(dst1.x * dst1.x + dst1.y * dst1.y);
        if (minor_length * max_anisotropy < major_length && minor_length > 0.0f) {
            Float scale = major_length / (minor_length * max_anisotropy);
            dst1 = Vec2(dst1.x * scale, dst1.y * scale);
            minor_length *= scale;
        }
        if (minor_length == 0.0f) return triangle(0, st);
        Float lod = fmax_(0.0f, (Float)levels() - 1.0f + std::log2(minor_length));
        size_t ilod = (size_t)std::floor(lod);
        Spectrum col2 = ewa(ilod + 1, st, dst0, dst1);
        Spectrum col1 = ewa(ilod, st, dst0, dst1);
        Float tt = lod - (Float)ilod;
        return col1 * (1.0f - tt) + col2 * tt;  // pbrt.rs lerp: (1 - t) * a + t * b
    }
    Spectrum ewa(size_t level, const Vec2& st_in, Vec2 dst0, Vec2 dst1) const {  // mipmap.rs:337-396
        if (level >= levels()) return texel(levels() - 1, 0, 0);
        const Float us = (Float)pyramid[level].us, vs = (Float)pyramid[level].vs;
        Vec2 st(st_in.x * us - 0.5f, st_in.y * vs - 0.5f);
        dst0 = Vec2(dst0.x * us, dst0.y * vs);
        dst1 = Vec2(dst1.x * us, dst1.y * vs);
        Float a = dst0.y * dst0.y + dst1.y * dst1.y + 1.0f;
        Float b = -2.0f * (dst0.x * dst0.y + dst1.x * dst1.y);
        Float c = dst0.x * dst0.x + dst1.x * dst1.x + 1.0f;
        Float inv_f = 1.0f / (a * c - b * b * 0.25f);
        a *= inv_f; b *= inv_f; c *= inv_f;
        Float det = -b * b + 4.0f * a * c;
        Float inv_det = 1.0f / det;
        Float u_sqrt = std::sqrt(det * c), v_sqrt = std::sqrt(a * det);
        const Float fs0 = std::ceil(st.x - 2.0f * inv_det * u_sqrt), fs1 = std::floor(st.x + 2.0f * inv_det * u_sqrt);
        const Float ft0 = std::ceil(st.y - 2.0f * inv_det * v_sqrt), ft1 = std::floor(st.y + 2.0f * inv_det * v_sqrt);
        // The level is chosen from the minor axis and the major one is at most max_anisotropy times longer, so a footprint this wide
        // (> 16 M texel visits) needs non-finite ellipse coefficients or a "maxanisotropy" in the thousands; the reference's loop
        // would not end in any useful time there.  Answer the level's first texel instead (the one deliberate deviation of this
        // function; the GPU path does the same).
        if (!(fs1 - fs0 <= 4096.0f) || !(ft1 - ft0 <= 4096.0f)) return texel(level, 0, 0);
        long s0 = (long)fs0, s1 = (long)fs1, t0 = (long)ft0, t1 = (long)ft1;
        Spectrum sum;
        Float sum_wts = 0.0f;
        for (long it = t0; it <= t1; ++it) {
            Float tt = (Float)it - st.y;
            for (long is = s0; is <= s1; ++is) {
                Float ss = (Float)is - st.x;
                Float r2 = a * ss * ss + b * ss * tt + c * tt * tt;
                if (r2 < 1.0f) {
                    size_t index = r2 <= 0.0f ? 0 : std::min((size_t)(r2 * 128.0f), (size_t)127);  // `as usize` saturates at 0
                    Float weight = weight_lut[index];
                    sum += texel(level, is, it) * weight;
                    sum_wts += weight;
                }
            }
        }
        return sum / sum_wts;
    }
};

// ImageTexture<Spectrum> with its UVMapping2D (imagemap.rs:17-150, texture.rs:93-122)
struct ImageTexture {
    MipMapRGB mipmap;
    Float su, sv, du, dv;
    // ImageTexture<Float> (channels == 1) runs the same arithmetic on one value; it is carried in three equal channels here
    static std::vector<float> as_rgb(const PbrtTexture& t) {
        const size_t n = (size_t)t.res[0] * t.res[1];
        std::vector<float> v(3 * n);
        for (size_t i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) v[3 * i + c] = t.channels == 1 ? t.texels[i] : t.texels[3 * i + c];
        return v;
    }
    uint32_t kind = PBRT_TEX_IMAGE, channels = 3, child[3] = {0, 0, 0};
    Spectrum value;
    uint32_t mapping = PBRT_MAP_UV;
    float map_m[16] = {0};
    ImageTexture(const PbrtTexture& t, int)  // ConstantTexture / ScaleTexture / MixTexture
        : mipmap(1, 1, zero3(), PBRT_WRAP_REPEAT, true, 8.0f), su(1), sv(1), du(0), dv(0), kind(t.kind), channels(t.channels),
          value(t.value[0], t.channels == 1 ? t.value[0] : t.value[1], t.channels == 1 ? t.value[0] : t.value[2]) {
        for (int i = 0; i < 3; ++i) child[i] = t.child[i];
    }
    static const float* zero3() { static const float z[3] = {0, 0, 0}; return z; }
    ImageTexture(const PbrtTexture& t)
        : mipmap((int)t.res[0], (int)t.res[1], as_rgb(t).data(), t.wrap, t.trilinear != 0, t.max_anisotropy), su(t.su), sv(t.sv), du(t.du), dv(t.dv), channels(t.channels), mapping(t.mapping) {
        for (int i = 0; i < 16; ++i) map_m[i] = t.map_m[i];
    }
};

// InfiniteAreaLight's map and sampling distribution (infinite.rs:250-300 and the image branches above it)
struct EnvLight {
    MipMapRGB lmap;
    Distribution2D distribution;
    EnvLight(int w, int h, const float* rgb) : lmap(w, h, rgb) {
        int width = 2 * lmap.width(), height = 2 * lmap.height();
        std::vector<Float> img;
        Float fwidth = 0.5f / std::min((Float)width, (Float)height);
        for (int v = 0; v < height; ++v) {
            Float vp = ((Float)v + 0.5f) / (Float)height;
            Float sin_theta = std::sin(PI * ((Float)v + 0.5f) / (Float)height);
            for (int u = 0; u < width; ++u) {
                Float up = ((Float)u + 0.5f) / (Float)width;
                img.push_back(lmap.lookup(Vec2(up, vp), fwidth).y() * sin_theta);
            }
        }
        distribution = Distribution2D(img, width, height);
    }
};

struct InteractionCommon {
    Point3 p;
    Float time = 0.0f;
    Vec3 p_error, wo;
    Normal3 n;
};
// interaction.rs:58-94
inline Ray spawn_ray(const InteractionCommon& it, const Vec3& d) {
    return Ray(offset_ray_origin(it.p, it.p_error, it.n, d), d, INF, it.time);
}
inline Ray spawn_ray_to(const InteractionCommon& a, const InteractionCommon& b) {
    Point3 origin = offset_ray_origin(a.p, a.p_error, a.n, b.p - a.p);
    Point3 target = offset_ray_origin(b.p, b.p_error, b.n, origin - b.p);
    return Ray(origin, target - origin, 1.0f - SHADOW_EPSILON, a.time);
}

struct SurfaceInteraction {
    InteractionCommon common;
    Vec2 uv;
    Vec3 dpdu, dpdv;
    Normal3 shading_n;
    Vec3 shading_dpdu, shading_dpdv;
    Normal3 shading_dndu, shading_dndv;  // triangle.rs:389-421 (isect.dndu / dndv themselves stay zero: the outer variables of :332-333)
    bool shape_flips = false;             // isect.shape is Some and reverse_orientation ^ transform_swaps_handedness (set_shading_geometry)
    Float dudx = 0, dvdx = 0, dudy = 0, dvdy = 0;  // compute_differentials (interaction.rs:388-474)
    Vec3 dpdx, dpdy;
    int32_t prim = -1;  // index into Scene::tris (isect.primitive)
    bool primitive_lost = false;  // isect.primitive == None after transform_surface_interaction (quirk Q7)
    Float b[3] = {0, 0, 0};
};

struct AreaLight {  // Light enum, in-scope kinds: DiffuseAreaLight over one triangle (lights/diffuse.rs), PointLight,
                    // SpotLight, DistantLight (lights/{point,spot,distant}.rs)
    int kind = PBRT_LIGHT_DIFFUSE_AREA;
    Spectrum l_emit;  // l_emit | I | L
    uint32_t tri = 0;
    bool two_sided = false;
    Float area = 0.0f;
    uint32_t n_samples = 1;  // Light::get_n_samples
    Vec3 p;           // p_light | w_light
    Float w2l[9] = {0};
    Float cos_total_width = 0.0f, cos_falloff_start = 0.0f;
    Float l2w[9] = {0};
    std::shared_ptr<EnvLight> env;  // InfiniteAreaLight (lights/infinite.rs)
    bool is_delta() const { return kind != PBRT_LIGHT_DIFFUSE_AREA && kind != PBRT_LIGHT_INFINITE; }  // light.rs:178-190
};

// Transform::transform_point / transform_vector / transform_ray (transform.rs:490-550,662-708), row-major m[16]
inline Point3 xf_point(const float* m, const Point3& p) {
    Float x = p.x, y = p.y, z = p.z;
    Float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
    Float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
    Float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
    Float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (wp == 1.0f) return Point3(xp, yp, zp);
    Float inv = 1.0f / wp;
    return Point3(inv * xp, inv * yp, inv * zp);
}
inline Vec3 xf_vector(const float* m, const Vec3& v) {
    return Vec3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
inline Ray xf_ray(const float* m, const Ray& r) {
    Float x = r.o.x, y = r.o.y, z = r.o.z;
    Point3 o = xf_point(m, r.o);
    Vec3 o_error = Vec3(std::fabs(m[0] * x) + std::fabs(m[1] * y) + std::fabs(m[2] * z) + std::fabs(m[3]),
                        std::fabs(m[4] * x) + std::fabs(m[5] * y) + std::fabs(m[6] * z) + std::fabs(m[7]),
                        std::fabs(m[8] * x) + std::fabs(m[9] * y) + std::fabs(m[10] * z) + std::fabs(m[11])) * gamma(3);
    Vec3 d = xf_vector(m, r.d);
    Float ls = length_squared(d);
    Float t_max = r.t_max;
    if (ls > 0.0f) {
        Float dt = dot(vabs(d), o_error) / ls;
        o = o + d * dt;
        t_max -= dt;
    }
    Ray out(o, d, t_max, r.time);
    if (r.has_differential) {  // transform.rs:550-556
        out.has_differential = true;
        out.rx_origin = xf_point(m, r.rx_origin);
        out.ry_origin = xf_point(m, r.ry_origin);
        out.rx_direction = xf_vector(m, r.rx_direction);
        out.ry_direction = xf_vector(m, r.ry_direction);
    }
    return out;
}

// Transform::transform_point_with_abs_error (transform.rs:709-760)
inline Point3 xf_point_abs_error(const float* m, const Point3& pt, const Vec3& pe, Vec3& abs_error) {
    Float x = pt.x, y = pt.y, z = pt.z;
    Float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
    Float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
    Float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
    Float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    abs_error.x = (gamma(3) + 1.0f) * (std::fabs(m[0]) * pe.x + std::fabs(m[1]) * pe.y + std::fabs(m[2]) * pe.z) +
                  gamma(3) * (std::fabs(m[0] * x) + std::fabs(m[1] * y) + std::fabs(m[2] * z) + std::fabs(m[3]));
    abs_error.y = (gamma(3) + 1.0f) * (std::fabs(m[4]) * pe.x + std::fabs(m[5]) * pe.y + std::fabs(m[6]) * pe.z) +
                  gamma(3) * (std::fabs(m[4] * x) + std::fabs(m[5] * y) + std::fabs(m[6] * z) + std::fabs(m[7]));
    abs_error.z = (gamma(3) + 1.0f) * (std::fabs(m[8]) * pe.x + std::fabs(m[9]) * pe.y + std::fabs(m[10]) * pe.z) +
                  gamma(3) * (std::fabs(m[8] * x) + std::fabs(m[9] * y) + std::fabs(m[10] * z) + std::fabs(m[11]));
    if (wp == 1.0f) return Point3(xp, yp, zp);
    Float inv = 1.0f / wp;
    return Point3(inv * xp, inv * yp, inv * zp);
}
// Transform::transform_normal (transform.rs:528-537): the transpose of m_inv
inline Normal3 xf_normal(const float* mi, const Normal3& n) {
    return Normal3(mi[0] * n.x + mi[4] * n.y + mi[8] * n.z, mi[1] * n.x + mi[5] * n.y + mi[9] * n.z, mi[2] * n.x + mi[6] * n.y + mi[10] * n.z);
}

struct SurfaceInteraction;
struct ImageTexture;
inline Spectrum texture_evaluate(const std::vector<std::unique_ptr<ImageTexture>>& all, const ImageTexture& t, const SurfaceInteraction& si);

struct Scene {
    std::vector<PbrtBvhNode> nodes;
    std::vector<PbrtTri> tris;
    std::vector<Mesh> meshes;
    std::vector<MaterialLobes> materials;
    std::vector<MaterialLobes> materials_single;           // built with allow_multiple_lobes = false (directlighting.rs:77, whitted.rs:64)
    mutable bool allow_multiple_lobes = true;              // of the render in progress: PathIntegrator true, Direct / Whitted false
    std::vector<PbrtMaterial> material_src;                // as described, for materials with image textures (evaluated per hit)
    std::vector<std::unique_ptr<ImageTexture>> textures;
    std::vector<AreaLight> lights;
    std::vector<PbrtInstance> instances;  // TransformedPrimitives (primitive.rs:198-272)
    mutable uint32_t instancing = PBRT_INSTANCING_REFERENCE;  // PbrtRenderParams.instancing of the render in progress
    PbrtCamera camera;
    Bounds3 world_bound;

    void tri_verts(const PbrtTri& t, Point3& p0, Point3& p1, Point3& p2) const {
        const Mesh& m = meshes[t.mesh];
        p0 = m.P(t.v[0]); p1 = m.P(t.v[1]); p2 = m.P(t.v[2]);
    }
    void get_uvs(const PbrtTri& t, Vec2 uv[3]) const {  // triangle.rs:96-110
        const Mesh& m = meshes[t.mesh];
        if (m.uv.empty()) { uv[0] = Vec2(0, 0); uv[1] = Vec2(1, 0); uv[2] = Vec2(1, 1); }
        else { uv[0] = m.UV(t.v[0]); uv[1] = m.UV(t.v[1]); uv[2] = m.UV(t.v[2]); }
    }

    // Everything in Triangle::intersect after the hit test (triangle.rs:274-448).
    void fill_interaction(const PbrtTri& tri, const Ray& ray, const TriHit& h, SurfaceInteraction& isect) const {
        const Mesh& mesh = meshes[tri.mesh];
        Point3 p0, p1, p2;
        tri_verts(tri, p0, p1, p2);
        const Float b0 = h.b0, b1 = h.b1, b2 = h.b2;
        Vec2 uv[3];
        get_uvs(tri, uv);
        Vec2 duv02(uv[0].x - uv[2].x, uv[0].y - uv[2].y), duv12(uv[1].x - uv[2].x, uv[1].y - uv[2].y);
        Vec3 dp02 = p0 - p2, dp12 = p1 - p2;
        Float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
        bool degenerate_uv = std::fabs(determinant) < 1e-8f;
        Vec3 dpdu, dpdv;
        if (!degenerate_uv) {
            Float invdet = 1.0f / determinant;
            dpdu = (dp02 * duv12.y - dp12 * duv02.y) * invdet;
            dpdv = (dp02 * -duv12.x + dp12 * duv02.x) * invdet;
        }
        if (degenerate_uv || length_squared(cross(dpdu, dpdv)) == 0.0f)
            coordinate_system(normalize(cross(p2 - p0, p1 - p0)), dpdu, dpdv);
        Float x_abs_sum = std::fabs(b0 * p0.x) + std::fabs(b1 * p1.x) + std::fabs(b2 * p2.x);
        Float y_abs_sum = std::fabs(b0 * p0.y) + std::fabs(b1 * p1.y) + std::fabs(b2 * p2.y);
        Float z_abs_sum = std::fabs(b0 * p0.z) + std::fabs(b1 * p1.z) + std::fabs(b2 * p2.z);
        Vec3 p_error = Vec3(x_abs_sum, y_abs_sum, z_abs_sum) * gamma(7);
        Point3 p_hit = p0 * b0 + p1 * b1 + p2 * b2;
        Vec2 uv_hit(uv[0].x * b0 + uv[1].x * b1 + uv[2].x * b2, uv[0].y * b0 + uv[1].y * b1 + uv[2].y * b2);
        Vec3 wo = -ray.d;  // NOT normalised (quirk Q5)
        Normal3 surface_normal = normalize(cross(dp02, dp12));
        if (mesh.reverse_orientation ^ mesh.swaps_handedness) surface_normal = -surface_normal;
        Normal3 sh_n = surface_normal;
        Vec3 sh_dpdu = dpdu, sh_dpdv = dpdv;
        if (!mesh.n.empty() || !mesh.s.empty()) {
            Normal3 ns;
            if (!mesh.n.empty()) {
                ns = mesh.N(tri.v[0]) * b0 + mesh.N(tri.v[1]) * b1 + mesh.N(tri.v[2]) * b2;
                if (length_squared(ns) > 0.0f) ns = normalize(ns);
                else ns = surface_normal;
            } else ns = surface_normal;
            Vec3 ss;
            if (!mesh.s.empty()) {
                ss = mesh.S(tri.v[0]) * b0 + mesh.S(tri.v[1]) * b1 + mesh.S(tri.v[2]) * b2;
                if (length_squared(ss) > 0.0f) ss = normalize(ss);
                else ss = normalize(dpdu);
            } else ss = normalize(dpdu);
            Vec3 ts = cross(ss, ns);
            if (length_squared(ts) > 0.0f) { ts = normalize(ts); ss = cross(ts, ns); }
            else coordinate_system(ns, ss, ts);
            if (!mesh.n.empty()) {  // dndu / dndv of the shading geometry (triangle.rs:392-411)
                Normal3 dn1 = mesh.N(tri.v[0]) - mesh.N(tri.v[2]), dn2 = mesh.N(tri.v[1]) - mesh.N(tri.v[2]);
                if (!degenerate_uv) {
                    Float inv_det = 1.0f / determinant;
                    isect.shading_dndu = (dn1 * duv12.y - dn2 * duv02.y) * inv_det;
                    isect.shading_dndv = (dn1 * -duv12.x + dn2 * duv02.x) * inv_det;
                }
            }
            sh_n = normalize(cross(ss, ts));
            surface_normal = faceforward(surface_normal, sh_n);
            sh_dpdu = ss;
            sh_dpdv = ts;
        }
        isect.shape_flips = mesh.reverse_orientation ^ mesh.swaps_handedness;
        isect.common.p = p_hit;
        isect.common.time = ray.time;
        isect.common.p_error = p_error;
        isect.common.wo = wo;
        isect.common.n = surface_normal;
        isect.uv = uv_hit;
        isect.dpdu = dpdu;
        isect.dpdv = dpdv;
        isect.shading_n = sh_n;
        isect.shading_dpdu = sh_dpdu;
        isect.shading_dpdv = sh_dpdv;
        isect.b[0] = b0; isect.b[1] = b1; isect.b[2] = b2;
    }

    // The alpha tests of Triangle::intersect (alpha_mask only, triangle.rs:313-330) and Triangle::intersect_p (alpha_mask and
    // shadow_alpha_mask, after its own dpdu / dpdv block that rejects a degenerate triangle, triangle.rs:593-654).  The local interaction
    // carries p_hit, uv_hit and no differentials, which is all a texture lookup reads.  true = the candidate hit is rejected.
    bool alpha_rejects(const PbrtTri& tri, const Point3& p0, const Point3& p1, const Point3& p2, const TriHit& h, bool any_hit) const;

    // The candidate that BVHAccel::intersect last wrote into `isect`: the reference builds the full interaction for EVERY accepted
    // candidate and only the last one survives, so it is rebuilt once at the end from this record.
    struct HitRec {
        int32_t prim = -1;     // index into tris
        TriHit h;
        int32_t inst = -1;     // instance the candidate was found in (its triangle test ran in object space)
        Ray ray;               // the ray of that test (object-space ray for instance candidates)
    };
    // BVHAccel::intersect (bvh.rs:401-462) over the tree rooted at `root` -> GeometricPrimitive::intersect (primitive.rs:150-186) ->
    // Triangle::intersect, or TransformedPrimitive::intersect (primitive.rs:216-253) for an instance.  Returns the reference's hit flag.
    bool bvh_intersect(uint32_t root, const Ray& ray, HitRec& rec, int32_t inst, Counters* cnt) const {
        bool hit = false;
        Vec3 inv_dir(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        int dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
        uint32_t to_visit = 0, cur = root;
        uint32_t stack[64];
        for (;;) {
            const PbrtBvhNode& node = nodes[cur];
            if (cnt) cnt->nodes_visited++;
            if (bounds_intersect_p(node.pmin, node.pmax, ray, inv_dir, dir_is_neg)) {
                if (node.n_prims > 0) {
                    for (uint32_t i = 0; i < node.n_prims; ++i) {
                        const PbrtTri& tri = tris[node.offset + i];
                        if (tri.mesh == PBRT_MESH_INSTANCE) {
                            const PbrtInstance& I = instances[tri.v[0]];
                            // Transform::inverse(prim_to_world).transform_ray(r): the inverse carries m_inv as its matrix
                            Ray ro = xf_ray(I.m_inv, ray);
                            if (bvh_intersect(I.root, ro, rec, (int32_t)tri.v[0], cnt)) {
                                ray.t_max = ro.t_max;  // r.t_max.set(ray.t_max.get()) -- the object ray's parameter, as written
                                // REFERENCE: an identity instance has by now overwritten the interaction and shortened the ray,
                                // but reports no hit (primitive.rs:221-253); FIXED (pbrt-v3) reports every instance hit
                                if (instancing == PBRT_INSTANCING_FIXED || !I.identity) hit = true;
                            }
                            continue;
                        }
                        Point3 p0, p1, p2;
                        tri_verts(tri, p0, p1, p2);
                        TriHit h;
                        if (cnt) cnt->tris_tested++;
                        if (triangle_test(p0, p1, p2, ray, h) && !alpha_rejects(tri, p0, p1, p2, h, false)) {
                            ray.t_max = h.t;
                            rec.prim = node.offset + (int32_t)i;
                            rec.h = h;
                            rec.inst = inst;
                            rec.ray = ray;
                            hit = true;
                        }
                    }
                    if (to_visit == 0) break;
                    cur = stack[--to_visit];
                } else if (dir_is_neg[node.axis]) {
                    stack[to_visit++] = cur + 1;
                    cur = (uint32_t)node.offset;
                } else {
                    stack[to_visit++] = (uint32_t)node.offset;
                    cur = cur + 1;
                }
            } else {
                if (to_visit == 0) break;
                cur = stack[--to_visit];
            }
        }
        return hit;
    }
    // Transform::transform_surface_interaction (transform.rs:815-860) with instance_to_world
    void instance_to_world(const PbrtInstance& I, SurfaceInteraction& si) const {
        SurfaceInteraction r;
        r.common.p = xf_point_abs_error(I.m, si.common.p, si.common.p_error, r.common.p_error);
        r.common.n = normalize(xf_normal(I.m_inv, si.common.n));
        r.common.wo = normalize(xf_vector(I.m, si.common.wo));
        r.common.time = si.common.time;
        r.uv = si.uv;
        r.dpdu = xf_vector(I.m, si.dpdu);
        r.dpdv = xf_vector(I.m, si.dpdv);
        r.shading_n = normalize(xf_normal(I.m_inv, si.shading_n));
        r.shading_dpdu = xf_vector(I.m, si.shading_dpdu);
        r.shading_dpdv = xf_vector(I.m, si.shading_dpdv);
        r.shading_dndu = xf_normal(I.m_inv, si.shading_dndu);
        r.shading_dndv = xf_normal(I.m_inv, si.shading_dndv);
        r.shape_flips = false;  // ret.shape = None (transform.rs:830)
        r.shading_n = faceforward(r.shading_n, r.common.n);
        r.b[0] = si.b[0]; r.b[1] = si.b[1]; r.b[2] = si.b[2];
        r.prim = si.prim;
        si = r;
    }
    // Scene::intersect (scene.rs:55-66)
    bool intersect(const Ray& ray, SurfaceInteraction& isect, Counters* cnt, Float* t_hit_out = nullptr) const {
        if (cnt) cnt->closest_rays++;
        if (nodes.empty()) return false;
        HitRec rec;
        bool hit = bvh_intersect(0, ray, rec, -1, cnt);
        if (hit) {
            fill_interaction(tris[rec.prim], rec.ray, rec.h, isect);
            isect.prim = rec.prim;
            isect.primitive_lost = false;
            if (rec.inst >= 0) {
                const PbrtInstance& I = instances[rec.inst];
                if (instancing == PBRT_INSTANCING_FIXED) {
                    if (!I.identity) instance_to_world(I, isect);  // pbrt-v3: the primitive (material) is kept
                } else if (!I.identity) {
                    instance_to_world(I, isect);
                    isect.primitive_lost = true;                   // ret.primitive = None (transform.rs:856)
                }  // an identity instance leaves the interaction (and its primitive) as the inner intersect wrote it
            }
            if (t_hit_out) *t_hit_out = ray.t_max;
        }
        return hit;
    }
    // BVHAccel::intersect_p (bvh.rs:463-514); TransformedPrimitive::intersect_p (primitive.rs:254-261)
    bool bvh_intersect_p(uint32_t root, const Ray& ray, Counters* cnt) const {
        Vec3 inv_dir(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        int dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
        uint32_t to_visit = 0, cur = root;
        uint32_t stack[64];
        for (;;) {
            const PbrtBvhNode& node = nodes[cur];
            if (cnt) cnt->nodes_visited++;
            if (bounds_intersect_p(node.pmin, node.pmax, ray, inv_dir, dir_is_neg)) {
                if (node.n_prims > 0) {
                    for (uint32_t i = 0; i < node.n_prims; ++i) {
                        const PbrtTri& tri = tris[node.offset + i];
                        if (tri.mesh == PBRT_MESH_INSTANCE) {
                            const PbrtInstance& I = instances[tri.v[0]];
                            if (bvh_intersect_p(I.root, xf_ray(I.m_inv, ray), cnt)) return true;
                            continue;
                        }
                        Point3 p0, p1, p2;
                        tri_verts(tri, p0, p1, p2);
                        TriHit h;
                        if (cnt) cnt->tris_tested++;
                        if (triangle_test(p0, p1, p2, ray, h) && !alpha_rejects(tri, p0, p1, p2, h, true)) return true;
                    }
                    if (to_visit == 0) break;
                    cur = stack[--to_visit];
                } else if (dir_is_neg[node.axis]) {
                    stack[to_visit++] = cur + 1;
                    cur = (uint32_t)node.offset;
                } else {
                    stack[to_visit++] = (uint32_t)node.offset;
                    cur = cur + 1;
                }
            } else {
                if (to_visit == 0) break;
                cur = stack[--to_visit];
            }
        }
        return false;
    }
    bool intersect_p(const Ray& ray, Counters* cnt) const {
        if (cnt) cnt->shadow_rays++;
        if (nodes.empty()) return false;
        return bvh_intersect_p(0, ray, cnt);
    }

    // Triangle::sample + sample_with_ref_point (triangle.rs:676-744)
    InteractionCommon tri_sample(const PbrtTri& tri, const InteractionCommon& iref, const Vec2& u, Float& pdf) const {
        const Mesh& mesh = meshes[tri.mesh];
        Point3 p0, p1, p2;
        tri_verts(tri, p0, p1, p2);
        Float su0 = std::sqrt(u.x);
        Float bx = 1.0f - su0, by = u.y * su0;
        InteractionCommon it;
        it.p = p0 * bx + p1 * by + p2 * (1.0f - bx - by);
        it.n = normalize(cross(p1 - p0, p2 - p0));
        if (!mesh.n.empty()) {
            Normal3 ns = mesh.N(tri.v[0]) * bx + mesh.N(tri.v[1]) * by + mesh.N(tri.v[2]) * (1.0f - bx - by);
            it.n = faceforward(it.n, ns);
        } else if (mesh.reverse_orientation ^ mesh.swaps_handedness) it.n = it.n * -1.0f;
        Point3 p_abs_sum = vabs(p0 * bx) + vabs(p1 * by) + vabs(p2 * (1.0f - bx - by));
        it.p_error = p_abs_sum * gamma(6);
        Float area = 0.5f * length(cross(p1 - p0, p2 - p0));
        pdf = 1.0f / area;
        it.time = 0.0f;
        // sample_with_ref_point
        Vec3 wi = it.p - iref.p;
        if (length_squared(wi) == 0.0f) pdf = 0.0f;
        else {
            wi = normalize(wi);
            pdf *= length_squared(iref.p - it.p) / abs_dot(it.n, -wi);
            if (std::isinf(pdf)) pdf = 0.0f;
        }
        return it;
    }
    Float tri_area(const PbrtTri& tri) const {  // triangle.rs:667-675
        Point3 p0, p1, p2;
        tri_verts(tri, p0, p1, p2);
        return 0.5f * length(cross(p1 - p0, p2 - p0));
    }
    Spectrum light_l(const AreaLight& l, const Normal3& n, const Vec3& w) const {  // diffuse.rs:164-170
        return (l.two_sided || dot(n, w) > 0.0f) ? l.l_emit : Spectrum(0.0f);
    }
    // DiffuseAreaLight::sample_li (diffuse.rs:64-84)
    Float world_radius() const {  // Bounds3f::bounding_sphere (geometry.rs:2079-2091), used by DistantLight::preprocess
        Point3 c = (world_bound.p_min + world_bound.p_max) / 2.0f;
        bool inside = c.x >= world_bound.p_min.x && c.x <= world_bound.p_max.x && c.y >= world_bound.p_min.y && c.y <= world_bound.p_max.y &&
                      c.z >= world_bound.p_min.z && c.z <= world_bound.p_max.z;
        return inside ? length(c - world_bound.p_max) : 0.0f;
    }
    static Vec3 rot(const Float* m, const Vec3& v) {  // Transform::transform_vector, upper 3x3
        return Vec3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
    }
    static Float spherical_theta(const Vec3& v) { return std::acos(clamp_t(v.z, -1.0f, 1.0f)); }  // geometry.rs:1584-1596
    static Float spherical_phi(const Vec3& v) {
        Float p = std::atan2(v.y, v.x);
        return p < 0.0f ? p + 2.0f * PI : p;
    }
    // Light::le of a ray that left the scene: non-zero for InfiniteAreaLight only (infinite.rs le)
    Spectrum light_le(const AreaLight& l, const Vec3& ray_d) const {
        if (l.kind != PBRT_LIGHT_INFINITE) return Spectrum();
        Vec3 w = normalize(rot(l.w2l, ray_d));
        return l.env->lmap.lookup(Vec2(spherical_phi(w) * INV_2_PI, spherical_theta(w) * INV_PI), 0.0f);
    }
    Float infinite_pdf_li(const AreaLight& l, const Vec3& w) const {  // infinite.rs pdf_li
        Vec3 wi = rot(l.w2l, w);
        Float theta = spherical_theta(wi), phi = spherical_phi(wi);
        Float sin_theta = std::sin(theta);
        if (sin_theta == 0.0f) return 0.0f;
        return l.env->distribution.pdf(Vec2(phi * INV_2_PI, theta * INV_PI)) / (2.0f * PI * PI * sin_theta);
    }
    Float spot_falloff(const AreaLight& l, const Vec3& w) const {  // spot.rs SpotLight::falloff
        Vec3 wl = normalize(Vec3(l.w2l[0] * w.x + l.w2l[1] * w.y + l.w2l[2] * w.z, l.w2l[3] * w.x + l.w2l[4] * w.y + l.w2l[5] * w.z,
                                 l.w2l[6] * w.x + l.w2l[7] * w.y + l.w2l[8] * w.z));
        Float cos_theta = wl.z;
        if (cos_theta < l.cos_total_width) return 0.0f;
        if (cos_theta >= l.cos_falloff_start) return 1.0f;
        Float delta = (cos_theta - l.cos_total_width) / (l.cos_falloff_start - l.cos_total_width);
        return (delta * delta) * (delta * delta);
    }
    Spectrum sample_li(const AreaLight& l, const InteractionCommon& iref, const Vec2& u, Vec3& wi, Float& pdf, InteractionCommon& light_intr) const {
        if (l.kind == PBRT_LIGHT_POINT || l.kind == PBRT_LIGHT_SPOT) {  // point.rs / spot.rs sample_li
            wi = normalize(l.p - iref.p);
            pdf = 1.0f;
            light_intr = InteractionCommon();
            light_intr.p = l.p;
            light_intr.time = iref.time;
            Float d2 = length_squared(l.p - iref.p);
            if (l.kind == PBRT_LIGHT_POINT) return l.l_emit / d2;
            return l.l_emit * spot_falloff(l, -wi) / d2;
        }
        if (l.kind == PBRT_LIGHT_DISTANT) {  // distant.rs sample_li
            wi = l.p;
            pdf = 1.0f;
            light_intr = InteractionCommon();
            light_intr.p = iref.p + l.p * (2.0f * world_radius());
            light_intr.time = iref.time;
            return l.l_emit;
        }
        if (l.kind == PBRT_LIGHT_INFINITE) {  // infinite.rs sample_li
            Float map_pdf = 0.0f;
            Vec2 uv = l.env->distribution.sample_continuous(u, map_pdf);
            if (map_pdf == 0.0f) return Spectrum();
            Float theta = uv.y * PI, phi = uv.x * 2.0f * PI;
            Float cos_theta = std::cos(theta), sin_theta = std::sin(theta);
            Float sin_phi = std::sin(phi), cos_phi = std::cos(phi);
            wi = rot(l.l2w, Vec3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
            pdf = map_pdf / (2.0f * PI * PI * sin_theta);
            if (sin_theta == 0.0f) pdf = 0.0f;
            light_intr = InteractionCommon();
            light_intr.p = iref.p + wi * (2.0f * world_radius());
            light_intr.time = iref.time;
            return l.env->lmap.lookup(uv, 0.0f);
        }
        light_intr = tri_sample(tris[l.tri], iref, u, pdf);
        if (pdf == 0.0f || length_squared(light_intr.p - iref.p) == 0.0f) { pdf = 0.0f; return Spectrum(); }
        wi = normalize(light_intr.p - iref.p);
        return light_l(l, light_intr.n, -wi);
    }
    // DiffuseAreaLight::pdf_li -> Triangle::pdf_with_ref_point (triangle.rs:745-764)
    Float pdf_li(const AreaLight& l, const SurfaceInteraction& iref, const Vec3& wi, Counters* cnt) const {
        if (l.kind == PBRT_LIGHT_INFINITE) return infinite_pdf_li(l, wi);
        Ray ray = spawn_ray(iref.common, wi);
        const PbrtTri& tri = tris[l.tri];
        Point3 p0, p1, p2;
        tri_verts(tri, p0, p1, p2);
        TriHit h;
        if (cnt) cnt->light_tri_tests++;
        if (!triangle_test(p0, p1, p2, ray, h)) return 0.0f;
        SurfaceInteraction isect_light;
        fill_interaction(tri, ray, h, isect_light);
        Float pdf = length_squared(iref.common.p - isect_light.common.p) / (abs_dot(isect_light.common.n, -wi) * tri_area(tri));
        if (std::isinf(pdf)) pdf = 0.0f;
        return pdf;
    }
};

// ---------------------------------------------------------------------------------------------
// Light distributions (lightdistrib.rs)
struct LightDistribution {
    int strategy;  // effective: PBRT_LIGHTS_UNIFORM / POWER / SPATIAL
    const Scene* scene;
    std::shared_ptr<Distribution1D> fixed;
    int n_voxels[3];
    std::vector<std::atomic<Distribution1D*>> voxels;  // dense stand-in for the hash table (a pure cache)

    LightDistribution(const Scene* sc, int requested) : scene(sc) {
        size_t nl = sc->lights.size();
        if (requested == PBRT_LIGHTS_UNIFORM || nl == 1) {  // lightdistrib.rs:397-400
            strategy = PBRT_LIGHTS_UNIFORM;
            fixed = std::make_shared<Distribution1D>(std::vector<Float>(nl, 1.0f));
        } else if (requested == PBRT_LIGHTS_POWER) {
            strategy = PBRT_LIGHTS_POWER;
            std::vector<Float> power;  // integrator.rs:574-584, diffuse.rs:85-93
            for (const AreaLight& l : sc->lights) {
                Spectrum pw;
                if (l.kind == PBRT_LIGHT_POINT) pw = l.l_emit * (4.0f * PI);                                     // point.rs power
                else if (l.kind == PBRT_LIGHT_SPOT) pw = l.l_emit * 2.0f * PI * (1.0f - 0.5f * (l.cos_falloff_start + l.cos_total_width));  // spot.rs
                else if (l.kind == PBRT_LIGHT_DISTANT) { Float r = sc->world_radius(); pw = l.l_emit * PI * r * r; }  // distant.rs
                else if (l.kind == PBRT_LIGHT_INFINITE) { Float r = sc->world_radius(); pw = l.env->lmap.lookup(Vec2(0.5f, 0.5f), 0.5f) * Spectrum(PI * r * r); }  // infinite.rs
                else pw = l.l_emit * (l.two_sided ? 2.0f : 1.0f) * l.area * PI;
                power.push_back(pw.y());
            }
            fixed = std::make_shared<Distribution1D>(power);
        } else {
            strategy = PBRT_LIGHTS_SPATIAL;
            const Bounds3& b = sc->world_bound;  // lightdistrib.rs:127-150, max_voxels = 64
            Vec3 diag = b.diagonal();
            Float bmax = diag[b.maximum_extent()];
            size_t total = 1;
            for (int i = 0; i < 3; ++i) {
                n_voxels[i] = std::max(1, f2i(std::round(diag[i] / bmax * 64.0f)));
                total *= (size_t)n_voxels[i];
            }
            voxels = std::vector<std::atomic<Distribution1D*>>(total);
            for (auto& v : voxels) v.store(nullptr);
        }
    }
    ~LightDistribution() { for (auto& v : voxels) delete v.load(); }

    Distribution1D compute_distribution(const int pi[3]) const {  // lightdistrib.rs:169-269
        Point3 p0((Float)pi[0] / (Float)n_voxels[0], (Float)pi[1] / (Float)n_voxels[1], (Float)pi[2] / (Float)n_voxels[2]);
        Point3 p1((Float)(pi[0] + 1) / (Float)n_voxels[0], (Float)(pi[1] + 1) / (Float)n_voxels[1], (Float)(pi[2] + 1) / (Float)n_voxels[2]);
        Bounds3 vb(scene->world_bound.lerp(p0), scene->world_bound.lerp(p1));
        const size_t n_samples = 128, nl = scene->lights.size();
        std::vector<Float> contrib(nl, 0.0f);
        for (size_t i = 0; i < n_samples; ++i) {
            Point3 po = vb.lerp(Point3(radical_inverse(0, i), radical_inverse(1, i), radical_inverse(2, i)));
            InteractionCommon intr;
            intr.p = po;
            intr.wo = Vec3(1.0f, 0.0f, 0.0f);
            Vec2 u(radical_inverse(3, i), radical_inverse(4, i));
            for (size_t j = 0; j < nl; ++j) {
                Float pdf = 0.0f;
                Vec3 wi;
                InteractionCommon li_intr;
                Spectrum li = scene->sample_li(scene->lights[j], intr, u, wi, pdf, li_intr);
                if (pdf > 0.0f) contrib[j] += li.y() / pdf;
            }
        }
        Float sum = 0.0f;
        for (Float c : contrib) sum += c;
        Float avg = sum / (Float)(n_samples * nl);
        Float min_contrib = (avg > 0.0f) ? 0.001f * avg : 1.0f;
        for (Float& c : contrib) c = fmax_(c, min_contrib);
        return Distribution1D(contrib);
    }
    void voxel_of(const Point3& p, int pi[3]) const {  // lightdistrib.rs:282-294
        Vec3 off = scene->world_bound.offset(p);
        for (int i = 0; i < 3; ++i) pi[i] = clamp_t(f2i(off[i] * (Float)n_voxels[i]), 0, n_voxels[i] - 1);
    }
    const Distribution1D* lookup(const Point3& p) {
        if (strategy != PBRT_LIGHTS_SPATIAL) return fixed.get();
        int pi[3];
        voxel_of(p, pi);
        size_t idx = ((size_t)pi[2] * n_voxels[1] + pi[1]) * n_voxels[0] + pi[0];
        Distribution1D* d = voxels[idx].load(std::memory_order_acquire);
        if (d) return d;
        Distribution1D* fresh = new Distribution1D(compute_distribution(pi));
        Distribution1D* expected = nullptr;
        if (voxels[idx].compare_exchange_strong(expected, fresh, std::memory_order_acq_rel)) return fresh;
        delete fresh;  // another thread computed the (identical, deterministic) distribution first
        return expected;
    }
};

// sampling.rs:229-233
inline Float power_heuristic(int nf, Float f_pdf, int ng, Float g_pdf) {
    Float f = (Float)nf * f_pdf, g = (Float)ng * g_pdf;
    return (f * f) / (f * f + g * g);
}

struct ShadeCtx {
    const Scene* scene;
    Sampler* sampler;
    LightDistribution* light_distrib;
    Counters* cnt;
};

// pbrt.rs solve_linear_system_2x2 (transform.rs:219-235)
inline bool solve_linear_system_2x2(const Float a[2][2], const Float b[2], Float& x0, Float& x1) {
    Float det = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    if (std::fabs(det) < 1e-10f) return false;
    x0 = (a[1][1] * b[0] - a[0][1] * b[1]) / det;
    x1 = (a[0][0] * b[1] - a[1][0] * b[0]) / det;
    if (std::isnan(x0) || std::isnan(x1)) return false;
    return true;
}
// SurfaceInteraction::compute_differentials (interaction.rs:388-474)
inline void compute_differentials(SurfaceInteraction& si, const Ray& ray) {
    si.dudx = si.dvdx = si.dudy = si.dvdy = 0.0f;
    si.dpdx = si.dpdy = Vec3();
    if (!ray.has_differential) return;
    const Normal3& n = si.common.n;
    const Point3& p = si.common.p;
    Float d = dot(n, Vec3(p.x, p.y, p.z));
    Float tx = -(dot(n, ray.rx_origin) - d) / dot(n, ray.rx_direction);
    if (std::isinf(tx) || std::isnan(tx)) return;
    Point3 px = ray.rx_origin + ray.rx_direction * tx;
    Float ty = -(dot(n, ray.ry_origin) - d) / dot(n, ray.ry_direction);
    if (std::isinf(ty) || std::isnan(ty)) return;
    Point3 py = ray.ry_origin + ray.ry_direction * ty;
    si.dpdx = px - p;
    si.dpdy = py - p;
    int dim[2];
    if (std::fabs(n.x) > std::fabs(n.y) && std::fabs(n.x) > std::fabs(n.z)) { dim[0] = 1; dim[1] = 2; }
    else if (std::fabs(n.y) > std::fabs(n.z)) { dim[0] = 0; dim[1] = 2; }
    else { dim[0] = 0; dim[1] = 1; }
    auto comp = [](const Vec3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); };
    const Float a[2][2] = {{comp(si.dpdu, dim[0]), comp(si.dpdv, dim[0])}, {comp(si.dpdu, dim[1]), comp(si.dpdv, dim[1])}};
    const Float bx[2] = {comp(px, dim[0]) - comp(p, dim[0]), comp(px, dim[1]) - comp(p, dim[1])};
    const Float by[2] = {comp(py, dim[0]) - comp(p, dim[0]), comp(py, dim[1]) - comp(p, dim[1])};
    if (!solve_linear_system_2x2(a, bx, si.dudx, si.dvdx)) { si.dudx = 0.0f; si.dvdx = 0.0f; }
    if (!solve_linear_system_2x2(a, by, si.dudy, si.dvdy)) { si.dudy = 0.0f; si.dvdy = 0.0f; }
}
// ImageTexture<Spectrum>::evaluate (imagemap.rs:133-148) through UVMapping2D::map (texture.rs:101-121); convert_out is the identity on RGB
inline Spectrum texture_evaluate(const std::vector<std::unique_ptr<ImageTexture>>& all, const ImageTexture& t, const SurfaceInteraction& si) {
    switch (t.kind) {
        case PBRT_TEX_CONSTANT: return t.value;  // constant.rs:17-20
        case PBRT_TEX_SCALE:    // scale.rs: tex1 * tex2
            return texture_evaluate(all, *all[t.child[0] - 1], si) * texture_evaluate(all, *all[t.child[1] - 1], si);
        case PBRT_TEX_MIX: {    // mix.rs: t1 * (1 - amt) + t2 * amt
            Spectrum t1 = texture_evaluate(all, *all[t.child[0] - 1], si), t2 = texture_evaluate(all, *all[t.child[1] - 1], si);
            Float amt = texture_evaluate(all, *all[t.child[2] - 1], si).c[0];
            return t1 * Spectrum(1.0f - amt) + t2 * Spectrum(amt);
        }
        default: break;
    }
    Vec2 dstdx, dstdy, st;
    auto fix = [](Float& d) { if (d > 0.5f) d = 1.0f - d; else if (d < -0.5f) d = -(d + 1.0f); };
    if (t.mapping == PBRT_MAP_SPHERICAL) {  // SphericalMapping2D::map (texture.rs:136-172)
        auto sphere = [&](const Point3& p) {
            Vec3 v = normalize(xf_point(t.map_m, p) - Point3(0, 0, 0));
            return Vec2(Scene::spherical_theta(v) * INV_PI, Scene::spherical_phi(v) * INV_2_PI);
        };
        st = sphere(si.common.p);
        const Float delta = 0.1f, inv = 1.0f / delta;  // Vector2f / Float multiplies by the reciprocal (geometry.rs:1281-1288)
        Vec2 sx = sphere(si.common.p + si.dpdx * delta), sy = sphere(si.common.p + si.dpdy * delta);
        dstdx = Vec2((sx.x - st.x) * inv, (sx.y - st.y) * inv);
        dstdy = Vec2((sy.x - st.x) * inv, (sy.y - st.y) * inv);
        fix(dstdx.y); fix(dstdy.y);
    } else if (t.mapping == PBRT_MAP_CYLINDRICAL) {  // CylindricalMapping2D::map (texture.rs:174-215)
        auto cylinder = [&](const Point3& p) {
            Vec3 v = normalize(xf_point(t.map_m, p) - Point3(0, 0, 0));
            return Vec2(PI + std::atan2(v.y, v.x) * INV_2_PI, v.z);
        };
        st = cylinder(si.common.p);
        const Float delta = 0.01f, inv = 1.0f / delta;
        Vec2 sx = cylinder(si.common.p + si.dpdx * delta);
        dstdx = Vec2((sx.x - st.x) * inv, (sx.y - st.y) * inv);
        fix(dstdx.y);
        Vec2 sy = cylinder(si.common.p + si.dpdy * delta);
        dstdy = Vec2((sy.x - st.x) * inv, (sy.y - st.y) * inv);
        fix(dstdy.y);
    } else if (t.mapping == PBRT_MAP_PLANAR) {  // PlanarMapping2D::map (texture.rs:226-252)
        const Vec3 vs(t.map_m[0], t.map_m[1], t.map_m[2]), vt(t.map_m[3], t.map_m[4], t.map_m[5]);
        const Vec3 vec(si.common.p.x, si.common.p.y, si.common.p.z);
        dstdx = Vec2(dot(si.dpdx, vs), dot(si.dpdx, vt));
        dstdy = Vec2(dot(si.dpdy, vs), dot(si.dpdy, vt));
        st = Vec2(t.du + dot(vec, vs), t.dv + dot(vec, vt));
    } else {
        dstdx = Vec2(si.dudx * t.su, si.dvdx * t.sv);
        dstdy = Vec2(si.dudy * t.su, si.dvdy * t.sv);
        st = Vec2(si.uv.x * t.su + t.du, si.uv.y * t.sv + t.dv);
    }
    return t.mipmap.lookup(st, dstdx, dstdy);
}
inline bool material_textured(const PbrtMaterial& m) {
    for (int g = 0; g < PBRT_MAX_TEX_GROUPS; ++g) if (m.tex[g]) return true;
    return m.bump != 0;
}
// Material::bump (material.rs:116-219) followed by SurfaceInteraction::set_shading_geometry (interaction.rs:345-370)
inline void material_bump(const Scene& sc, const ImageTexture& d, SurfaceInteraction& si);

// SurfaceInteraction::compute_scattering_functions (interaction.rs:371-387: compute_differentials, then the material's) followed by
// Bsdf::new (reflection.rs:235-245).  `local` receives the lobes of a material with image textures (they depend on the hit).
inline Bsdf make_bsdf(const Scene& sc, SurfaceInteraction& si, const Ray& ray, MaterialLobes& local) {
    const uint32_t mi = sc.tris[si.prim].material;
    const MaterialLobes* mlp = sc.allow_multiple_lobes ? &sc.materials[mi] : &sc.materials_single[mi];
    if (!sc.textures.empty()) compute_differentials(si, ray);  // (unconditional in the reference; only textures and the specular rays of direct / whitted read it)
    if (material_textured(sc.material_src[mi])) {
        PbrtMaterial m = sc.material_src[mi];
        if (m.bump) material_bump(sc, *sc.textures[m.bump - 1], si);
        for (int g = 0; g < PBRT_MAX_TEX_GROUPS; ++g)
            if (m.tex[g]) {
                Spectrum v = texture_evaluate(sc.textures, *sc.textures[m.tex[g] - 1], si);
                int nv = 0;
                const int o = pbrt_material_tex_offset(m.kind, g, &nv);
                if (o < 0) continue;  // rejected at scene creation
                m.params[o] = v.c[0];
                if (nv == 3) { m.params[o + 1] = v.c[1]; m.params[o + 2] = v.c[2]; }
            }
        compile_material(m, local, sc.allow_multiple_lobes);
        mlp = &local;
    }
    const MaterialLobes& ml = *mlp;
    Bsdf b;
    b.eta = ml.eta;
    b.ns = si.shading_n;
    b.ng = si.common.n;
    b.ss = normalize(si.shading_dpdu);
    b.ts = cross(si.shading_n, b.ss);
    b.bxdfs = &ml.bxdfs;
    return b;
}
inline bool Scene::alpha_rejects(const PbrtTri& tri, const Point3& p0, const Point3& p1, const Point3& p2, const TriHit& h, bool any_hit) const {
    const Mesh& mesh = meshes[tri.mesh];
    if (!mesh.alpha && !(any_hit && mesh.shadow_alpha)) return false;
    Vec2 uv[3];
    get_uvs(tri, uv);
    if (any_hit) {  // triangle.rs:594-627
        Vec2 duv02(uv[0].x - uv[2].x, uv[0].y - uv[2].y), duv12(uv[1].x - uv[2].x, uv[1].y - uv[2].y);
        Vec3 dp02 = p0 - p2, dp12 = p1 - p2;
        Float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
        bool degenerate_uv = std::fabs(determinant) < 1e-8f;
        Vec3 dpdu, dpdv;
        if (!degenerate_uv) {
            Float invdet = 1.0f / determinant;
            dpdu = (dp02 * duv12.y - dp12 * duv02.y) * invdet;
            dpdv = (dp02 * -duv12.x + dp12 * duv02.x) * invdet;
        }
        if (degenerate_uv || length_squared(cross(dpdu, dpdv)) == 0.0f) {
            if (length_squared(cross(p2 - p0, p1 - p0)) == 0.0f) return true;  // "the intersection is bogus"
        }
    }
    SurfaceInteraction local;
    local.common.p = p0 * h.b0 + p1 * h.b1 + p2 * h.b2;
    local.uv = Vec2(uv[0].x * h.b0 + uv[1].x * h.b1 + uv[2].x * h.b2, uv[0].y * h.b0 + uv[1].y * h.b1 + uv[2].y * h.b2);
    if (mesh.alpha && texture_evaluate(textures, *textures[mesh.alpha - 1], local).c[0] == 0.0f) return true;
    if (any_hit && mesh.shadow_alpha && texture_evaluate(textures, *textures[mesh.shadow_alpha - 1], local).c[0] == 0.0f) return true;
    return false;
}
inline Spectrum isect_le(const Scene& sc, const SurfaceInteraction& si, const Vec3& w) {  // interaction.rs:475-483
    if (si.primitive_lost) return Spectrum();  // no primitive => no area light (interaction.rs:475-483)
    int32_t al = sc.tris[si.prim].area_light;
    if (al < 0) return Spectrum();
    return sc.light_l(sc.lights[al], si.common.n, w);
}

inline std::unique_ptr<Sampler> make_sampler(const PbrtRenderParams& rp) {  // api.rs make_sampler for the samplers in scope
    if (rp.sampler == PBRT_SAMPLER_HALTON) return std::unique_ptr<Sampler>(new HaltonSampler((int64_t)rp.spp, rp.sample_bounds, rp.sample_at_pixel_center != 0));
    return std::unique_ptr<Sampler>(new SobolSampler((int64_t)rp.spp, rp.sample_bounds));
}

// integrator.rs:406-570, handle_media = false, specular = false
inline Spectrum estimate_direct(ShadeCtx& cx, const SurfaceInteraction& it, const Bsdf& bsdf, const Vec2& u_scattering, int light_num, const Vec2& u_light) {
    const Scene& sc = *cx.scene;
    const AreaLight& light = sc.lights[light_num];
    const int bsdf_flags = BSDF_ALL & ~BSDF_SPECULAR;
    Spectrum ld(0.0f);
    Vec3 wi;
    Float light_pdf = 0.0f, scattering_pdf = 0.0f;
    InteractionCommon light_intr;
    Spectrum li = sc.sample_li(light, it.common, u_light, wi, light_pdf, light_intr);
    if (light_pdf > 0.0f && !li.is_black()) {
        Spectrum f = bsdf.f(it.common.wo, wi, bsdf_flags) * Spectrum(abs_dot(wi, it.shading_n));
        scattering_pdf = bsdf.pdf(it.common.wo, wi, bsdf_flags);
        if (!f.is_black()) {
            Ray sray = spawn_ray_to(it.common, light_intr);  // VisibilityTester::unoccluded light.rs:199-206
            if (sc.intersect_p(sray, cx.cnt)) li = Spectrum(0.0f);
            if (!li.is_black()) {
                if (light.is_delta()) ld += f * li / light_pdf;
                else {
                    Float weight = power_heuristic(1, light_pdf, 1, scattering_pdf);
                    ld += f * li * Spectrum(weight) / light_pdf;
                }
            }
        }
    }
    if (!light.is_delta()) {  // sample BSDF with multiple importance sampling
        int sampled_type = 0;  // quirk Q8
        Spectrum f = bsdf.sample_f(it.common.wo, wi, u_scattering, scattering_pdf, bsdf_flags, sampled_type);
        f *= Spectrum(abs_dot(wi, it.shading_n));
        bool sampled_specular = (sampled_type & BSDF_SPECULAR) != 0;
        if (!f.is_black() && scattering_pdf > 0.0f) {
            Float weight = 1.0f;
            if (!sampled_specular) {
                light_pdf = sc.pdf_li(light, it, wi, cx.cnt);
                if (light_pdf == 0.0f) return ld;
                weight = power_heuristic(1, scattering_pdf, 1, light_pdf);
            }
            Ray ray = spawn_ray(it.common, wi);
            Spectrum tr(1.0f);
            Spectrum li2;
            SurfaceInteraction light_isect;
            if (sc.intersect(ray, light_isect, cx.cnt)) {
                if (!light_isect.primitive_lost && sc.tris[light_isect.prim].area_light == light_num) li2 = isect_le(sc, light_isect, -wi);
            } else li2 = sc.light_le(light, ray.d);  // zero unless the light is infinite
            if (!li2.is_black()) ld += f * li2 * tr * weight / scattering_pdf;
        }
    }
    return ld;
}

// integrator.rs:359-403 with a light distribution
inline Spectrum uniform_sample_one_light(ShadeCtx& cx, const SurfaceInteraction& it, const Bsdf& bsdf, const Distribution1D* distrib) {
    size_t n_lights = cx.scene->lights.size();
    if (n_lights == 0) return Spectrum();
    Float pdf = 0.0f;
    size_t light_num = distrib->sample_discrete(cx.sampler->get_1d(), pdf);
    if (pdf == 0.0f) return Spectrum();
    Vec2 u_light = cx.sampler->get_2d();
    Vec2 u_scattering = cx.sampler->get_2d();
    return estimate_direct(cx, it, bsdf, u_scattering, (int)light_num, u_light) / pdf;
}

// PathIntegrator::li, integrators/path.rs:59-282 (no BSSRDF)
inline Spectrum path_li(ShadeCtx& cx, const Ray& r, uint32_t max_depth, Float rr_threshold) {
    const Scene& sc = *cx.scene;
    Spectrum l, beta(1.0f);
    Ray ray = r;
    bool specular_bounce = false;
    uint32_t bounces = 0;
    Float eta_scale = 1.0f;
    for (;;) {
        SurfaceInteraction isect;
        if (sc.intersect(ray, isect, cx.cnt)) {
            if (bounces == 0 || specular_bounce) l += beta * isect_le(sc, isect, -ray.d);
            if (bounces >= max_depth) break;
            if (isect.primitive_lost || sc.tris[isect.prim].material == PBRT_NO_MATERIAL) {  // null BSDF: path.rs:109-116
                ray = spawn_ray(isect.common, ray.d);
                continue;
            }
            MaterialLobes local_lobes;
            Bsdf bsdf = make_bsdf(sc, isect, ray, local_lobes);
            const Distribution1D* distrib = cx.light_distrib->lookup(isect.common.p);
            if (bsdf.num_components(BSDF_ALL & ~BSDF_SPECULAR) > 0) {
                Spectrum ld = beta * uniform_sample_one_light(cx, isect, bsdf, distrib);
                l += ld;
            }
            Vec3 wo = -ray.d, wi;
            Float pdf = 0.0f;
            int sampled_type = 255;
            Spectrum f = bsdf.sample_f(wo, wi, cx.sampler->get_2d(), pdf, BSDF_ALL, sampled_type);
            if (f.is_black() || pdf == 0.0f) break;
            beta *= (f * abs_dot(wi, isect.shading_n)) / pdf;
            specular_bounce = (sampled_type & BSDF_SPECULAR) != 0;
            if ((sampled_type & BSDF_SPECULAR) && (sampled_type & BSDF_TRANSMISSION)) {
                Float eta = bsdf.eta;
                if (dot(wo, isect.common.n) > 0.0f) eta_scale *= eta * eta;
                else eta_scale *= 1.0f / (eta * eta);
            }
            ray = spawn_ray(isect.common, wi);
            Spectrum rr_beta = beta * eta_scale;
            if (rr_beta.max_component_value() < rr_threshold && bounces > 3) {
                Float q = fmax_(0.05f, 1.0f - rr_beta.max_component_value());
                if (cx.sampler->get_1d() < q) break;
                beta = beta / (1.0f - q);
            }
        } else {
            if (bounces == 0 || specular_bounce)  // environment emission, path.rs:267-275
                for (const AreaLight& light : sc.lights)
                    if (light.kind == PBRT_LIGHT_INFINITE) l += beta * sc.light_le(light, ray.d);
            break;
        }
        bounces += 1;
    }
    return l;
}

inline void material_bump(const Scene& sc, const ImageTexture& d, SurfaceInteraction& si) {
    SurfaceInteraction si_eval = si;
    Float du = 0.5f * (std::fabs(si.dudx) + std::fabs(si.dudy));
    if (du == 0.0f) du = 0.0005f;
    si_eval.common.p = si.common.p + si.shading_dpdu * du;
    si_eval.uv = Vec2(si.uv.x + du, si.uv.y + 0.0f);
    // (si_eval.n is set too, from dndu; nothing a UV-mapped image reads)
    const Float u_displace = texture_evaluate(sc.textures, d, si_eval).c[0];
    Float dv = 0.5f * (std::fabs(si.dvdx) + std::fabs(si.dvdy));
    if (dv == 0.0f) dv = 0.0005f;
    si_eval.common.p = si.common.p + si.shading_dpdv * dv;
    si_eval.uv = Vec2(si.uv.x + 0.0f, si.uv.y + dv);
    const Float v_displace = texture_evaluate(sc.textures, d, si_eval).c[0];
    const Float displace = texture_evaluate(sc.textures, d, si).c[0];
    const Vec3 dpdu = si.shading_dpdu + si.shading_n * ((u_displace - displace) / du) + si.shading_dndu * displace;
    const Vec3 dpdv = si.shading_dpdv + si.shading_n * ((v_displace - displace) / dv) + si.shading_dndv * displace;
    // set_shading_geometry(dpdu, dpdv, dndu, dndv, false)
    si.shading_n = normalize(cross(dpdu, dpdv));
    if (si.shape_flips) si.shading_n = -si.shading_n;
    si.shading_n = faceforward(si.shading_n, si.common.n);
    si.shading_dpdu = dpdu;
    si.shading_dpdv = dpdv;
}

// ---- DirectLightingIntegrator (integrators/directlighting.rs) and WhittedIntegrator (integrators/whitted.rs) ----
struct DirectCfg {
    bool whitted = false;
    bool sample_all = true;
    uint32_t max_depth = 5;
    std::vector<int32_t> n_light_samples;  // DirectLightingIntegrator::preprocess (directlighting.rs:52-59)
};
// uniform_sample_all_lights (integrator.rs:300-355)
inline Spectrum uniform_sample_all_lights(ShadeCtx& cx, const SurfaceInteraction& it, const Bsdf& bsdf, const std::vector<int32_t>& n_light_samples) {
    const Scene& sc = *cx.scene;
    Sampler& sampler = *cx.sampler;
    Spectrum l;
    for (size_t j = 0; j < n_light_samples.size() && j < sc.lights.size(); ++j) {
        const int32_t n_samples = n_light_samples[j];
        size_t li_idx = 0, li_start = 0, sc_idx = 0, sc_start = 0;
        const bool have_light = sampler.get_2d_array_idxs(n_samples, li_idx, li_start);
        const bool have_scat = sampler.get_2d_array_idxs(n_samples, sc_idx, sc_start);
        if (!have_light || !have_scat) {  // fall back to a single sample
            Vec2 u_light = sampler.get_2d();
            Vec2 u_scattering = sampler.get_2d();
            l += estimate_direct(cx, it, bsdf, u_scattering, (int)j, u_light);
        } else {
            Spectrum ld;
            for (int32_t k = 0; k < n_samples; ++k) {
                const Vec2 u_scattering = sampler.sample_array_2d[sc_idx][sc_start + (size_t)k];
                const Vec2 u_light = sampler.sample_array_2d[li_idx][li_start + (size_t)k];
                ld += estimate_direct(cx, it, bsdf, u_scattering, (int)j, u_light);
            }
            l += ld / (Float)n_samples;
        }
    }
    return l;
}
// uniform_sample_one_light without a light distribution (integrator.rs:383-403)
inline Spectrum uniform_sample_one_light_uniform(ShadeCtx& cx, const SurfaceInteraction& it, const Bsdf& bsdf) {
    const size_t n_lights = cx.scene->lights.size();
    if (n_lights == 0) return Spectrum();
    const size_t light_num = std::min((size_t)f2i(cx.sampler->get_1d() * (Float)n_lights), n_lights - 1);
    const Float pdf = 1.0f / (Float)n_lights;
    Vec2 u_light = cx.sampler->get_2d();
    Vec2 u_scattering = cx.sampler->get_2d();
    return estimate_direct(cx, it, bsdf, u_scattering, (int)light_num, u_light) / pdf;
}
inline Spectrum direct_li(ShadeCtx& cx, const DirectCfg& cfg, const Ray& ray, int depth);
// specular_reflect / specular_transmit (directlighting.rs:124-260, whitted.rs:125-254), including the child ray's differential
// (it feeds the texture filtering of what the mirror or the glass shows).
inline Spectrum specular_bounce(ShadeCtx& cx, const DirectCfg& cfg, const Ray& ray, const SurfaceInteraction& isect, const Bsdf& bsdf, bool transmit, int depth) {
    const Vec3 wo = isect.common.wo;
    Vec3 wi;
    Float pdf = 0.0f;
    const Normal3 ns = isect.shading_n;
    int sampled_type = 0;
    const int flags = (transmit ? BSDF_TRANSMISSION : BSDF_REFLECTION) | BSDF_SPECULAR;
    Spectrum f = bsdf.sample_f(wo, wi, cx.sampler->get_2d(), pdf, flags, sampled_type);
    if (pdf > 0.0f && !f.is_black() && abs_dot(wi, ns) != 0.0f) {
        Ray rd = spawn_ray(isect.common, wi);
        if (ray.has_differential) {
            const Normal3 dndx = isect.shading_dndu * isect.dudx + isect.shading_dndv * isect.dvdx;
            const Normal3 dndy = isect.shading_dndu * isect.dudy + isect.shading_dndv * isect.dvdy;
            const Vec3 dwodx = -ray.rx_direction - wo, dwody = -ray.ry_direction - wo;
            const Float ddndx = dot(dwodx, ns) + dot(wo, dndx), ddndy = dot(dwody, ns) + dot(wo, dndy);
            rd.has_differential = true;
            rd.rx_origin = isect.common.p + isect.dpdx;
            rd.ry_origin = isect.common.p + isect.dpdy;
            if (!transmit) {  // directlighting.rs:148-172
                rd.rx_direction = wi - dwodx + (dndx * dot(wo, ns) + ns * ddndx) * 2.0f;
                rd.ry_direction = wi - dwody + (dndy * dot(wo, ns) + ns * ddndy) * 2.0f;
            } else {          // directlighting.rs:219-249
                Float eta = bsdf.eta;
                const Vec3 w = -wo;
                if (dot(wo, ns) < 0.0f) eta = 1.0f / eta;
                const Float mu = eta * dot(w, ns) - dot(wi, ns);
                const Float dmudx = (eta - (eta * eta * dot(w, ns)) / dot(wi, ns)) * ddndx;
                const Float dmudy = (eta - (eta * eta * dot(w, ns)) / dot(wi, ns)) * ddndy;
                rd.rx_direction = wi + dwodx * eta - (dndx * mu + ns * dmudx);
                rd.ry_direction = wi + dwody * eta - (dndy * mu + ns * dmudy);
            }
        }
        return f * direct_li(cx, cfg, rd, depth + 1) * Spectrum(abs_dot(wi, ns) / pdf);
    }
    return Spectrum(0.0f);
}
// DirectLightingIntegrator::li (directlighting.rs:70-123) / WhittedIntegrator::li (whitted.rs:43-124)
inline Spectrum direct_li(ShadeCtx& cx, const DirectCfg& cfg, const Ray& ray, int depth) {
    const Scene& sc = *cx.scene;
    Spectrum l;
    SurfaceInteraction isect;
    if (sc.intersect(ray, isect, cx.cnt)) {
        if (isect.primitive_lost || sc.tris[isect.prim].material == PBRT_NO_MATERIAL)  // no BSDF: continue through, same depth
            return direct_li(cx, cfg, spawn_ray(isect.common, ray.d), depth);
        const Normal3 n = isect.shading_n;
        const Vec3 wo = isect.common.wo;
        MaterialLobes local_lobes;
        Bsdf bsdf = make_bsdf(sc, isect, ray, local_lobes);
        l += isect_le(sc, isect, wo);
        if (cfg.whitted) {
            for (size_t j = 0; j < sc.lights.size(); ++j) {  // whitted.rs:74-98
                const AreaLight& light = sc.lights[j];
                Vec3 wi;
                Float pdf = 0.0f;
                InteractionCommon light_intr;
                Spectrum li = sc.sample_li(light, isect.common, cx.sampler->get_2d(), wi, pdf, light_intr);
                if (li.is_black() || pdf == 0.0f) continue;
                Spectrum f = bsdf.f(wo, wi, BSDF_ALL);
                if (!f.is_black() && !sc.intersect_p(spawn_ray_to(isect.common, light_intr), cx.cnt)) l += f * li * abs_dot(wi, n) / pdf;
            }
        } else if (!sc.lights.empty()) {
            if (cfg.sample_all) l += uniform_sample_all_lights(cx, isect, bsdf, cfg.n_light_samples);
            else l += uniform_sample_one_light_uniform(cx, isect, bsdf);
        }
        if ((uint32_t)(depth + 1) < cfg.max_depth) {
            l += specular_bounce(cx, cfg, ray, isect, bsdf, false, depth);
            l += specular_bounce(cx, cfg, ray, isect, bsdf, true, depth);
        }
    } else {
        for (const AreaLight& light : sc.lights) l += sc.light_le(light, ray.d);  // Light::le(ray): zero unless infinite
    }
    return l;
}

// PerspectiveCamera::generate_ray_differential (perspective.rs:190-280); the differentials feed texture filtering only.
inline Ray camera_ray(const PbrtCamera& cam, const Vec2& p_film, Float time, const Vec2& p_lens) {
    Point3 p_camera = xf_point(cam.raster_to_camera, Point3(p_film.x, p_film.y, 0.0f));
    // dx_camera / dy_camera, PerspectiveCamera::new (perspective.rs:82-99)
    const Point3 r0 = xf_point(cam.raster_to_camera, Point3(0.0f, 0.0f, 0.0f));
    const Vec3 dx_camera = xf_point(cam.raster_to_camera, Point3(1.0f, 0.0f, 0.0f)) - r0;
    const Vec3 dy_camera = xf_point(cam.raster_to_camera, Point3(0.0f, 1.0f, 0.0f)) - r0;
    Ray in_ray(Point3(0, 0, 0), normalize(p_camera), INF, lerp(time, cam.shutter_open, cam.shutter_close));
    in_ray.has_differential = true;
    in_ray.rx_origin = in_ray.ry_origin = Point3(0, 0, 0);
    in_ray.rx_direction = normalize(Vec3(p_camera.x, p_camera.y, p_camera.z) + dx_camera);
    in_ray.ry_direction = normalize(Vec3(p_camera.x, p_camera.y, p_camera.z) + dy_camera);
    if (cam.lens_radius > 0.0f) {  // perspective.rs:246-271
        Vec2 pl = concentric_sample_disk(p_lens);
        pl = Vec2(pl.x * cam.lens_radius, pl.y * cam.lens_radius);
        Vec3 dx = normalize(Vec3(p_camera.x, p_camera.y, p_camera.z) + dx_camera);
        Float ftx = cam.focal_distance / dx.z;
        Point3 pfx = Point3(0, 0, 0) + dx * ftx;
        in_ray.rx_origin = Point3(pl.x, pl.y, 0.0f);
        in_ray.rx_direction = normalize(pfx - in_ray.rx_origin);
        Vec3 dy = normalize(Vec3(p_camera.x, p_camera.y, p_camera.z) + dy_camera);
        Float fty = cam.focal_distance / dy.z;
        Point3 pfy = Point3(0, 0, 0) + dy * fty;
        in_ray.ry_origin = Point3(pl.x, pl.y, 0.0f);
        in_ray.ry_direction = normalize(pfy - in_ray.ry_origin);
    }
    if (cam.lens_radius > 0.0f) {
        Vec2 pl = concentric_sample_disk(p_lens);
        pl = Vec2(pl.x * cam.lens_radius, pl.y * cam.lens_radius);
        Float ft = cam.focal_distance / in_ray.d.z;
        Point3 p_focus = in_ray.o + in_ray.d * ft;
        in_ray.o = Point3(pl.x, pl.y, 0.0f);
        in_ray.d = normalize(p_focus - in_ray.o);
    }
    return xf_ray(cam.camera_to_world, in_ray);
}

// FilmTile::add_sample (film.rs:94-147) applied directly to the cropped film (merge is a plain +=)
struct Film {
    int32_t bounds[4];
    std::vector<Float> rgbw;  // area * 4
    void init(const int32_t cropped[4]) {
        for (int i = 0; i < 4; ++i) bounds[i] = cropped[i];
        size_t w = (size_t)std::max(0, bounds[2] - bounds[0]), h = (size_t)std::max(0, bounds[3] - bounds[1]);
        rgbw.assign(w * h * 4, 0.0f);
    }
};
inline void film_add_sample(const PbrtRenderParams& rp, Float* rgbw, const Vec2& p_film, Spectrum l, Float sample_weight) {
    if (l.y() > rp.max_sample_luminance) l *= Spectrum(rp.max_sample_luminance / l.y());
    Vec2 pd(p_film.x - 0.5f, p_film.y - 0.5f);
    int32_t p0x = f2i(std::ceil(pd.x - rp.filter_radius[0])), p0y = f2i(std::ceil(pd.y - rp.filter_radius[1]));
    int32_t p1x = f2i(std::floor(pd.x + rp.filter_radius[0])) + 1, p1y = f2i(std::floor(pd.y + rp.filter_radius[1])) + 1;
    const int32_t* cb = rp.cropped_pixel_bounds;
    p0x = std::max(p0x, cb[0]); p0y = std::max(p0y, cb[1]);
    p1x = std::min(p1x, cb[2]); p1y = std::min(p1y, cb[3]);
    Float inv_rx = 1.0f / rp.filter_radius[0], inv_ry = 1.0f / rp.filter_radius[1];
    const Float ts = 16.0f;
    int32_t width = cb[2] - cb[0];
    for (int32_t y = p0y; y < p1y; ++y) {
        Float fy = std::fabs(((Float)y - pd.y) * inv_ry * ts);
        int iy = f2i(fmin_(std::floor(fy), ts - 1.0f));
        for (int32_t x = p0x; x < p1x; ++x) {
            Float fx = std::fabs(((Float)x - pd.x) * inv_rx * ts);
            int ix = f2i(fmin_(std::floor(fx), ts - 1.0f));
            Float fw = rp.filter_table[iy * 16 + ix];
            Float* px = rgbw + 4 * ((size_t)(y - cb[1]) * width + (x - cb[0]));
            Spectrum c = l * Spectrum(sample_weight) * Spectrum(fw);
            px[0] += c.c[0]; px[1] += c.c[1]; px[2] += c.c[2];
            px[3] += fw;
        }
    }
}

// sampling.rs:309-324
inline Vec3 uniform_sample_hemisphere(const Vec2& u) {
    Float z = u.x;
    Float r = std::sqrt(fmax_(0.0f, 1.0f - z * z));
    Float phi = 2.0f * PI * u.y;
    return Vec3(r * std::cos(phi), r * std::sin(phi), z);
}
// AOIntegrator::li (integrators/ao.rs:47-97): n_samples hemisphere rays from the first hit, drawn from the sampler's 2D array
inline Spectrum ao_li(ShadeCtx& cx, const Ray& ray, int32_t n_samples, bool cos_sample) {
    const Scene& sc = *cx.scene;
    Spectrum l;
    SurfaceInteraction isect;
    if (sc.intersect(ray, isect, cx.cnt)) {
        Normal3 n = faceforward(isect.common.n, -ray.d);
        Vec3 s = normalize(isect.dpdu);
        Vec3 t = cross(isect.common.n, s);
        const Vec2* u = cx.sampler->get_2d_array(n_samples);
        if (u) {
            for (int32_t i = 0; i < n_samples; ++i) {
                Vec3 wi;
                Float pdf;
                if (cos_sample) { wi = cosine_sample_hemisphere(u[i]); pdf = std::fabs(wi.z) * INV_PI; }
                else { wi = uniform_sample_hemisphere(u[i]); pdf = INV_2_PI; }
                wi = Vec3(s.x * wi.x + t.x * wi.y + n.x * wi.z, s.y * wi.x + t.y * wi.y + n.y * wi.z, s.z * wi.x + t.z * wi.y + n.z * wi.z);
                if (pdf != 0.0f && !sc.intersect_p(spawn_ray(isect.common, wi), cx.cnt)) l += Spectrum(dot(wi, n) / (pdf * (Float)n_samples));
            }
        }
    }
    return l;
}

// One camera sample: integrator.rs:134-197 (quirk Q1: only NaN is rejected)
inline Spectrum render_sample(ShadeCtx& cx, const PbrtRenderParams& rp, int32_t px, int32_t py, Vec2& p_film_out, const DirectCfg* direct = nullptr) {
    Sampler& s = *cx.sampler;
    Vec2 u = s.get_2d();
    Vec2 p_film((Float)px + u.x, (Float)py + u.y);
    Float time = s.get_1d();
    Vec2 p_lens = s.get_2d();
    Ray ray = camera_ray(cx.scene->camera, p_film, time, p_lens);
    ray.scale_differentials(1.0f / std::sqrt((Float)rp.spp));  // integrator.rs:140-144
    if (cx.cnt) cx.cnt->camera_rays++;
    Spectrum l = rp.integrator == PBRT_INTEGRATOR_AO ? ao_li(cx, ray, (int32_t)rp.ao_samples, rp.ao_cos_sample != 0)
                 : direct                           ? direct_li(cx, *direct, ray, 0)
                                                    : path_li(cx, ray, rp.max_depth, rp.rr_threshold);
    if (l.has_nans()) l = Spectrum(0.0f);
    p_film_out = p_film;
    return l;
}

// SamplerIntegrator::render (integrator.rs:70-220): 16x16 tiles pulled from an atomic cursor by
// n_threads workers.  Each tile accumulates into a private buffer (FilmTile, with its 1-px apron)
// that is merged under a lock, like film.merge_film_tile.  sample_rgb (optional) receives every
// sample's radiance [pixel in rect row-major][sample][3].
inline void render(const Scene& sc, const PbrtRenderParams& rp, const int32_t rect[4], Float* film_rgbw, Float* sample_rgb, int n_threads,
                   Counters* total) {
    sc.instancing = rp.instancing;
    LightDistribution ld(&sc, (int)rp.light_strategy);
    const bool is_direct = rp.integrator == PBRT_INTEGRATOR_DIRECT || rp.integrator == PBRT_INTEGRATOR_WHITTED;
    DirectCfg dcfg;
    dcfg.whitted = rp.integrator == PBRT_INTEGRATOR_WHITTED;
    dcfg.sample_all = rp.direct_strategy == PBRT_DIRECT_SAMPLE_ALL;
    dcfg.max_depth = rp.max_depth;
    for (const AreaLight& l : sc.lights) dcfg.n_light_samples.push_back((int32_t)std::max(1u, l.n_samples));  // round_count is the identity (sobol.rs:213, halton.rs:308)
    sc.allow_multiple_lobes = !is_direct;
    const int tile = 16;
    int32_t x0 = rect[0], y0 = rect[1], x1 = rect[2], y1 = rect[3];
    int ntx = (x1 - x0 + tile - 1) / tile, nty = (y1 - y0 + tile - 1) / tile;
    std::atomic<int> cursor(0);
    std::mutex film_mutex, cnt_mutex;
    const int32_t* cb = rp.cropped_pixel_bounds;
    const int32_t fw = cb[2] - cb[0], fh = cb[3] - cb[1];
    auto worker = [&]() {
        Counters local;
        std::unique_ptr<Sampler> sampler_owner = make_sampler(rp);
        Sampler& sampler = *sampler_owner;
        if (rp.integrator == PBRT_INTEGRATOR_AO) sampler.request_2d_array((int32_t)rp.ao_samples);  // AOIntegrator::preprocess ao.rs:44-46
        if (is_direct && dcfg.sample_all && !dcfg.whitted)  // DirectLightingIntegrator::preprocess (directlighting.rs:52-66)
            for (uint32_t i = 0; i < dcfg.max_depth; ++i)
                for (size_t j = 0; j < sc.lights.size(); ++j) {
                    sampler.request_2d_array(dcfg.n_light_samples[j]);
                    sampler.request_2d_array(dcfg.n_light_samples[j]);
                }
        ShadeCtx cx{&sc, &sampler, &ld, &local};
        std::vector<Float> tilebuf;
        for (;;) {
            int t = cursor.fetch_add(1);
            if (t >= ntx * nty) break;
            int tx = t % ntx, ty = t / ntx;
            int32_t tx0 = x0 + tx * tile, ty0 = y0 + ty * tile;
            int32_t tx1 = std::min(tx0 + tile, x1), ty1 = std::min(ty0 + tile, y1);
            // tile pixel bounds incl. filter apron (film.rs:308-345), as a private cropped film
            PbrtRenderParams trp = rp;
            int32_t ax0 = f2i(std::ceil((Float)tx0 - 0.5f - rp.filter_radius[0])), ay0 = f2i(std::ceil((Float)ty0 - 0.5f - rp.filter_radius[1]));
            int32_t ax1 = f2i(std::floor((Float)tx1 - 0.5f + rp.filter_radius[0])) + 1, ay1 = f2i(std::floor((Float)ty1 - 0.5f + rp.filter_radius[1])) + 1;
            trp.cropped_pixel_bounds[0] = std::max(ax0, cb[0]); trp.cropped_pixel_bounds[1] = std::max(ay0, cb[1]);
            trp.cropped_pixel_bounds[2] = std::min(ax1, cb[2]); trp.cropped_pixel_bounds[3] = std::min(ay1, cb[3]);
            const int32_t* tb = trp.cropped_pixel_bounds;
            int32_t tw = std::max(0, tb[2] - tb[0]), th = std::max(0, tb[3] - tb[1]);
            tilebuf.assign((size_t)tw * th * 4, 0.0f);
            for (int32_t py = ty0; py < ty1; ++py)
                for (int32_t px = tx0; px < tx1; ++px) {
                    sampler.start_pixel(px, py);
                    if (!(px >= rp.pixel_bounds[0] && px < rp.pixel_bounds[2] && py >= rp.pixel_bounds[1] && py < rp.pixel_bounds[3])) continue;
                    bool more = true;
                    while (more) {
                        Vec2 p_film;
                        int64_t si = sampler.current_pixel_sample_index;
                        Spectrum l = render_sample(cx, rp, px, py, p_film, is_direct ? &dcfg : nullptr);
                        if (sample_rgb) {
                            size_t pi = (size_t)(py - y0) * (size_t)(x1 - x0) + (size_t)(px - x0);
                            Float* o = sample_rgb + (pi * rp.spp + (size_t)si) * 3;
                            o[0] = l.c[0]; o[1] = l.c[1]; o[2] = l.c[2];
                        }
                        if (tw > 0 && th > 0) film_add_sample(trp, tilebuf.data(), p_film, l, 1.0f);
                        more = sampler.start_next_sample();
                    }
                }
            if (film_rgbw && tw > 0 && th > 0) {
                std::lock_guard<std::mutex> g(film_mutex);
                for (int32_t y = tb[1]; y < tb[3]; ++y)
                    for (int32_t x = tb[0]; x < tb[2]; ++x) {
                        const Float* s = &tilebuf[4 * ((size_t)(y - tb[1]) * tw + (x - tb[0]))];
                        Float* d = film_rgbw + 4 * ((size_t)(y - cb[1]) * fw + (x - cb[0]));
                        d[0] += s[0]; d[1] += s[1]; d[2] += s[2]; d[3] += s[3];
                    }
            }
        }
        std::lock_guard<std::mutex> g(cnt_mutex);
        if (total) total->add(local);
    };
    (void)fh;
    if (n_threads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
}

}  // namespace orc
