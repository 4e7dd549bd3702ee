// ORACLE -- TEST INFRASTRUCTURE ONLY (see o_math.hpp).
// o_reflection.hpp: BxDFs, Bsdf, TrowbridgeReitz distribution, material -> lobe lists.
// Follows src/core/reflection.rs, src/core/microfacet.rs, src/materials/*.rs.  A BxDF's sc_opt is
// Some(scale) only on the lobes of a MixMaterial's children (src/materials/mixmat.rs).
#pragma once
#include <vector>

#include "../include/pbrt_gpu.h"
#include "o_math.hpp"

namespace orc {

enum BxdfType { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };

// reflection.rs:1803-1976 -- shading-frame helpers
inline Float cos_theta(const Vec3& w) { return w.z; }
inline Float cos_2_theta(const Vec3& w) { return w.z * w.z; }
inline Float abs_cos_theta(const Vec3& w) { return std::fabs(w.z); }
inline Float sin_2_theta(const Vec3& w) { return fmax_(0.0f, 1.0f - cos_2_theta(w)); }
inline Float sin_theta(const Vec3& w) { return std::sqrt(sin_2_theta(w)); }
inline Float tan_theta(const Vec3& w) { return sin_theta(w) / cos_theta(w); }
inline Float tan_2_theta(const Vec3& w) { return sin_2_theta(w) / cos_2_theta(w); }
inline Float cos_phi(const Vec3& w) { Float st = sin_theta(w); return (st == 0.0f) ? 1.0f : clamp_t(w.x / st, -1.0f, 1.0f); }
inline Float sin_phi(const Vec3& w) { Float st = sin_theta(w); return (st == 0.0f) ? 0.0f : clamp_t(w.y / st, -1.0f, 1.0f); }
inline Float cos_2_phi(const Vec3& w) { return cos_phi(w) * cos_phi(w); }
inline Float sin_2_phi(const Vec3& w) { return sin_phi(w) * sin_phi(w); }
inline bool same_hemisphere(const Vec3& w, const Vec3& wp) { return w.z * wp.z > 0.0f; }
inline Vec3 reflect(const Vec3& wo, const Vec3& n) { return -wo + n * 2.0f * dot(wo, n); }
inline bool refract(const Vec3& wi, const Normal3& n, Float eta, Vec3& wt) {
    Float cos_theta_i = dot(n, wi);
    Float sin2_theta_i = fmax_(0.0f, 1.0f - cos_theta_i * cos_theta_i);
    Float sin2_theta_t = eta * eta * sin2_theta_i;
    if (sin2_theta_t >= 1.0f) return false;
    Float cos_theta_t = std::sqrt(1.0f - sin2_theta_t);
    wt = -wi * eta + n * (eta * cos_theta_i - cos_theta_t);
    return true;
}
inline Float pow5(Float v) { return (v * v) * (v * v) * v; }

// reflection.rs:1920-1951
inline Float fr_dielectric(Float cos_theta_i, Float eta_i, Float eta_t) {
    cos_theta_i = clamp_t(cos_theta_i, -1.0f, 1.0f);
    bool entering = cos_theta_i > 0.0f;
    if (!entering) { std::swap(eta_i, eta_t); cos_theta_i = std::fabs(cos_theta_i); }
    Float sin_theta_i = std::sqrt(fmax_(0.0f, 1.0f - cos_theta_i * cos_theta_i));
    Float sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0f) return 1.0f;
    Float cos_theta_t = std::sqrt(fmax_(0.0f, 1.0f - sin_theta_t * sin_theta_t));
    Float r_parl = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
    Float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
// reflection.rs:1953-1976
inline Spectrum fr_conductor(Float cos_theta_i, const Spectrum& eta_i, const Spectrum& eta_t, const Spectrum& k) {
    cos_theta_i = clamp_t(cos_theta_i, -1.0f, 1.0f);
    Spectrum eta = eta_t / eta_i;
    Spectrum eta_k = k / eta_i;
    Float cos_theta_i2 = cos_theta_i * cos_theta_i;
    Float sin_theta_i2 = 1.0f - cos_theta_i2;
    Spectrum eta_2 = eta * eta;
    Spectrum eta_k2 = eta_k * eta_k;
    Spectrum t0 = eta_2 - eta_k2 - Spectrum(sin_theta_i2);
    Spectrum a2_plus_b2 = sqrt(t0 * t0 + eta_2 * eta_k2 * Spectrum(4.0f));
    Spectrum t1 = a2_plus_b2 + Spectrum(cos_theta_i2);
    Spectrum a = sqrt((a2_plus_b2 + t0) * 0.5f);
    Spectrum t2 = a * 2.0f * cos_theta_i;
    Spectrum rs = (t1 - t2) / (t1 + t2);
    Spectrum t3 = a2_plus_b2 * cos_theta_i2 + Spectrum(sin_theta_i2 * sin_theta_i2);
    Spectrum t4 = t2 * sin_theta_i2;
    Spectrum rp = rs * (t3 - t4) / (t3 + t4);
    return (rp + rs) * Spectrum(0.5f);
}

// sampling.rs:344-365, :215-221
inline Vec2 concentric_sample_disk(const Vec2& u) {
    Vec2 uo(u.x * 2.0f - 1.0f, u.y * 2.0f - 1.0f);
    if (uo.x == 0.0f && uo.y == 0.0f) return Vec2(0.0f, 0.0f);
    Float theta, r;
    if (std::fabs(uo.x) > std::fabs(uo.y)) { r = uo.x; theta = PI_OVER_4 * (uo.y / uo.x); }
    else { r = uo.y; theta = PI_OVER_2 - PI_OVER_4 * (uo.x / uo.y); }
    return Vec2(std::cos(theta) * r, std::sin(theta) * r);
}
inline Vec3 cosine_sample_hemisphere(const Vec2& u) {
    Vec2 d = concentric_sample_disk(u);
    Float z = std::sqrt(fmax_(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    return Vec3(d.x, d.y, z);
}

// TrowbridgeReitzDistribution, sample_visible_area = true   microfacet.rs:224-353,475-569
struct TRDist {
    Float alpha_x, alpha_y;
    static Float roughness_to_alpha(Float roughness) {  // microfacet.rs:243-255
        if (1e-3f > roughness) roughness = 1e-3f;
        Float x = std::log(roughness);
        return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
    }
    Float d(const Vec3& wh) const {
        Float t2 = tan_2_theta(wh);
        if (std::isinf(t2)) return 0.0f;
        Float cos_4_theta = cos_2_theta(wh) * cos_2_theta(wh);
        Float e = (cos_2_phi(wh) / (alpha_x * alpha_x) + sin_2_phi(wh) / (alpha_y * alpha_y)) * t2;
        return 1.0f / (PI * alpha_x * alpha_y * cos_4_theta * (1.0f + e) * (1.0f + e));
    }
    Float lambda(const Vec3& w) const {
        Float abs_tan_theta = std::fabs(tan_theta(w));
        if (std::isinf(abs_tan_theta)) return 0.0f;
        Float alpha = std::sqrt(cos_2_phi(w) * alpha_x * alpha_x + sin_2_phi(w) * alpha_y * alpha_y);
        Float a2t2 = (alpha * abs_tan_theta) * (alpha * abs_tan_theta);
        return (-1.0f + std::sqrt(1.0f + a2t2)) / 2.0f;
    }
    Float g1(const Vec3& w) const { return 1.0f / (1.0f + lambda(w)); }
    Float g(const Vec3& wo, const Vec3& wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
    Float pdf(const Vec3& wo, const Vec3& wh) const { return d(wh) * g1(wo) * abs_dot(wo, wh) / abs_cos_theta(wo); }
    static void sample_11(Float cos_theta_, Float u1, Float u2, Float& slope_x, Float& slope_y) {  // microfacet.rs:475-531
        if (cos_theta_ > 0.9999f) {
            Float r = std::sqrt(u1 / (1.0f - u1));
            Float phi = TAU * u2;
            slope_x = r * std::cos(phi);
            slope_y = r * std::sin(phi);
            return;
        }
        Float sin_theta_ = std::sqrt(fmax_(0.0f, 1.0f - cos_theta_ * cos_theta_));
        Float tan_theta_ = sin_theta_ / cos_theta_;
        Float a = 1.0f / tan_theta_;
        Float g1 = 2.0f / (1.0f + std::sqrt(1.0f + 1.0f / (a * a)));
        a = 2.0f * u1 / g1 - 1.0f;
        Float tmp = 1.0f / (a * a - 1.0f);
        if (tmp > 1e10f) tmp = 1e10f;
        Float b = tan_theta_;
        Float d = std::sqrt(fmax_(b * b * tmp * tmp - (a * a - b * b) * tmp, 0.0f));
        Float slope_x_1 = b * tmp - d, slope_x_2 = b * tmp + d;
        slope_x = (a < 0.0f || slope_x_2 > 1.0f / tan_theta_) ? slope_x_1 : slope_x_2;
        Float s, nu2;
        if (u2 > 0.5f) { s = 1.0f; nu2 = 2.0f * (u2 - 0.5f); }
        else { s = -1.0f; nu2 = 2.0f * (0.5f - u2); }
        Float z = (nu2 * (nu2 * (nu2 * 0.27385f - 0.73369f) + 0.46341f)) / (nu2 * (nu2 * (nu2 * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
        slope_y = s * z * std::sqrt(1.0f + slope_x * slope_x);
    }
    static Vec3 sample(const Vec3& wi, Float ax, Float ay, Float u1, Float u2) {  // microfacet.rs:533-569
        Vec3 wis = normalize(Vec3(ax * wi.x, ay * wi.y, wi.z));
        Float slope_x = 0.0f, slope_y = 0.0f;
        sample_11(cos_theta(wis), u1, u2, slope_x, slope_y);
        Float tmp = cos_phi(wis) * slope_x - sin_phi(wis) * slope_y;
        slope_y = sin_phi(wis) * slope_x + cos_phi(wis) * slope_y;
        slope_x = tmp;
        slope_x *= ax;
        slope_y *= ay;
        return normalize(Vec3(-slope_x, -slope_y, 1.0f));
    }
    Vec3 sample_wh(const Vec3& wo, const Vec2& u) const {  // microfacet.rs:298-349 (visible-area branch)
        if (wo.z < 0.0f) return -sample(-wo, alpha_x, alpha_y, u.x, u.y);
        return sample(wo, alpha_x, alpha_y, u.x, u.y);
    }
};
inline TRDist make_tr(Float ax, Float ay) { TRDist d; d.alpha_x = fmax_(ax, 0.001f); d.alpha_y = fmax_(ay, 0.001f); return d; }

enum BxdfKind { BX_SPEC_REFL, BX_SPEC_TRANS, BX_FRESNEL_SPEC, BX_LAMBERT_REFL, BX_OREN_NAYAR, BX_MF_REFL, BX_MF_TRANS, BX_FRESNEL_BLEND, BX_LAMBERT_TRANS };
enum FresnelKind { FR_NOOP, FR_CONDUCTOR, FR_DIELECTRIC };

struct Fresnel {
    int kind = FR_NOOP;
    Spectrum c_eta_i, c_eta_t, c_k;  // conductor
    Float d_eta_i = 1.0f, d_eta_t = 1.0f;  // dielectric
    Spectrum evaluate(Float cos_i) const {
        if (kind == FR_CONDUCTOR) return fr_conductor(cos_i, c_eta_i, c_eta_t, c_k);
        if (kind == FR_DIELECTRIC) return Spectrum(fr_dielectric(cos_i, d_eta_i, d_eta_t));
        return Spectrum(1.0f);
    }
};

struct Bxdf {
    int kind;
    Spectrum r, t;        // r: R / Kd / rd ; t: T / rs
    Float eta_a = 1.0f, eta_b = 1.0f;
    Fresnel fresnel;
    TRDist dist;
    Float on_a = 0.0f, on_b = 0.0f;  // Oren-Nayar A, B
    bool has_sc = false;  // sc_opt: Some(scale) on the lobes a MixMaterial's children produce (mixmat.rs:52-72), None everywhere else
    Spectrum sc;

    int type() const {
        switch (kind) {
            case BX_SPEC_REFL: return BSDF_REFLECTION | BSDF_SPECULAR;
            case BX_SPEC_TRANS: return BSDF_TRANSMISSION | BSDF_SPECULAR;
            case BX_FRESNEL_SPEC: return BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
            case BX_LAMBERT_REFL: case BX_OREN_NAYAR: return BSDF_DIFFUSE | BSDF_REFLECTION;
            case BX_MF_REFL: case BX_FRESNEL_BLEND: return BSDF_REFLECTION | BSDF_GLOSSY;
            case BX_LAMBERT_TRANS: return BSDF_DIFFUSE | BSDF_TRANSMISSION;  // reflection.rs:1043-1045
            default: return BSDF_TRANSMISSION | BSDF_GLOSSY;  // BX_MF_TRANS
        }
    }
    bool matches_flags(int f) const { return (type() & f) == type(); }

    Spectrum f(const Vec3& wo, const Vec3& wi) const {
        switch (kind) {
            case BX_SPEC_REFL: case BX_SPEC_TRANS: case BX_FRESNEL_SPEC: return Spectrum(0.0f);
            case BX_LAMBERT_REFL: return has_sc ? sc * r * Spectrum(INV_PI) : r * Spectrum(INV_PI);  // reflection.rs:962-968
            case BX_LAMBERT_TRANS: return has_sc ? sc * t * INV_PI : t * Spectrum(INV_PI);  // reflection.rs:1010-1016
            case BX_OREN_NAYAR: {  // reflection.rs:1067-1095
                Float sin_theta_i = sin_theta(wi), sin_theta_o = sin_theta(wo);
                Float max_cos = 0.0f;
                if (sin_theta_i > 1.0e-4f && sin_theta_o > 1.0e-4f) {
                    Float d_cos = cos_phi(wi) * cos_phi(wo) + sin_phi(wi) * sin_phi(wo);
                    max_cos = fmax_(d_cos, 0.0f);
                }
                Float sin_alpha, tan_beta;
                if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sin_theta_o; tan_beta = sin_theta_i / abs_cos_theta(wi); }
                else { sin_alpha = sin_theta_i; tan_beta = sin_theta_o / abs_cos_theta(wo); }
                if (has_sc) return sc * r * Spectrum(INV_PI * (on_a + on_b * max_cos * sin_alpha * tan_beta));  // reflection.rs:1090-1091
                return r * Spectrum(INV_PI * (on_a + on_b * max_cos * sin_alpha * tan_beta));
            }
            case BX_MF_REFL: {  // reflection.rs:1149-1170
                Float cos_theta_o = abs_cos_theta(wo), cos_theta_i = abs_cos_theta(wi);
                Vec3 wh = wi + wo;
                if (cos_theta_i == 0.0f || cos_theta_o == 0.0f) return Spectrum(0.0f);
                if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spectrum(0.0f);
                wh = normalize(wh);
                Spectrum F = fresnel.evaluate(dot(wi, wh));
                if (has_sc) return sc * r * dist.d(wh) * dist.g(wo, wi) * F / (4.0f * cos_theta_i * cos_theta_o);  // reflection.rs:1163-1165
                return r * dist.d(wh) * dist.g(wo, wi) * F / (4.0f * cos_theta_i * cos_theta_o);
            }
            case BX_MF_TRANS: {  // reflection.rs:1246-1312
                if (same_hemisphere(wo, wi)) return Spectrum(0.0f);
                Float cos_theta_o = cos_theta(wo), cos_theta_i = cos_theta(wi);
                if (cos_theta_o == 0.0f || cos_theta_i == 0.0f) return Spectrum(0.0f);
                Float eta = (cos_theta_o > 0.0f) ? (eta_b / eta_a) : (eta_a / eta_b);
                Vec3 wh = normalize(wo + wi * eta);
                if (wh.z < 0.0f) wh = -wh;
                if (dot(wo, wh) * dot(wi, wh) > 0.0f) return Spectrum(0.0f);
                Spectrum F = Spectrum(fr_dielectric(dot(wo, wh), eta_a, eta_b));
                Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                Float factor = 1.0f / eta;  // TransportMode::Radiance
                const Float g = std::fabs(dist.d(wh) * dist.g(wo, wi) * eta * eta * abs_dot(wi, wh) * abs_dot(wo, wh) * factor * factor /
                                          (cos_theta_i * cos_theta_o * sqrt_denom * sqrt_denom));
                if (has_sc) return sc * (Spectrum(1.0f) - F) * t * g;  // reflection.rs:1283-1296
                return (Spectrum(1.0f) - F) * t * g;
            }
            default: {  // BX_FRESNEL_BLEND  reflection.rs:1398-1427  (r = rd, t = rs)
                Spectrum diffuse = r * (Spectrum(1.0f) - t) * (28.0f / (23.0f * PI)) * (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wi))) *
                                   (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wo)));
                Vec3 wh = wi + wo;
                if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spectrum(0.0f);
                wh = normalize(wh);
                Spectrum schlick = t + (Spectrum(1.0f) - t) * pow5(1.0f - dot(wi, wh));
                Spectrum specular = schlick * (dist.d(wh) / (4.0f * std::fabs(dot(wi, wh)) * fmax_(abs_cos_theta(wi), abs_cos_theta(wo))));
                if (has_sc) return sc * (diffuse + specular);  // reflection.rs:1417-1418
                return diffuse + specular;
            }
        }
    }
    Float pdf(const Vec3& wo, const Vec3& wi) const {
        switch (kind) {
            case BX_SPEC_REFL: return 0.0f;
            case BX_SPEC_TRANS: case BX_FRESNEL_SPEC:  // reflection.rs:828-834, :938-944 (sic: cosine pdf)
            case BX_LAMBERT_REFL: case BX_OREN_NAYAR:
                return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * INV_PI : 0.0f;
            case BX_LAMBERT_TRANS:  // reflection.rs:1036-1042
                return !same_hemisphere(wo, wi) ? abs_cos_theta(wi) * INV_PI : 0.0f;
            case BX_MF_REFL: {
                if (!same_hemisphere(wo, wi)) return 0.0f;
                Vec3 wh = normalize(wo + wi);
                return dist.pdf(wo, wh) / (4.0f * dot(wo, wh));
            }
            case BX_MF_TRANS: {  // reflection.rs:1348-1370
                if (same_hemisphere(wo, wi)) return 0.0f;
                Float eta = (cos_theta(wo) > 0.0f) ? (eta_b / eta_a) : (eta_a / eta_b);
                Vec3 wh = normalize(wo + wi * eta);
                Float wo_dot_wh = dot(wo, wh), wi_dot_wh = dot(wi, wh);
                if (wo_dot_wh * wi_dot_wh > 0.0f) return 0.0f;
                Float sqrt_denom = wo_dot_wh + eta * wi_dot_wh;
                Float dwh_dwi = std::fabs((eta * eta * wi_dot_wh) / (sqrt_denom * sqrt_denom));
                return dist.pdf(wo, wh) * dwh_dwi;
            }
            default: {  // BX_FRESNEL_BLEND reflection.rs:1462-1474
                if (!same_hemisphere(wo, wi)) return 0.0f;
                Vec3 wh = normalize(wo + wi);
                Float pdf_wh = dist.pdf(wo, wh);
                return 0.5f * (abs_cos_theta(wi) * INV_PI + pdf_wh / (4.0f * dot(wo, wh)));
            }
        }
    }
    // *sampled_type is only rewritten by FresnelSpecular (and only when non-zero on entry)
    Spectrum sample_f(const Vec3& wo, Vec3& wi, const Vec2& u, Float& pdf_, int& sampled_type) const {
        switch (kind) {
            case BX_SPEC_REFL: {  // reflection.rs:724-745
                wi = Vec3(-wo.x, -wo.y, wo.z);
                pdf_ = 1.0f;
                if (has_sc) return sc * fresnel.evaluate(cos_theta(wi)) * r / abs_cos_theta(wi);  // reflection.rs:740-741
                return fresnel.evaluate(cos_theta(wi)) * r / abs_cos_theta(wi);
            }
            case BX_SPEC_TRANS: {  // reflection.rs:787-827
                bool entering = cos_theta(wo) > 0.0f;
                Float eta_i = entering ? eta_a : eta_b, eta_t = entering ? eta_b : eta_a;
                if (!refract(wo, faceforward(Normal3(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, wi)) return Spectrum();
                pdf_ = 1.0f;
                Spectrum ft = t * (Spectrum(1.0f) - Spectrum(fr_dielectric(cos_theta(wi), eta_a, eta_b)));
                ft *= Spectrum((eta_i * eta_i) / (eta_t * eta_t));
                if (has_sc) return sc * ft / abs_cos_theta(wi);  // reflection.rs:822-823
                return ft / abs_cos_theta(wi);
            }
            case BX_FRESNEL_SPEC: {  // reflection.rs:871-937
                Float F = fr_dielectric(cos_theta(wo), eta_a, eta_b);
                if (u.x < F) {
                    wi = Vec3(-wo.x, -wo.y, wo.z);
                    if (sampled_type != 0) sampled_type = BSDF_REFLECTION | BSDF_SPECULAR;
                    pdf_ = F;
                    if (has_sc) return sc * r * F / abs_cos_theta(wi);  // reflection.rs:894-895
                    return r * F / abs_cos_theta(wi);
                }
                bool entering = cos_theta(wo) > 0.0f;
                Float eta_i = entering ? eta_a : eta_b, eta_t = entering ? eta_b : eta_a;
                if (!refract(wo, faceforward(Normal3(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, wi)) return Spectrum();
                Spectrum ft = t * (1.0f - F);
                ft *= Spectrum((eta_i * eta_i) / (eta_t * eta_t));
                if (sampled_type != 0) sampled_type = BSDF_TRANSMISSION | BSDF_SPECULAR;
                pdf_ = 1.0f - F;
                if (has_sc) return sc * ft / abs_cos_theta(wi);  // reflection.rs:931-932
                return ft / abs_cos_theta(wi);
            }
            case BX_LAMBERT_REFL: case BX_OREN_NAYAR: {  // reflection.rs:969-987, :1096-1114
                wi = cosine_sample_hemisphere(u);
                if (wo.z < 0.0f) wi.z *= -1.0f;
                pdf_ = pdf(wo, wi);
                return has_sc ? sc * f(wo, wi) : f(wo, wi);  // (sic: scaled twice, reflection.rs:982-983, :1109-1110; Bsdf::sample_f recomputes f, quirk Q9)
            }
            case BX_LAMBERT_TRANS: {  // reflection.rs:1017-1035
                wi = cosine_sample_hemisphere(u);
                if (wo.z > 0.0f) wi.z *= -1.0f;
                pdf_ = pdf(wo, wi);
                return has_sc ? sc * f(wo, wi) : f(wo, wi);  // reflection.rs:1030-1031
            }
            case BX_MF_REFL: {  // reflection.rs:1172-1196
                if (wo.z == 0.0f) return Spectrum();
                Vec3 wh = dist.sample_wh(wo, u);
                wi = reflect(wo, wh);
                if (!same_hemisphere(wo, wi)) return Spectrum();
                pdf_ = dist.pdf(wo, wh) / (4.0f * dot(wo, wh));
                return has_sc ? sc * f(wo, wi) : f(wo, wi);  // reflection.rs:1191-1192
            }
            case BX_MF_TRANS: {  // reflection.rs:1318-1347
                if (wo.z == 0.0f) return Spectrum();
                Vec3 wh = dist.sample_wh(wo, u);
                Float eta = (cos_theta(wo) > 0.0f) ? (eta_a / eta_b) : (eta_b / eta_a);
                if (refract(wo, wh, eta, wi)) { pdf_ = pdf(wo, wi); return has_sc ? sc * f(wo, wi) : f(wo, wi); }  // reflection.rs:1339-1340
                return Spectrum();
            }
            default: {  // BX_FRESNEL_BLEND reflection.rs:1428-1461
                Vec2 uu = u;
                if (uu.x < 0.5f) {
                    uu.x = fmin_(2.0f * uu.x, FLOAT_ONE_MINUS_EPSILON);
                    wi = cosine_sample_hemisphere(uu);
                    if (wo.z < 0.0f) wi.z *= -1.0f;
                } else {
                    uu.x = fmin_(2.0f * (uu.x - 0.5f), FLOAT_ONE_MINUS_EPSILON);
                    Vec3 wh = dist.sample_wh(wo, uu);
                    wi = reflect(wo, wh);
                    if (!same_hemisphere(wo, wi)) return Spectrum(0.0f);
                }
                pdf_ = pdf(wo, wi);
                return has_sc ? sc * f(wo, wi) : f(wo, wi);  // reflection.rs:1456-1457
            }
        }
    }
};

// One material with constant textures always yields the same lobe list; compute it once.
struct MaterialLobes {
    Float eta = 1.0f;
    std::vector<Bxdf> bxdfs;
};
inline Spectrum spec3(const float* p) { return Spectrum(p[0], p[1], p[2]); }
inline Spectrum clamp_pos(const Spectrum& s) { return clamp_spectrum(s, 0.0f, INF); }

inline bool compile_material(const PbrtMaterial& m, MaterialLobes& out, bool allow_multiple_lobes = true) {
    const float* p = m.params;
    out.bxdfs.clear();
    out.eta = 1.0f;
    switch (m.kind) {
        case PBRT_MAT_MATTE: {  // matte.rs:43-86
            Spectrum r = clamp_pos(spec3(p));
            Float sig = clamp_t(p[3], 0.0f, 90.0f);
            if (!r.is_black()) {
                Bxdf b;
                b.r = r;
                if (sig == 0.0f) b.kind = BX_LAMBERT_REFL;
                else {  // OrenNayar::new reflection.rs:1057-1066
                    b.kind = BX_OREN_NAYAR;
                    Float sigma = radians(sig);
                    Float sigma2 = sigma * sigma;
                    b.on_a = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.33f)));
                    b.on_b = 0.45f * sigma2 / (sigma2 + 0.09f);
                }
                out.bxdfs.push_back(b);
            }
            return true;
        }
        case PBRT_MAT_PLASTIC: {  // plastic.rs:57-125
            Spectrum kd = clamp_pos(spec3(p)), ks = clamp_pos(spec3(p + 3));
            Float rough = p[6];
            if (!kd.is_black()) { Bxdf b; b.kind = BX_LAMBERT_REFL; b.r = kd; out.bxdfs.push_back(b); }
            if (!ks.is_black()) {
                Bxdf b; b.kind = BX_MF_REFL; b.r = ks;
                b.fresnel.kind = FR_DIELECTRIC; b.fresnel.d_eta_i = 1.5f; b.fresnel.d_eta_t = 1.0f;
                if (p[7] != 0.0f) rough = TRDist::roughness_to_alpha(rough);
                b.dist = make_tr(rough, rough);
                out.bxdfs.push_back(b);
            }
            return true;
        }
        case PBRT_MAT_METAL: {  // metal.rs:144-205
            Float ur = p[6], vr = p[7];
            if (p[8] != 0.0f) { ur = TRDist::roughness_to_alpha(ur); vr = TRDist::roughness_to_alpha(vr); }
            Bxdf b; b.kind = BX_MF_REFL; b.r = Spectrum(1.0f);
            b.fresnel.kind = FR_CONDUCTOR; b.fresnel.c_eta_i = Spectrum(1.0f); b.fresnel.c_eta_t = spec3(p); b.fresnel.c_k = spec3(p + 3);
            b.dist = make_tr(ur, vr);
            out.bxdfs.push_back(b);
            return true;
        }
        case PBRT_MAT_MIRROR: {  // mirror.rs:34-70
            Bxdf b; b.kind = BX_SPEC_REFL; b.r = clamp_pos(spec3(p)); b.fresnel.kind = FR_NOOP;
            out.bxdfs.push_back(b);
            return true;
        }
        case PBRT_MAT_GLASS: {  // glass.rs:83-211, allow_multiple_lobes = true (path.rs:108)
            Float urough = p[7], vrough = p[8];
            Spectrum r = clamp_pos(spec3(p)), t = clamp_pos(spec3(p + 3));
            bool is_specular = urough == 0.0f && vrough == 0.0f;
            Float eta = p[6];
            out.eta = eta;
            if (is_specular && allow_multiple_lobes) {
                Bxdf b; b.kind = BX_FRESNEL_SPEC; b.r = r; b.t = t; b.eta_a = 1.0f; b.eta_b = eta;
                out.bxdfs.push_back(b);
            } else {
                if (p[9] != 0.0f) { urough = TRDist::roughness_to_alpha(urough); vrough = TRDist::roughness_to_alpha(vrough); }
                if (!r.is_black()) {
                    Bxdf b; b.kind = is_specular ? BX_SPEC_REFL : BX_MF_REFL; b.r = r;
                    b.fresnel.kind = FR_DIELECTRIC; b.fresnel.d_eta_i = 1.0f; b.fresnel.d_eta_t = eta;
                    if (!is_specular) b.dist = make_tr(urough, vrough);
                    out.bxdfs.push_back(b);
                }
                if (!t.is_black()) {
                    Bxdf b; b.kind = is_specular ? BX_SPEC_TRANS : BX_MF_TRANS; b.t = t; b.eta_a = 1.0f; b.eta_b = eta;
                    if (!is_specular) b.dist = make_tr(urough, vrough);
                    out.bxdfs.push_back(b);
                }
            }
            return true;
        }
        case PBRT_MAT_UBER: {  // uber.rs:114-259
            Float e = p[17];
            Spectrum op = clamp_pos(spec3(p + 12));
            Spectrum t = clamp_pos(Spectrum(1.0f) - op);
            Spectrum kd = op * clamp_pos(spec3(p)), ks = op * clamp_pos(spec3(p + 3));
            Float ur = p[15], vr = p[16];
            Spectrum kr = op * clamp_pos(spec3(p + 6)), kt = op * clamp_pos(spec3(p + 9));
            out.eta = t.is_black() ? e : 1.0f;
            if (!t.is_black()) { Bxdf b; b.kind = BX_SPEC_TRANS; b.t = t; b.eta_a = 1.0f; b.eta_b = 1.0f; out.bxdfs.push_back(b); }
            if (!kd.is_black()) { Bxdf b; b.kind = BX_LAMBERT_REFL; b.r = kd; out.bxdfs.push_back(b); }
            if (!ks.is_black()) {
                Bxdf b; b.kind = BX_MF_REFL; b.r = ks;
                b.fresnel.kind = FR_DIELECTRIC; b.fresnel.d_eta_i = 1.0f; b.fresnel.d_eta_t = e;
                if (p[18] != 0.0f) { ur = TRDist::roughness_to_alpha(ur); vr = TRDist::roughness_to_alpha(vr); }
                b.dist = make_tr(ur, vr);
                out.bxdfs.push_back(b);
            }
            if (!kr.is_black()) {
                Bxdf b; b.kind = BX_SPEC_REFL; b.r = kr;
                b.fresnel.kind = FR_DIELECTRIC; b.fresnel.d_eta_i = 1.0f; b.fresnel.d_eta_t = e;
                out.bxdfs.push_back(b);
            }
            if (!kt.is_black()) { Bxdf b; b.kind = BX_SPEC_TRANS; b.t = kt; b.eta_a = 1.0f; b.eta_b = e; out.bxdfs.push_back(b); }
            return true;
        }
        case PBRT_MAT_SUBSTRATE: {  // substrate.rs:62-114
            Spectrum d = clamp_pos(spec3(p)), s = clamp_pos(spec3(p + 3));
            Float ru = p[6], rv = p[7];
            if (!d.is_black() || !s.is_black()) {
                if (p[8] != 0.0f) { ru = TRDist::roughness_to_alpha(ru); rv = TRDist::roughness_to_alpha(rv); }
                Bxdf b; b.kind = BX_FRESNEL_BLEND; b.r = d; b.t = s; b.dist = make_tr(ru, rv);
                out.bxdfs.push_back(b);
            }
            return true;
        }
        case PBRT_MAT_TRANSLUCENT: {  // translucent.rs:48-189 (TransportMode::Radiance; eta fixed at 1.5)
            const Float eta = 1.5f;
            out.eta = eta;
            Spectrum r = clamp_pos(spec3(p + 6)), t = clamp_pos(spec3(p + 9));
            if (r.is_black() && t.is_black()) return true;
            Spectrum kd = clamp_pos(spec3(p)), ks = clamp_pos(spec3(p + 3));
            Float rough = p[12];
            if (!kd.is_black()) {
                if (!r.is_black()) { Bxdf b; b.kind = BX_LAMBERT_REFL; b.r = r * kd; out.bxdfs.push_back(b); }
                if (!t.is_black()) { Bxdf b; b.kind = BX_LAMBERT_TRANS; b.t = t * kd; out.bxdfs.push_back(b); }
            }
            if (!ks.is_black() && (!r.is_black() || !t.is_black())) {
                if (p[13] != 0.0f) rough = TRDist::roughness_to_alpha(rough);
                if (!r.is_black()) {
                    Bxdf b; b.kind = BX_MF_REFL; b.r = r * ks;
                    b.fresnel.kind = FR_DIELECTRIC; b.fresnel.d_eta_i = 1.0f; b.fresnel.d_eta_t = eta;
                    b.dist = make_tr(rough, rough);
                    out.bxdfs.push_back(b);
                }
                if (!t.is_black()) { Bxdf b; b.kind = BX_MF_TRANS; b.t = t * ks; b.eta_a = 1.0f; b.eta_b = eta; b.dist = make_tr(rough, rough); out.bxdfs.push_back(b); }
            }
            return true;
        }
        default: return false;
    }
}

// Material number `index` of a scene's material array, MixMaterial included (mixmat.rs:41-98): the lobes of "namedmaterial1" built with
// scale s1 = clamp(amount), then those of "namedmaterial2" built with s2 = clamp(1 - s1) on a second SurfaceInteraction and added to the
// first Bsdf (whose eta and frame stay m1's).  `scale`: the Option<Spectrum> a parent mix hands down -- every other material passes it to
// each BxDF it creates (matte.rs:74-82, plastic.rs:88-120, ...), a MixMaterial ignores it (`_scale`, mixmat.rs:48).  Bsdf::add asserts
// fewer than MAX_BXDFS = 8 lobes (reflection.rs:246-249): a ninth is an error here.  Children have lower indices (pbrt_gpu.h).
inline bool compile_material_at(const PbrtMaterial* mats, uint32_t n_mats, uint32_t index, MaterialLobes& out, bool allow_multiple_lobes = true,
                                const Spectrum* scale = nullptr) {
    if (index >= n_mats) return false;
    const PbrtMaterial& m = mats[index];
    if (m.kind != PBRT_MAT_MIX) {
        if (!compile_material(m, out, allow_multiple_lobes)) return false;
        if (scale) for (Bxdf& b : out.bxdfs) { b.has_sc = true; b.sc = *scale; }
        return true;
    }
    const Float i1 = m.params[3], i2 = m.params[4];
    if (!(i1 >= 0.0f && i1 < (Float)index && i2 >= 0.0f && i2 < (Float)index) || i1 != std::floor(i1) || i2 != std::floor(i2)) return false;
    for (uint32_t c : {(uint32_t)i1, (uint32_t)i2}) {  // this version: constant amount, untextured children
        if (mats[c].bump) return false;
        for (int g = 0; g < PBRT_MAX_TEX_GROUPS; ++g) if (mats[c].tex[g]) return false;
    }
    for (int g = 0; g < PBRT_MAX_TEX_GROUPS; ++g) if (m.tex[g]) return false;
    if (m.bump) return false;
    const Spectrum s1 = clamp_pos(spec3(m.params));
    const Spectrum s2 = clamp_pos(Spectrum(1.0f) - s1);
    MaterialLobes second;
    if (!compile_material_at(mats, n_mats, (uint32_t)i1, out, allow_multiple_lobes, &s1)) return false;
    if (!compile_material_at(mats, n_mats, (uint32_t)i2, second, allow_multiple_lobes, &s2)) return false;
    for (const Bxdf& b : second.bxdfs) {
        if (out.bxdfs.size() >= 8) return false;
        out.bxdfs.push_back(b);
    }
    return true;
}

// Bsdf, reflection.rs:223-446
struct Bsdf {
    Float eta;
    Normal3 ns, ng;
    Vec3 ss, ts;
    const std::vector<Bxdf>* bxdfs;

    int num_components(int flags) const {
        int n = 0;
        for (const Bxdf& b : *bxdfs) if (b.matches_flags(flags)) ++n;
        return n;
    }
    Vec3 world_to_local(const Vec3& v) const { return Vec3(dot(v, ss), dot(v, ts), dot(v, ns)); }
    Vec3 local_to_world(const Vec3& v) const {
        return Vec3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z);
    }
    Spectrum f(const Vec3& wo_w, const Vec3& wi_w, int flags) const {
        Vec3 wi = world_to_local(wi_w), wo = world_to_local(wo_w);
        if (wo.z == 0.0f) return Spectrum(0.0f);
        bool reflect_ = (dot(wi_w, ng) * dot(wo_w, ng)) > 0.0f;
        Spectrum f_(0.0f);
        for (const Bxdf& b : *bxdfs)
            if (b.matches_flags(flags) && ((reflect_ && (b.type() & BSDF_REFLECTION)) || (!reflect_ && (b.type() & BSDF_TRANSMISSION))))
                f_ += b.f(wo, wi);
        return f_;
    }
    Float pdf(const Vec3& wo_w, const Vec3& wi_w, int flags) const {
        if (bxdfs->empty()) return 0.0f;
        Vec3 wo = world_to_local(wo_w), wi = world_to_local(wi_w);
        if (wo.z == 0.0f) return 0.0f;
        Float pdf_ = 0.0f;
        int matching = 0;
        for (const Bxdf& b : *bxdfs)
            if (b.matches_flags(flags)) { ++matching; pdf_ += b.pdf(wo, wi); }
        return matching > 0 ? pdf_ / (Float)matching : 0.0f;
    }
    // reflection.rs:298-420.  `pdf_` is left untouched on the wo.z == 0 early-out, as in the reference.
    Spectrum sample_f(const Vec3& wo_world, Vec3& wi_world, const Vec2& u, Float& pdf_, int flags, int& sampled_type) const {
        int matching = num_components(flags);
        if (matching == 0) { pdf_ = 0.0f; sampled_type = 0; return Spectrum(); }
        int32_t ci = f2i(std::floor(u.x * (Float)matching));
        int comp = std::min(clamp_t(ci, 0, 255), matching - 1);  // `as u8` saturates
        const Bxdf* bxdf = nullptr;
        int count = comp, index = 0;
        for (size_t i = 0; i < bxdfs->size(); ++i) {
            bool m = (*bxdfs)[i].matches_flags(flags);
            if (m && count == 0) { bxdf = &(*bxdfs)[i]; index = (int)i; break; }
            if (m) --count;
        }
        if (!bxdf) return Spectrum();
        Vec2 ur(fmin_(u.x * (Float)matching - (Float)comp, FLOAT_ONE_MINUS_EPSILON), u.y);
        Vec3 wi;
        Vec3 wo = world_to_local(wo_world);
        if (wo.z == 0.0f) return Spectrum();
        pdf_ = 0.0f;
        if (sampled_type != 0) sampled_type = bxdf->type();
        Spectrum f_ = bxdf->sample_f(wo, wi, ur, pdf_, sampled_type);
        if (pdf_ == 0.0f) { if (sampled_type != 0) sampled_type = 0; return Spectrum(); }
        wi_world = local_to_world(wi);
        if (!(bxdf->type() & BSDF_SPECULAR) && matching > 1)
            for (size_t i = 0; i < bxdfs->size(); ++i)
                if ((int)i != index && (*bxdfs)[i].matches_flags(flags)) pdf_ += (*bxdfs)[i].pdf(wo, wi);
        if (matching > 1) pdf_ /= (Float)matching;
        if (!(bxdf->type() & BSDF_SPECULAR)) {
            bool reflect_ = dot(wi_world, ng) * dot(wo_world, ng) > 0.0f;
            f_ = Spectrum();
            for (const Bxdf& b : *bxdfs)
                if (b.matches_flags(flags) && ((reflect_ && (b.type() & BSDF_REFLECTION)) || (!reflect_ && (b.type() & BSDF_TRANSMISSION))))
                    f_ += b.f(wo, wi);
        }
        return f_;
    }
};

}  // namespace orc
