// pb_bsdf.cuh -- Bsdf::f / pdf / sample_f and the in-scope BxDF lobes on the device.
//   Bsdf                         src/core/reflection.rs:223-446
//   SpecularReflection :711  SpecularTransmission :755  FresnelSpecular :841  LambertianReflection :953
//   OrenNayar :1049  MicrofacetReflection :1128  MicrofacetTransmission :1214  FresnelBlend :1374
//   fr_dielectric :1920  fr_conductor :1953  refract :1897  reflect :1890  frame helpers :1803-1886
//   TrowbridgeReitzDistribution  src/core/microfacet.rs:224-353,475-569 (sample_visible_area = true)
//   concentric_sample_disk / cosine_sample_hemisphere   src/core/sampling.rs:344-365,215-221
// A lobe's sc_opt is Some(scale) only under a MixMaterial (DLobe::has_sc; general instantiations only).  TransportMode is Radiance.
#pragma once
#include "pb_scene.cuh"

namespace pb {

PB_D float cos_theta(V3 w) { return w.z; }
PB_D float cos2_theta(V3 w) { return w.z * w.z; }
PB_D float abs_cos_theta(V3 w) { return fabsf(w.z); }
PB_D float sin2_theta(V3 w) { return fmaxf(0.0f, 1.0f - cos2_theta(w)); }
PB_D float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
PB_D float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
PB_D float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
PB_D float cos_phi(V3 w) { float st = sin_theta(w); return (st == 0.0f) ? 1.0f : clampf(w.x / st, -1.0f, 1.0f); }
PB_D float sin_phi(V3 w) { float st = sin_theta(w); return (st == 0.0f) ? 0.0f : clampf(w.y / st, -1.0f, 1.0f); }
PB_D float cos2_phi(V3 w) { float c = cos_phi(w); return c * c; }
PB_D float sin2_phi(V3 w) { float s = sin_phi(w); return s * s; }
PB_D bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
PB_D V3 reflect3(V3 wo, V3 n) { return -wo + n * 2.0f * dot3(wo, n); }
PB_D bool refract3(V3 wi, V3 n, float eta, V3& wt) {
    float cos_theta_i = dot3(n, wi);
    float sin2_theta_i = fmaxf(0.0f, 1.0f - cos_theta_i * cos_theta_i);
    float sin2_theta_t = eta * eta * sin2_theta_i;
    if (sin2_theta_t >= 1.0f) return false;
    float cos_theta_t = sqrtf(1.0f - sin2_theta_t);
    wt = -wi * eta + n * (eta * cos_theta_i - cos_theta_t);
    return true;
}
PB_D float pow5(float v) { return (v * v) * (v * v) * v; }

PB_D float fr_dielectric(float cos_theta_i, float eta_i, float eta_t) {
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    if (!(cos_theta_i > 0.0f)) {
        float tmp = eta_i; eta_i = eta_t; eta_t = tmp;
        cos_theta_i = fabsf(cos_theta_i);
    }
    float sin_theta_i = sqrtf(fmaxf(0.0f, 1.0f - cos_theta_i * cos_theta_i));
    float sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0f) return 1.0f;
    float cos_theta_t = sqrtf(fmaxf(0.0f, 1.0f - sin_theta_t * sin_theta_t));
    float r_parl = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
    float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
PB_D Sp fr_conductor(float cos_theta_i, Sp eta_i, Sp eta_t, Sp k) {
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    Sp eta = eta_t / eta_i;
    Sp eta_k = k / eta_i;
    float cos_theta_i2 = cos_theta_i * cos_theta_i;
    float sin_theta_i2 = 1.0f - cos_theta_i2;
    Sp eta_2 = eta * eta;
    Sp eta_k2 = eta_k * eta_k;
    Sp t0 = eta_2 - eta_k2 - sp1(sin_theta_i2);
    Sp a2_plus_b2 = sqrtsp(t0 * t0 + eta_2 * eta_k2 * sp1(4.0f));
    Sp t1 = a2_plus_b2 + sp1(cos_theta_i2);
    Sp a = sqrtsp((a2_plus_b2 + t0) * 0.5f);
    Sp t2 = a * 2.0f * cos_theta_i;
    Sp rs = (t1 - t2) / (t1 + t2);
    Sp t3 = a2_plus_b2 * cos_theta_i2 + sp1(sin_theta_i2 * sin_theta_i2);
    Sp t4 = t2 * sin_theta_i2;
    Sp rp = rs * (t3 - t4) / (t3 + t4);
    return (rp + rs) * sp1(0.5f);
}
PB_D Sp fresnel_eval(const DLobe& L, float cos_i) {
    if (L.fresnel == FRESNEL_CONDUCTOR)
        return fr_conductor(cos_i, sp1(1.0f), mksp(L.fr_a[0], L.fr_a[1], L.fr_a[2]), mksp(L.fr_k[0], L.fr_k[1], L.fr_k[2]));
    if (L.fresnel == FRESNEL_DIELECTRIC) return sp1(fr_dielectric(cos_i, L.fr_a[0], L.fr_a[1]));
    return sp1(1.0f);
}

PB_D float2 concentric_sample_disk(float2 u) {
    float ox = u.x * 2.0f - 1.0f, oy = u.y * 2.0f - 1.0f;
    if (ox == 0.0f && oy == 0.0f) return make_float2(0.0f, 0.0f);
    float theta, r;
    if (fabsf(ox) > fabsf(oy)) { r = ox; theta = PB_PI_OVER_4 * (oy / ox); }
    else { r = oy; theta = PB_PI_OVER_2 - PB_PI_OVER_4 * (ox / oy); }
    float st, ct;
    sincos_rn(theta, st, ct);
    return make_float2(ct * r, st * r);
}
PB_D V3 cosine_sample_hemisphere(float2 u) {
    float2 d = concentric_sample_disk(u);
    float z = sqrtf(fmaxf(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    return mk3(d.x, d.y, z);
}

// ---- Trowbridge-Reitz -------------------------------------------------------------------------
PB_D float tr_d(float ax, float ay, V3 wh) {
    float t2 = tan2_theta(wh);
    if (isinf(t2)) return 0.0f;
    float cos4 = cos2_theta(wh) * cos2_theta(wh);
    float e = (cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay)) * t2;
    return 1.0f / (PB_PI * ax * ay * cos4 * (1.0f + e) * (1.0f + e));
}
PB_D float tr_lambda(float ax, float ay, V3 w) {
    float att = fabsf(tan_theta(w));
    if (isinf(att)) return 0.0f;
    float alpha = sqrtf(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
    float a2t2 = (alpha * att) * (alpha * att);
    return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
}
PB_D float tr_g1(float ax, float ay, V3 w) { return 1.0f / (1.0f + tr_lambda(ax, ay, w)); }
PB_D float tr_g(float ax, float ay, V3 wo, V3 wi) { return 1.0f / (1.0f + tr_lambda(ax, ay, wo) + tr_lambda(ax, ay, wi)); }
PB_D float tr_pdf(float ax, float ay, V3 wo, V3 wh) { return tr_d(ax, ay, wh) * tr_g1(ax, ay, wo) * absdot3(wo, wh) / abs_cos_theta(wo); }
PB_D void tr_sample_11(float cos_t, float u1, float u2, float& slope_x, float& slope_y) {
    if (cos_t > 0.9999f) {
        float r = sqrtf(u1 / (1.0f - u1));
        float phi = PB_TAU * u2;
        float sp, cp;
        sincos_rn(phi, sp, cp);
        slope_x = r * cp;
        slope_y = r * sp;
        return;
    }
    float sin_t = sqrtf(fmaxf(0.0f, 1.0f - cos_t * cos_t));
    float tan_t = sin_t / cos_t;
    float a = 1.0f / tan_t;
    float g1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
    a = 2.0f * u1 / g1 - 1.0f;
    float tmp = 1.0f / (a * a - 1.0f);
    if (tmp > 1e10f) tmp = 1e10f;
    float b = tan_t;
    float d = sqrtf(fmaxf(b * b * tmp * tmp - (a * a - b * b) * tmp, 0.0f));
    float s1 = b * tmp - d, s2 = b * tmp + d;
    slope_x = (a < 0.0f || s2 > 1.0f / tan_t) ? s1 : s2;
    float s, v;
    if (u2 > 0.5f) { s = 1.0f; v = 2.0f * (u2 - 0.5f); }
    else { s = -1.0f; v = 2.0f * (0.5f - u2); }
    float z = (v * (v * (v * 0.27385f - 0.73369f) + 0.46341f)) / (v * (v * (v * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    slope_y = s * z * sqrtf(1.0f + slope_x * slope_x);
}
PB_D V3 tr_sample(V3 wi, float ax, float ay, float u1, float u2) {
    V3 wis = norm3(mk3(ax * wi.x, ay * wi.y, wi.z));
    float sx = 0.0f, sy = 0.0f;
    tr_sample_11(cos_theta(wis), u1, u2, sx, sy);
    float tmp = cos_phi(wis) * sx - sin_phi(wis) * sy;
    sy = sin_phi(wis) * sx + cos_phi(wis) * sy;
    sx = tmp;
    sx *= ax;
    sy *= ay;
    return norm3(mk3(-sx, -sy, 1.0f));
}
PB_D V3 tr_sample_wh(float ax, float ay, V3 wo, float2 u) {
    if (wo.z < 0.0f) return -tr_sample(-wo, ax, ay, u.x, u.y);
    return tr_sample(wo, ax, ay, u.x, u.y);
}

// ---- lobes (local shading frame) ---------------------------------------------------------------
PB_D Sp lobe_r(const DLobe& L) { return mksp(L.r[0], L.r[1], L.r[2]); }
PB_D Sp lobe_t(const DLobe& L) { return mksp(L.t[0], L.t[1], L.t[2]); }
PB_D Sp lobe_sc(const DLobe& L) { const float* p = lobe_sc_slot(L); return mksp(p[0], p[1], p[2]); }

// SPEC: what the caller knows about the material at compile time (k_shade's specialised instantiations).
//   0        nothing: kinds, types and the number of lobes are read from the material;
//   1 + k    a SINGLE lobe of kind k (untextured matte with sigma = 0 / > 0, metal, substrate, mirror, smooth glass);
//   9        two lobes, Lambert then microfacet reflection (plastic, and an uber material that comes down to the same list).
// With SPEC != 0 lobe kind / type / count are constants, the switches and the loops over lobes fold away and what is left is the same
// arithmetic in the same order.  The per-lobe functions take the same number for ONE lobe: 0 = read L.kind, 1 + k = kind k.
PB_HD constexpr int pb_spec_type(int kind) {
    return kind == LOBE_SPEC_REFL ? (BSDF_REFLECTION | BSDF_SPECULAR)
         : kind == LOBE_SPEC_TRANS ? (BSDF_TRANSMISSION | BSDF_SPECULAR)
         : kind == LOBE_FRESNEL_SPEC ? (BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR)
         : (kind == LOBE_LAMBERT || kind == LOBE_OREN_NAYAR) ? (BSDF_DIFFUSE | BSDF_REFLECTION)
         : (kind == LOBE_MF_REFL || kind == LOBE_FRESNEL_BLEND) ? (BSDF_REFLECTION | BSDF_GLOSSY)
         : kind == LOBE_LAMBERT_TRANS ? (BSDF_DIFFUSE | BSDF_TRANSMISSION)
         : (BSDF_TRANSMISSION | BSDF_GLOSSY);
}
#define PB_SPEC_PLASTIC 9
PB_HD constexpr int pb_spec_n(int spec) { return spec == 0 ? -1 : (spec <= 1 + LOBE_FRESNEL_BLEND ? 1 : 2); }
PB_HD constexpr int pb_spec_lobe(int spec, int i) {  // lobe i of signature `spec`, as the per-lobe functions' template argument
    return spec == 0 ? 0 : (spec <= 1 + LOBE_FRESNEL_BLEND ? spec : (i == 0 ? 1 + LOBE_LAMBERT : 1 + LOBE_MF_REFL));
}
PB_HD constexpr bool pb_spec_has_nonspecular(int spec) {  // DMaterial::nonspecular > 0 for this signature
    return (pb_spec_type(pb_spec_lobe(spec, 0) - 1) & BSDF_SPECULAR) == 0 || (pb_spec_n(spec) > 1 && (pb_spec_type(pb_spec_lobe(spec, 1) - 1) & BSDF_SPECULAR) == 0);
}
#define PB_LOBE_KIND(L) (SPEC >= 1 ? (int)(SPEC - 1) : (L).kind)
#define PB_LOBE_TYPE(L) (SPEC >= 1 ? pb_spec_type(SPEC - 1) : (L).type)
#define PB_SPEC_OF_KIND(kind) ((kind) + 1)
// sc_opt (reflection.rs:714 ff.): `sc * <the unscaled expression>`, evaluated left to right, in f() of the non-specular lobes and in sample_f()
// of the specular ones.  A material with a scaled lobe is never in a specialised class (pbrt_gpu_scene_create), so SPEC != 0 code has no trace of
// it.  The non-specular lobes' sample_f scales f() a second time (reflection.rs:982-983, quirk Q9) -- and Bsdf::sample_f then replaces that value
// by the sum of f() over the matching lobes (reflection.rs:393-410), so it is not computed here.
#define PB_SCALED(L) (SPEC == 0 && (L).has_sc != 0)
template <int SPEC = 0>
PB_D Sp lobe_f(const DLobe& L, V3 wo, V3 wi) {
    switch (PB_LOBE_KIND(L)) {
        case LOBE_LAMBERT: return (PB_SCALED(L) ? lobe_sc(L) * lobe_r(L) : lobe_r(L)) * sp1(PB_INV_PI);
        case LOBE_LAMBERT_TRANS: return (PB_SCALED(L) ? lobe_sc(L) * lobe_t(L) : lobe_t(L)) * sp1(PB_INV_PI);  // reflection.rs:1010-1016
        case LOBE_OREN_NAYAR: {
            float sin_i = sin_theta(wi), sin_o = sin_theta(wo);
            float max_cos = 0.0f;
            if (sin_i > 1.0e-4f && sin_o > 1.0e-4f) {
                float d_cos = cos_phi(wi) * cos_phi(wo) + sin_phi(wi) * sin_phi(wo);
                max_cos = fmaxf(d_cos, 0.0f);
            }
            float sin_alpha, tan_beta;
            if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sin_o; tan_beta = sin_i / abs_cos_theta(wi); }
            else { sin_alpha = sin_i; tan_beta = sin_o / abs_cos_theta(wo); }
            return (PB_SCALED(L) ? lobe_sc(L) * lobe_r(L) : lobe_r(L)) * sp1(PB_INV_PI * (L.on_a + L.on_b * max_cos * sin_alpha * tan_beta));
        }
        case LOBE_MF_REFL: {
            float cos_o = abs_cos_theta(wo), cos_i = abs_cos_theta(wi);
            V3 wh = wi + wo;
            if (cos_i == 0.0f || cos_o == 0.0f) return sp1(0.0f);
            if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return sp1(0.0f);
            wh = norm3(wh);
            Sp F = fresnel_eval(L, dot3(wi, wh));
            return (PB_SCALED(L) ? lobe_sc(L) * lobe_r(L) : lobe_r(L)) * tr_d(L.alpha_x, L.alpha_y, wh) * tr_g(L.alpha_x, L.alpha_y, wo, wi) * F / (4.0f * cos_i * cos_o);
        }
        case LOBE_MF_TRANS: {
            if (same_hemisphere(wo, wi)) return sp1(0.0f);
            float cos_o = cos_theta(wo), cos_i = cos_theta(wi);
            if (cos_o == 0.0f || cos_i == 0.0f) return sp1(0.0f);
            float eta = (cos_o > 0.0f) ? (L.eta_b / L.eta_a) : (L.eta_a / L.eta_b);
            V3 wh = norm3(wo + wi * eta);
            if (wh.z < 0.0f) wh = -wh;
            if (dot3(wo, wh) * dot3(wi, wh) > 0.0f) return sp1(0.0f);
            Sp F = sp1(fr_dielectric(dot3(wo, wh), L.eta_a, L.eta_b));
            float sqrt_denom = dot3(wo, wh) + eta * dot3(wi, wh);
            float factor = 1.0f / eta;
            return (PB_SCALED(L) ? lobe_sc(L) * (sp1(1.0f) - F) : sp1(1.0f) - F) * lobe_t(L) *
                   fabsf(tr_d(L.alpha_x, L.alpha_y, wh) * tr_g(L.alpha_x, L.alpha_y, wo, wi) * eta * eta * absdot3(wi, wh) * absdot3(wo, wh) *
                         factor * factor / (cos_i * cos_o * sqrt_denom * sqrt_denom));
        }
        case LOBE_FRESNEL_BLEND: {
            Sp rd = lobe_r(L), rs = lobe_t(L);
            Sp diffuse = rd * (sp1(1.0f) - rs) * (28.0f / (23.0f * PB_PI)) * (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wi))) *
                         (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wo)));
            V3 wh = wi + wo;
            if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return sp1(0.0f);
            wh = norm3(wh);
            Sp schlick = rs + (sp1(1.0f) - rs) * pow5(1.0f - dot3(wi, wh));
            Sp specular = schlick * (tr_d(L.alpha_x, L.alpha_y, wh) / (4.0f * fabsf(dot3(wi, wh)) * fmaxf(abs_cos_theta(wi), abs_cos_theta(wo))));
            return PB_SCALED(L) ? lobe_sc(L) * (diffuse + specular) : diffuse + specular;
        }
        default: return sp1(0.0f);  // specular lobes
    }
}

template <int SPEC = 0>
PB_D float lobe_pdf(const DLobe& L, V3 wo, V3 wi) {
    switch (PB_LOBE_KIND(L)) {
        case LOBE_SPEC_REFL: return 0.0f;
        case LOBE_SPEC_TRANS: case LOBE_FRESNEL_SPEC:  // sic: cosine pdf (reflection.rs:828-834, :938-944)
        case LOBE_LAMBERT: case LOBE_OREN_NAYAR:
            return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * PB_INV_PI : 0.0f;
        case LOBE_LAMBERT_TRANS: return !same_hemisphere(wo, wi) ? abs_cos_theta(wi) * PB_INV_PI : 0.0f;  // reflection.rs:1036-1042
        case LOBE_MF_REFL: {
            if (!same_hemisphere(wo, wi)) return 0.0f;
            V3 wh = norm3(wo + wi);
            return tr_pdf(L.alpha_x, L.alpha_y, wo, wh) / (4.0f * dot3(wo, wh));
        }
        case LOBE_MF_TRANS: {
            if (same_hemisphere(wo, wi)) return 0.0f;
            float eta = (cos_theta(wo) > 0.0f) ? (L.eta_b / L.eta_a) : (L.eta_a / L.eta_b);
            V3 wh = norm3(wo + wi * eta);
            float wo_wh = dot3(wo, wh), wi_wh = dot3(wi, wh);
            if (wo_wh * wi_wh > 0.0f) return 0.0f;
            float sqrt_denom = wo_wh + eta * wi_wh;
            float dwh_dwi = fabsf((eta * eta * wi_wh) / (sqrt_denom * sqrt_denom));
            return tr_pdf(L.alpha_x, L.alpha_y, wo, wh) * dwh_dwi;
        }
        default: {  // LOBE_FRESNEL_BLEND
            if (!same_hemisphere(wo, wi)) return 0.0f;
            V3 wh = norm3(wo + wi);
            float pdf_wh = tr_pdf(L.alpha_x, L.alpha_y, wo, wh);
            return 0.5f * (abs_cos_theta(wi) * PB_INV_PI + pdf_wh / (4.0f * dot3(wo, wh)));
        }
    }
}

// sampled_type is only rewritten by FresnelSpecular, and only when non-zero on entry
template <int SPEC = 0>
PB_D Sp lobe_sample_f(const DLobe& L, V3 wo, V3& wi, float2 u, float& pdf, int& sampled_type) {
    switch (PB_LOBE_KIND(L)) {
        case LOBE_SPEC_REFL: {
            wi = mk3(-wo.x, -wo.y, wo.z);
            pdf = 1.0f;
            const Sp Fr = fresnel_eval(L, cos_theta(wi));
            return (PB_SCALED(L) ? lobe_sc(L) * Fr : Fr) * lobe_r(L) / abs_cos_theta(wi);  // reflection.rs:739-744
        }
        case LOBE_SPEC_TRANS: {
            bool entering = cos_theta(wo) > 0.0f;
            float eta_i = entering ? L.eta_a : L.eta_b, eta_t = entering ? L.eta_b : L.eta_a;
            if (!refract3(wo, faceforward3(mk3(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, wi)) return sp1(0.0f);
            pdf = 1.0f;
            Sp ft = lobe_t(L) * (sp1(1.0f) - sp1(fr_dielectric(cos_theta(wi), L.eta_a, L.eta_b)));
            ft = ft * sp1((eta_i * eta_i) / (eta_t * eta_t));
            return (PB_SCALED(L) ? lobe_sc(L) * ft : ft) / abs_cos_theta(wi);  // reflection.rs:822-826
        }
        case LOBE_FRESNEL_SPEC: {
            float F = fr_dielectric(cos_theta(wo), L.eta_a, L.eta_b);
            if (u.x < F) {
                wi = mk3(-wo.x, -wo.y, wo.z);
                if (sampled_type != 0) sampled_type = BSDF_REFLECTION | BSDF_SPECULAR;
                pdf = F;
                return (PB_SCALED(L) ? lobe_sc(L) * lobe_r(L) : lobe_r(L)) * F / abs_cos_theta(wi);  // reflection.rs:894-898
            }
            bool entering = cos_theta(wo) > 0.0f;
            float eta_i = entering ? L.eta_a : L.eta_b, eta_t = entering ? L.eta_b : L.eta_a;
            if (!refract3(wo, faceforward3(mk3(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, wi)) return sp1(0.0f);
            Sp ft = lobe_t(L) * (1.0f - F);
            ft = ft * sp1((eta_i * eta_i) / (eta_t * eta_t));
            if (sampled_type != 0) sampled_type = BSDF_TRANSMISSION | BSDF_SPECULAR;
            pdf = 1.0f - F;
            return (PB_SCALED(L) ? lobe_sc(L) * ft : ft) / abs_cos_theta(wi);  // reflection.rs:931-935
        }
        case LOBE_LAMBERT: case LOBE_OREN_NAYAR: {
            wi = cosine_sample_hemisphere(u);
            if (wo.z < 0.0f) wi.z *= -1.0f;
            pdf = lobe_pdf<SPEC>(L, wo, wi);
            return lobe_f<SPEC>(L, wo, wi);
        }
        case LOBE_LAMBERT_TRANS: {  // reflection.rs:1017-1035
            wi = cosine_sample_hemisphere(u);
            if (wo.z > 0.0f) wi.z *= -1.0f;
            pdf = lobe_pdf<SPEC>(L, wo, wi);
            return lobe_f<SPEC>(L, wo, wi);
        }
        case LOBE_MF_REFL: {
            if (wo.z == 0.0f) return sp1(0.0f);
            V3 wh = tr_sample_wh(L.alpha_x, L.alpha_y, wo, u);
            wi = reflect3(wo, wh);
            if (!same_hemisphere(wo, wi)) return sp1(0.0f);
            pdf = tr_pdf(L.alpha_x, L.alpha_y, wo, wh) / (4.0f * dot3(wo, wh));
            return lobe_f<SPEC>(L, wo, wi);
        }
        case LOBE_MF_TRANS: {
            if (wo.z == 0.0f) return sp1(0.0f);
            V3 wh = tr_sample_wh(L.alpha_x, L.alpha_y, wo, u);
            float eta = (cos_theta(wo) > 0.0f) ? (L.eta_a / L.eta_b) : (L.eta_b / L.eta_a);
            if (refract3(wo, wh, eta, wi)) { pdf = lobe_pdf<SPEC>(L, wo, wi); return lobe_f<SPEC>(L, wo, wi); }
            return sp1(0.0f);
        }
        default: {  // LOBE_FRESNEL_BLEND
            if (u.x < 0.5f) {
                u.x = fminf(2.0f * u.x, PB_ONE_MINUS_EPSILON);
                wi = cosine_sample_hemisphere(u);
                if (wo.z < 0.0f) wi.z *= -1.0f;
            } else {
                u.x = fminf(2.0f * (u.x - 0.5f), PB_ONE_MINUS_EPSILON);
                V3 wh = tr_sample_wh(L.alpha_x, L.alpha_y, wo, u);
                wi = reflect3(wo, wh);
                if (!same_hemisphere(wo, wi)) return sp1(0.0f);
            }
            pdf = lobe_pdf<SPEC>(L, wo, wi);
            return lobe_f<SPEC>(L, wo, wi);
        }
    }
}

// ---- Bsdf ---------------------------------------------------------------------------------------
struct BsdfFrame {
    V3 ns, ng, ss, ts;
    const DMaterial* mat;
};
template <int SPEC = 0>
PB_D bool lobe_matches(const DLobe& L, int flags) { return (PB_LOBE_TYPE(L) & flags) == PB_LOBE_TYPE(L); }
PB_D V3 to_local(const BsdfFrame& B, V3 v) { return mk3(dot3(v, B.ss), dot3(v, B.ts), dot3(v, B.ns)); }
PB_D V3 to_world(const BsdfFrame& B, V3 v) {
    return mk3(B.ss.x * v.x + B.ts.x * v.y + B.ns.x * v.z, B.ss.y * v.x + B.ts.y * v.y + B.ns.y * v.z, B.ss.z * v.x + B.ts.z * v.y + B.ns.z * v.z);
}
template <int N> struct LobeTag { static constexpr int value = N; };
// body(tag, lobe, index) for every lobe of the material, in order; tag.value is the lobe's compile-time kind number (0: not known)
template <int SPEC, typename F>
PB_D void for_each_lobe(const BsdfFrame& B, F&& body) {
    if constexpr (SPEC == 0) {
        for (int i = 0; i < B.mat->n_lobes; ++i) body(LobeTag<0>{}, B.mat->lobes[i], i);
    } else {
        body(LobeTag<pb_spec_lobe(SPEC, 0)>{}, B.mat->lobes[0], 0);
        if constexpr (pb_spec_n(SPEC) > 1) body(LobeTag<pb_spec_lobe(SPEC, 1)>{}, B.mat->lobes[1], 1);
    }
}
template <int SPEC = 0>
PB_D int bsdf_num_components(const BsdfFrame& B, int flags) {
    int n = 0;
    if constexpr (SPEC == 0) {  // (plain loops for the general instantiation: the lambdas cost it 176 B of stack per thread)
        for (int i = 0; i < B.mat->n_lobes; ++i) n += lobe_matches<0>(B.mat->lobes[i], flags) ? 1 : 0;
    } else for_each_lobe<SPEC>(B, [&](auto tag, const DLobe& L, int) { n += lobe_matches<decltype(tag)::value>(L, flags) ? 1 : 0; });
    return n;
}
template <int SPEC = 0>
PB_D Sp bsdf_sum_f(const BsdfFrame& B, V3 wo_w, V3 wi_w, V3 wo, V3 wi, int flags) {
    bool refl = dot3(wi_w, B.ng) * dot3(wo_w, B.ng) > 0.0f;
    Sp f = sp1(0.0f);
    if constexpr (SPEC == 0) {
        for (int i = 0; i < B.mat->n_lobes; ++i) {
            const DLobe& L = B.mat->lobes[i];
            if (lobe_matches<0>(L, flags) && ((refl && (L.type & BSDF_REFLECTION)) || (!refl && (L.type & BSDF_TRANSMISSION)))) f = f + lobe_f<0>(L, wo, wi);
        }
    } else
        for_each_lobe<SPEC>(B, [&](auto tag, const DLobe& L, int) {
            constexpr int K = decltype(tag)::value;
            const int type = pb_spec_type(K - 1);
            if (lobe_matches<K>(L, flags) && ((refl && (type & BSDF_REFLECTION)) || (!refl && (type & BSDF_TRANSMISSION)))) f = f + lobe_f<K>(L, wo, wi);
        });
    return f;
}
template <int SPEC = 0>
PB_D Sp bsdf_f(const BsdfFrame& B, V3 wo_w, V3 wi_w, int flags) {
    V3 wi = to_local(B, wi_w), wo = to_local(B, wo_w);
    if (wo.z == 0.0f) return sp1(0.0f);
    return bsdf_sum_f<SPEC>(B, wo_w, wi_w, wo, wi, flags);
}
template <int SPEC = 0>
PB_D float bsdf_pdf(const BsdfFrame& B, V3 wo_w, V3 wi_w, int flags) {
    if (SPEC == 0 && B.mat->n_lobes == 0) return 0.0f;
    V3 wo = to_local(B, wo_w), wi = to_local(B, wi_w);
    if (wo.z == 0.0f) return 0.0f;
    float pdf = 0.0f;
    int matching = 0;
    if constexpr (SPEC == 0) {
        for (int i = 0; i < B.mat->n_lobes; ++i) {
            const DLobe& L = B.mat->lobes[i];
            if (lobe_matches<0>(L, flags)) { ++matching; pdf += lobe_pdf<0>(L, wo, wi); }
        }
    } else
        for_each_lobe<SPEC>(B, [&](auto tag, const DLobe& L, int) {
            constexpr int K = decltype(tag)::value;
            if (lobe_matches<K>(L, flags)) { ++matching; pdf += lobe_pdf<K>(L, wo, wi); }
        });
    return matching > 0 ? fdiv0(pdf, (float)matching) : 0.0f;  // pdf is 0 for every wi below the horizon
}
// reflection.rs:298-420.  `pdf` is left untouched by the wo.z == 0 early-out, as in the reference.
template <int SPEC = 0>
PB_D Sp bsdf_sample_f(const BsdfFrame& B, V3 wo_w, V3& wi_w, float2 u, float& pdf, int flags, int& sampled_type) {
    int matching = bsdf_num_components<SPEC>(B, flags);
    if (matching == 0) { pdf = 0.0f; sampled_type = 0; return sp1(0.0f); }
    int ci = f2i_sat(floorf(u.x * (float)matching));
    ci = ci < 0 ? 0 : (ci > 255 ? 255 : ci);  // `as u8`
    int comp_i = min(ci, matching - 1);
    int index = -1, count = comp_i;
    if constexpr (SPEC == 0) {
        for (int i = 0; i < B.mat->n_lobes; ++i) {
            bool m = lobe_matches<0>(B.mat->lobes[i], flags);
            if (m && count == 0) { index = i; break; }
            if (m) --count;
        }
    } else
        for_each_lobe<SPEC>(B, [&](auto tag, const DLobe& L, int i) {
            if (index >= 0) return;
            const bool m = lobe_matches<decltype(tag)::value>(L, flags);
            if (m && count == 0) index = i;
            else if (m) --count;
        });
    if (index < 0) return sp1(0.0f);
    float2 ur = make_float2(fminf(u.x * (float)matching - (float)comp_i, PB_ONE_MINUS_EPSILON), u.y);
    V3 wi = mk3(0.0f, 0.0f, 0.0f);
    V3 wo = to_local(B, wo_w);
    if (wo.z == 0.0f) return sp1(0.0f);
    pdf = 0.0f;
    Sp f = sp1(0.0f);
    int type = 0;
    if constexpr (SPEC == 0) {
        const DLobe& L = B.mat->lobes[index];
        type = L.type;
        if (sampled_type != 0) sampled_type = type;
        f = lobe_sample_f<0>(L, wo, wi, ur, pdf, sampled_type);
    } else
        for_each_lobe<SPEC>(B, [&](auto tag, const DLobe& L, int i) {  // the chosen lobe
            constexpr int K = decltype(tag)::value;
            if (i != index) return;
            type = pb_spec_type(K - 1);
            if (sampled_type != 0) sampled_type = type;
            f = lobe_sample_f<K>(L, wo, wi, ur, pdf, sampled_type);
        });
    if (pdf == 0.0f) { if (sampled_type != 0) sampled_type = 0; return sp1(0.0f); }
    wi_w = to_world(B, wi);
    if (!(type & BSDF_SPECULAR) && matching > 1) {
        if constexpr (SPEC == 0) {
            for (int i = 0; i < B.mat->n_lobes; ++i)
                if (i != index && lobe_matches<0>(B.mat->lobes[i], flags)) pdf += lobe_pdf<0>(B.mat->lobes[i], wo, wi);
        } else
            for_each_lobe<SPEC>(B, [&](auto tag, const DLobe& L, int i) {
                constexpr int K = decltype(tag)::value;
                if (i != index && lobe_matches<K>(L, flags)) pdf += lobe_pdf<K>(L, wo, wi);
            });
    }
    if (matching > 1) pdf /= (float)matching;
    if (!(type & BSDF_SPECULAR)) f = bsdf_sum_f<SPEC>(B, wo_w, wi_w, wo, wi, flags);
    return f;
}

}  // namespace pb
