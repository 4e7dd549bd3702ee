// pb_material.cuh -- material -> lobe list: src/materials/*.rs compute_scattering_functions for the in-scope kinds (MixMaterial: compile_mix below).
// Host code runs it once per material with constant textures; k_texture runs it per hit for materials with image textures
// (the lobe list depends on the texel: a black Kd drops its lobe).  alpha_u / alpha_v: the material's Trowbridge-Reitz alphas
// (roughness through roughness_to_alpha when "remaproughness"; microfacet.rs:243-255 needs logf, so the host supplies them).
#pragma once
#include <string.h>
#include "../../include/pbrt_gpu.h"
#include "pb_scene.cuh"

namespace pb {

PB_HD Sp sp3(const float* p) { return mksp(p[0], p[1], p[2]); }
PB_HD Sp clamp_pos(Sp s) { return mksp(clampf(s.r, 0.0f, INFINITY), clampf(s.g, 0.0f, INFINITY), clampf(s.b, 0.0f, INFINITY)); }
PB_HD DLobe blank_lobe(int kind) {
    DLobe l;
    memset(&l, 0, sizeof l);
    l.kind = kind;
    l.eta_a = l.eta_b = 1.0f;
    switch (kind) {
        case LOBE_SPEC_REFL: l.type = BSDF_REFLECTION | BSDF_SPECULAR; break;
        case LOBE_SPEC_TRANS: l.type = BSDF_TRANSMISSION | BSDF_SPECULAR; break;
        case LOBE_FRESNEL_SPEC: l.type = BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR; break;
        case LOBE_LAMBERT: case LOBE_OREN_NAYAR: l.type = BSDF_DIFFUSE | BSDF_REFLECTION; break;
        case LOBE_MF_REFL: case LOBE_FRESNEL_BLEND: l.type = BSDF_REFLECTION | BSDF_GLOSSY; break;
        case LOBE_LAMBERT_TRANS: l.type = BSDF_DIFFUSE | BSDF_TRANSMISSION; break;
        default: l.type = BSDF_TRANSMISSION | BSDF_GLOSSY; break;
    }
    return l;
}
PB_HD void set3(float* d, Sp s) { d[0] = s.r; d[1] = s.g; d[2] = s.b; }
PB_HD void set_tr(DLobe& l, float ax, float ay) {  // TrowbridgeReitzDistribution::new microfacet.rs:232-238
    l.alpha_x = fmaxf(ax, 0.001f);
    l.alpha_y = fmaxf(ay, 0.001f);
}
PB_HD void set_dielectric(DLobe& l, float ei, float et) { l.fresnel = FRESNEL_DIELECTRIC; l.fr_a[0] = ei; l.fr_a[1] = et; }

// TrowbridgeReitzDistribution::roughness_to_alpha (microfacet.rs:243-255) and the alphas of a material on the device (for a textured
// roughness; the host computes them with libm's logf, which log_rn restates bit for bit)
PB_D float roughness_to_alpha_dev(float roughness) {
    if (1e-3f > roughness) roughness = 1e-3f;
    const float x = log_rn(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
PB_D void material_alphas_dev(uint32_t kind, const float* p, float& au, float& av) {
    int iu = -1, iv = -1, ir = -1;
    switch (kind) {
        case PBRT_MAT_PLASTIC: iu = iv = 6; ir = 7; break;
        case PBRT_MAT_METAL: case PBRT_MAT_SUBSTRATE: iu = 6; iv = 7; ir = 8; break;
        case PBRT_MAT_GLASS: iu = 7; iv = 8; ir = 9; break;
        case PBRT_MAT_UBER: iu = 15; iv = 16; ir = 18; break;
        case PBRT_MAT_TRANSLUCENT: iu = iv = 12; ir = 13; break;
        default: break;
    }
    au = av = 0.0f;
    if (iu < 0) return;
    au = p[iu]; av = p[iv];
    if (p[ir] != 0.0f) { au = roughness_to_alpha_dev(au); av = roughness_to_alpha_dev(av); }
}

PB_HD bool compile_material_core(uint32_t kind, const float* p, float alpha_u, float alpha_v, DMaterial& out, bool allow_multiple_lobes = true) {
    memset(&out, 0, sizeof out);
    out.eta = 1.0f;
    int n = 0;
    auto push = [&](const DLobe& l) { if (n < PB_MAX_LOBES) out.lobes[n++] = l; };
    switch (kind) {
        case PBRT_MAT_MATTE: {  // matte.rs:43-86
            Sp r = clamp_pos(sp3(p));
            float sig = clampf(p[3], 0.0f, 90.0f);
            if (!is_black(r)) {
                DLobe l = blank_lobe(sig == 0.0f ? LOBE_LAMBERT : LOBE_OREN_NAYAR);
                set3(l.r, r);
                if (sig != 0.0f) {  // OrenNayar::new reflection.rs:1057-1066
                    float sigma = (PB_PI / 180.0f) * sig;
                    float sigma2 = sigma * sigma;
                    l.on_a = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.33f)));
                    l.on_b = 0.45f * sigma2 / (sigma2 + 0.09f);
                }
                push(l);
            }
            break;
        }
        case PBRT_MAT_PLASTIC: {  // plastic.rs:57-125
            Sp kd = clamp_pos(sp3(p)), ks = clamp_pos(sp3(p + 3));
            float rough = p[6];
            if (!is_black(kd)) { DLobe l = blank_lobe(LOBE_LAMBERT); set3(l.r, kd); push(l); }
            if (!is_black(ks)) {
                DLobe l = blank_lobe(LOBE_MF_REFL);
                set3(l.r, ks);
                set_dielectric(l, 1.5f, 1.0f);
                rough = alpha_u;
                set_tr(l, rough, rough);
                push(l);
            }
            break;
        }
        case PBRT_MAT_METAL: {  // metal.rs:144-205
            float ur = p[6], vr = p[7];
            ur = alpha_u; vr = alpha_v;
            DLobe l = blank_lobe(LOBE_MF_REFL);
            set3(l.r, sp1(1.0f));
            l.fresnel = FRESNEL_CONDUCTOR;
            set3(l.fr_a, sp3(p));
            set3(l.fr_k, sp3(p + 3));
            set_tr(l, ur, vr);
            push(l);
            break;
        }
        case PBRT_MAT_MIRROR: {  // mirror.rs:34-70
            DLobe l = blank_lobe(LOBE_SPEC_REFL);
            set3(l.r, clamp_pos(sp3(p)));
            l.fresnel = FRESNEL_NOOP;
            push(l);
            break;
        }
        case PBRT_MAT_GLASS: {  // glass.rs:83-211 with allow_multiple_lobes = true (path.rs:108)
            float ur = p[7], vr = p[8];
            Sp r = clamp_pos(sp3(p)), t = clamp_pos(sp3(p + 3));
            bool is_specular = ur == 0.0f && vr == 0.0f;
            float eta = p[6];
            out.eta = eta;
            if (is_specular && allow_multiple_lobes) {  // PathIntegrator (path.rs:108); Direct / Whitted pass false (directlighting.rs:77)
                DLobe l = blank_lobe(LOBE_FRESNEL_SPEC);
                set3(l.r, r); set3(l.t, t);
                l.eta_a = 1.0f; l.eta_b = eta;
                push(l);
            } else {
                ur = alpha_u; vr = alpha_v;
                if (!is_black(r)) {
                    DLobe l = blank_lobe(is_specular ? LOBE_SPEC_REFL : LOBE_MF_REFL);
                    set3(l.r, r); set_dielectric(l, 1.0f, eta);
                    if (!is_specular) set_tr(l, ur, vr);
                    push(l);
                }
                if (!is_black(t)) {
                    DLobe l = blank_lobe(is_specular ? LOBE_SPEC_TRANS : LOBE_MF_TRANS);
                    set3(l.t, t); l.eta_a = 1.0f; l.eta_b = eta;
                    if (!is_specular) set_tr(l, ur, vr);
                    push(l);
                }
            }
            break;
        }
        case PBRT_MAT_UBER: {  // uber.rs:114-259
            float e = p[17];
            Sp op = clamp_pos(sp3(p + 12));
            Sp t = clamp_pos(sp1(1.0f) - op);
            Sp kd = op * clamp_pos(sp3(p)), ks = op * clamp_pos(sp3(p + 3));
            float ur = p[15], vr = p[16];
            Sp kr = op * clamp_pos(sp3(p + 6)), kt = op * clamp_pos(sp3(p + 9));
            out.eta = is_black(t) ? e : 1.0f;
            if (!is_black(t)) { DLobe l = blank_lobe(LOBE_SPEC_TRANS); set3(l.t, t); l.eta_a = 1.0f; l.eta_b = 1.0f; push(l); }
            if (!is_black(kd)) { DLobe l = blank_lobe(LOBE_LAMBERT); set3(l.r, kd); push(l); }
            if (!is_black(ks)) {
                DLobe l = blank_lobe(LOBE_MF_REFL);
                set3(l.r, ks);
                set_dielectric(l, 1.0f, e);
                ur = alpha_u; vr = alpha_v;
                set_tr(l, ur, vr);
                push(l);
            }
            if (!is_black(kr)) { DLobe l = blank_lobe(LOBE_SPEC_REFL); set3(l.r, kr); set_dielectric(l, 1.0f, e); push(l); }
            if (!is_black(kt)) { DLobe l = blank_lobe(LOBE_SPEC_TRANS); set3(l.t, kt); l.eta_a = 1.0f; l.eta_b = e; push(l); }
            break;
        }
        case PBRT_MAT_SUBSTRATE: {  // substrate.rs:62-114
            Sp d = clamp_pos(sp3(p)), s = clamp_pos(sp3(p + 3));
            float ru = p[6], rv = p[7];
            if (!is_black(d) || !is_black(s)) {
                ru = alpha_u; rv = alpha_v;
                DLobe l = blank_lobe(LOBE_FRESNEL_BLEND);
                set3(l.r, d); set3(l.t, s);
                set_tr(l, ru, rv);
                push(l);
            }
            break;
        }
        case PBRT_MAT_TRANSLUCENT: {  // translucent.rs:48-189 (eta fixed at 1.5, TransportMode::Radiance)
            const float eta = 1.5f;
            out.eta = eta;
            Sp r = clamp_pos(sp3(p + 6)), t = clamp_pos(sp3(p + 9));
            if (is_black(r) && is_black(t)) break;
            Sp kd = clamp_pos(sp3(p)), ks = clamp_pos(sp3(p + 3));
            if (!is_black(kd)) {
                if (!is_black(r)) { DLobe l = blank_lobe(LOBE_LAMBERT); set3(l.r, r * kd); push(l); }
                if (!is_black(t)) { DLobe l = blank_lobe(LOBE_LAMBERT_TRANS); set3(l.t, t * kd); push(l); }
            }
            if (!is_black(ks) && (!is_black(r) || !is_black(t))) {
                const float rough = alpha_u;
                if (!is_black(r)) { DLobe l = blank_lobe(LOBE_MF_REFL); set3(l.r, r * ks); set_dielectric(l, 1.0f, eta); set_tr(l, rough, rough); push(l); }
                if (!is_black(t)) { DLobe l = blank_lobe(LOBE_MF_TRANS); set3(l.t, t * ks); l.eta_a = 1.0f; l.eta_b = eta; set_tr(l, rough, rough); push(l); }
            }
            break;
        }
        default: return false;
    }
    out.n_lobes = n;
    const int nonspec = BSDF_ALL & ~BSDF_SPECULAR;
    for (int i = 0; i < n; ++i)
        if ((out.lobes[i].type & nonspec) == out.lobes[i].type) out.nonspecular++;
    return true;
}

// Every lobe of a MixMaterial's child carries the Option<Spectrum> the mix handed down (matte.rs:74-82, plastic.rs:88-120, glass.rs:116-203, ...)
PB_HD void scale_lobes(DMaterial& m, Sp sc) {
    for (int i = 0; i < m.n_lobes; ++i) { m.lobes[i].has_sc = 1; set3(lobe_sc_slot(m.lobes[i]), sc); }
}
// MixMaterial::compute_scattering_functions (mixmat.rs:41-98) over two compiled children: `first` was built with scale s1 = clamp(amount),
// `second` with s2 = clamp(1 - s1); the second Bsdf's BxDFs are added to the first, whose eta (and shading frame) stay.  false: more
// lobes than a DMaterial holds.
PB_HD bool append_lobes(DMaterial& first, const DMaterial& second) {
    if (first.n_lobes + second.n_lobes > PB_MAX_LOBES) return false;
    for (int i = 0; i < second.n_lobes; ++i) first.lobes[first.n_lobes++] = second.lobes[i];
    first.nonspecular += second.nonspecular;
    return true;
}

}  // namespace pb
