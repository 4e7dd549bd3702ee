// pb_texture.cuh -- ImageTexture<Spectrum>::evaluate on the device: UVMapping2D::map (texture.rs:101-121), MipMap::lookup
// (mipmap.rs:233-296: trilinear or EWA), SurfaceInteraction::compute_differentials (interaction.rs:388-474).
#pragma once
#include "pb_interaction.cuh"

namespace pb {

// MipMap::texel (mipmap.rs:208-232).  Repeat wraps ((s as usize) mod size, sizes are powers of two); Clamp clamps; Black answers the
// clamped texel as well (the reference's "TMP" branch), so it differs from Clamp only in the host-side resampling.
PB_D Sp tex_texel(const DTexture& T, int level, long long s, long long t) {
    const int us = max(1, T.w >> level), vs = max(1, T.h >> level);
    long long ss, tt;
    if (T.wrap == 0u) { ss = s & (long long)(us - 1); tt = t & (long long)(vs - 1); }
    else { ss = min(max(s, 0ll), (long long)us - 1); tt = min(max(t, 0ll), (long long)vs - 1); }
    const float4 v = __ldg(T.texels + T.off[level] + (size_t)tt * (size_t)us + (size_t)ss);
    return mksp(v.x, v.y, v.z);
}
PB_D Sp tex_triangle(const DTexture& T, int level, float2 st) {  // mipmap.rs:323-336
    level = min(max(level, 0), T.n_levels - 1);
    const int us = max(1, T.w >> level), vs = max(1, T.h >> level);
    const float s = st.x * (float)us - 0.5f, t = st.y * (float)vs - 0.5f;
    const float fs = floorf(s), ft = floorf(t);
    const long long s0 = (long long)fs, t0 = (long long)ft;
    const float ds = s - (float)s0, dt = t - (float)t0;
    const Sp tmp1 = tex_texel(T, level, s0 + 1, t0 + 1) * (ds * dt);
    const Sp tmp2 = tex_texel(T, level, s0 + 1, t0) * (ds * (1.0f - dt));
    const Sp tmp3 = tex_texel(T, level, s0, t0 + 1) * ((1.0f - ds) * dt);
    const Sp tmp4 = tex_texel(T, level, s0, t0) * ((1.0f - ds) * (1.0f - dt));
    return tmp4 + tmp3 + tmp2 + tmp1;
}
PB_D Sp tex_lookup_width(const DTexture& T, float2 st, float width) {  // lookup_pnt_flt, mipmap.rs:233-252
    const float nl = (float)T.n_levels;
    const float level = nl - 1.0f + log2_rn(fmaxf(width, 1e-8f));
    if (level < 0.0f) return tex_triangle(T, 0, st);
    if (level >= nl - 1.0f) return tex_texel(T, T.n_levels - 1, 0, 0);
    const int il = (int)floorf(level);
    const float delta = level - (float)il;
    return tex_triangle(T, il, st) * (1.0f - delta) + tex_triangle(T, il + 1, st) * delta;
}
// mipmap.rs:337-396.  The level is chosen from the minor axis and the major one is at most max_anisotropy times longer, so a footprint
// wider than PB_EWA_MAX_SPAN texels needs non-finite ellipse coefficients or a "maxanisotropy" in the thousands, where the
// reference's loop would not end in any useful time; the oracle and this code answer the level's first texel instead.
#define PB_EWA_MAX_SPAN 4096
PB_D Sp tex_ewa(const DTexture& T, const float* __restrict__ lut, int level, float2 st_in, float2 dst0, float2 dst1) {
    if (level >= T.n_levels) return tex_texel(T, T.n_levels - 1, 0, 0);
    const float us = (float)max(1, T.w >> level), vs = (float)max(1, T.h >> level);
    const float sx = st_in.x * us - 0.5f, sy = st_in.y * vs - 0.5f;
    dst0 = make_float2(dst0.x * us, dst0.y * vs);
    dst1 = make_float2(dst1.x * us, dst1.y * vs);
    float a = dst0.y * dst0.y + dst1.y * dst1.y + 1.0f;
    float b = -2.0f * (dst0.x * dst0.y + dst1.x * dst1.y);
    float c = dst0.x * dst0.x + dst1.x * dst1.x + 1.0f;
    const float inv_f = 1.0f / (a * c - b * b * 0.25f);
    a *= inv_f; b *= inv_f; c *= inv_f;
    const float det = -b * b + 4.0f * a * c;
    const float inv_det = 1.0f / det;
    const float u_sqrt = sqrtf(det * c), v_sqrt = sqrtf(a * det);
    const float fs0 = ceilf(sx - 2.0f * inv_det * u_sqrt), fs1 = floorf(sx + 2.0f * inv_det * u_sqrt);
    const float ft0 = ceilf(sy - 2.0f * inv_det * v_sqrt), ft1 = floorf(sy + 2.0f * inv_det * v_sqrt);
    if (!(fs1 - fs0 <= (float)PB_EWA_MAX_SPAN) || !(ft1 - ft0 <= (float)PB_EWA_MAX_SPAN)) return tex_texel(T, level, 0, 0);
    const long long s0 = (long long)fs0, s1 = (long long)fs1, t0 = (long long)ft0, t1 = (long long)ft1;
    Sp sum = sp1(0.0f);
    float sum_wts = 0.0f;
    for (long long it = t0; it <= t1; ++it) {
        const float tt = (float)it - sy;
        for (long long is = s0; is <= s1; ++is) {
            const float ss = (float)is - sx;
            const float r2 = a * ss * ss + b * ss * tt + c * tt * tt;
            if (r2 < 1.0f) {
                const int index = r2 <= 0.0f ? 0 : min((int)(r2 * 128.0f), 127);
                const float weight = __ldg(lut + index);
                sum = sum + tex_texel(T, level, is, it) * weight;
                sum_wts += weight;
            }
        }
    }
    return sum / sum_wts;
}
PB_D Sp tex_lookup(const DTexture& T, const float* __restrict__ lut, float2 st, float2 dst0, float2 dst1) {  // lookup_pnt_vec_vec, mipmap.rs:253-296
    if (T.trilinear) {
        const float width = fmaxf(fmaxf(fabsf(dst0.x), fabsf(dst0.y)), fmaxf(fabsf(dst1.x), fabsf(dst1.y)));
        return tex_lookup_width(T, st, width);
    }
    if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) { const float2 sw = dst0; dst0 = dst1; dst1 = sw; }
    const float major_length = sqrtf(dst0.x * dst0.x + dst0.y * dst0.y);
    float minor_length = sqrtf(dst1.x * dst1.x + dst1.y * dst1.y);
    if (minor_length * T.max_anisotropy < major_length && minor_length > 0.0f) {
        const float scale = major_length / (minor_length * T.max_anisotropy);
        dst1 = make_float2(dst1.x * scale, dst1.y * scale);
        minor_length *= scale;
    }
    if (minor_length == 0.0f) return tex_triangle(T, 0, st);
    const float lod = fmaxf(0.0f, (float)T.n_levels - 1.0f + log2_rn(minor_length));
    const int ilod = (int)floorf(lod);
    const Sp col2 = tex_ewa(T, lut, ilod + 1, st, dst0, dst1);
    const Sp col1 = tex_ewa(T, lut, ilod, st, dst0, dst1);
    const float tt = lod - (float)ilod;
    return col1 * (1.0f - tt) + col2 * tt;
}

// transform.rs:219-235
PB_D bool solve_2x2(float a00, float a01, float a10, float a11, float b0, float b1, float& x0, float& x1) {
    const float det = a00 * a11 - a01 * a10;
    if (fabsf(det) < 1e-10f) return false;
    x0 = (a11 * b0 - a01 * b1) / det;
    x1 = (a00 * b1 - a10 * b0) / det;
    if (x0 != x0 || x1 != x1) return false;
    return true;
}
struct UvDiff { float dudx, dvdx, dudy, dvdy; V3 dpdx, dpdy; };
PB_D float comp3(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// interaction.rs:388-474 for a ray that carries a differential
PB_D UvDiff compute_differentials(const Isect& is, V3 rx_o, V3 ry_o, V3 rx_d, V3 ry_d) {
    UvDiff r;
    r.dudx = r.dvdx = r.dudy = r.dvdy = 0.0f;
    r.dpdx = r.dpdy = mk3(0.0f, 0.0f, 0.0f);
    const V3 n = is.n, p = is.p;
    const float d = dot3(n, p);
    const float tx = -(dot3(n, rx_o) - d) / dot3(n, rx_d);
    if (isinf(tx) || tx != tx) return r;
    const V3 px = rx_o + rx_d * tx;
    const float ty = -(dot3(n, ry_o) - d) / dot3(n, ry_d);
    if (isinf(ty) || ty != ty) return r;
    const V3 py = ry_o + ry_d * ty;
    r.dpdx = px - p;
    r.dpdy = py - p;
    int d0, d1;
    if (fabsf(n.x) > fabsf(n.y) && fabsf(n.x) > fabsf(n.z)) { d0 = 1; d1 = 2; }
    else if (fabsf(n.y) > fabsf(n.z)) { d0 = 0; d1 = 2; }
    else { d0 = 0; d1 = 1; }
    const float a00 = comp3(is.dpdu, d0), a01 = comp3(is.dpdv, d0), a10 = comp3(is.dpdu, d1), a11 = comp3(is.dpdv, d1);
    const float bx0 = comp3(px, d0) - comp3(p, d0), bx1 = comp3(px, d1) - comp3(p, d1);
    const float by0 = comp3(py, d0) - comp3(p, d0), by1 = comp3(py, d1) - comp3(p, d1);
    if (!solve_2x2(a00, a01, a10, a11, bx0, bx1, r.dudx, r.dvdx)) { r.dudx = 0.0f; r.dvdx = 0.0f; }
    if (!solve_2x2(a00, a01, a10, a11, by0, by1, r.dudy, r.dvdy)) { r.dudy = 0.0f; r.dvdy = 0.0f; }
    return r;
}
PB_D V3 map_point(const float* m, V3 p) {  // Transform::transform_point
    const float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    const float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11], wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (wp == 1.0f) return mk3(xp, yp, zp);
    const float inv = 1.0f / wp;
    return mk3(inv * xp, inv * yp, inv * zp);
}
PB_D float2 map_sphere(const float* m, V3 p) {  // SphericalMapping2D::sphere (texture.rs:136-146)
    const V3 v = norm3(map_point(m, p) - mk3(0.0f, 0.0f, 0.0f));
    return make_float2(spherical_theta(v) * PB_INV_PI, spherical_phi(v) * PB_INV_2_PI);
}
PB_D float2 map_cylinder(const float* m, V3 p) {  // CylindricalMapping2D::cylinder (texture.rs:184-191)
    const V3 v = norm3(map_point(m, p) - mk3(0.0f, 0.0f, 0.0f));
    return make_float2(PB_PI + atan2_rn(v.y, v.x) * PB_INV_2_PI, v.z);
}
PB_D float map_fix(float d) { return d > 0.5f ? 1.0f - d : (d < -0.5f ? -(d + 1.0f) : d); }
PB_D Sp texture_evaluate_image(const DTexture& T, const float* __restrict__ lut, const Isect& is, const UvDiff& dd) {  // imagemap.rs:133-148
    float2 dstdx, dstdy, st;
    if (T.mapping == 1u) {         // SphericalMapping2D::map (texture.rs:148-172)
        st = map_sphere(T.map_m, is.p);
        const float delta = 0.1f, inv = 1.0f / delta;  // Vector2f / Float multiplies by the reciprocal (geometry.rs:1281-1288)
        const float2 sx = map_sphere(T.map_m, is.p + dd.dpdx * delta), sy = map_sphere(T.map_m, is.p + dd.dpdy * delta);
        dstdx = make_float2((sx.x - st.x) * inv, map_fix((sx.y - st.y) * inv));
        dstdy = make_float2((sy.x - st.x) * inv, map_fix((sy.y - st.y) * inv));
    } else if (T.mapping == 2u) {  // CylindricalMapping2D::map (texture.rs:193-215)
        st = map_cylinder(T.map_m, is.p);
        const float delta = 0.01f, inv = 1.0f / delta;
        const float2 sx = map_cylinder(T.map_m, is.p + dd.dpdx * delta), sy = map_cylinder(T.map_m, is.p + dd.dpdy * delta);
        dstdx = make_float2((sx.x - st.x) * inv, map_fix((sx.y - st.y) * inv));
        dstdy = make_float2((sy.x - st.x) * inv, map_fix((sy.y - st.y) * inv));
    } else if (T.mapping == 3u) {  // PlanarMapping2D::map (texture.rs:226-252)
        const V3 vs = mk3(T.map_m[0], T.map_m[1], T.map_m[2]), vt = mk3(T.map_m[3], T.map_m[4], T.map_m[5]);
        dstdx = make_float2(dot3(dd.dpdx, vs), dot3(dd.dpdx, vt));
        dstdy = make_float2(dot3(dd.dpdy, vs), dot3(dd.dpdy, vt));
        st = make_float2(T.du + dot3(is.p, vs), T.dv + dot3(is.p, vt));
    } else {                       // UVMapping2D::map (texture.rs:101-121)
        dstdx = make_float2(dd.dudx * T.su, dd.dvdx * T.sv);
        dstdy = make_float2(dd.dudy * T.su, dd.dvdy * T.sv);
        st = make_float2(is.uv.x * T.su + T.du, is.uv.y * T.sv + T.dv);
    }
    return tex_lookup(T, lut, st, dstdx, dstdy);
}
// Texture::evaluate over the small texture graph: constant.rs:17-20, scale.rs (tex1 * tex2), mix.rs (t1 * (1 - amt) + t2 * amt).
// Operands have lower indices than the node and the host bounds the depth (PBRT_MAX_TEXTURE_DEPTH), so a four-entry stack walks it
// in post order -- one call site for the image lookup instead of an inlined copy per path through the graph.
#define PB_TEX_STACK 4
__device__ PB_NOINLINE Sp texture_evaluate(const DTexture* __restrict__ all, uint32_t index, const float* __restrict__ lut, const Isect& is, const UvDiff& dd) {
    uint32_t node[PB_TEX_STACK];
    int next[PB_TEX_STACK];
    Sp val[PB_TEX_STACK][3];
    int sp = 0;
    node[0] = index; next[0] = 0;
    for (;;) {
        const DTexture& T = all[node[sp]];
        Sp result;
        if (T.kind == 0u) result = texture_evaluate_image(T, lut, is, dd);
        else if (T.kind == 1u) result = mksp(T.value[0], T.value[1], T.value[2]);
        else {
            const int nc = T.kind == 2u ? 2 : 3;
            if (next[sp] < nc && sp + 1 < PB_TEX_STACK) {  // descend into the next operand
                node[sp + 1] = T.child[next[sp]] - 1u;
                next[sp + 1] = 0;
                ++sp;
                continue;
            }
            if (T.kind == 2u) result = val[sp][0] * val[sp][1];
            else {
                const float amt = val[sp][2].r;
                result = val[sp][0] * sp1(1.0f - amt) + val[sp][1] * sp1(amt);
            }
        }
        if (sp == 0) return result;
        --sp;
        val[sp][next[sp]] = result;
        next[sp] += 1;
    }
}

}  // namespace pb
