// ORACLE -- TEST INFRASTRUCTURE ONLY (see o_math.hpp).
// o_geom.hpp: Bounds3f, Ray, Triangle, BVHAccel build + traversal.
#pragma once
#include <algorithm>
#include <vector>

#include "../include/pbrt_gpu.h"
#include "o_math.hpp"

namespace orc {

// src/core/geometry.rs:1980-2078
struct Bounds3 {
    Point3 p_min, p_max;
    Bounds3() {  // Default: inverted box (geometry.rs:1993-2011)
        Float mn = -std::numeric_limits<float>::max(), mx = std::numeric_limits<float>::max();
        p_min = Point3(mx, mx, mx);
        p_max = Point3(mn, mn, mn);
    }
    Bounds3(const Point3& a, const Point3& b) : p_min(a), p_max(b) {}
    Vec3 diagonal() const { return p_max - p_min; }
    Float surface_area() const {
        Vec3 d = diagonal();
        Float r = d.x * d.y + d.x * d.z + d.y * d.z;
        return r + r;
    }
    int maximum_extent() const {
        Vec3 d = diagonal();
        if (d.x > d.y && d.x > d.z) return 0;
        if (d.y > d.z) return 1;
        return 2;
    }
    Vec3 offset(const Point3& p) const {
        Vec3 o = p - p_min;
        if (p_max.x > p_min.x) o.x /= p_max.x - p_min.x;
        if (p_max.y > p_min.y) o.y /= p_max.y - p_min.y;
        if (p_max.z > p_min.z) o.z /= p_max.z - p_min.z;
        return o;
    }
    Point3 lerp(const Point3& t) const {  // geometry.rs:2176-2182
        return Point3(orc::lerp(t.x, p_min.x, p_max.x), orc::lerp(t.y, p_min.y, p_max.y), orc::lerp(t.z, p_min.z, p_max.z));
    }
};
inline Bounds3 bnd_union(const Bounds3& b, const Point3& p) {
    return Bounds3(Point3(fmin_(b.p_min.x, p.x), fmin_(b.p_min.y, p.y), fmin_(b.p_min.z, p.z)),
                   Point3(fmax_(b.p_max.x, p.x), fmax_(b.p_max.y, p.y), fmax_(b.p_max.z, p.z)));
}
inline Bounds3 bnd_union(const Bounds3& a, const Bounds3& b) {
    return Bounds3(Point3(fmin_(a.p_min.x, b.p_min.x), fmin_(a.p_min.y, b.p_min.y), fmin_(a.p_min.z, b.p_min.z)),
                   Point3(fmax_(a.p_max.x, b.p_max.x), fmax_(a.p_max.y, b.p_max.y), fmax_(a.p_max.z, b.p_max.z)));
}

struct Ray {
    Point3 o;
    Vec3 d;
    mutable Float t_max;  // Cell<f32> in the reference
    Float time;
    // Option<RayDifferential> (geometry.rs:2408-2414): only camera rays carry one
    bool has_differential = false;
    Point3 rx_origin, ry_origin;
    Vec3 rx_direction, ry_direction;
    Ray() : t_max(INF), time(0) {}
    Ray(const Point3& o_, const Vec3& d_, Float tm = INF, Float t = 0.0f) : o(o_), d(d_), t_max(tm), time(t) {}
    void scale_differentials(Float s) {  // geometry.rs:2398-2405
        if (!has_differential) return;
        rx_origin = o + (rx_origin - o) * s;
        ry_origin = o + (ry_origin - o) * s;
        rx_direction = d + (rx_direction - d) * s;
        ry_direction = d + (ry_direction - d) * s;
    }
};

// Bounds3f::intersect_p(ray, inv_dir, dir_is_neg)  src/core/geometry.rs:2211-2268
inline bool bounds_intersect_p(const Float* pmin, const Float* pmax, const Ray& ray, const Vec3& inv_dir, const int dir_is_neg[3]) {
    const Float g = 1.0f + 2.0f * gamma(3);
    Float t_min = ((dir_is_neg[0] ? pmax[0] : pmin[0]) - ray.o.x) * inv_dir.x;
    Float t_max = ((dir_is_neg[0] ? pmin[0] : pmax[0]) - ray.o.x) * inv_dir.x;
    Float ty_min = ((dir_is_neg[1] ? pmax[1] : pmin[1]) - ray.o.y) * inv_dir.y;
    Float ty_max = ((dir_is_neg[1] ? pmin[1] : pmax[1]) - ray.o.y) * inv_dir.y;
    t_max *= g;
    ty_max *= g;
    if (t_min > ty_max || ty_min > t_max) return false;
    if (ty_min > t_min) t_min = ty_min;
    if (ty_max < t_max) t_max = ty_max;
    Float tz_min = ((dir_is_neg[2] ? pmax[2] : pmin[2]) - ray.o.z) * inv_dir.z;
    Float tz_max = ((dir_is_neg[2] ? pmin[2] : pmax[2]) - ray.o.z) * inv_dir.z;
    tz_max *= g;
    if (t_min > tz_max || tz_min > t_max) return false;
    if (tz_min > t_min) t_min = tz_min;
    if (tz_max < t_max) t_max = tz_max;
    return (t_min < ray.t_max) && (t_max > 0.0f);
}

// ---------------------------------------------------------------------------------------------
// Scene storage (copied from the flat PbrtSceneDesc)
struct Mesh {
    std::vector<Float> p, n, s, uv;
    bool reverse_orientation, swaps_handedness;
    uint32_t alpha = 0, shadow_alpha = 0;  // TriangleMesh.alpha_mask / shadow_alpha_mask as 1 + float texture index (triangle.rs:39-40)
    Point3 P(uint32_t i) const { return Point3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
    Normal3 N(uint32_t i) const { return Normal3(n[3 * i], n[3 * i + 1], n[3 * i + 2]); }
    Vec3 S(uint32_t i) const { return Vec3(s[3 * i], s[3 * i + 1], s[3 * i + 2]); }
    Vec2 UV(uint32_t i) const { return Vec2(uv[2 * i], uv[2 * i + 1]); }
};

// The part of Triangle::intersect shared by intersect / intersect_p:
// watertight test + conservative t bound.  src/shapes/triangle.rs:134-273 (== :450-591)
struct TriHit { Float t, b0, b1, b2; };
inline bool triangle_test(const Point3& p0, const Point3& p1, const Point3& p2, const Ray& ray, TriHit& h) {
    Point3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
    int kz = max_dimension(vabs(ray.d));
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    Vec3 d = permute(ray.d, kx, ky, kz);
    p0t = permute(p0t, kx, ky, kz);
    p1t = permute(p1t, kx, ky, kz);
    p2t = permute(p2t, kx, ky, kz);
    Float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
    p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
    p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
    Float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    Float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    Float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // f64 fallback :189-200
        double p2txp1ty = (double)p2t.x * (double)p1t.y;
        double p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (Float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y;
        double p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (Float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y;
        double p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (Float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    Float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0t.z *= sz; p1t.z *= sz; p2t.z *= sz;
    Float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if ((det < 0.0f && (t_scaled >= 0.0f || t_scaled < ray.t_max * det)) ||
        (det > 0.0f && (t_scaled <= 0.0f || t_scaled > ray.t_max * det)))
        return false;
    Float inv_det = 1.0f / det;
    Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    Float t = t_scaled * inv_det;
    Float max_zt = max_component(vabs(Vec3(p0t.z, p1t.z, p2t.z)));
    Float delta_z = gamma(3) * max_zt;
    Float max_xt = max_component(vabs(Vec3(p0t.x, p1t.x, p2t.x)));
    Float max_yt = max_component(vabs(Vec3(p0t.y, p1t.y, p2t.y)));
    Float delta_x = gamma(5) * (max_xt + max_zt);
    Float delta_y = gamma(5) * (max_yt + max_zt);
    Float delta_e = 2.0f * (gamma(2) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
    Float max_e = max_component(vabs(Vec3(e0, e1, e2)));
    Float delta_t = 3.0f * (gamma(3) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * std::fabs(inv_det);
    if (t <= delta_t) return false;
    h.t = t; h.b0 = b0; h.b1 = b1; h.b2 = b2;
    return true;
}

// ---------------------------------------------------------------------------------------------
// BVHAccel::new / recursive_build / flatten_bvh_tree   src/accelerators/bvh.rs:96-152,178-392
struct BuildNode {
    Bounds3 bounds;
    int child[2];
    int split_axis, first_prim_offset, n_primitives;
};
struct PrimInfo { uint32_t primitive_number; Bounds3 bounds; Point3 centroid; };

struct BvhBuilder {
    std::vector<PrimInfo> info;
    std::vector<BuildNode> arena;
    std::vector<uint32_t> ordered;
    size_t max_prims_in_node;
    int total_nodes = 0;

    int make_leaf(int node, size_t start, size_t end, const Bounds3& bounds) {
        int first = (int)ordered.size();
        for (size_t i = start; i < end; ++i) ordered.push_back(info[i].primitive_number);
        arena[node].first_prim_offset = first;
        arena[node].n_primitives = (int)(end - start);
        arena[node].bounds = bounds;
        arena[node].child[0] = arena[node].child[1] = -1;
        return node;
    }
    static size_t bucket_of(const Bounds3& cb, const Point3& c, int dim) {
        const size_t n_buckets = 12;
        int32_t bi = f2i((Float)n_buckets * cb.offset(c)[dim]);  // Rust `as usize` saturates at 0
        size_t b = bi < 0 ? 0 : (size_t)bi;
        if (b == n_buckets) b = n_buckets - 1;
        return b;
    }
    int build(size_t start, size_t end) {
        int node = (int)arena.size();
        arena.push_back(BuildNode());
        total_nodes += 1;
        Bounds3 bounds;
        for (size_t i = start; i < end; ++i) bounds = bnd_union(bounds, info[i].bounds);
        size_t n = end - start;
        if (n == 1) return make_leaf(node, start, end, bounds);
        Bounds3 cb;
        for (size_t i = start; i < end; ++i) cb = bnd_union(cb, info[i].centroid);
        int dim = cb.maximum_extent();
        size_t mid = (start + end) / 2;
        if (cb.p_max[dim] == cb.p_min[dim]) return make_leaf(node, start, end, bounds);
        // SplitMethod::SAH (HLBVH is an alias; Middle/EqualCounts are unimplemented, quirk Q3)
        if (n <= 2) {
            mid = (start + end) / 2;
            if (start != end - 1 && info[end - 1].centroid[dim] < info[start].centroid[dim]) std::swap(info[start], info[end - 1]);
        } else {
            const size_t n_buckets = 12;
            size_t count[12] = {0};
            Bounds3 bb[12];
            for (size_t i = start; i < end; ++i) {
                size_t b = bucket_of(cb, info[i].centroid, dim);
                count[b] += 1;
                bb[b] = bnd_union(bb[b], info[i].bounds);
            }
            Float cost[11];
            for (size_t i = 0; i < n_buckets - 1; ++i) {
                Bounds3 b0, b1;
                size_t c0 = 0, c1 = 0;
                for (size_t j = 0; j <= i; ++j) { b0 = bnd_union(b0, bb[j]); c0 += count[j]; }
                for (size_t j = i + 1; j < n_buckets; ++j) { b1 = bnd_union(b1, bb[j]); c1 += count[j]; }
                cost[i] = 1.0f + ((Float)c0 * b0.surface_area() + (Float)c1 * b1.surface_area()) / bounds.surface_area();
            }
            Float min_cost = cost[0];
            size_t min_bucket = 0;
            for (size_t i = 0; i < n_buckets - 1; ++i)
                if (cost[i] < min_cost) { min_cost = cost[i]; min_bucket = i; }
            Float leaf_cost = (Float)n;
            if (n > max_prims_in_node || min_cost < leaf_cost) {
                // Iterator::partition is stable (bvh.rs:297-320)
                auto it = std::stable_partition(info.begin() + start, info.begin() + end,
                                                [&](const PrimInfo& pi) { return bucket_of(cb, pi.centroid, dim) <= min_bucket; });
                mid = (size_t)(it - info.begin());
            } else {
                return make_leaf(node, start, end, bounds);
            }
        }
        // the second child is built first, so its primitives come first in ordered_prims (bvh.rs:334-352)
        int c1 = build(mid, end);
        int c0 = build(start, mid);
        arena[node].n_primitives = 0;
        arena[node].bounds = bnd_union(arena[c0].bounds, arena[c1].bounds);
        arena[node].child[0] = c0;
        arena[node].child[1] = c1;
        arena[node].split_axis = dim;
        return node;
    }
    int flatten(int node, std::vector<PbrtBvhNode>& out, int& offset) {
        int my = offset++;
        const BuildNode& b = arena[node];
        PbrtBvhNode ln;
        std::memset(&ln, 0, sizeof ln);
        for (int k = 0; k < 3; ++k) { ln.pmin[k] = b.bounds.p_min[k]; ln.pmax[k] = b.bounds.p_max[k]; }
        if (b.n_primitives > 0) {
            ln.offset = b.first_prim_offset;
            ln.n_prims = (uint16_t)b.n_primitives;
            ln.axis = 0;
            out[my] = ln;
        } else {
            flatten(b.child[0], out, offset);
            int second = flatten(b.child[1], out, offset);
            ln.offset = second;
            ln.n_prims = 0;
            ln.axis = (uint8_t)b.split_axis;
            out[my] = ln;
        }
        return my;
    }
};

// bounds: n*6 floats (pmin, pmax) per primitive, in declaration order.
inline void bvh_build(const Float* bounds, uint32_t n, uint32_t max_prims_in_node, std::vector<PbrtBvhNode>& nodes,
                      std::vector<uint32_t>& ordered) {
    nodes.clear();
    ordered.clear();
    if (n == 0) return;
    BvhBuilder bl;
    bl.max_prims_in_node = std::min<size_t>(max_prims_in_node, 255);
    bl.info.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        const Float* b = bounds + 6 * (size_t)i;
        bl.info[i].primitive_number = i;
        bl.info[i].bounds = Bounds3(Point3(b[0], b[1], b[2]), Point3(b[3], b[4], b[5]));
        // BVHPrimitiveInfo::new: centroid = 0.5*p_min + 0.5*p_max   (bvh.rs:33-40)
        bl.info[i].centroid = bl.info[i].bounds.p_min * 0.5f + bl.info[i].bounds.p_max * 0.5f;
    }
    bl.arena.reserve(2 * (size_t)n);
    bl.ordered.reserve(n);
    int root = bl.build(0, n);
    nodes.resize(bl.total_nodes);
    int off = 0;
    bl.flatten(root, nodes, off);
    ordered.swap(bl.ordered);
}

}  // namespace orc
