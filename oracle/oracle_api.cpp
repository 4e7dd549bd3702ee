// ORACLE -- TEST INFRASTRUCTURE ONLY (see o_math.hpp).  C entry points for ctypes.
#include <cstdio>
#include <cstring>
#include <string>

#include "o_render.hpp"

using namespace orc;

namespace {
std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }

Scene* build_scene(const PbrtSceneDesc* d) {
    std::unique_ptr<Scene> sc(new Scene());
    sc->nodes.assign(d->nodes, d->nodes + d->n_nodes);
    sc->tris.assign(d->tris, d->tris + d->n_tris);
    sc->meshes.resize(d->n_meshes);
    for (uint32_t i = 0; i < d->n_meshes; ++i) {
        const PbrtMesh& m = d->meshes[i];
        Mesh& o = sc->meshes[i];
        o.p.assign(m.p, m.p + 3 * (size_t)m.n_verts);
        if (m.n) o.n.assign(m.n, m.n + 3 * (size_t)m.n_verts);
        if (m.s) o.s.assign(m.s, m.s + 3 * (size_t)m.n_verts);
        if (m.uv) o.uv.assign(m.uv, m.uv + 2 * (size_t)m.n_verts);
        o.reverse_orientation = m.reverse_orientation != 0;
        o.swaps_handedness = m.transform_swaps_handedness != 0;
        o.alpha = m.alpha; o.shadow_alpha = m.shadow_alpha;
        for (uint32_t a : {m.alpha, m.shadow_alpha})
            if (a && (a > d->n_textures || d->textures[a - 1].channels != 1)) return nullptr;
    }
    sc->materials.resize(d->n_materials);
    sc->material_src.assign(d->materials, d->materials + d->n_materials);
    sc->materials_single.resize(d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        if (!compile_material_at(d->materials, d->n_materials, i, sc->materials[i])) return nullptr;
        compile_material_at(d->materials, d->n_materials, i, sc->materials_single[i], false);
        for (int g = 0; g < PBRT_MAX_TEX_GROUPS; ++g) {
            const uint32_t t = d->materials[i].tex[g];
            int nv = 0;
            if (t && (t > d->n_textures || pbrt_material_tex_offset(d->materials[i].kind, g, &nv) < 0 || (uint32_t)nv != d->textures[t - 1].channels)) return nullptr;
        }
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const uint32_t b = d->materials[i].bump;
        if (b && (b > d->n_textures || d->textures[b - 1].channels != 1)) return nullptr;
    }
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const PbrtTexture& t = d->textures[i];
        if (t.channels != 1 && t.channels != 3) return nullptr;
        if (t.kind == PBRT_TEX_IMAGE) {
            if (!t.texels || t.res[0] == 0 || t.res[1] == 0 || t.wrap > PBRT_WRAP_CLAMP || t.mapping > PBRT_MAP_PLANAR) return nullptr;
            sc->textures.emplace_back(new ImageTexture(t));
        } else {
            if (t.kind > PBRT_TEX_MIX) return nullptr;
            const int nc = t.kind == PBRT_TEX_CONSTANT ? 0 : (t.kind == PBRT_TEX_SCALE ? 2 : 3);
            for (int c = 0; c < nc; ++c) {  // children: earlier textures of the right type
                if (t.child[c] == 0 || t.child[c] > i) return nullptr;
                if (d->textures[t.child[c] - 1].channels != (c == 2 ? 1u : t.channels)) return nullptr;
            }
            sc->textures.emplace_back(new ImageTexture(t, 0));
        }
    }
    sc->lights.resize(d->n_lights);
    for (uint32_t i = 0; i < d->n_lights; ++i) {
        const PbrtLight& l = d->lights[i];
        if (l.kind > PBRT_LIGHT_INFINITE || (l.kind == PBRT_LIGHT_DIFFUSE_AREA && l.tri >= d->n_tris)) return nullptr;
        if (l.kind == PBRT_LIGHT_INFINITE) {
            const uint32_t w = l.env_res[0], h = l.env_res[1];
            if (!l.env_texels || w == 0 || h == 0) return nullptr;
            sc->lights[i].env = std::make_shared<EnvLight>((int)w, (int)h, l.env_texels);
            for (int k = 0; k < 9; ++k) sc->lights[i].l2w[k] = l.l2w[k];
        }
        sc->lights[i].kind = (int)l.kind;
        sc->lights[i].p = Vec3(l.p[0], l.p[1], l.p[2]);
        for (int k = 0; k < 9; ++k) sc->lights[i].w2l[k] = l.w2l[k];
        sc->lights[i].cos_total_width = l.cos_total_width;
        sc->lights[i].cos_falloff_start = l.cos_falloff_start;
        sc->lights[i].l_emit = Spectrum(l.L[0], l.L[1], l.L[2]);
        sc->lights[i].tri = l.tri;
        sc->lights[i].two_sided = l.two_sided != 0;
        sc->lights[i].area = l.area;
        sc->lights[i].n_samples = l.n_samples ? l.n_samples : 1u;
    }
    if (d->n_instances) sc->instances.assign(d->instances, d->instances + d->n_instances);
    for (const PbrtTri& t : sc->tris)
        if (t.mesh == PBRT_MESH_INSTANCE && t.v[0] >= d->n_instances) return nullptr;
    sc->camera = d->camera;
    sc->world_bound = Bounds3(Point3(d->world_bound[0], d->world_bound[1], d->world_bound[2]),
                              Point3(d->world_bound[3], d->world_bound[4], d->world_bound[5]));
    return sc.release();
}
void fill_stats(PbrtStats* st, const Counters& c) {
    if (!st) return;
    std::memset(st, 0, sizeof *st);
    st->camera_rays = c.camera_rays;
    st->closest_rays = c.closest_rays;
    st->shadow_rays = c.shadow_rays;
    st->rays = c.closest_rays + c.shadow_rays;
    st->nodes_visited = c.nodes_visited;
    st->tris_tested = c.tris_tested;
    st->light_tri_tests = c.light_tri_tests;
}
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_init(const char* sobol_table_path) {
    try {
        if (!sobol_tables().loaded) sobol_tables().load(sobol_table_path);
    } catch (const std::exception& e) { return fail(e.what()); }
    return 0;
}

void* orc_scene_create(const PbrtSceneDesc* d) {
    Scene* s = build_scene(d);
    if (!s) g_err = "unsupported scene description";
    return s;
}
void orc_scene_destroy(void* s) { delete (Scene*)s; }
void orc_set_instancing(void* s, uint32_t mode) { ((Scene*)s)->instancing = mode; }

int orc_render(void* scene, const PbrtRenderParams* rp, const int32_t rect[4], float* film_rgbw, float* sample_rgb, int n_threads,
               PbrtStats* stats) {
    if (!sobol_tables().loaded) return fail("orc_init not called");
    try {
        Counters c;
        const Scene& sc = *(Scene*)scene;
        if (rp->integrator > PBRT_INTEGRATOR_WHITTED) return fail("unknown integrator");
        render(sc, *rp, rect, film_rgbw, sample_rgb, n_threads, &c);
        fill_stats(stats, c);
    } catch (const std::exception& e) { return fail(e.what()); }
    return 0;
}

int orc_intersect(void* scene, uint32_t n, const float* o, const float* d, const float* t_max, int32_t* prim, float* t, float* b,
                  PbrtStats* stats) {
    const Scene& sc = *(Scene*)scene;
    Counters c;
    for (uint32_t i = 0; i < n; ++i) {
        Ray ray(Point3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), t_max[i]);
        SurfaceInteraction si;
        Float th = 0.0f;
        if (sc.intersect(ray, si, &c, &th)) {
            prim[i] = si.prim; t[i] = th; b[3 * i] = si.b[0]; b[3 * i + 1] = si.b[1]; b[3 * i + 2] = si.b[2];
        } else {
            prim[i] = -1; t[i] = 0.0f; b[3 * i] = b[3 * i + 1] = b[3 * i + 2] = 0.0f;
        }
    }
    fill_stats(stats, c);
    return 0;
}

int orc_intersect_p(void* scene, uint32_t n, const float* o, const float* d, const float* t_max, uint8_t* occluded, PbrtStats* stats) {
    const Scene& sc = *(Scene*)scene;
    Counters c;
    for (uint32_t i = 0; i < n; ++i) {
        Ray ray(Point3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), t_max[i]);
        occluded[i] = sc.intersect_p(ray, &c) ? 1 : 0;
    }
    fill_stats(stats, c);
    return 0;
}

// Full surface interaction of a hit, for shading-geometry parity: out = p[3], p_error[3], n[3], shading_n[3], ss[3] (normalised dpdu), uv[2]
int orc_interaction(void* scene, int32_t prim, const float* o, const float* d, float* out17) {
    const Scene& sc = *(Scene*)scene;
    Ray ray(Point3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2]));
    Point3 p0, p1, p2;
    sc.tri_verts(sc.tris[prim], p0, p1, p2);
    TriHit h;
    if (!triangle_test(p0, p1, p2, ray, h)) return 1;
    SurfaceInteraction si;
    sc.fill_interaction(sc.tris[prim], ray, h, si);
    Vec3 ss = normalize(si.shading_dpdu);
    float v[17] = {si.common.p.x, si.common.p.y, si.common.p.z, si.common.p_error.x, si.common.p_error.y, si.common.p_error.z,
                   si.common.n.x, si.common.n.y, si.common.n.z, si.shading_n.x, si.shading_n.y, si.shading_n.z, ss.x, ss.y, ss.z, si.uv.x, si.uv.y};
    std::memcpy(out17, v, sizeof v);
    return 0;
}

// BVHAccel::new: nodes_out must hold 2*n entries, ordered_out n entries.
int orc_bvh_build(const float* bounds, uint32_t n, uint32_t max_prims_in_node, PbrtBvhNode* nodes_out, uint32_t* n_nodes_out,
                  uint32_t* ordered_out) {
    std::vector<PbrtBvhNode> nodes;
    std::vector<uint32_t> ordered;
    bvh_build(bounds, n, max_prims_in_node, nodes, ordered);
    if (!nodes.empty()) std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(PbrtBvhNode));
    if (!ordered.empty()) std::memcpy(ordered_out, ordered.data(), ordered.size() * sizeof(uint32_t));
    *n_nodes_out = (uint32_t)nodes.size();
    return 0;
}

// ---- function-level known-answer entry points ------------------------------------------------
float orc_gamma(int n) { return gamma(n); }
float orc_next_float_up(float v) { return next_float_up(v); }
float orc_next_float_down(float v) { return next_float_down(v); }
void orc_offset_ray_origin(const float* p, const float* perr, const float* n, const float* w, float* out) {
    Point3 r = offset_ray_origin(Point3(p[0], p[1], p[2]), Vec3(perr[0], perr[1], perr[2]), Normal3(n[0], n[1], n[2]), Vec3(w[0], w[1], w[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
uint64_t orc_sobol_interval_to_index(uint32_t m, uint64_t frame, int32_t x, int32_t y) { return sobol_interval_to_index(m, frame, x, y); }
float orc_sobol_sample_float(int64_t a, int dim, uint32_t scramble) { return sobol_sample_float(a, dim, scramble); }
float orc_radical_inverse(int base_index, uint64_t a) { return radical_inverse(base_index, a); }
// GlobalSampler::sample_dimension of (pixel, sample, dim) for the sampler rp selects; also returns the global sample index
float orc_sampler_dimension(const PbrtRenderParams* rp, int32_t px, int32_t py, int64_t sample, int dim, uint64_t* index_out) {
    std::unique_ptr<Sampler> sp = make_sampler(*rp);
    sp->start_pixel(px, py);
    sp->set_sample_number(sample);
    if (index_out) *index_out = sp->interval_sample_index;
    return sp->sample_dimension(sp->interval_sample_index, dim);
}
// RADICAL_INVERSE_PERMUTATIONS slice of one dimension (prime p: p entries)
int orc_halton_permutation(int dim, uint16_t* out, int cap) {
    const uint32_t p = prime_tables().primes[dim];
    if ((int)p > cap) return -1;
    const uint16_t* src = radical_inverse_permutations().data() + prime_tables().sums[dim];
    for (uint32_t i = 0; i < p; ++i) out[i] = src[i];
    return (int)p;
}
uint32_t orc_prime(int i) { return prime_tables().primes[i]; }
uint32_t orc_prime_sum(int i) { return prime_tables().sums[i]; }
float orc_scrambled_radical_inverse(int base_index, uint64_t a, const uint16_t* perm) { return scrambled_radical_inverse(base_index, a, perm); }
// camera sample of (pixel, sample): out = p_film[2], time, p_lens[2], ray o[3], d[3]
void orc_camera_sample(void* scene, const PbrtRenderParams* rp, int32_t px, int32_t py, int64_t sample, float* out11) {
    const Scene& sc = *(Scene*)scene;
    std::unique_ptr<Sampler> sp = make_sampler(*rp);
    Sampler& s = *sp;
    s.start_pixel(px, py);
    s.set_sample_number(sample);
    Vec2 u = s.get_2d();
    Vec2 pf((Float)px + u.x, (Float)py + u.y);
    Float time = s.get_1d();
    Vec2 pl = s.get_2d();
    Ray r = camera_ray(sc.camera, pf, time, pl);
    float v[11] = {pf.x, pf.y, time, pl.x, pl.y, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z};
    std::memcpy(out11, v, sizeof v);
}
// Bsdf::f / pdf / sample_f of material `mat` in the frame (ns, ng, ss): all world vectors.
// out = f[3], pdf, sample f[3], sample pdf, wi[3], sampled_type
// (material `index` of an array: MixMaterial names its children by index)
int orc_bsdf_at(const PbrtMaterial* mats, uint32_t n_mats, uint32_t index, const float* ns, const float* ng, const float* ss_in, const float* wo,
                const float* wi, const float* u, int flags, float* out12) {
    MaterialLobes ml;
    if (!compile_material_at(mats, n_mats, index, ml)) return -1;
    Bsdf b;
    b.eta = ml.eta;
    b.ns = Normal3(ns[0], ns[1], ns[2]);
    b.ng = Normal3(ng[0], ng[1], ng[2]);
    b.ss = Vec3(ss_in[0], ss_in[1], ss_in[2]);
    b.ts = cross(b.ns, b.ss);
    b.bxdfs = &ml.bxdfs;
    Vec3 wov(wo[0], wo[1], wo[2]), wiv(wi[0], wi[1], wi[2]);
    Spectrum f = b.f(wov, wiv, flags);
    Float pdf = b.pdf(wov, wiv, flags);
    Vec3 wis;
    Float spdf = 0.0f;
    int st = 255;
    Spectrum sf = b.sample_f(wov, wis, Vec2(u[0], u[1]), spdf, flags, st);
    float v[12] = {f.c[0], f.c[1], f.c[2], pdf, sf.c[0], sf.c[1], sf.c[2], spdf, wis.x, wis.y, wis.z, (float)st};
    std::memcpy(out12, v, sizeof v);
    return 0;
}
int orc_bsdf(const PbrtMaterial* mat, const float* ns, const float* ng, const float* ss_in, const float* wo, const float* wi, const float* u,
             int flags, float* out12) {
    return orc_bsdf_at(mat, 1, 0, ns, ng, ss_in, wo, wi, u, flags, out12);
}
// Spatial (or uniform/power) light distribution at point p: writes n_lights func values then n_lights+1 cdf values, returns func_int
float orc_light_distribution(void* scene, int strategy, const float* p, float* func_out, float* cdf_out) {
    const Scene& sc = *(Scene*)scene;
    LightDistribution ld(&sc, strategy);
    const Distribution1D* d = ld.lookup(Point3(p[0], p[1], p[2]));
    for (size_t i = 0; i < d->func.size(); ++i) func_out[i] = d->func[i];
    for (size_t i = 0; i < d->cdf.size(); ++i) cdf_out[i] = d->cdf[i];
    return d->func_int;
}
// MipMap::lookup of an image texture (after its UVMapping2D): n lookups at st[2i..] with differentials dst0[2i..], dst1[2i..] -> rgb[3i..]
int orc_texture_lookup(const PbrtTexture* t, uint32_t n, const float* st, const float* dst0, const float* dst1, float* rgb) {
    if (!t || !t->texels || t->res[0] == 0 || t->res[1] == 0 || (t->channels != 1 && t->channels != 3)) return fail("bad texture");
    ImageTexture tex(*t);
    for (uint32_t i = 0; i < n; ++i) {
        Spectrum v = tex.mipmap.lookup(Vec2(st[2 * i], st[2 * i + 1]), Vec2(dst0[2 * i], dst0[2 * i + 1]), Vec2(dst1[2 * i], dst1[2 * i + 1]));
        rgb[3 * i] = v.c[0]; rgb[3 * i + 1] = v.c[1]; rgb[3 * i + 2] = v.c[2];
    }
    return 0;
}
// level `level` of the texture's MIP pyramid: writes us*vs*3 floats when rgb != NULL; returns us | vs << 16, or -1 past the last level
int orc_texture_level(const PbrtTexture* t, uint32_t level, float* rgb) {
    ImageTexture tex(*t);
    if (level >= tex.mipmap.levels()) return -1;
    const MipMapRGB::Level& l = tex.mipmap.pyramid[level];
    if (rgb) for (size_t i = 0; i < l.t.size(); ++i) { rgb[3 * i] = l.t[i].c[0]; rgb[3 * i + 1] = l.t[i].c[1]; rgb[3 * i + 2] = l.t[i].c[2]; }
    return l.us | (l.vs << 16);
}
// film.add_sample of one sample onto a film of cropped_pixel_bounds
void orc_film_add_sample(const PbrtRenderParams* rp, float* film_rgbw, const float* p_film, const float* rgb, float weight) {
    film_add_sample(*rp, film_rgbw, Vec2(p_film[0], p_film[1]), Spectrum(rgb[0], rgb[1], rgb[2]), weight);
}

}  // extern "C"
