"""Delta lights (PointLight, SpotLight, DistantLight -- src/lights/{point,spot,distant}.rs) through the
oracle and the host mirror, on the CPU.  With maxdepth = 1 a path's radiance is exactly the direct lighting of its first vertex, and
for a delta light on a Lambertian floor that has the closed form  Kd/pi * Li * cos(theta)  -- the analytic anchor the oracle's
estimate_direct delta branch (src/core/integrator.rs:470-480) is pinned to here."""
import math

import numpy as np
import pytest

from rs_pbrt_b200 import HostScene, _abi, scenes

KD = 0.5


def floor_scene(add_lights, occluder=False, spp=4, res=16, strategy="uniform"):
    """Matte floor (Kd) seen from above; maxdepth = 1 => radiance = direct lighting of the first vertex."""
    h = HostScene()
    m = h.material(_abi.MAT_MATTE, [KD, KD, KD, 0.0])
    add_lights(h)
    P = np.array([[-4, 0, -4], [4, 0, -4], [4, 0, 4], [-4, 0, 4]], np.float32)
    h.trianglemesh(np.array([0, 2, 1, 0, 3, 2], np.uint32), P, material=m)
    if occluder:  # a small quad hovering over the +x half
        Q = np.array([[0.5, 1, -1], [2.5, 1, -1], [2.5, 1, 1], [0.5, 1, 1]], np.float32)
        h.trianglemesh(np.array([0, 2, 1, 0, 3, 2], np.uint32), Q, material=m)
    h.look_at([0, 6, 0], [0, 0, 0], [0, 0, 1])
    h.film(res, res)
    h.camera(fov=50.0)
    h.sampler(spp)
    h.integrator(maxdepth=1, lightsamplestrategy=strategy)
    h.world_end()
    return h


def first_hits(h, orc):
    """Floor hit point of every camera sample (y = 0 plane, or the occluder at y = 1 when the ray crosses it)."""
    rp = h.params.contents
    sb = list(rp.sample_bounds)
    pts = np.zeros((sb[3] - sb[1], sb[2] - sb[0], rp.spp, 3))
    for y in range(sb[1], sb[3]):
        for x in range(sb[0], sb[2]):
            for s in range(rp.spp):
                c = orc.camera_sample(h.params, x, y, s).astype(np.float64)
                o, d = c[5:8], c[8:11]
                pts[y - sb[1], x - sb[0], s] = o + d * (-o[1] / d[1])
    return pts


def test_point_light_inverse_square(oracle):
    pl, I = np.array([0.5, 2.0, -0.25]), np.array([30.0, 20.0, 10.0])
    h = floor_scene(lambda h: h.light_point(pl, I))
    orc = oracle.OracleScene(h.desc)
    _, samples, st = orc.render(h.params, want_samples=True, n_threads=2)
    p = first_hits(h, orc)
    v = pl - p
    d2 = (v ** 2).sum(-1)
    expect = (KD / math.pi) * I[None, None, None, :] * (v[..., 1] / np.sqrt(d2) / d2)[..., None]
    assert np.allclose(samples, expect, rtol=2e-5, atol=1e-7)
    # one light: every first vertex casts one shadow ray and nothing else is sampled from the light
    assert st["shadow_rays"] == st["camera_rays"]


def test_point_light_scale_and_order(product_lib):
    h = HostScene()
    h.light_point([1, 2, 3], [2, 4, 8], scale=[0.5, 0.25, 2.0])
    m = h.material(_abi.MAT_MATTE, [KD, KD, KD, 0.0])
    P = np.array([[-1, 0, -1], [1, 0, -1], [1, 0, 1]], np.float32)
    h.trianglemesh(np.array([0, 1, 2], np.uint32), P, material=m, emit=[1, 1, 1])
    h.light_distant([0, 2, 0], [0, 0, 0], [3, 3, 3])
    h.trianglemesh(np.array([0, 1, 2], np.uint32), P + 1, material=m, emit=[2, 2, 2])
    h.light_spot([0, 5, 0], [0, 0, 0], [7, 7, 7], coneangle=40.0, conedeltaangle=10.0)
    h.look_at([0, 6, 0], [0, 0, 0], [0, 0, 1])
    h.film(8, 8)
    h.camera(fov=50.0)
    h.world_end()
    d = h.desc.contents
    L = [d.lights[i] for i in range(d.n_lights)]
    # scene.lights keeps declaration order: LightSource directives interleaved with the emissive shapes (api.rs)
    assert [l.kind for l in L] == [_abi.LIGHT_POINT, _abi.LIGHT_DIFFUSE_AREA, _abi.LIGHT_DISTANT, _abi.LIGHT_DIFFUSE_AREA, _abi.LIGHT_SPOT]
    assert list(L[0].L) == [1.0, 1.0, 16.0] and list(L[0].p) == [1.0, 2.0, 3.0]
    assert np.allclose(list(L[2].p), [0, 1, 0])
    assert [d.tris[L[i].tri].area_light for i in (1, 3)] == [1, 3]
    s = L[4]
    assert list(s.p) == [0.0, 5.0, 0.0]
    R = np.array(list(s.w2l), np.float64).reshape(3, 3)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)
    assert np.allclose(R[2], [0, -1, 0])  # the spot axis maps to +z in light space
    assert s.cos_total_width == pytest.approx(math.cos(math.radians(40.0)), abs=1e-7)
    assert s.cos_falloff_start == pytest.approx(math.cos(math.radians(30.0)), abs=1e-7)


def test_spot_light_cone_and_falloff(oracle):
    pl, to, I = np.array([0.0, 3.0, 0.0]), np.array([0.5, 0.0, 0.0]), np.array([40.0, 40.0, 40.0])
    cone, delta = 25.0, 10.0
    h = floor_scene(lambda h: h.light_spot(pl, to, I, coneangle=cone, conedeltaangle=delta), res=24)
    orc = oracle.OracleScene(h.desc)
    _, samples, _ = orc.render(h.params, want_samples=True, n_threads=2)
    p = first_hits(h, orc)
    v = pl - p
    d2 = (v ** 2).sum(-1)
    axis = (to - pl) / np.linalg.norm(to - pl)
    cos_t = (-(v / np.sqrt(d2)[..., None]) * axis).sum(-1)
    ct, cf = math.cos(math.radians(cone)), math.cos(math.radians(cone - delta))
    fall = np.clip((cos_t - ct) / (cf - ct), 0.0, 1.0) ** 4
    expect = (KD / math.pi) * I[None, None, None, :] * (fall * v[..., 1] / np.sqrt(d2) / d2)[..., None]
    edge = (np.abs(cos_t - ct) < 1e-5) | (np.abs(cos_t - cf) < 1e-5)
    assert (fall == 0).any() and (fall == 1).any() and ((fall > 0) & (fall < 1)).any()
    assert np.allclose(samples[~edge], expect[~edge], rtol=3e-4, atol=1e-6)


def test_distant_light_cosine_and_shadow(oracle):
    frm, Lr = np.array([1.0, 2.0, 0.0]), np.array([3.0, 2.0, 1.0])
    h = floor_scene(lambda h: h.light_distant(frm, [0, 0, 0], Lr), occluder=True, res=24)
    orc = oracle.OracleScene(h.desc)
    _, samples, _ = orc.render(h.params, want_samples=True, n_threads=2)
    rp = h.params.contents
    sb = list(rp.sample_bounds)
    w = frm / np.linalg.norm(frm)
    n_lit = n_shadow = 0
    for y in range(sb[1], sb[3]):
        for x in range(sb[0], sb[2]):
            for s in range(rp.spp):
                c = orc.camera_sample(h.params, x, y, s).astype(np.float64)
                o, d = c[5:8], c[8:11]
                prim, t, _, _ = orc.intersect(o, d)
                if prim[0] < 0:
                    assert not samples[y - sb[1], x - sb[0], s].any()
                    continue
                p = o + d * float(t[0])
                on_top = abs(p[1] - 1.0) < 1e-4
                # shadow: a floor point is dark iff the ray towards the light crosses the occluder at y = 1
                q = p + w * ((1.0 - p[1]) / w[1])
                shadowed = (not on_top) and 0.5 < q[0] < 2.5 and -1 < q[2] < 1
                near_edge = (not on_top) and (min(abs(q[0] - 0.5), abs(q[0] - 2.5)) < 1e-3 or min(abs(q[2] + 1), abs(q[2] - 1)) < 1e-3)
                if near_edge:
                    continue
                expect = np.zeros(3) if shadowed else (KD / math.pi) * Lr * w[1]
                assert np.allclose(samples[y - sb[1], x - sb[0], s], expect, rtol=2e-5, atol=1e-7)
                n_shadow += shadowed
                n_lit += not shadowed
    assert n_lit > 100 and n_shadow > 10


def test_delta_lights_in_all_light_distributions(oracle):
    """power(): 4 pi I (point), 2 pi I (1 - (cos_falloff + cos_total)/2) (spot), pi r^2 L (distant); the spatial distribution samples
    every light kind through sample_li (lightdistrib.rs:169-269)."""
    h = scenes.cornell_box(xres=16, yres=16, spp=4, lights="delta", strategy="power")
    orc = oracle.OracleScene(h.desc)
    d = h.desc.contents
    n = d.n_lights
    func, cdf, fi = orc.light_distribution(1, [278, 273, 279], n)
    lum = lambda c: 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]
    L = [d.lights[i] for i in range(n)]
    wb = np.array(list(d.world_bound), np.float64)
    r = np.linalg.norm((wb[3:] - wb[:3]) / 2)
    expect = []
    for l in L:
        c = np.array(list(l.L), np.float64)
        if l.kind == _abi.LIGHT_POINT:
            expect.append(lum(c) * 4 * math.pi)
        elif l.kind == _abi.LIGHT_SPOT:
            expect.append(lum(c) * 2 * math.pi * (1 - 0.5 * (l.cos_falloff_start + l.cos_total_width)))
        elif l.kind == _abi.LIGHT_DISTANT:
            expect.append(lum(c) * math.pi * r * r)
        else:
            expect.append(lum(c) * l.area * math.pi)
    assert np.allclose(func, expect, rtol=1e-5)
    func_s, cdf_s, _ = orc.light_distribution(2, [278, 100, 279], n)
    assert np.all(func_s > 0) and cdf_s[-1] == pytest.approx(1.0)
    for strat in ("uniform", "power", "spatial"):
        hh = scenes.cornell_box(xres=16, yres=16, spp=4, lights="delta", strategy=strat)
        film, samples, st = oracle.OracleScene(hh.desc).render(hh.params, want_samples=True, n_threads=4)
        assert np.isfinite(samples).all() and samples.mean() > 0.05


# ---- InfiniteAreaLight (src/lights/infinite.rs) ----------------------------------------------------------------------------
Y_UP = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], np.float32)  # light +z (the map's pole) -> world +y


def env_lookup_np(tex, phi, theta):
    """MipMap::triangle(0, st) with Repeat wrap, in float64 (mipmap.rs:323-336)."""
    h, w = tex.shape[:2]
    s = phi / (2 * np.pi) * w - 0.5
    t = theta / np.pi * h - 0.5
    s0, t0 = np.floor(s).astype(int), np.floor(t).astype(int)
    ds, dt = (s - s0)[..., None], (t - t0)[..., None]
    g = lambda a, b: tex[b % h, a % w].astype(np.float64)
    return g(s0, t0) * (1 - ds) * (1 - dt) + g(s0, t0 + 1) * (1 - ds) * dt + g(s0 + 1, t0) * ds * (1 - dt) + g(s0 + 1, t0 + 1) * ds * dt


def test_constant_infinite_light_empty_scene(oracle):
    """No geometry: every camera ray escapes and collects Le = the single texel (bilinear weights sum to 1 up to rounding)."""
    h = HostScene()
    h.light_infinite([0.5, 1.0, 2.0], scale=[2.0, 1.0, 0.5])
    h.look_at([0, 0, 0], [0, 0, 1], [0, 1, 0])
    h.film(8, 8)
    h.camera(fov=60.0)
    h.sampler(4)
    h.integrator(maxdepth=3)
    h.world_end()
    d = h.desc.contents
    assert d.n_lights == 1 and d.lights[0].kind == _abi.LIGHT_INFINITE and list(d.lights[0].env_res) == [1, 1]
    assert [d.lights[0].env_texels[k] for k in range(3)] == [1.0, 1.0, 1.0]
    _, samples, st = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=2)
    assert np.allclose(samples, 1.0, rtol=1e-6)
    assert st["rays"] == st["camera_rays"] == 8 * 8 * 4


def test_constant_sky_on_lambertian_floor_is_unbiased(oracle):
    """E[L] = Kd * L_sky for a diffuse floor under a uniform sky (white-furnace half): checks sample_li's pdf, pdf_li and the
    MIS weights of both strategies against each other."""
    sky = np.array([1.0, 2.0, 0.5])
    h = floor_scene(lambda h: h.light_infinite(sky), spp=64, res=16)
    _, samples, _ = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=4)
    assert np.allclose(samples.reshape(-1, 3).mean(0), KD * sky, rtol=0.02)


def test_image_infinite_light_matches_numerical_integral(oracle):
    rng = np.random.default_rng(5)
    hh, ww = 16, 32
    tex = (rng.random((hh, ww, 3)) ** 3 * 4).astype(np.float32)
    tex[:3, 5:9] += 30.0  # a bright patch near the pole (world +y): importance sampling matters
    h = floor_scene(lambda h: h.light_infinite([1, 1, 1], texels=tex, light_to_world=Y_UP), spp=256, res=12)
    d = h.desc.contents
    assert list(d.lights[0].env_res) == [ww, hh]
    _, samples, _ = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=8)
    got = samples.reshape(-1, 3).mean(0)
    # irradiance on a y-up floor: world dir (x, y, z) = Y_UP @ light dir; cos = world y = light z = cos(theta_light)
    n_t, n_p = 512, 1024
    theta = (np.arange(n_t) + 0.5) / n_t * (np.pi / 2)
    phi = (np.arange(n_p) + 0.5) / n_p * 2 * np.pi
    T, P = np.meshgrid(theta, phi, indexing="ij")
    Lw = env_lookup_np(tex, P, T)
    w = (np.sin(T) * np.cos(T))[..., None] * (np.pi / 2 / n_t) * (2 * np.pi / n_p)
    expect = KD / np.pi * (Lw * w).sum((0, 1))
    assert np.allclose(got, expect, rtol=0.03)


def test_infinite_light_power_and_mirror_escape(oracle):
    """power() = lookup((.5,.5), width .5) * pi r^2 feeds the power distribution; a mirror hit is a specular bounce, so the
    escaping ray adds Le again (path.rs:267-275)."""
    sky = np.array([0.8, 0.9, 1.0], np.float32)

    def build(strategy):
        h = HostScene()
        mirror = h.material(_abi.MAT_MIRROR, [1.0, 1.0, 1.0])
        h.light_infinite(sky)
        h.light_point([0, 3, 0], [5, 5, 5])
        P = np.array([[-2, 0, -2], [2, 0, -2], [2, 0, 2], [-2, 0, 2]], np.float32)
        h.trianglemesh(np.array([0, 2, 1, 0, 3, 2], np.uint32), P, material=mirror)
        h.look_at([0, 3, -3], [0, 0, 0], [0, 1, 0])
        h.film(12, 12)
        h.camera(fov=40.0)
        h.sampler(4)
        h.integrator(maxdepth=3, lightsamplestrategy=strategy)
        h.world_end()
        return h

    h = build("power")
    orc = oracle.OracleScene(h.desc)
    d = h.desc.contents
    func, _, _ = orc.light_distribution(1, [0, 0, 0], 2)
    wb = np.array(list(d.world_bound), np.float64)
    r = np.linalg.norm((wb[3:] - wb[:3]) / 2)
    lum = lambda c: 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]
    assert np.allclose(func, [lum(sky) * math.pi * r * r, lum([5, 5, 5]) * 4 * math.pi], rtol=1e-5)
    _, samples, _ = orc.render(h.params, want_samples=True, n_threads=2)
    # a perfect mirror has no non-specular lobe: no NEE at all, the camera ray reflects into the sky
    assert np.allclose(samples, sky[None, None, None, :], rtol=1e-5)


def test_non_power_of_two_map_is_resampled(oracle):
    """MipMap::new zooms a map to the next power of two with normalised 4-tap Lanczos weights: a constant map stays constant (so the
    render equals the constant light's up to rounding), and a power-of-two map is left alone."""
    sky = np.array([1.0, 2.0, 0.5], np.float32)
    base = floor_scene(lambda h: h.light_infinite(sky), spp=16, res=8)
    _, s0, _ = oracle.OracleScene(base.desc).render(base.params, want_samples=True, n_threads=2)
    for shape in [(5, 12), (6, 16), (8, 16)]:
        tex = np.broadcast_to(sky, shape + (3,)).copy()
        h = floor_scene(lambda hh: hh.light_infinite([1, 1, 1], texels=tex), spp=16, res=8)
        assert list(h.desc.contents.lights[0].env_res) == [shape[1], shape[0]]
        _, s1, _ = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=2)
        # the distribution has another resolution than the 1x1 light's, so individual samples differ; the estimate does not
        assert np.allclose(s1.reshape(-1, 3).mean(0), KD * sky, rtol=0.03)
        assert np.allclose(s1.reshape(-1, 3).mean(0), s0.reshape(-1, 3).mean(0), rtol=0.05)
