"""DirectLightingIntegrator and WhittedIntegrator in the oracle (integrators/directlighting.rs, whitted.rs): closed forms.  CPU only."""
import numpy as np
import pytest

import oracle_lib
from rs_pbrt_b200 import _abi, scenes
from rs_pbrt_b200.host import HostScene

f32 = np.float32


def _floor(h, mat, size=40.0):
    P = np.array([[-size, 0, -size], [size, 0, -size], [size, 0, size], [-size, 0, size]], f32)
    h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), P, material=mat)


def _camera(h, res=16, spp=4):
    h.look_at([0.0, 3.0, -9.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0])
    h.film(res, res)
    h.camera(fov=35.0)
    h.sampler(spp)


def test_whitted_point_light_on_a_matte_floor_is_the_closed_form():
    """whitted.rs:74-98 with a PointLight: L = Kd/pi * I / r^2 * cos(theta), no noise at all (sample_li of a delta light is
    deterministic), for every camera sample."""
    h = HostScene()
    kd, I, pl = 0.6, 50.0, np.array([1.0, 4.0, 2.0])
    m = h.material(_abi.MAT_MATTE, [kd, kd, kd, 0.0])
    h.light_point(pl.tolist(), [I, I, I])
    _floor(h, m)
    _camera(h)
    h.integrator_whitted(maxdepth=5)
    h.world_end(n_threads=1)
    o = oracle_lib.OracleScene(h.desc)
    _, samples, st = o.render(h.params, n_threads=4, want_samples=True)
    checked = 0
    for py in range(9, 16, 2):
        for px in range(0, 16, 3):
            for s in range(4):
                cs = o.camera_sample(h.params, px, py, s)
                ro, rd = cs[5:8].astype(np.float64), cs[8:11].astype(np.float64)
                P = ro + rd * (-ro[1] / rd[1])
                w = pl - P
                r2 = w @ w
                exp = kd / np.pi * I / r2 * (w[1] / np.sqrt(r2))
                np.testing.assert_allclose(samples[py, px, s], exp, rtol=2e-5)
                checked += 1
    assert checked > 50 and st["shadow_rays"] > 0


@pytest.mark.parametrize("integ", [("direct", "all"), ("direct", "one"), "whitted"])
def test_matte_floor_under_a_constant_sky_reflects_kd(integ):
    """Direct light of a Lambertian plane under a uniform white InfiniteAreaLight is Kd: an unbiased estimate for all three."""
    h = HostScene()
    kd = 0.5
    m = h.material(_abi.MAT_MATTE, [kd, kd, kd, 0.0])
    if integ == ("direct", "all"):
        h.light_samples(4)
    h.light_infinite([1.0, 1.0, 1.0])
    _floor(h, m)
    _camera(h, res=16, spp=16)
    scenes._set_integrator(h, integ, 5, "uniform")
    h.world_end(n_threads=1)
    film, _, _ = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=4)
    floor_rows = film[10:, :, :3] / film[10:, :, 3:4]
    assert abs(floor_rows.mean() - kd) < 0.02
    top = film[0, :, :3] / film[0, :, 3:4]
    sky = top[top.mean(-1) > 0.9]  # pixels whose camera rays all escaped: light.le(ray) of every light, exactly
    assert len(sky) > 0
    np.testing.assert_allclose(sky, 1.0, rtol=1e-6)


@pytest.mark.parametrize("integ", ["whitted", ("direct", "one")])
def test_mirror_shows_the_emitter_and_maxdepth_gates_the_recursion(integ):
    """specular_reflect (directlighting.rs:124-195): camera -> mirror floor -> emissive ceiling gives Kr * Le, but only while
    depth + 1 < max_depth."""
    def build(maxdepth):
        h = HostScene()
        mir = h.material(_abi.MAT_MIRROR, [0.8, 0.8, 0.8])
        blk = h.material(_abi.MAT_MATTE, [0.0, 0.0, 0.0, 0.0])
        _floor(h, mir)
        P = np.array([[-60, 12, -60], [-60, 12, 60], [60, 12, 60], [60, 12, -60]], f32)  # facing down
        h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), P, material=blk, emit=[2.0, 3.0, 4.0], two_sided=True)
        _camera(h, res=8, spp=2)
        scenes._set_integrator(h, integ, maxdepth, "uniform")
        h.world_end(n_threads=1)
        return oracle_lib.OracleScene(h.desc).render(h.params, n_threads=2, want_samples=True)[1]
    deep = build(5)
    np.testing.assert_allclose(deep[6:, :, :, :], np.broadcast_to(f32(0.8) * np.array([2.0, 3.0, 4.0], f32), deep[6:].shape), rtol=1e-6)
    assert np.all(build(1)[6:] == 0.0)


def test_glass_pane_splits_into_reflection_and_transmission():
    """allow_multiple_lobes = false (directlighting.rs:77): specular glass is SpecularReflection + SpecularTransmission and BOTH
    children are followed.  Under a uniform unit sky every escaping leaf returns 1, so the depth-limited sum stays below 1 and
    approaches it as max_depth grows (Fresnel weights of one interface sum to 1; the radiance scaling of entering and leaving cancels)."""
    def build(maxdepth):
        h = HostScene()
        g = h.material(_abi.MAT_GLASS, [1, 1, 1, 1, 1, 1, 1.5, 0.0, 0.0, 1.0])
        h.light_infinite([1.0, 1.0, 1.0])
        for z, flip in ((0.0, False), (0.5, True)):  # a slab: front face towards the camera, back face away from it
            P = np.array([[-30, -30, z], [30, -30, z], [30, 30, z], [-30, 30, z]], f32)
            idx = np.array([0, 2, 1, 0, 3, 2] if not flip else [0, 1, 2, 0, 2, 3], np.uint32)
            h.trianglemesh(idx, P, material=g)
        h.look_at([0.0, 0.0, -5.0], [0.3, 0.2, 0.0], [0.0, 1.0, 0.0])
        h.film(8, 8)
        h.camera(fov=30.0)
        h.sampler(2)
        h.integrator_whitted(maxdepth=maxdepth)
        h.world_end(n_threads=1)
        return oracle_lib.OracleScene(h.desc).render(h.params, n_threads=2, want_samples=True)[1]
    l2, l4, l8 = build(2), build(4), build(8)
    assert np.all(l2 <= l4 + 1e-6) and np.all(l4 <= l8 + 1e-6)
    assert np.all(l8 < 1.0 + 1e-5) and l8.min() > 0.985
    assert 0.02 < l2.mean() < 0.2  # depth 2 sees only the first interface's reflection (~4 %) -- the transmitted child ends inside the slab


def test_sample_arrays_run_out_and_the_single_sample_fallback_takes_over():
    """uniform_sample_all_lights (integrator.rs:312-330): preprocess asks for max_depth array pairs per light, a tree with both
    specular children has more vertices than that; the render stays finite, deterministic across thread counts, and close to the
    "one" strategy in the mean."""
    kw = dict(xres=24, yres=24, spp=8, materials="mixed", maxdepth=3)
    ha = scenes.cornell_box(integrator=("direct", "all"), lightsamples=2, **kw)
    fa1, _, _ = oracle_lib.OracleScene(ha.desc).render(ha.params, n_threads=1)
    fa8, _, _ = oracle_lib.OracleScene(ha.desc).render(ha.params, n_threads=8)
    assert np.array_equal(fa1, fa8) and np.isfinite(fa1).all()
    ho = scenes.cornell_box(integrator=("direct", "one"), **kw)
    fo, _, _ = oracle_lib.OracleScene(ho.desc).render(ho.params, n_threads=8)
    assert abs(fa1[..., :3].mean() - fo[..., :3].mean()) < 0.05 * fo[..., :3].mean()
