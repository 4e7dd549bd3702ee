"""The CUDA kernels' SOURCE, run on host threads (tests/emu: every CUDA thread is a host thread, warp collectives and __syncthreads
are barriers, the CUDA runtime is malloc/memcpy), against the oracle on tiny scenes.  This checks kernel LOGIC -- queues, compaction,
class sorting, the dimension ledger, indexing of kernels that have not seen a GPU yet -- without GPU minutes; it says nothing about the
real memory model, scheduling or performance, and the library it builds is test infrastructure that the product never loads
(GpuScene gets it through its explicit `lib=` argument here).  The -m gpu tests remain the parity tests proper."""
import ctypes as C
import os
import shutil
import sys
from pathlib import Path

import numpy as np
import pytest

from rs_pbrt_b200 import GpuScene, HostScene, _abi, scenes
from rs_pbrt_b200.host import PbrtError

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu
    return _abi.bind(C.CDLL(str(build_emu.build())))


def check(emu, oracle, h, rect=None, count_work=False, exact_weights=True):
    rp = h.params.contents
    r = rect or list(rp.sample_bounds)
    old = rp.flags
    if count_work:
        rp.flags = old | _abi.RENDER_COUNT_WORK
    g = GpuScene(h.desc, 0, lib=emu)
    try:
        gs, st = g.render_samples(h.params, r)
        film, _ = g.render(h.params, rect=r)
    finally:
        g.close()
        rp.flags = old
    film_o, os_, so = oracle.OracleScene(h.desc).render(h.params, rect=r, n_threads=4, want_samples=True)
    assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32)), float(np.abs(gs - os_).max())
    if exact_weights:  # (a filter wider than the pixel adds the neighbours' samples in an order of its own: the weight sums then agree to rounding)
        assert np.array_equal(film[..., 3], film_o[..., 3])
    assert np.allclose(film, film_o, rtol=1e-6, atol=1e-7)  # per-pixel sums: the merge order of neighbouring tiles may differ
    keys = ["camera_rays", "rays", "closest_rays", "shadow_rays"] + (["nodes_visited", "tris_tested", "light_tri_tests"] if count_work else [])
    assert {k: st[k] for k in keys} == {k: so[k] for k in keys}
    return st


def test_path_cornell_with_work_counters(emu, oracle):
    check(emu, oracle, scenes.cornell_box(xres=12, yres=12, spp=2), count_work=True)


def test_path_materials_lights_halton(emu, oracle):
    check(emu, oracle, scenes.cornell_box(xres=10, yres=10, spp=2, materials="mixed", lights="delta", strategy="power"))
    check(emu, oracle, scenes.sky_scene(xres=10, yres=10, spp=2, env="two", strategy="spatial"))
    check(emu, oracle, scenes.cornell_box(xres=10, yres=10, spp=3, sampler="halton", lensradius=5.0, focaldistance=900.0))


def test_non_power_of_two_environment_map(emu, oracle):
    """The library's host-side restatement of MipMap::new's Lanczos zoom (pbrt_gpu.cu build_env) against the oracle's."""
    rng = np.random.default_rng(2)
    for shape in [(5, 12), (16, 24), (7, 7)]:
        tex = (rng.random(shape + (3,)) ** 2 * 3).astype(np.float32)
        tex[0, 1] += 20
        h = HostScene()
        m = h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0])
        mir = h.material(_abi.MAT_MIRROR, [0.9, 0.9, 0.9])
        h.light_infinite([1, 1, 1], texels=tex, light_to_world=scenes.Y_UP)
        h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), np.array([[-4, 0, -4], [-4, 0, 4], [4, 0, 4], [4, 0, -4]], np.float32), material=m)
        h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), np.array([[-1, 0.5, -1], [-1, 1.5, 1], [1, 1.5, 1], [1, 0.5, -1]], np.float32), material=mir)
        h.look_at([0, 3, -6], [0, 0.5, 0], [0, 1, 0])
        h.film(10, 8)
        h.camera(fov=45.0)
        h.sampler(4)
        h.integrator(maxdepth=3, lightsamplestrategy="power")
        h.world_end()
        check(emu, oracle, h)


def test_global_memory_traversal_and_shading_normals(emu, oracle):
    """A scene too large for the shared-memory BVH: the global-memory k_trace variant with the shared-memory stack top."""
    h = scenes.statue(n_side=40, xres=10, yres=10, spp=2)
    assert h.n_tris * 48 > 49152
    check(emu, oracle, h, count_work=True)  # counting renders walk the reference-layout nodes (the counters are defined on them)
    check(emu, oracle, h)                   # the wide-record traversal (k_trace_wide): same samples, same ray counts
    h = scenes.conference(xres=16, yres=9, spp=2, n_chairs=6, detail=6, n_light_quads=4)
    assert h.n_tris * 48 > 49152
    check(emu, oracle, h)
    # 40 area lights: sample_discrete's two-round search (>= 32 lights) against the reference's binary search
    h = scenes.conference(xres=12, yres=8, spp=2, n_chairs=2, detail=4, n_light_quads=20)
    assert h.desc.contents.n_lights >= 32
    check(emu, oracle, h)


@pytest.mark.parametrize("mode", ["1", "2", "prep", "2+prep"])
def test_ray_coherence_order_modes(emu, oracle, mode, monkeypatch):
    """PB_RAY_SORT=1 and =2 (two-level scatter) and PB_RAY_PREP=1 (k_rayprep: per-ray traversal constants computed ahead of k_trace)
    must not change any result.  The switches are read once per process, so each mode runs in a child process."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import ctypes as C, numpy as np, oracle_lib\n"
            "from rs_pbrt_b200 import GpuScene, _abi, scenes\n"
            "E = _abi.bind(C.CDLL(%r))\n"
            "for h in (scenes.cornell_box(xres=12, yres=12, spp=2), scenes.statue(n_side=40, xres=8, yres=8, spp=2)):\n"
            "    g = GpuScene(h.desc, 0, lib=E); gs, st = g.render_samples(h.params, list(h.params.contents.sample_bounds)); g.close()\n"
            "    _, o, so = oracle_lib.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=4)\n"
            "    assert np.array_equal(gs.view(np.uint32), o.view(np.uint32)) and st['rays'] == so['rays']\n"
            "print('ok')\n") % (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "emu" / "_build" / "librs_pbrt_b200_emu.so"))
    knobs = {"1": dict(PB_RAY_SORT="1"), "2": dict(PB_RAY_SORT="2"), "prep": dict(PB_RAY_PREP="1"), "2+prep": dict(PB_RAY_SORT="2", PB_RAY_PREP="1")}[mode]
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **knobs), capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("nsamples,cossample,sampler", [(4, True, "sobol"), (6, False, "sobol"), (5, True, "halton")])
def test_ao_integrator(emu, oracle, nsamples, cossample, sampler):
    """k_ao_shade / k_ao_resolve and the AO branch of render_impl (written after round 1's GPU budget was spent)."""
    h = scenes.cornell_box(xres=10, yres=10, spp=2, sampler=sampler, integrator=("ao", nsamples, cossample))
    st = check(emu, oracle, h, count_work=True)
    assert st["shadow_rays"] > 0


def test_ao_with_shading_normals(emu, oracle):
    check(emu, oracle, scenes.statue(n_side=40, xres=8, yres=8, spp=2, integrator=("ao", 4, True)))


def test_ray_cast_api_and_host_render(emu, oracle, tmp_path):
    """pbrt_gpu_intersect / intersect_p (k_trace MODE 1 / 2) and the host mirror's end-to-end call (scene upload, render, film merge,
    image output) on the emulated kernels."""
    h = scenes.cornell_box(xres=8, yres=8, spp=2)
    g = GpuScene(h.desc, 0, lib=emu)
    rng = np.random.default_rng(3)
    n = 300
    o = np.tile(np.array([278.0, 273.0, -800.0], np.float32), (n, 1)) + rng.normal(0, 30, (n, 3)).astype(np.float32)
    d = (np.array([278.0, 273.0, 280.0]) + rng.normal(0, 250, (n, 3)) - o).astype(np.float32)
    prim, t, b, st = g.intersect(o, d)
    occ, st2 = g.intersect_p(o, d)
    g.close()
    osc = oracle.OracleScene(h.desc)
    po, to, bo, so = osc.intersect(o, d)
    oo, so2 = osc.intersect_p(o, d)
    assert np.array_equal(prim, po) and np.array_equal(t.view(np.uint32), to.view(np.uint32)) and np.array_equal(b.view(np.uint32), bo.view(np.uint32))
    assert np.array_equal(occ, oo) and (prim >= 0).sum() > 50
    assert (st["nodes_visited"], st["tris_tested"]) == (so["nodes_visited"], so["tris_tested"])
    assert (st2["nodes_visited"], st2["tris_tested"]) == (so2["nodes_visited"], so2["tris_tested"])
    # the host mirror of Integrator::render, built and run entirely on the emulation library
    hs = HostScene(lib=emu)
    m = hs.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0])
    hs.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), np.array([[-3, 0, -3], [-3, 0, 3], [3, 0, 3], [3, 0, -3]], np.float32), material=m)
    hs.trianglemesh(np.array([0, 2, 1], np.uint32), np.array([[-1, 2, -1], [1, 2, -1], [0, 2, 1]], np.float32), material=m, emit=[5, 5, 5])
    hs.look_at([0, 3, -6], [0, 0.5, 0], [0, 1, 0])
    hs.film(8, 6)
    hs.camera(fov=45.0)
    hs.sampler(2)
    hs.integrator(maxdepth=3)
    hs.world_end()
    st = hs.render(device=0)
    film = hs.film_rgbw()
    ref, _, so = oracle.OracleScene(hs.desc).render(hs.params, n_threads=2)
    assert st["rays"] == so["rays"] and np.array_equal(film[..., 3], ref[..., 3]) and np.allclose(film, ref, rtol=1e-6, atol=1e-7)
    hs.write_image(tmp_path / "emu.ppm")
    assert (tmp_path / "emu.ppm").stat().st_size > 8 * 6 * 3


@pytest.mark.parametrize("mode", ["fixed", "reference"])
def test_object_instances(emu, oracle, mode):
    """Two-level traversal (k_trace<.., INST>), interactions carried back through instance_to_world, and both reporting modes of quirk Q7:
    rotated / non-uniformly scaled instances, an identity instance in front of a wall under a sky, and the landscape stand-in."""
    import test_oracle_instancing as T
    tr = [T.rot_scale(30, [1.5, 0.7, 1.0], [-2, 0, 1]), T.translate(1.5, 0, 0.5), T.rot_scale(-50, [0.5, 2.0, 0.5], [0, 0, 2])]
    check(emu, oracle, T.scene(mode, tr, res=(10, 8), spp=2), count_work=True)
    check(emu, oracle, T.scene(mode, [np.eye(4, dtype=np.float32), T.translate(2, 0, 0)], wall=True, sky=np.array([0.25, 0.5, 1.0], np.float32), res=(10, 8), spp=2),
          count_work=True)
    check(emu, oracle, scenes.landscape(xres=14, yres=8, spp=2, n_trees=40, grid=12, detail=6, instancing=mode), count_work=True)


def test_ray_casts_into_instances(emu, oracle):
    """pbrt_gpu_intersect / intersect_p on an instanced scene (PBRT_INSTANCING_REFERENCE semantics: an identity instance reports no hit)."""
    import test_oracle_instancing as T
    h = T.scene("reference", [T.rot_scale(30, [1.5, 0.7, 1.0], [-2, 0, 1]), np.eye(4, dtype=np.float32)], wall=True, res=(8, 6), spp=1)
    g = GpuScene(h.desc, 0, lib=emu)
    rng = np.random.default_rng(5)
    n = 400
    o = np.tile(np.array([0.0, 3.0, -7.0], np.float32), (n, 1))
    d = (rng.uniform([-3.5, -0.2, -1], [3.5, 2.5, 3], (n, 3)) - o).astype(np.float32)
    prim, t, b, st = g.intersect(o, d)
    occ, st2 = g.intersect_p(o, d)
    g.close()
    osc = oracle.OracleScene(h.desc)
    po, to, bo, so = osc.intersect(o, d)
    oo, so2 = osc.intersect_p(o, d)
    assert np.array_equal(prim, po) and np.array_equal(t.view(np.uint32), to.view(np.uint32)) and np.array_equal(b.view(np.uint32), bo.view(np.uint32))
    assert np.array_equal(occ, oo)
    assert (st["nodes_visited"], st["tris_tested"], st2["nodes_visited"], st2["tris_tested"]) == (so["nodes_visited"], so["tris_tested"], so2["nodes_visited"], so2["tris_tested"])
    first_object_tri = h.desc.contents.n_tris - 12
    assert (prim >= first_object_tri).sum() > 10  # hits on the rotated instance's triangles


@pytest.mark.parametrize("kw", [dict(textures="ewa"), dict(textures="trilinear", lensradius=6.0, focaldistance=900.0),
                                dict(textures="ewa", sampler="halton", strategy="power"), dict(textures="ewa", lights="delta", spp=4),
                                dict(textures="ewa+float"), dict(textures="trilinear+float", sampler="halton"),
                                dict(textures="ewa+float+graph"), dict(textures="trilinear+graph")],
                         ids=["ewa", "trilinear-thin-lens", "ewa-halton", "ewa-delta-lights", "ewa-float", "trilinear-float-halton", "ewa-float-graph",
                              "trilinear-graph"])
def test_image_textures(emu, oracle, kw):
    """k_raygen's ray differentials, k_texture (compute_differentials, UVMapping2D, MIP pyramid of any wrap mode, trilinear / EWA
    lookups, log2_rn) and the per-hit lobe lists k_shade takes from it: matte, plastic and uber materials with textured Kd / Ks /
    opacity, a non-power-of-two image among them; "+float": ImageTexture<Float> on sigma and on roughness (roughness_to_alpha with
    log_rn per hit); "+graph": ConstantTexture / ScaleTexture / MixTexture nodes over the images, three levels deep."""
    a = dict(xres=14, yres=14, spp=2)
    a.update(kw)
    check(emu, oracle, scenes.cornell_box(**a), count_work=True)


def test_image_texture_on_an_instance(emu, oracle):
    """A textured object instance (pbrt-v3 instancing): the interaction k_texture filters is the one carried back to world space."""
    rng = np.random.default_rng(9)
    h = HostScene()
    tex = h.texture_image(rng.random((16, 16, 3)).astype(np.float32), uscale=2.0, vscale=2.0)
    m = h.material(_abi.MAT_PLASTIC, [0.5, 0.5, 0.5, 0.2, 0.2, 0.2, 0.2, 1.0], textures={0: tex})
    g = h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0])
    h.light_infinite([1.0, 1.0, 1.0])
    h.light_distant([0.3, 1.0, -0.4], [0, 0, 0], [2.0, 2.0, 2.0])
    P = np.array([[-6, 0, -6], [6, 0, -6], [6, 0, 6], [-6, 0, 6]], np.float32)
    h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), P, material=g)
    obj = h.object_begin()
    Q = np.array([[-1, 0, 0], [1, 0, 0], [1, 2, 0], [-1, 2, 0]], np.float32)
    h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), Q, UV=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32), material=m)
    h.object_end()
    c, s = np.cos(0.5), np.sin(0.5)
    h.object_instance(obj, [[1.3 * c, 0, s, 0.5], [0, 0.8, 0, 0.1], [-1.3 * s, 0, c, 0.3], [0, 0, 0, 1]])
    h.object_instance(obj, None)
    h.instancing("fixed")
    h.look_at([0.5, 2.0, -6.0], [0.3, 0.8, 0.0], [0, 1, 0])
    h.film(20, 20)
    h.camera(fov=40.0)
    h.sampler(2)
    h.integrator(maxdepth=3, lightsamplestrategy="uniform")
    h.world_end(n_threads=1)
    check(emu, oracle, h)


def test_log2_restatement(emu):
    """log2_rn (glibc's log2f, MIPMap level selection) through the emulated KAT hook, against the host libm.  (log_rn, its sibling for
    roughness_to_alpha, is exercised by the "+float" texture scenes and exhaustively by tools/checks/glibc_logf_check.c.)"""
    rng = np.random.default_rng(11)
    x = np.concatenate([np.exp(rng.uniform(-40, 10, 20000)), [1.0, 0.5, 2.0, 1e-8, 1e-38, 1e-42, 3.4e38]]).astype(np.float32)
    out = np.zeros_like(x)
    fp = C.POINTER(C.c_float)
    assert emu.pbrt_gpu_kat_log2(0, x.size, x.ctypes.data_as(fp), out.ctypes.data_as(fp)) == 0
    libm = C.CDLL("libm.so.6")
    libm.log2f.restype = C.c_float
    libm.log2f.argtypes = [C.c_float]
    ref = np.array([libm.log2f(float(v)) for v in x], np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("kw", [
    dict(integrator="whitted", materials="mixed", lights="delta"),
    dict(integrator=("direct", "one"), materials="mixed", lights="delta", sampler="halton"),
    dict(integrator=("direct", "all"), materials="mixed", lights="delta", lightsamples=3, xres=8, yres=8),
    dict(integrator=("direct", "all"), materials="mixed", lightsamples=2, maxdepth=3),  # the 2D arrays run out: single-sample fallback
    dict(integrator=("direct", "all"), materials="mixed", sampler="halton", lensradius=6.0, focaldistance=900.0),
], ids=["whitted", "one-halton", "all-n3", "all-arrays-exhausted", "all-halton-lens"])
def test_direct_and_whitted_integrators(emu, oracle, kw):
    """pb_direct.cuh: the depth-first walk of the reflection / transmission tree (glass block: both children), direct light from the
    sampler's 2D arrays or single samples, MIS against area lights, delta lights, Le at every vertex."""
    a = dict(xres=10, yres=10, spp=2)
    a.update(kw)
    check(emu, oracle, scenes.cornell_box(**a), count_work=True)


def test_direct_integrator_under_environment_light_and_through_null_materials(emu, oracle):
    """Escaped rays take light.le of every light; MIS rays that leave the scene take the environment; a Material "none" surface is
    walked through at the same depth (directlighting.rs:78-80)."""
    rng = np.random.default_rng(4)
    for integ in (("direct", "all"), "whitted"):
        h = HostScene()
        h.light_samples(2)
        h.light_infinite([1.0, 1.0, 1.0], scale=[0.8, 0.8, 0.8], texels=scenes.sky_map(16, 8, 2), light_to_world=scenes.Y_UP)
        floor = h.material(_abi.MAT_PLASTIC, [0.4, 0.3, 0.2, 0.3, 0.3, 0.3, 0.15, 1.0])
        glass = h.material(_abi.MAT_GLASS, [1, 1, 1, 1, 1, 1, 1.5, 0.0, 0.0, 1.0])
        P = np.array([[-6, 0, -6], [6, 0, -6], [6, 0, 6], [-6, 0, 6]], np.float32)
        h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), P, material=floor)
        Q = np.array([[-1.5, 0.2, 1], [1.5, 0.2, 1], [1.5, 2.5, 1], [-1.5, 2.5, 1]], np.float32)
        h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), Q, material=glass)
        h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), Q + np.float32([0.2, 0, -1.0]), material=-1)  # Material "none" in front of it
        h.look_at([0.5, 1.5, -6.0], [0.0, 1.0, 0.0], [0, 1, 0])
        h.film(14, 14)
        h.camera(fov=40.0)
        h.sampler(2)
        scenes._set_integrator(h, integ, 4, "uniform")
        h.world_end(n_threads=1)
        check(emu, oracle, h)


@pytest.mark.parametrize("integ", [("ao", 4, True), ("direct", "all"), "whitted"], ids=["ao", "direct-all", "whitted"])
@pytest.mark.parametrize("mode", ["fixed", "reference"])
def test_sibling_integrators_over_object_instances(emu, oracle, integ, mode):
    """The landscape stand-in (instanced trees, distant + infinite light) under the AO, direct-lighting and Whitted integrators: the
    two-level traversal, isect.wo of transformed hits, pass-through of instance hits in the reference's mode."""
    check(emu, oracle, scenes.landscape(xres=18, yres=10, spp=2, n_trees=50, grid=12, detail=6, instancing=mode, integrator=integ, maxdepth=3))


def test_bump_maps(emu, oracle):
    """Material::bump in k_texture and the bump-mapped shading frame in k_shade: Cornell with bump maps on matte / plastic materials
    (with and without other textures, through a scale node), then a curved mesh with vertex normals and UVs (shading.dndu / dndv),
    once directly with reverse_orientation (set_shading_geometry's flip) and once as a rotated instance (no flip: ret.shape = None)."""
    check(emu, oracle, scenes.cornell_box(xres=14, yres=14, spp=2, textures="ewa+bump"), count_work=True)
    check(emu, oracle, scenes.cornell_box(xres=10, yres=10, spp=2, textures="trilinear+float+graph+bump", sampler="halton"))
    rng = np.random.default_rng(31)
    n = 9
    u = np.linspace(0.0, 1.0, n)
    U, V = np.meshgrid(u, u, indexing="ij")
    H = 0.4 * np.sin(3.0 * U) * np.cos(2.0 * V)
    P = np.stack([(U - 0.5) * 4.0, H, (V - 0.5) * 4.0], -1).reshape(-1, 3).astype(np.float32)
    N = np.stack([-0.4 * 3.0 * np.cos(3.0 * U) * np.cos(2.0 * V) / 4.0, np.ones_like(U), 0.4 * 2.0 * np.sin(3.0 * U) * np.sin(2.0 * V) / 4.0], -1)
    N = (N / np.linalg.norm(N, axis=-1, keepdims=True)).reshape(-1, 3).astype(np.float32)
    UV = np.stack([U, V], -1).reshape(-1, 2).astype(np.float32)
    i0 = (np.arange(n - 1)[:, None] * n + np.arange(n - 1)[None, :]).reshape(-1)
    idx = np.stack([i0, i0 + 1, i0 + n + 1, i0, i0 + n + 1, i0 + n], -1).reshape(-1).astype(np.uint32)
    for instanced in (False, True):
        h = HostScene()
        bump = h.texture_image(rng.random((16, 16, 3)).astype(np.float32), float_valued=True, uscale=2.0, vscale=2.0, scale=0.3)
        kd = h.texture_image((0.2 + 0.7 * rng.random((8, 8, 3))).astype(np.float32))
        m = h.material(_abi.MAT_PLASTIC, [0.5, 0.5, 0.5, 0.3, 0.3, 0.3, 0.2, 1.0], textures={0: kd}, bump=bump)
        h.light_infinite([1.0, 1.0, 1.0], scale=[0.7, 0.7, 0.7])
        h.light_point([1.0, 3.0, -1.0], [20.0, 18.0, 15.0])
        if instanced:
            obj = h.object_begin()
            h.trianglemesh(idx, P, N=N, UV=UV, material=m)
            h.object_end()
            c, s = np.cos(0.6), np.sin(0.6)
            h.object_instance(obj, [[c, 0, s, 0.2], [0, 1.2, 0, 0.0], [-s, 0, c, 0.1], [0, 0, 0, 1]])
            h.instancing("fixed")
        else:
            h.trianglemesh(idx, P, N=N, UV=UV, material=m, reverse_orientation=True)
        h.look_at([0.0, 4.0, -5.0], [0.0, 0.0, 0.0], [0, 1, 0])
        h.film(16, 16)
        h.camera(fov=40.0)
        h.sampler(2)
        h.integrator(maxdepth=3, lightsamplestrategy="uniform")
        h.world_end(n_threads=1)
        check(emu, oracle, h)


@pytest.mark.parametrize("integ", ["whitted", ("direct", "all")], ids=["whitted", "direct-all"])
def test_textures_seen_through_specular_bounces(emu, oracle, integ):
    """The ray differentials of specular_reflect / specular_transmit (directlighting.rs:148-172, 219-249): a curved mirror with vertex
    normals (dndu / dndv != 0) and a glass pane in front of EWA- and trilinear-filtered walls; plus the textured Cornell variants."""
    rng = np.random.default_rng(41)
    n = 7
    u = np.linspace(0.0, 1.0, n)
    U, V = np.meshgrid(u, u, indexing="ij")
    H = 0.3 * np.sin(3.0 * U) * np.cos(2.0 * V)
    P = np.stack([(U - 0.5) * 4.0, H, (V - 0.5) * 4.0], -1).reshape(-1, 3).astype(np.float32)
    N = np.stack([-0.9 * np.cos(3.0 * U) * np.cos(2.0 * V) / 4.0, np.ones_like(U), 0.6 * np.sin(3.0 * U) * np.sin(2.0 * V) / 4.0], -1)
    N = (N / np.linalg.norm(N, axis=-1, keepdims=True)).reshape(-1, 3).astype(np.float32)
    UV = np.stack([U, V], -1).reshape(-1, 2).astype(np.float32)
    i0 = (np.arange(n - 1)[:, None] * n + np.arange(n - 1)[None, :]).reshape(-1)
    idx = np.stack([i0, i0 + 1, i0 + n + 1, i0, i0 + n + 1, i0 + n], -1).reshape(-1).astype(np.uint32)
    h = HostScene()
    if integ != "whitted":
        h.light_samples(2)
    wall_a = h.texture_image((0.1 + 0.8 * rng.random((16, 16, 3))).astype(np.float32), uscale=3.0, vscale=3.0)
    wall_b = h.texture_image((0.1 + 0.8 * rng.random((8, 8, 3))).astype(np.float32), trilinear=True, uscale=2.0, vscale=2.0)
    mirror = h.material(_abi.MAT_MIRROR, [0.9, 0.9, 0.9])
    glass = h.material(_abi.MAT_GLASS, [1, 1, 1, 1, 1, 1, 1.5, 0.0, 0.0, 1.0])
    ma = h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0], textures={0: wall_a})
    mb = h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0], textures={0: wall_b})
    h.light_point([0.0, 3.5, -1.0], [30.0, 28.0, 25.0])
    h.light_infinite([1.0, 1.0, 1.0], scale=[0.3, 0.3, 0.3])
    h.trianglemesh(idx, P, N=N, UV=UV, material=mirror)
    quad = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    uvq = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    h.trianglemesh(quad, np.array([[-3, 0, 3], [3, 0, 3], [3, 4, 3], [-3, 4, 3]], np.float32), UV=uvq, material=ma)       # back wall
    h.trianglemesh(quad, np.array([[-3, 0, -2], [-3, 0, 3], [-3, 4, 3], [-3, 4, -2]], np.float32), UV=uvq, material=mb)   # side wall
    h.trianglemesh(quad, np.array([[0.5, 0.3, 0.5], [2.0, 0.3, 1.0], [2.0, 2.2, 1.0], [0.5, 2.2, 0.5]], np.float32), material=glass)
    h.look_at([0.5, 3.0, -5.0], [0.0, 0.5, 0.5], [0, 1, 0])
    h.film(14, 14)
    h.camera(fov=42.0)
    h.sampler(2)
    scenes._set_integrator(h, integ, 4, "uniform")
    h.world_end(n_threads=1)
    check(emu, oracle, h)
    check(emu, oracle, scenes.cornell_box(xres=10, yres=10, spp=2, integrator=integ, textures="trilinear+float+graph+bump", lightsamples=2 if integ != "whitted" else 1))


def test_sibling_integrators_in_many_small_batches(emu, oracle, monkeypatch):
    """The AO and direct / whitted render loops split a frame into batches of pixels and of samples per pixel; force tiny batches (the
    sample loop, the pixel loop, per-batch state re-initialisation) and require the same bits."""
    monkeypatch.setenv("PB_SIBLING_BATCH_LOG2", "6")  # 64 light samples (direct) / any-hit rays (AO) in flight
    check(emu, oracle, scenes.cornell_box(xres=8, yres=8, spp=4, integrator=("direct", "all"), materials="mixed", lightsamples=2, maxdepth=3))
    check(emu, oracle, scenes.cornell_box(xres=8, yres=8, spp=3, integrator="whitted", textures="ewa+bump", sampler="halton"))
    check(emu, oracle, scenes.cornell_box(xres=8, yres=8, spp=8, integrator=("ao", 16, True)))  # 8 spp x 16 rays > one batch: the sample loop
    check(emu, oracle, scenes.landscape(xres=10, yres=6, spp=2, n_trees=20, grid=8, detail=6, instancing="fixed", integrator=("direct", "one"), maxdepth=3))


def test_path_integrator_batches_in_flight_with_textures_and_instances(emu, tmp_path):
    """Several batches on two streams (PB_BATCH_LOG2 is read once per process, hence the subprocess): each batch context has its own
    differential / per-hit material / instance-record buffers."""
    import subprocess
    script = tmp_path / "run.py"
    script.write_text('''
import sys, ctypes as C, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from rs_pbrt_b200 import GpuScene, _abi, scenes
import oracle_lib
emu = _abi.bind(C.CDLL(%r))
for h in (scenes.cornell_box(xres=24, yres=24, spp=4, textures="ewa+float+graph+bump"),
          scenes.landscape(xres=32, yres=20, spp=4, n_trees=30, grid=10, detail=6, instancing="fixed")):
    g = GpuScene(h.desc, 0, lib=emu)
    gs, st = g.render_samples(h.params, list(h.params.contents.sample_bounds))
    film, st2 = g.render(h.params)
    fo, so, sto = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=4, want_samples=True)
    assert np.array_equal(gs.view(np.uint32), so.view(np.uint32)), float(np.abs(gs - so).max())
    assert np.allclose(film, fo, rtol=1e-6, atol=1e-7) and st["rays"] == sto["rays"]
    assert st2["trace_launches"] > 12, st2["trace_launches"]  # more than one batch went through
print("ok")
''' % (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "emu" / "_build" / "librs_pbrt_b200_emu.so")))
    env = dict(os.environ, PB_BATCH_LOG2="10", PB_STREAMS="2")  # two batches in flight whatever the scene (render_impl takes one stream for few-launch frames)
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]


def test_spherical_cylindrical_and_planar_texture_mappings(emu, oracle):
    """The other TextureMapping2D kinds (texture.rs:123-252): st and its screen-space derivatives from finite differences of the hit
    point through world_to_texture (acos / atan2 restated), the seam fix-up, planar projection -- also as a bump map, whose offset
    points move in space, not in uv."""
    check(emu, oracle, scenes.mapped_walls(16, 16, 2))


@pytest.mark.parametrize("peer", ["peer", "staged"])
def test_tile_interleaved_shares_and_multi_device_render(emu, oracle, peer):
    """pbrt_gpu_render_tiles_device: the Morton-interleaved tile shares of a frame (an edge-tile frame: 40x28) add up to the frame a
    single call renders; pbrt_gpu_render_multi does the same over three (emulated) devices from one process, with the peer-access sum
    and with the staged copy.  Run in a child process: the emulated device count is an environment switch."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import ctypes as C, numpy as np, oracle_lib\n"
            "from rs_pbrt_b200 import GpuScene, _abi, scenes\n"
            "from rs_pbrt_b200.host import render_multi\n"
            "E = _abi.bind(C.CDLL(%r))\n"
            "for h in (scenes.cornell_box(xres=40, yres=28, spp=4, materials='mixed'), scenes.cornell_box(xres=33, yres=17, spp=2, filter='gaussian', xwidth=1.5, ywidth=1.5)):\n"
            "    gs = [GpuScene(h.desc, d, lib=E) for d in range(3)]\n"
            "    full, st = gs[0].render(h.params)\n"
            "    ref, _, so = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=4)\n"
            "    same = lambda a, b: np.allclose(a, b, rtol=2e-6, atol=1e-6)  # (a wide filter adds neighbours' samples with atomics: the order varies)\n"
            "    assert st['rays'] == so['rays'] and same(full, ref)\n"
            "    parts = np.zeros_like(full); rays = 0\n"
            "    for k in range(3): rays += gs[0].render_tiles_device(h.params, parts.ctypes.data, k, 3)['rays']\n"
            "    assert rays == st['rays'] and same(parts, full)\n"
            "    one = np.zeros_like(full); gs[0].render_tiles_device(h.params, one.ctypes.data, 0, 1)\n"
            "    assert same(one, full)\n"
            "    multi, sm = render_multi(gs, h.params)\n"
            "    assert sm['rays'] == st['rays'] and sm['camera_rays'] == st['camera_rays']\n"
            "    assert same(multi, full)\n"
            "    m1, s1 = render_multi(gs[:1], h.params)\n"
            "    assert s1['rays'] == st['rays'] and same(m1, full)\n"
            "    for g in gs: g.close()\n"
            "print('ok')\n") % (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "emu" / "_build" / "librs_pbrt_b200_emu.so"))
    env = dict(os.environ, PB_EMU_DEVICES="3")
    if peer == "staged":
        env["PB_EMU_NO_PEER"] = "1"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]


@pytest.mark.parametrize("kw", [dict(alpha="masks"), dict(alpha="masks", materials="mixed", lights="delta", sampler="halton"),
                                dict(alpha="masks", integrator=("direct", "all"), lightsamples=2), dict(alpha="masks", integrator=("ao", 6, True))],
                         ids=["path", "mixed-halton-delta", "direct-all", "ao"])
def test_alpha_and_shadow_alpha_masks(emu, oracle, kw):
    """Shape "alpha" / "shadowalpha" (triangle.rs:313-330, 593-654; k_trace<.., ALPHA>): an image cut-out, a shadow-alpha-only card and
    a `float alpha 0` card in the Cornell box -- samples, film and ray / node / triangle counters equal the oracle's."""
    a = dict(xres=20, yres=20, spp=4)
    a.update(kw)
    check(emu, oracle, scenes.cornell_box(**a), count_work=True)


def test_alpha_masks_through_the_ray_cast_api(emu, oracle):
    h = scenes.cornell_box(xres=8, yres=8, spp=1, alpha="masks")
    rng = np.random.default_rng(4)
    n = 3000
    o = rng.uniform(20, 530, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    g = GpuScene(h.desc, 0, lib=emu)
    pg, tg, bg, sg = g.intersect(o, d)
    og, so_g = g.intersect_p(o, d)
    g.close()
    osc = oracle.OracleScene(h.desc)
    po, to, bo, so = osc.intersect(o, d)
    oo, so_p = osc.intersect_p(o, d)
    assert np.array_equal(pg, po) and np.array_equal(tg.view(np.uint32), to.view(np.uint32)) and np.array_equal(bg.view(np.uint32), bo.view(np.uint32))
    assert np.array_equal(og, oo)
    assert (sg["nodes_visited"], sg["tris_tested"]) == (so["nodes_visited"], so["tris_tested"])
    assert (so_g["nodes_visited"], so_g["tris_tested"]) == (so_p["nodes_visited"], so_p["tris_tested"])
    # the `float alpha 0` card (triangles with the red material, in front of everything) is never reported
    card = [i for i in range(h.desc.contents.n_tris) if h.desc.contents.meshes[h.desc.contents.tris[i].mesh].alpha and h.desc.contents.meshes[h.desc.contents.tris[i].mesh].shadow_alpha]
    assert card and not np.isin(pg, card).any()


def test_spatial_light_tables_as_rows_handed_out_on_first_touch(emu, oracle, monkeypatch):
    """Above its byte budget the spatial light distribution keeps rows for the voxels the paths reach instead of a dense
    nvox x n_lights table (lightdistrib.rs:271-377 fills its hash lazily too); the render is the same, and a budget that cannot hold
    the touched voxels fails the render instead of the allocation."""
    h = scenes.cornell_box(xres=12, yres=12, spp=2)
    monkeypatch.setenv("PB_LIGHTGRID_BYTES", str(64 * 1024))
    check(emu, oracle, h, count_work=True)
    monkeypatch.setenv("PB_LIGHTGRID_BYTES", "128")
    g = GpuScene(h.desc, 0, lib=emu)
    try:
        with pytest.raises(RuntimeError, match="spatial light distribution"):
            g.render(h.params)
    finally:
        g.close()


@pytest.mark.parametrize("kw", [dict(), dict(sampler="halton", lights="delta", strategy="power"), dict(integrator=("direct", "all"), lightsamples=2),
                                dict(integrator="whitted")], ids=["path", "path-halton-delta", "direct", "whitted"])
def test_translucent_material(emu, oracle, kw):
    """TranslucentMaterial (translucent.rs:48-189): Lambertian reflection + LambertianTransmission + microfacet reflection /
    transmission at eta 1.5; the scene has the four-lobe, the diffuse-only and the reflect-only variants."""
    check(emu, oracle, scenes.cornell_box(xres=10, yres=10, spp=3, materials="translucent", **kw))


@pytest.mark.parametrize("kw", [dict(), dict(sampler="halton", lights="delta", strategy="power"), dict(integrator=("direct", "all"), lightsamples=2),
                                dict(integrator="whitted")], ids=["path", "path-halton-delta", "direct", "whitted"])
def test_mix_material(emu, oracle, kw):
    """MixMaterial (mixmat.rs:41-98): sc_opt on every lobe kind (Lambert + mirror under an amount outside [0, 1], plastic + glass,
    a mix of a mix over FresnelBlend, translucent + Oren-Nayar = five lobes), in the general k_shade class and in the direct / whitted kernels."""
    check(emu, oracle, scenes.cornell_box(xres=10, yres=10, spp=3, materials="mix", **kw))


def test_mix_material_outside_the_gpu_path(emu):
    """More than five lobes in all, a textured child, a textured amount: PBRT_E_UNSUPPORTED (the caller keeps its CPU loop); a child index
    that is not an earlier material: PBRT_E_INVALID."""
    def scene(build):
        h = HostScene()
        m = build(h)
        h.trianglemesh(np.array([0, 1, 2], np.uint32), np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), material=m)
        h.light_point([0.0, 0.0, 2.0], [1.0, 1.0, 1.0])
        h.look_at([0, 0, 3], [0, 0, 0], [0, 1, 0]); h.film(4, 4); h.camera(fov=45.0); h.sampler(1); h.integrator(maxdepth=2); h.world_end()
        return h

    def create(h):
        handle = C.c_void_p()
        rc = emu.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle))
        if rc == 0:
            emu.pbrt_gpu_scene_destroy(handle)
        return rc

    uber = [0.3, 0.3, 0.3, 0.2, 0.2, 0.2, 0.1, 0.1, 0.1, 0.2, 0.2, 0.2, 0.5, 0.5, 0.5, 0.1, 0.1, 1.5, 1.0]  # five lobes on its own
    assert create(scene(lambda h: h.material_mix(h.material(_abi.MAT_UBER, uber), h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0])))) == _abi.PBRT_E_UNSUPPORTED
    assert "five lobes" in emu.pbrt_gpu_last_error().decode()
    # glass is one lobe for the path integrator and two for direct / whitted (allow_multiple_lobes): 4 + 1 fits, 4 + 2 does not
    tr = [0.6, 0.5, 0.3, 0.3, 0.3, 0.3, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.15, 1.0]
    assert create(scene(lambda h: h.material_mix(h.material(_abi.MAT_TRANSLUCENT, tr), h.material(_abi.MAT_GLASS, [1, 1, 1, 1, 1, 1, 1.5, 0, 0, 1])))) == _abi.PBRT_E_UNSUPPORTED

    def textured_child(h):
        t = h.texture_constant([0.2, 0.3, 0.4])
        return h.material_mix(h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0], textures={0: t}), h.material(_abi.MAT_MIRROR, [0.9, 0.9, 0.9]))
    assert create(scene(textured_child)) == _abi.PBRT_E_UNSUPPORTED
    assert "textured" in emu.pbrt_gpu_last_error().decode()

    h = scene(lambda h: h.material_mix(h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0]), h.material(_abi.MAT_MIRROR, [0.9, 0.9, 0.9])))
    assert create(h) == 0
    mats = h.desc.contents.materials
    mats[2].params[4] = 2.0  # names itself
    assert create(h) == _abi.PBRT_E_INVALID
    mats[2].params[4] = 0.5  # not a whole number
    assert create(h) == _abi.PBRT_E_INVALID
    mats[2].params[4] = 1.0
    mats[2].tex[0] = 1  # a textured amount (no such texture either: rejected before it is looked up)
    assert create(h) in (_abi.PBRT_E_UNSUPPORTED, _abi.PBRT_E_INVALID)


def _random_material(h, rng, depth=0):
    """One of the nine material kinds with random parameters -- black components (a lobe drops out), zero / tiny / large roughness with and
    without "remaproughness", Oren-Nayar sigma, partial and full opacity, amounts outside [0, 1] -- and the number of lobes it can have at most."""
    col = lambda: [0.0, 0.0, 0.0] if rng.random() < 0.15 else [float(x) for x in rng.random(3) * rng.choice([0.3, 0.9, 1.0])]
    rough = lambda: float(rng.choice([0.0, 0.0005, 0.02, 0.1, 0.6, 1.3]))
    remap = lambda: float(rng.integers(0, 2))
    kind = int(rng.integers(0, 9 if depth < 2 else 8))
    if kind == _abi.MAT_MATTE:
        return h.material(kind, col() + [float(rng.choice([0.0, 0.0, 20.0, 75.0, 120.0]))]), 1
    if kind == _abi.MAT_PLASTIC:
        return h.material(kind, col() + col() + [rough(), remap()]), 2
    if kind == _abi.MAT_METAL:
        return h.material(kind, [float(x) for x in 0.1 + rng.random(3) * 2] + [float(x) for x in 1 + rng.random(3) * 4] + [rough(), rough(), remap()]), 1
    if kind == _abi.MAT_MIRROR:
        return h.material(kind, col()), 1
    if kind == _abi.MAT_GLASS:
        r = rough()
        return h.material(kind, col() + col() + [float(rng.choice([1.0, 1.33, 1.5, 2.4])), r, r if rng.random() < 0.7 else rough(), remap()]), 2
    if kind == _abi.MAT_UBER:
        op = [1.0, 1.0, 1.0] if rng.random() < 0.5 else col()
        return h.material(kind, col() + col() + col() + col() + op + [rough(), rough(), float(rng.choice([1.0, 1.5])), remap()]), 5
    if kind == _abi.MAT_SUBSTRATE:
        return h.material(kind, col() + col() + [rough(), rough(), remap()]), 1
    if kind == _abi.MAT_TRANSLUCENT:
        return h.material(kind, col() + col() + col() + col() + [rough(), remap()]), 4
    for _ in range(20):  # MIX over two earlier materials, at most five lobes in all
        (m1, n1), (m2, n2) = _random_material(h, rng, depth + 1), _random_material(h, rng, depth + 1)
        if n1 + n2 <= 5:
            return h.material_mix(m1, m2, [float(x) for x in rng.choice([0.0, 0.3, 0.5, 1.0, 1.4], 3)]), n1 + n2
    return h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0]), 1


@pytest.mark.parametrize("seed", range(int(os.environ.get("RS_PBRT_FUZZ_SEEDS", "48"))))  # (RS_PBRT_FUZZ_SEEDS=3000: the long hunt, ~3 min)
def test_randomised_materials_and_settings(emu, oracle, seed):
    """Cornell boxes whose seven surfaces carry random materials (every kind, MixMaterial up to two levels deep, degenerate parameters) under
    random lights / sampler / integrator / light strategy / depth / lens: samples, film and ray counters equal the oracle's bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    mats = lambda h: {k: _random_material(h, rng)[0] for k in ("floor", "short", "tall", "back", "ceiling", "left", "right")}
    integrator = [("path",) * 1, "path", "path", ("direct", "all"), ("direct", "one"), "whitted"][int(rng.integers(0, 6))]
    integrator = "path" if integrator == ("path",) else integrator
    kw = dict(lights=str(rng.choice(["area", "area", "delta", "point", "spot", "distant"])), sampler=str(rng.choice(["sobol", "halton"])),
              maxdepth=int(rng.integers(1, 7)), integrator=integrator)
    if integrator == "path":
        kw["strategy"] = str(rng.choice(["spatial", "power", "uniform"]))
    if rng.random() < 0.3:
        kw.update(lensradius=float(rng.choice([2.0, 8.0])), focaldistance=float(rng.choice([700.0, 1000.0])))
    if isinstance(integrator, tuple):
        kw["lightsamples"] = int(rng.integers(1, 4))
    h = scenes.cornell_box(xres=9, yres=7, spp=int(rng.integers(1, 5)), materials=mats, **kw)
    rp = h.params.contents
    rp.rr_threshold = float(rng.choice([1.0, 1.0, 0.05, 4.0]))           # "rrthreshold" (path.rs:31, :253)
    rp.max_sample_luminance = float(rng.choice([np.inf, np.inf, 0.5, 8.0]))  # "maxsampleluminance" (film.rs:96, :118-121)
    if rng.random() < 0.2:  # integrator "pixelbounds": samples outside are skipped (integrator.rs:125)
        rp.pixel_bounds[0], rp.pixel_bounds[1], rp.pixel_bounds[2], rp.pixel_bounds[3] = 2, 1, 7, 6
    check(emu, oracle, h)


@pytest.mark.parametrize("seed", range(int(os.environ.get("RS_PBRT_FUZZ_FAMILIES", "40"))))
def test_randomised_scene_families(emu, oracle, seed):
    """The scene generators (textured / alpha-masked Cornell boxes, sky scenes, instanced landscapes in both readings of quirk Q7, fBm statues large
    enough for the global-memory and wide-record traversals, conference rooms) under random sizes, seeds, samplers, filters, integrators and light
    strategies, with the work counters on: samples, film, ray / node / triangle counts equal the oracle's."""
    rng = np.random.default_rng(5000 + seed)
    pick = lambda *a: a[int(rng.integers(0, len(a)))]
    integ = pick("path", "path", "path", ("direct", "all"), ("direct", "one"), "whitted", ("ao", int(rng.integers(1, 7)), bool(rng.integers(0, 2))))
    fam = pick("cornell", "cornell", "sky", "landscape", "landscape", "statue", "conference")
    xres, yres, spp = int(rng.integers(5, 15)), int(rng.integers(5, 13)), int(rng.integers(1, 4))
    grow = int(os.environ.get("RS_PBRT_FUZZ_GROW", "1"))  # larger frames for runs under PB_BATCH_LOG2=10 [PB_STREAMS=2]: several batches per frame, two in flight
    xres, yres = xres * grow, yres * grow
    if fam == "cornell":
        tex = pick(None, "ewa", "trilinear", "ewa+float", "trilinear+float+graph", "ewa+float+graph+bump", "trilinear+bump")
        # (with textures the floor, the back wall and the blocks take the textured materials; ceiling and side walls then carry random ones, mixes included)
        random_mats = lambda hh: {k: _random_material(hh, rng)[0] for k in ("ceiling", "left", "right")}
        kw = dict(textures=tex, alpha=pick(None, None, "masks"), materials=pick("matte", "mixed", "translucent", "mix") if tex is None else pick("matte", random_mats),
                  lights=pick("area", "delta", "point", "spot", "distant"), sampler=pick("sobol", "halton"), integrator=integ, maxdepth=int(rng.integers(1, 6)),
                  strategy=pick("spatial", "power", "uniform"), samplepixelcenter=bool(rng.integers(0, 2)))
        if rng.random() < 0.4:
            f = pick("gaussian", "triangle", "box")
            w = float(pick(0.5, 1.0, 1.7)) if f != "box" else float(pick(0.5, 1.3))
            kw.update(filter=f, xwidth=w, ywidth=float(pick(w, 0.8)))
        if rng.random() < 0.3:
            kw.update(lensradius=float(pick(3.0, 9.0)), focaldistance=float(pick(600.0, 1100.0)))
        if isinstance(integ, tuple) and integ[0] == "direct":
            kw["lightsamples"] = int(rng.integers(1, 4))
        if rng.random() < 0.3:
            kw["crop"] = [0.1, 0.9, 0.2, 0.75]
        h = scenes.cornell_box(xres=xres, yres=yres, spp=spp, **kw)
    elif fam == "sky":
        h = scenes.sky_scene(xres=xres, yres=yres, spp=spp, env=pick("constant", "image", "two"), extra_lights=bool(rng.integers(0, 2)), sampler=pick("sobol", "halton"),
                             strategy=pick("spatial", "power", "uniform"), maxdepth=int(rng.integers(1, 6)))
    elif fam == "landscape":
        h = scenes.landscape(xres=xres + 4, yres=yres, spp=spp, n_trees=int(rng.integers(1, 60)), grid=int(rng.integers(4, 20)), detail=int(rng.integers(3, 8)),
                             seed=int(rng.integers(0, 1000)), instancing=pick("fixed", "reference"), sampler=pick("sobol", "halton"), integrator=integ,
                             n_prototypes=int(rng.integers(1, 5)), sky=pick("map", "constant"), strategy=pick("spatial", "power"), maxdepth=int(rng.integers(1, 5)))
    elif fam == "statue":
        h = scenes.statue(n_side=int(rng.integers(8, 110)), xres=xres, yres=yres, spp=spp, seed=int(rng.integers(0, 1000)), with_normals=bool(rng.integers(0, 2)),
                          integrator=integ, maxdepth=int(rng.integers(1, 5)))
    else:
        h = scenes.conference(xres=xres + 4, yres=yres, spp=spp, seed=int(rng.integers(0, 1000)), n_chairs=int(rng.integers(1, 7)), detail=int(rng.integers(3, 8)),
                              n_light_quads=int(rng.integers(1, 20)), maxdepth=int(rng.integers(1, 5)))
    check(emu, oracle, h, count_work=True, exact_weights=not (fam == "cornell" and "filter" in kw))


@pytest.mark.parametrize("scene", ["cornell", "alpha", "statue-small", "statue-wide", "landscape-fixed", "landscape-reference", "conference"])
def test_degenerate_rays_through_the_ray_cast_entry_points(emu, oracle, scene):
    """pbrt_gpu_intersect / intersect_p with the rays a renderer never sends but Bounds3f::intersect_p and Triangle::intersect must still answer
    like the reference: axis-parallel directions (infinite reciprocals, 0 * inf in the slab test), origins on vertices, edges and faces, origins
    inside boxes, zero and denormal direction components, tiny and huge t_max, unnormalised and very long / very short directions -- on the
    shared-memory tree, the global-memory walk, the wide two-box records, instances (both readings of Q7) and alpha-masked meshes.
    Hit primitive, t, barycentrics, occlusion and the node / triangle counters equal the oracle's bit for bit."""
    h = {"cornell": lambda: scenes.cornell_box(xres=4, yres=4, spp=1, materials="mixed"),
         "alpha": lambda: scenes.cornell_box(xres=4, yres=4, spp=1, alpha="masks"),
         "statue-small": lambda: scenes.statue(n_side=12, xres=4, yres=4, spp=1),
         "statue-wide": lambda: scenes.statue(n_side=100, xres=4, yres=4, spp=1),
         "landscape-fixed": lambda: scenes.landscape(xres=8, yres=4, spp=1, n_trees=25, grid=10, detail=5, instancing="fixed"),
         "landscape-reference": lambda: scenes.landscape(xres=8, yres=4, spp=1, n_trees=25, grid=10, detail=5, instancing="reference"),
         "conference": lambda: scenes.conference(xres=8, yres=4, spp=1, n_chairs=3, detail=4, n_light_quads=4)}[scene]()
    d_ = h.desc.contents
    rng = np.random.default_rng(sum(map(ord, scene)))
    # world-space vertices of the top-level meshes, and the world bound
    verts = np.concatenate([np.ctypeslib.as_array(d_.meshes[i].p, (d_.meshes[i].n_verts * 3,)).reshape(-1, 3) for i in range(d_.n_meshes)]).astype(np.float32)
    lo, hi = verts.min(0), verts.max(0)
    n = 1500
    ctr = ((lo + hi) / 2).astype(np.float32)
    o = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    k = np.arange(n)
    axis = rng.integers(0, 3, n)
    # a third: axis-parallel (one or two zero components, some negative zero)
    m = k % 3 == 0
    z = d.copy(); z[np.arange(n), axis] = np.where(rng.random(n) < 0.5, 0.0, -0.0); d[m] = z[m]
    m2 = k % 9 == 0
    z = d.copy(); z[np.arange(n), (axis + 1) % 3] = 0.0; d[m2] = z[m2]
    # origins exactly on vertices / edge midpoints / just outside the bound looking in
    on_v = k % 5 == 1
    o[on_v] = verts[rng.integers(0, len(verts), on_v.sum())]
    on_e = k % 5 == 2
    a, b = verts[rng.integers(0, len(verts), on_e.sum())], verts[rng.integers(0, len(verts), on_e.sum())]
    o[on_e] = (a + b) * np.float32(0.5)
    out = k % 7 == 3
    o[out] = (ctr + (o[out] - ctr) * np.float32(3.0)).astype(np.float32)
    d[out] = (ctr - o[out] + rng.normal(0, 0.05, (out.sum(), 3)) * (hi - lo)).astype(np.float32)
    # scale: denormal-length, tiny, huge directions
    scale = np.where(k % 11 == 4, 1e-30, np.where(k % 11 == 5, 1e-4, np.where(k % 11 == 6, 1e12, 1.0))).astype(np.float32)
    d = (d * scale[:, None]).astype(np.float32)
    d[np.all(d == 0, axis=1)] = np.array([0.0, 1.0, 0.0], np.float32)
    t_max = np.where(k % 13 == 7, 1e-6, np.where(k % 13 == 8, 0.5, np.inf)).astype(np.float32)
    g = GpuScene(h.desc, 0, lib=emu)
    try:
        prim, t, b, st = g.intersect(o, d, t_max)
        occ, st2 = g.intersect_p(o, d, t_max)
    finally:
        g.close()
    osc = oracle.OracleScene(h.desc)
    po, to, bo, so = osc.intersect(o, d, t_max)
    oo, so2 = osc.intersect_p(o, d, t_max)
    assert np.array_equal(prim, po), np.argwhere(prim != po)[:5]
    hit = po >= 0
    assert np.array_equal(t.view(np.uint32)[hit], to.view(np.uint32)[hit]) and np.array_equal(b.view(np.uint32)[hit], bo.view(np.uint32)[hit])
    assert np.array_equal(occ, oo), np.argwhere(occ != oo)[:5]
    assert (st["nodes_visited"], st["tris_tested"]) == (so["nodes_visited"], so["tris_tested"])
    assert (st2["nodes_visited"], st2["tris_tested"]) == (so2["nodes_visited"], so2["tris_tested"])
    assert hit.sum() > 50


@pytest.mark.parametrize("maxdepth", [65, 200, 0xffffffff])
def test_large_maxdepth_ends_when_the_paths_do(emu, oracle, maxdepth):
    """A "maxdepth" beyond 64 (paths end by Russian roulette long before, path.rs:253-262) takes the queue-polled loop instead of queueing
    max_depth + 1 iterations: the render returns, with the oracle's samples and ray counts."""
    h = scenes.cornell_box(xres=8, yres=8, spp=2, materials="mixed")
    h.params.contents.max_depth = maxdepth
    st = check(emu, oracle, h)
    assert st["trace_launches"] < 200


def test_golden_fixtures_through_the_kernels(emu):
    """tests/golden/widened_16.npz and round2_16.npz through the kernels' source, with no oracle in the process: fixed inputs, frozen outputs."""
    from golden_cases import round2_cases, widened_cases
    for fixture, cases in (("widened_16.npz", widened_cases), ("round2_16.npz", round2_cases)):
        g = np.load(ROOT / "tests" / "golden" / fixture)
        for name, h in cases():
            gpu = GpuScene(h.desc, 0, lib=emu)
            try:
                samples, st = gpu.render_samples(h.params, list(h.params.contents.sample_bounds))
            finally:
                gpu.close()
            assert st["rays"] == int(g[name + "_rays"]), name
            assert np.array_equal(samples, g[name + "_samples"]), name


@pytest.mark.parametrize("seed", range(24))
def test_randomised_tile_shares_add_up(emu, seed):
    """pbrt_gpu_render_tiles_device for random frame sizes (edge tiles, frames smaller than a tile, crop windows), filters and part counts (more parts
    than tiles included): the parts' films add up to the frame of one call, their ray counts to its ray count, and every camera sample is drawn once."""
    rng = np.random.default_rng(9000 + seed)
    pick = lambda *a: a[int(rng.integers(0, len(a)))]
    kw = dict(xres=int(rng.integers(3, 70)), yres=int(rng.integers(3, 50)), spp=int(pick(1, 2, 4)), materials=pick("matte", "mixed", "mix"), sampler=pick("sobol", "halton"))
    if rng.random() < 0.5:
        kw.update(filter=pick("gaussian", "triangle"), xwidth=float(pick(1.0, 2.0)), ywidth=float(pick(1.0, 1.5)))
    if rng.random() < 0.4:
        kw["crop"] = [float(rng.uniform(0.0, 0.4)), float(rng.uniform(0.6, 1.0)), float(rng.uniform(0.0, 0.4)), float(rng.uniform(0.6, 1.0))]
    h = scenes.cornell_box(**kw)
    g = GpuScene(h.desc, 0, lib=emu)
    try:
        full, st = g.render(h.params)
        n_parts = int(pick(1, 2, 3, 5, 8, 13, 40))
        parts = np.zeros_like(full)
        rays = cams = 0
        for k in range(n_parts):
            s = g.render_tiles_device(h.params, parts.ctypes.data, k, n_parts)
            rays += s["rays"]; cams += s["camera_rays"]
    finally:
        g.close()
    assert rays == st["rays"] and cams == st["camera_rays"]
    assert np.allclose(parts, full, rtol=2e-6, atol=1e-6)
    if "filter" not in kw:
        assert np.array_equal(parts[..., 3], full[..., 3])


@pytest.mark.parametrize("seed", range(int(os.environ.get("RS_PBRT_FUZZ_TEXTURES", "32"))))
def test_randomised_textures(emu, oracle, seed):
    """Image textures with random resolutions (1 x 1, one texel wide, non-powers of two -> the Lanczos zoom), wrap modes, trilinear / EWA with
    random anisotropy limits, uv scales and offsets (zero, negative, huge), the three non-uv mappings with random matrices, float textures on
    sigma / roughness, constant / scale / mix nodes on top, bump maps -- on a floor seen at a grazing angle, a facing wall and a mirror that shows
    both without ray differentials, optionally through a thin lens: samples, film and counters equal the oracle's."""
    rng = np.random.default_rng(7000 + seed)
    pick = lambda *a: a[int(rng.integers(0, len(a)))]
    h = HostScene()

    def image(float_valued=False):
        res = (int(pick(1, 2, 3, 5, 8, 16, 17, 31)), int(pick(1, 2, 4, 7, 8, 16, 20)))
        img = rng.random(res + (3,)).astype(np.float32) ** float(pick(1.0, 3.0))
        if rng.random() < 0.2:
            img[rng.random(res) < 0.3] = 0.0  # black texels: a lobe drops out at those hits
        t = h.texture_image(img, trilinear=bool(rng.integers(0, 2)), max_anisotropy=float(pick(1.0, 2.0, 8.0, 64.0)), wrap=int(pick(_abi.WRAP_REPEAT, _abi.WRAP_BLACK, _abi.WRAP_CLAMP)),
                            scale=float(pick(1.0, 0.5, 2.0)), gamma=bool(rng.integers(0, 2)), uscale=float(pick(1.0, 0.0, -2.0, 37.5, 0.01)), vscale=float(pick(1.0, 3.0, -0.5, 1e-3)),
                            udelta=float(pick(0.0, 0.25, -7.5)), vdelta=float(pick(0.0, 0.5)), float_valued=float_valued)
        m = pick("uv", "uv", "spherical", "cylindrical", "planar")
        if m == "planar":
            h.texture_mapping(t, "planar", [float(x) for x in rng.normal(0, 0.4, 6)])
        elif m != "uv":
            a = rng.normal(size=(3, 3)); q, _ = np.linalg.qr(a)
            w2t = np.eye(4, dtype=np.float32); w2t[:3, :3] = (q * float(pick(1.0, 0.3))).astype(np.float32); w2t[:3, 3] = rng.normal(0, 1, 3).astype(np.float32)
            h.texture_mapping(t, m, w2t)
        return t

    def spectrum_tex(depth=0):
        k = pick("image", "image", "constant", "scale", "mix") if depth < 2 else "image"
        if k == "image":
            return image()
        if k == "constant":
            return h.texture_constant([float(x) for x in rng.random(3)])
        if k == "scale":
            return h.texture_scale(spectrum_tex(depth + 1), spectrum_tex(depth + 1))
        return h.texture_mix(spectrum_tex(depth + 1), spectrum_tex(depth + 1), float_tex(depth + 1))

    def float_tex(depth=0):
        k = pick("image", "constant", "scale") if depth < 2 else "image"
        if k == "image":
            return image(float_valued=True)
        if k == "constant":
            return h.texture_constant([float(rng.random())], float_valued=True)
        return h.texture_scale(float_tex(depth + 1), float_tex(depth + 1))

    def material():
        bump = float_tex() if rng.random() < 0.3 else None
        k = pick("matte", "plastic", "uber", "substrate", "translucent", "glass", "metal")
        if k == "matte":
            return h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, float(pick(0.0, 30.0))], textures={0: spectrum_tex(), **({1: float_tex()} if rng.random() < 0.3 else {})}, bump=bump)
        if k == "plastic":
            return h.material(_abi.MAT_PLASTIC, [0.5, 0.5, 0.5, 0.3, 0.3, 0.3, 0.1, float(rng.integers(0, 2))],
                              textures={int(g): (spectrum_tex() if g < 2 else float_tex()) for g in rng.choice(3, int(rng.integers(1, 4)), replace=False)}, bump=bump)
        if k == "uber":
            return h.material(_abi.MAT_UBER, [0.4, 0.4, 0.4, 0.2, 0.2, 0.2, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 0.1, 0.2, 1.5, 1.0],
                              textures={int(g): (spectrum_tex() if g < 5 else float_tex()) for g in rng.choice(8, int(rng.integers(1, 4)), replace=False)}, bump=bump)
        if k == "substrate":
            return h.material(_abi.MAT_SUBSTRATE, [0.4, 0.3, 0.2, 0.1, 0.1, 0.1, 0.1, 0.15, 1.0], textures={int(pick(0, 1)): spectrum_tex(), int(pick(2, 3)): float_tex()}, bump=bump)
        if k == "translucent":
            return h.material(_abi.MAT_TRANSLUCENT, [0.6, 0.5, 0.3, 0.3, 0.3, 0.3, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.15, 1.0], textures={int(pick(0, 1, 2, 3)): spectrum_tex(), 4: float_tex()}, bump=bump)
        if k == "glass":
            return h.material(_abi.MAT_GLASS, [1, 1, 1, 1, 1, 1, 1.5, 0.0, 0.0, 1.0], textures={int(pick(0, 1)): spectrum_tex(), **({3: float_tex(), 4: float_tex()} if rng.random() < 0.4 else {})}, bump=bump)
        return h.material(_abi.MAT_METAL, [0.2, 0.92, 1.1, 3.9, 2.45, 2.14, 0.05, 0.05, 1.0], textures={int(pick(0, 1)): spectrum_tex(), int(pick(2, 3)): float_tex()}, bump=bump)

    quad = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32) * np.float32(pick(1.0, 4.0))
    h.light_infinite([1.0, 1.0, 1.0], scale=[0.5, 0.5, 0.5])
    h.light_point([0.0, 4.0, -2.0], [25.0, 25.0, 25.0])
    h.trianglemesh(quad, np.array([[-50, 0, -5], [50, 0, -5], [50, 0, 200], [-50, 0, 200]], np.float32), UV=uv, material=material())   # floor to the horizon
    h.trianglemesh(quad, np.array([[-3, 0, 3], [3, 0, 3], [3, 4, 3], [-3, 4, 3]], np.float32), UV=uv if rng.random() < 0.7 else None, material=material())
    h.trianglemesh(quad, np.array([[-3, 0, -2], [-3, 0, 3], [-3, 4, 3], [-3, 4, -2]], np.float32), material=h.material(_abi.MAT_MIRROR, [0.9, 0.9, 0.9]))
    h.look_at([1.5, float(pick(0.3, 2.5)), -6.0], [0.0, 1.0, 1.0], [0, 1, 0])
    h.film(int(rng.integers(6, 14)), int(rng.integers(5, 11)))
    h.camera(fov=float(pick(30.0, 45.0, 80.0)), **(dict(lensradius=0.05, focaldistance=7.0) if rng.random() < 0.25 else {}))
    h.sampler(int(pick(1, 2, 4)), name=pick("sobol", "halton"))
    integ = pick("path", "path", ("direct", "all"), "whitted")
    scenes._set_integrator(h, integ, int(rng.integers(1, 5)), pick("uniform", "power", "spatial"))
    h.world_end(n_threads=1)
    check(emu, oracle, h, count_work=True)


@pytest.mark.parametrize("seed", range(int(os.environ.get("RS_PBRT_FUZZ_SOUPS", "32"))))
def test_randomised_triangle_soups_and_lights(emu, oracle, seed):
    """Triangle soups in a box: random meshes with or without per-vertex normals (not necessarily consistent with the geometry), tangents and uvs
    (degenerate uvs included: the shading frame falls back to coordinate_system), ReverseOrientation and handedness flags, slivers, coincident and
    intersecting triangles, one- and two-sided emitters (a few up to dozens: the spatial light grid), point / spot (narrow, wide, zero falloff) /
    distant / constant and image infinite lights in random order, every material kind -- under every integrator, sampler and light strategy."""
    rng = np.random.default_rng(11000 + seed)
    pick = lambda *a: a[int(rng.integers(0, len(a)))]
    h = HostScene()
    mats = [_random_material(h, rng)[0] for _ in range(int(rng.integers(1, 5)))]
    light_m = h.material(_abi.MAT_MATTE, [0.7, 0.7, 0.7, 0.0])

    def delta_light():
        k = pick("point", "spot", "distant", "infinite", "infinite-map")
        if k == "point":
            h.light_point([float(x) for x in rng.uniform(-1.5, 1.5, 3)], [float(x) for x in rng.uniform(0.0, 8.0, 3)], scale=pick(None, [0.5, 2.0, 1.0]))
        elif k == "spot":
            h.light_spot([float(x) for x in rng.uniform(-1.5, 1.5, 3)], [float(x) for x in rng.uniform(-1.0, 1.0, 3)], [float(x) for x in rng.uniform(0.0, 20.0, 3)],
                         coneangle=float(pick(5.0, 30.0, 80.0)), conedeltaangle=float(pick(0.0, 5.0, 30.0)))
        elif k == "distant":
            h.light_distant([float(x) for x in rng.normal(size=3)], [0.0, 0.0, 0.0], [float(x) for x in rng.uniform(0.0, 2.0, 3)])
        elif k == "infinite":
            h.light_infinite([float(x) for x in rng.uniform(0.0, 1.0, 3)], scale=pick(None, [0.3, 0.3, 0.3]))
        else:
            shape = (int(pick(1, 4, 5, 16)), int(pick(1, 8, 12)))
            a = rng.normal(size=(3, 3)); q, _ = np.linalg.qr(a)
            h.light_infinite([1.0, 1.0, 1.0], texels=(rng.random(shape + (3,)) ** 3 * 2).astype(np.float32), light_to_world=q.astype(np.float32))

    n_delta = int(rng.integers(0, 4))
    for _ in range(n_delta // 2):
        delta_light()
    n_emit = 0
    for _ in range(int(rng.integers(1, 6))):
        nt = int(rng.integers(1, 14))
        nv = int(rng.integers(3, 3 * nt + 1))
        P = rng.uniform(-2.0, 2.0, (nv, 3)).astype(np.float32)
        if rng.random() < 0.2:
            P[:, int(rng.integers(0, 3))] = np.float32(rng.uniform(-2, 2))  # a flat mesh: coplanar, overlapping triangles
        idx = np.stack([rng.permutation(nv)[:3] for _ in range(nt)]).astype(np.uint32)  # three distinct vertices per triangle (slivers welcome)
        N = (rng.normal(size=(nv, 3)).astype(np.float32) if rng.random() < 0.5 else None)
        S = (rng.normal(size=(nv, 3)).astype(np.float32) if rng.random() < 0.3 else None)
        UV = None
        if rng.random() < 0.6:
            UV = rng.random((nv, 2)).astype(np.float32)
            if rng.random() < 0.3:
                UV[:] = UV[0]  # degenerate parameterisation
        emit = None
        if rng.random() < 0.35 and N is None:
            emit = [float(x) for x in rng.uniform(0.5, 6.0, 3)]
            n_emit += nt
        h.trianglemesh(idx.reshape(-1), P, N=N, S=S, UV=UV, material=light_m if emit else pick(*mats), emit=emit, two_sided=bool(rng.integers(0, 2)),
                       reverse_orientation=bool(rng.integers(0, 2)), swaps_handedness=bool(rng.random() < 0.2))
    for _ in range(n_delta - n_delta // 2):
        delta_light()
    if n_emit == 0 and n_delta == 0:
        h.light_point([0.0, 0.0, 0.0], [5.0, 5.0, 5.0])
    h.look_at([float(x) for x in rng.uniform(-3, 3, 3)], [0.0, 0.0, 0.0], [0, 1, 0])
    h.film(int(rng.integers(5, 12)), int(rng.integers(5, 10)))
    h.camera(fov=float(pick(40.0, 70.0)))
    h.sampler(int(pick(1, 2, 4)), name=pick("sobol", "halton"))
    integ = pick("path", "path", "path", ("direct", "all"), ("direct", "one"), "whitted", ("ao", 3, True))
    if isinstance(integ, tuple) and integ[0] == "direct":
        pass
    scenes._set_integrator(h, integ, int(rng.integers(1, 6)), pick("uniform", "power", "spatial"))
    h.world_end(n_threads=1)
    try:
        check(emu, oracle, h, count_work=True)
    except PbrtError as e:
        # directlighting "all" over dozens of emitters asks for more sample-array dimensions than the sampler has (1024 Sobol', 1000 Halton): the
        # reference panics (sobol.rs:119-124, halton.rs:256-262), the library refuses -- and the oracle has to refuse the same scene
        assert e.code == _abi.PBRT_E_UNSUPPORTED and "dimensions" in str(e), e
        with pytest.raises(RuntimeError, match="dimensions"):
            oracle.OracleScene(h.desc).render(h.params, n_threads=1)


@pytest.mark.parametrize("seed", range(int(os.environ.get("RS_PBRT_FUZZ_INSTANCES", "32"))))
def test_randomised_object_instances(emu, oracle, seed):
    """ObjectInstances of small random objects (one to three meshes each, with or without normals and uvs) under random instance transforms --
    rotations, non-uniform and NEGATIVE scales (the transform swaps handedness), near-flat scales, large translations, the identity (which the
    reference treats as its own case) and overlapping copies -- among ordinary world-space meshes and emitters, in both readings of quirk Q7, under
    every integrator: samples, film and node / triangle counters equal the oracle's."""
    rng = np.random.default_rng(13000 + seed)
    pick = lambda *a: a[int(rng.integers(0, len(a)))]
    h = HostScene()
    h.instancing(pick("fixed", "reference"))
    mats = [_random_material(h, rng)[0] for _ in range(3)]
    white = h.material(_abi.MAT_MATTE, [0.7, 0.7, 0.7, 0.0])
    h.light_point([0.5, 3.0, -1.0], [30.0, 30.0, 30.0])
    if rng.random() < 0.5:
        h.light_infinite([0.4, 0.5, 0.7])
    quad = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    h.trianglemesh(quad, np.array([[-8, -1, -8], [8, -1, -8], [8, -1, 8], [-8, -1, 8]], np.float32), material=white)
    if rng.random() < 0.6:
        h.trianglemesh(quad, np.array([[-1, 4, -1], [1, 4, -1], [1, 4, 1], [-1, 4, 1]], np.float32), material=white, emit=[6.0, 5.0, 4.0], two_sided=True)
    objs = []
    for _ in range(int(rng.integers(1, 4))):
        o = h.object_begin()
        for _ in range(int(rng.integers(1, 4))):
            nt = int(rng.integers(1, 10))
            nv = int(rng.integers(3, 3 * nt + 1))
            P = rng.uniform(-0.6, 0.6, (nv, 3)).astype(np.float32)
            idx = np.stack([rng.permutation(nv)[:3] for _ in range(nt)]).astype(np.uint32)
            h.trianglemesh(idx.reshape(-1), P, N=rng.normal(size=(nv, 3)).astype(np.float32) if rng.random() < 0.4 else None,
                           UV=rng.random((nv, 2)).astype(np.float32) if rng.random() < 0.4 else None, material=pick(*mats), reverse_orientation=bool(rng.integers(0, 2)))
        h.object_end()
        objs.append(o)
    for _ in range(int(rng.integers(1, 9))):
        k = pick("identity", "rigid", "scaled", "scaled", "mirrored", "flat", "far")
        if k == "identity":
            h.object_instance(pick(*objs), None if rng.random() < 0.5 else np.eye(4, dtype=np.float32))
            continue
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        sc = {"rigid": np.ones(3), "scaled": rng.uniform(0.3, 2.5, 3), "mirrored": rng.uniform(0.5, 1.5, 3) * np.array([-1.0, 1.0, 1.0]),
              "flat": np.array([1.0, 1e-3, 1.0]), "far": np.ones(3)}[k]
        M = np.eye(4)
        M[:3, :3] = q @ np.diag(sc)
        M[:3, 3] = rng.uniform(-2.0, 2.0, 3) + (np.array([0.0, 0.0, 40.0]) if k == "far" else 0.0)
        h.object_instance(pick(*objs), M.astype(np.float32))
    h.look_at([float(rng.uniform(-1, 1)), 1.5, -6.0], [0.0, 0.5, 0.0], [0, 1, 0])
    h.film(int(rng.integers(6, 13)), int(rng.integers(5, 10)))
    h.camera(fov=float(pick(40.0, 60.0)))
    h.sampler(int(pick(1, 2, 4)), name=pick("sobol", "halton"))
    scenes._set_integrator(h, pick("path", "path", ("direct", "all"), "whitted", ("ao", 3, True)), int(rng.integers(1, 5)), pick("uniform", "power", "spatial"))
    h.world_end(n_threads=1)
    check(emu, oracle, h, count_work=True)


@pytest.mark.parametrize("kw", [dict(xres=1920, yres=1080, spp=64, crop=[0.9965, 1.0, 0.994, 1.0]), dict(xres=2048, yres=2048, spp=1024, crop=[0.5, 0.501, 0.9995, 1.0]),
                                dict(xres=1500, yres=700, spp=16, crop=[0.0, 0.004, 0.0, 0.01], sampler="halton"),
                                dict(xres=4096, yres=4096, spp=4, crop=[0.99975, 1.0, 0.0, 0.0005], filter="gaussian", xwidth=2.0, ywidth=2.0)],
                         ids=["1920x1080x64-corner", "2048x2048x1024", "halton-1500x700", "4096x4096-gaussian"])
def test_large_frames_through_a_small_crop_window(emu, oracle, kw):
    """Frame resolutions and sample counts of the BASELINE configs (and beyond: 4096 squared) rendered through a crop window of a few pixels: the
    Sobol' index of a sample then carries the full resolution's bits (sobol_interval_to_index with m = 11 / 12, sample numbers up to 1023), Halton's
    pixel strata wrap (128 x 243), tile numbers and Morton codes are those of the large frame -- at a cost the emulation can pay."""
    st = check(emu, oracle, scenes.cornell_box(materials="mixed", **kw), exact_weights="filter" not in kw)
    assert st["camera_rays"] > 0
