"""GPU parity of the rows SURVEY.md section 8(f) widened into: the AO, direct-lighting and Whitted integrators, object instancing,
image textures (trilinear / EWA, float textures, constant / scale / mix nodes, bump maps, the non-UV mappings) and their
combinations, against the oracle and a frozen fixture.

All of these were seen green on a B200 at the end of round 1 (GPUTEST_r01.json: 31 xpassed) and are ordinary, strict tests since
round 2: a regression in any of them fails the suite."""
import numpy as np
import pytest

from rs_pbrt_b200 import HostScene, _abi, scenes
from test_gpu_parity_materials import compare

pytestmark = pytest.mark.gpu


def ao_cornell(nsamples, cossample, spp, sampler="sobol", res=32):
    h = scenes.cornell_box(xres=res, yres=res, spp=spp, sampler=sampler, integrator=("ao", nsamples, cossample))
    assert h.params.contents.integrator == _abi.INTEGRATOR_AO
    return h


@pytest.mark.parametrize("nsamples,cossample,spp,sampler", [(16, True, 4, "sobol"), (64, True, 2, "sobol"), (8, False, 4, "sobol"), (12, True, 3, "halton")])
def test_ao_cornell(oracle, nsamples, cossample, spp, sampler):
    compare(ao_cornell(nsamples, cossample, spp, sampler), oracle)


def test_ao_shading_normals_scene(oracle):
    h = scenes.statue(n_side=24, xres=32, yres=32, spp=4, integrator=("ao", 16, True))
    compare(h, oracle)


# ---- object instancing (k_trace<.., INST>, isect_to_world) ----
@pytest.mark.parametrize("mode", ["fixed", "reference"])
def test_object_instances(oracle, mode):
    import test_oracle_instancing as T
    tr = [T.rot_scale(30, [1.5, 0.7, 1.0], [-2, 0, 1]), T.translate(1.5, 0, 0.5), T.rot_scale(-50, [0.5, 2.0, 0.5], [0, 0, 2])]
    compare(T.scene(mode, tr, res=(48, 36), spp=8), oracle)
    compare(T.scene(mode, [np.eye(4, dtype=np.float32), T.translate(2, 0, 0)], wall=True, sky=np.array([0.25, 0.5, 1.0], np.float32), res=(48, 36), spp=8), oracle)


@pytest.mark.parametrize("mode", ["fixed", "reference"])
def test_landscape_stand_in(oracle, mode):
    compare(scenes.landscape(xres=64, yres=36, spp=8, n_trees=300, grid=48, detail=8, instancing=mode), oracle)


# ---- image textures (k_raygen differentials, k_texture, log2_rn) ----
@pytest.mark.parametrize("kw", [dict(textures="ewa"), dict(textures="trilinear", lensradius=6.0, focaldistance=900.0),
                                dict(textures="ewa", sampler="halton"), dict(textures="ewa", lights="delta"), dict(textures="ewa+float"),
                                dict(textures="trilinear+float", sampler="halton"), dict(textures="ewa+float+graph"), dict(textures="ewa+bump"),
                                dict(textures="trilinear+float+graph+bump", sampler="halton")],
                         ids=["ewa", "trilinear-thin-lens", "ewa-halton", "ewa-delta-lights", "ewa-float", "trilinear-float-halton", "ewa-float-graph",
                              "ewa-bump", "trilinear-everything"])
def test_image_textures(oracle, kw):
    a = dict(xres=64, yres=64, spp=8)
    a.update(kw)
    compare(scenes.cornell_box(**a), oracle)


def test_log2_restatement_matches_host_libm(product_lib):
    """log2_rn restates glibc's log2f (MIPMap level selection, mipmap.rs:236,288): bit-identical on a sweep of widths."""
    import ctypes as C
    rng = np.random.default_rng(11)
    x = np.concatenate([np.exp(rng.uniform(-40, 10, 1 << 18)), [1.0, 0.5, 2.0, 1e-8, 1e-38, 1e-42, 3.4e38]]).astype(np.float32)
    out = np.zeros_like(x)
    fp = C.POINTER(C.c_float)
    assert product_lib.pbrt_gpu_kat_log2(0, x.size, x.ctypes.data_as(fp), out.ctypes.data_as(fp)) == 0
    libm = C.CDLL("libm.so.6")  # numpy's float32 log2 may use its own SIMD kernel
    libm.log2f.restype = C.c_float
    libm.log2f.argtypes = [C.c_float]
    idx = np.concatenate([rng.choice(x.size - 7, 100_000, replace=False), np.arange(x.size - 7, x.size)])
    ref = np.array([libm.log2f(float(v)) for v in x[idx]], np.float32)
    assert np.array_equal(out[idx].view(np.uint32), ref.view(np.uint32))


# ---- DirectLighting / Whitted integrators (pb_direct.cuh) ----
@pytest.mark.parametrize("kw", [
    dict(integrator="whitted", materials="mixed", lights="delta"),
    dict(integrator=("direct", "one"), materials="mixed", lights="delta", sampler="halton"),
    dict(integrator=("direct", "all"), materials="mixed", lights="delta", lightsamples=3),
    dict(integrator=("direct", "all"), materials="mixed", lightsamples=2, maxdepth=3),
    dict(integrator=("direct", "all")),
    dict(integrator="whitted", textures="trilinear+float+graph+bump"),
    dict(integrator=("direct", "all"), textures="ewa+bump", lightsamples=2),
], ids=["whitted", "one-halton", "all-n3", "all-arrays-exhausted", "all-matte", "whitted-textures", "all-textures"])
def test_direct_and_whitted_integrators(oracle, kw):
    a = dict(xres=64, yres=64, spp=8)
    a.update(kw)
    compare(scenes.cornell_box(**a), oracle)


@pytest.mark.parametrize("integ", [("ao", 8, True), ("direct", "all"), "whitted"], ids=["ao", "direct-all", "whitted"])
def test_sibling_integrators_over_object_instances(oracle, integ):
    for mode in ("fixed", "reference"):
        compare(scenes.landscape(xres=48, yres=27, spp=4, n_trees=200, grid=32, detail=8, instancing=mode, integrator=integ, maxdepth=3), oracle)


def test_widened_golden_fixture():
    """The same fixture through the GPU library: fixed inputs and outputs that do not need the oracle at run time."""
    from pathlib import Path
    from golden_cases import widened_cases
    from rs_pbrt_b200 import GpuScene
    g = np.load(Path(__file__).resolve().parent / "golden" / "widened_16.npz")
    for name, h in widened_cases():
        gpu = GpuScene(h.desc, 0)
        try:
            samples, st = gpu.render_samples(h.params, list(h.params.contents.sample_bounds))
        finally:
            gpu.close()
        assert st["rays"] == int(g[name + "_rays"]), name
        assert np.array_equal(samples, g[name + "_samples"]), name


def test_spherical_cylindrical_and_planar_texture_mappings(oracle):
    compare(scenes.mapped_walls(64, 64, 8), oracle)


# ---- alpha / shadow-alpha masks (k_trace<.., ALPHA>, triangle.rs:313-330, 593-654) ----
@pytest.mark.parametrize("kw", [dict(alpha="masks"), dict(alpha="masks", materials="mixed", lights="delta", sampler="halton"),
                                dict(alpha="masks", integrator=("direct", "all"), lightsamples=2), dict(alpha="masks", integrator=("ao", 8, True)),
                                dict(alpha="masks", textures="ewa+bump")],
                         ids=["path", "mixed-halton-delta", "direct-all", "ao", "with-material-textures"])
def test_alpha_and_shadow_alpha_masks(oracle, kw):
    a = dict(xres=64, yres=64, spp=8)
    a.update(kw)
    compare(scenes.cornell_box(**a), oracle)


def test_alpha_masks_ray_casts(oracle):
    from rs_pbrt_b200 import GpuScene
    h = scenes.cornell_box(xres=8, yres=8, spp=1, alpha="masks")
    rng = np.random.default_rng(4)
    n = 200_000
    o = rng.uniform(20, 530, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    g = GpuScene(h.desc, 0)
    pg, tg, bg, sg = g.intersect(o, d)
    og, sp = g.intersect_p(o, d)
    g.close()
    osc = oracle.OracleScene(h.desc)
    po, to, bo, so = osc.intersect(o, d)
    oo, sq = osc.intersect_p(o, d)
    assert np.array_equal(pg, po) and np.array_equal(tg.view(np.uint32), to.view(np.uint32)) and np.array_equal(bg.view(np.uint32), bo.view(np.uint32))
    assert np.array_equal(og, oo)
    assert (sg["nodes_visited"], sg["tris_tested"], sp["nodes_visited"], sp["tris_tested"]) == (so["nodes_visited"], so["tris_tested"], sq["nodes_visited"], sq["tris_tested"])
