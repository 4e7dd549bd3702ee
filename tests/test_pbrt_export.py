"""The .pbrt emitter (rs_pbrt_b200/pbrt_export.py; SURVEY.md 8d: every synthetic scene also leaves as a file a real rs_pbrt build can
render): an exported scene, read back by tests/pbrt_reader.py, must be the same scene -- same flat description, same oracle render."""
import numpy as np
import pytest

import oracle_lib
import pbrt_reader
from rs_pbrt_b200 import pbrt_export, scenes


def q8(h):
    return h


def desc_equal(a, b):
    da, db = a.desc.contents, b.desc.contents
    # (materials are written per shape and re-declared on reading: their count may differ, what they are may not -- the renders below)
    assert (da.n_nodes, da.n_tris, da.n_meshes, da.n_lights, da.n_instances, da.n_textures) == (db.n_nodes, db.n_tris, db.n_meshes, db.n_lights, db.n_instances, db.n_textures)
    na = np.frombuffer((np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (32 * da.n_nodes)).from_address(np.ctypeslib.ctypes.addressof(da.nodes.contents)))), np.uint8)
    nb = np.frombuffer((np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (32 * db.n_nodes)).from_address(np.ctypeslib.ctypes.addressof(db.nodes.contents)))), np.uint8)
    assert np.array_equal(na, nb), "BVH differs"


@pytest.mark.parametrize("make", [
    lambda: scenes.cornell_box(xres=24, yres=24, spp=4),
    lambda: scenes.cornell_box(xres=20, yres=16, spp=4, materials="mixed", lights="delta", sampler="halton", filter="gaussian", xwidth=1.5, ywidth=1.5, lensradius=4.0, focaldistance=900.0),
    lambda: scenes.cornell_box(xres=16, yres=16, spp=2, integrator=("direct", "all"), lightsamples=2, materials="mixed"),
    lambda: scenes.statue(n_side=30, xres=16, yres=16, spp=2),
    lambda: scenes.landscape(xres=32, yres=18, spp=2, n_trees=20, grid=16, detail=6, sky="constant", instancing="reference", n_prototypes=3),
    lambda: scenes.cornell_box(xres=16, yres=16, spp=2, materials="translucent"),
    lambda: scenes.cornell_box(xres=16, yres=16, spp=2, materials="mix"),
], ids=["cornell", "mixed-delta-halton-thinlens", "direct-all", "statue-ply", "landscape-instances", "translucent", "mix"])
def test_export_read_back_renders_identically(tmp_path, make):
    h = make()
    notes = pbrt_export.write(h, tmp_path / "scene.pbrt", ply_threshold=500)
    assert (tmp_path / "scene.pbrt").read_text().count("WorldEnd") == 1
    h2 = pbrt_reader.read(tmp_path / "scene.pbrt")
    desc_equal(h, h2)
    f1, s1, st1 = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=4, want_samples=True)
    f2, s2, st2 = oracle_lib.OracleScene(h2.desc).render(h2.params, n_threads=4, want_samples=True)
    assert st1["rays"] == st2["rays"] and np.array_equal(s1, s2), notes


def test_export_of_textures_and_alpha_masks_is_exact_for_8_bit_texels(tmp_path):
    """Images leave as 8-bit PNGs (what rs_pbrt reads, imagemap.rs:44-57): with texels that are multiples of 1/255 the round trip is exact."""
    h = scenes.cornell_box(xres=20, yres=20, spp=2, textures="ewa+float+graph+bump", alpha="masks", quantize_textures=True)
    notes = pbrt_export.write(h, tmp_path / "tex.pbrt")
    assert not [n for n in notes if "quantised" in n], notes
    h2 = pbrt_reader.read(tmp_path / "tex.pbrt")
    desc_equal(h, h2)
    _, s1, st1 = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=4, want_samples=True)
    _, s2, st2 = oracle_lib.OracleScene(h2.desc).render(h2.params, n_threads=4, want_samples=True)
    assert st1["rays"] == st2["rays"] and np.array_equal(s1, s2)
