"""CPU tests of the product's host side: the C ABI library loads and exports every declared symbol, the C++
host mirror's BVHAccel::new matches the oracle's restatement bit for bit, and the GPU entry points fail loudly
without a device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from rs_pbrt_b200 import _abi, bvh_build, scenes

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(product_lib):
    names = set()
    for hdr in ("pbrt_gpu.h", "pbrt_host.h"):
        txt = (ROOT / "include" / hdr).read_text()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(pbrt_(?:gpu|host)_[a-z_0-9]+)\s*\(", txt))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(product_lib, n), n
    assert set(_abi.GPU_SYMBOLS + _abi.HOST_SYMBOLS) == names
    assert product_lib.pbrt_gpu_abi_version() == 4


def test_struct_sizes_match_the_header(tmp_path):
    """The ctypes mirror against include/pbrt_gpu.h itself: a C program prints sizeof of every struct of the ABI."""
    import shutil
    import subprocess
    assert C.sizeof(_abi.PbrtBvhNode) == 32  # == LinearBVHNode, copied verbatim
    assert C.sizeof(_abi.PbrtTri) == 24
    names = ["PbrtBvhNode", "PbrtTri", "PbrtMesh", "PbrtTexture", "PbrtMaterial", "PbrtLight", "PbrtCamera", "PbrtInstance", "PbrtSceneDesc",
             "PbrtRenderParams", "PbrtStats"]
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "pbrt_gpu.h"\nint main(void) {\n' +
                   "".join('printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in names) + "return 0; }\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for n in names:
        assert C.sizeof(getattr(_abi, n)) == int(out[n]), n


def _bounds_of(tris):
    return np.concatenate([tris.min(axis=1), tris.max(axis=1)], axis=1).astype(np.float32)


@pytest.mark.parametrize("n,seed,threads", [(1, 0, 1), (2, 1, 1), (3, 2, 1), (5, 3, 2), (64, 4, 1), (1000, 5, 4), (20000, 6, 8), (70000, 7, 8)])
def test_bvh_build_matches_oracle(oracle, product_lib, n, seed, threads):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-5, 5, (n, 1, 3))
    tris = (c + rng.normal(scale=0.2, size=(n, 3, 3))).astype(np.float32)
    b = _bounds_of(tris)
    nodes_p, ord_p = bvh_build(b, 4, threads)
    nodes_o, ord_o = oracle.bvh_build(b, 4)
    assert np.array_equal(ord_p, ord_o)
    assert np.array_equal(nodes_p, nodes_o)
    assert sorted(ord_p.tolist()) == list(range(n))


def test_bvh_build_degenerate_inputs(oracle, product_lib):
    # identical centroids -> one leaf holding everything; coplanar duplicates; max_prims 1
    b = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (9, 1))
    for mp in (1, 4, 255):
        a, oa = bvh_build(b, mp, 1)
        o, oo = oracle.bvh_build(b, mp)
        assert np.array_equal(a, o) and np.array_equal(oa, oo) and len(a) == 1
    rng = np.random.default_rng(1)
    pts = rng.uniform(0, 1, (300, 3)).astype(np.float32)
    pts[:, 1] = 0.5
    b = np.concatenate([pts, pts], 1)
    for mp in (1, 2, 4):
        a, oa = bvh_build(b, mp, 2)
        o, oo = oracle.bvh_build(b, mp)
        assert np.array_equal(a, o) and np.array_equal(oa, oo)
    a, oa = bvh_build(np.zeros((0, 6), np.float32), 4, 1)
    assert len(a) == 0 and len(oa) == 0


def test_host_scene_description_is_consistent(product_lib):
    h = scenes.cornell_box(xres=20, yres=12, spp=5)
    d, rp = h.desc.contents, h.params.contents
    assert d.n_tris == 32 and d.n_lights == 2 and d.n_materials == 4
    assert rp.spp == 8  # SobolSampler rounds up to a power of two (sobol.rs:39-45)
    assert list(rp.sample_bounds) == [0, 0, 20, 12] and list(rp.cropped_pixel_bounds) == [0, 0, 20, 12]
    lights = [d.lights[i] for i in range(2)]
    for i, l in enumerate(lights):
        assert d.tris[l.tri].area_light == i
        assert abs(l.area - 0.5 * 130 * 105) < 1e-2
    assert list(d.world_bound) == [0, 0, 0, 555, 555, 555]
    leaves = sum(d.nodes[i].n_prims for i in range(d.n_nodes))
    assert leaves == 32


def test_gpu_entry_points_fail_loudly_without_a_device(product_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = scenes.cornell_box(xres=8, yres=8, spp=1)
    handle = C.c_void_p()
    rc = product_lib.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle))
    assert rc == _abi.PBRT_E_NO_DEVICE and not handle
    assert b"no CPU fallback" in product_lib.pbrt_gpu_last_error()
    with pytest.raises(Exception):
        h.render(device=0)


def test_scene_validation_rejects_bad_input(product_lib):
    h = scenes.cornell_box(xres=8, yres=8, spp=1)
    d = h.desc.contents
    handle = C.c_void_p()
    old = d.tris[0].mesh
    d.tris[0].mesh = 99
    try:
        assert product_lib.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle)) == _abi.PBRT_E_INVALID
    finally:
        d.tris[0].mesh = old
    old = d.materials[0].kind
    d.materials[0].kind = 42
    try:
        assert product_lib.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle)) == _abi.PBRT_E_UNSUPPORTED
    finally:
        d.materials[0].kind = old
    assert product_lib.pbrt_gpu_scene_create(None, 0, C.byref(handle)) == _abi.PBRT_E_INVALID


def test_every_scene_builder_runs_at_tiny_size(product_lib):
    """The GPU-only tests are the only other users of the large scene builders: keep them importable and runnable on the CPU."""
    from rs_pbrt_b200 import scenes
    assert scenes.statue(n_side=16, xres=16, yres=16, spp=2).n_tris > 0
    assert scenes.conference(xres=16, yres=16, spp=2, n_chairs=2, detail=4, n_light_quads=4).n_tris > 0
    assert scenes.sky_scene(xres=8, yres=8, spp=2, sampler="halton", env="two").params.contents.sampler == 1
    assert scenes.cornell_box(xres=8, yres=8, spp=3, sampler="halton", lights="delta").params.contents.spp == 3


def test_device_math_header_on_the_host(tmp_path):
    """pb_math.cuh compiled for the host (intrinsics shimmed): fdiv0 -- the zero-numerator shortcut around div.rn's slow path -- equals the
    IEEE division bit for bit, and the libm restatements in the header itself (not only their C twins under tools/checks) equal libm."""
    import shutil
    import subprocess
    if not shutil.which("g++") or not Path("/usr/local/cuda/include/cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    exe = tmp_path / "fdiv0_check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I/usr/local/cuda/include", str(ROOT / "tools" / "checks" / "fdiv0_check.cpp"),
                    "-o", str(exe), "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "fdiv0: 0 mismatches" in r.stdout and "vs libm: 0 mismatches" in r.stdout


def test_bench_cornell_variants_build():
    """bench.py's Cornell workloads for the rows that are still to be measured describe what they claim."""
    import bench
    expect = {"cornell": (_abi.INTEGRATOR_PATH, 0), "cornell-textured": (_abi.INTEGRATOR_PATH, 1), "cornell-direct": (_abi.INTEGRATOR_DIRECT, 0),
              "cornell-whitted": (_abi.INTEGRATOR_WHITTED, 0), "cornell-ao": (_abi.INTEGRATOR_AO, 0)}
    for name, (integ, textured) in expect.items():
        h = bench.make_scene(name)
        rp, d = h.params.contents, h.desc.contents
        assert rp.integrator == integ and (d.n_textures > 0) == bool(textured), name
        assert rp.spp == bench.WORKLOADS[name]["spp"] and d.n_tris >= 32
