"""The C ABI's validation of caller-supplied scene descriptions (pbrt_gpu_scene_create): hostile values in one field at a time are
rejected or harmless -- never a crash, never an out-of-range index handed to a kernel.  Runs against the kernel emulation library
(the host side of pbrt_gpu.cu is the product's own code there), in a child process per scene."""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.mark.parametrize("scene", ["cornell", "textured", "landscape"])
@pytest.mark.parametrize("seed", [1, 2])
def test_hostile_descriptions_are_rejected_or_harmless(scene, seed):
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu

    build_emu.build()
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "emu" / "mutate_desc.py"), scene, str(seed)], capture_output=True, text=True, timeout=900)
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and lines and lines[-1] == "done", (r.returncode, lines[-3:], r.stderr[-2000:])
    accepted = [l for l in lines if l.startswith("ACCEPTED")]
    rendered = [l for l in lines if l.startswith("  rendered")]
    rejected = [l for l in lines if l.startswith("rejected")]
    assert len(accepted) == len(rendered) and "ACCEPTED unmodified" in accepted
    assert all(l.split()[-1] in ("0", "-1", "-2") for l in rendered), rendered
    assert len(rejected) >= 50, len(rejected)
    # what must never get through: indices past the caller's arrays (a TransformedPrimitive record of the landscape scene reads only v[0], so there
    # a wild v[1] / v[2] / material is one of the harmless cases)
    for needle in () if scene == "landscape" else (".material=", ".mesh=", "tris[", "].offset=%d" % 0x7fffffff):
        # (area_light = -2 means "none" like -1; mesh = 0xffffffff makes the record a TransformedPrimitive, legitimate when v[0] names an instance)
        bad = [l for l in accepted if needle in l and "area_light" not in l and not l.endswith(".mesh=4294967295")]
        assert not bad, bad


def test_hostile_render_parameters_return():
    """PbrtRenderParams with zero / non-power-of-two sample counts, unknown enumerators, NaN / zero / negative / infinite filter radii, bounds far outside
    the film, a "maxdepth" of 2^32 - 1: every pbrt_gpu_render call returns (an error code or a film) -- none crashes or runs away."""
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu

    build_emu.build()
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "emu" / "mutate_params.py")], capture_output=True, text=True, timeout=600)
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and lines and lines[-1] == "done", (r.returncode, lines[-3:], r.stderr[-2000:])
    out = {l.split()[1]: int(l.split()[2]) for l in lines if l.startswith("returned")}
    assert out["unmodified"] == 0 and len(out) >= 55
    assert out["spp=0"] == -1 and out["spp=3"] == -1 and out["sampler=99"] == -2 and out["integrator=99"] == -2 and out["light_strategy=99"] == -1
    assert out["filter_radius[0]=nan"] == -1 and out["filter_radius[1]=0.0"] == -1
    assert out["max_depth=4294967295"] == 0


def _emu():
    import ctypes as C

    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu
    from rs_pbrt_b200 import _abi

    return _abi.bind(C.CDLL(str(build_emu.build())))


def test_bvh_deeper_than_the_traversal_stack_is_refused():
    """A caller-built BVH whose far children nest deeper than 64 levels would run past the kernels' traversal stack (the reference's nodes_to_visit
    has 64 entries as well and panics, bvh.rs:420): PBRT_E_UNSUPPORTED at scene creation.  A chain of exactly 64 pending far children is fine."""
    import ctypes as C

    from rs_pbrt_b200 import _abi, scenes

    emu = _emu()
    h = scenes.cornell_box(xres=4, yres=4, spp=1)
    d = h.desc.contents
    keep = (d.nodes, d.n_nodes)

    def chain(n_interior):  # I0 L I1 L ... : every interior node's first child is the next interior, its far child a leaf => the stack grows by one per level
        nodes = (_abi.PbrtBvhNode * (2 * n_interior + 1))()
        for i, n in enumerate(nodes):
            for k in range(3):
                n.pmin[k], n.pmax[k] = d.world_bound[k], d.world_bound[3 + k]
        # layout: interior k at index k (first child = k + 1), leaves after the chain: the far child of interior k is leaf n_interior + 1 + (n_interior - 1 - k)
        for k in range(n_interior):
            nodes[k].n_prims = 0
            nodes[k].axis = k % 3
            nodes[k].offset = 2 * n_interior - k  # second child
        for j in range(n_interior, 2 * n_interior + 1):
            nodes[j].n_prims = 1
            nodes[j].offset = j % d.n_tris
        return nodes

    try:
        for depth, want in ((40, 0), (64, 0), (65, _abi.PBRT_E_UNSUPPORTED), (300, _abi.PBRT_E_UNSUPPORTED)):
            nodes = chain(depth)
            d.nodes, d.n_nodes = C.cast(nodes, C.POINTER(_abi.PbrtBvhNode)), len(nodes)
            handle = C.c_void_p()
            rc = emu.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle))
            assert rc == want, (depth, rc, emu.pbrt_gpu_last_error())
            if rc == 0:
                emu.pbrt_gpu_scene_destroy(handle)
            else:
                assert b"64-entry traversal stack" in emu.pbrt_gpu_last_error()
    finally:
        d.nodes, d.n_nodes = keep


def test_nested_object_instance_is_refused():
    """A TransformedPrimitive record inside an object's own tree (nested instancing): the traversal keeps one current instance, so this is refused at
    scene creation rather than rendered wrongly."""
    import ctypes as C

    from rs_pbrt_b200 import _abi, scenes

    emu = _emu()
    h = scenes.landscape(xres=8, yres=4, spp=1, n_trees=6, grid=6, detail=4, instancing="fixed", n_prototypes=2)
    d = h.desc.contents
    root = d.instances[0].root
    i = root
    while d.nodes[i].n_prims == 0:  # first leaf of the object's tree
        i += 1
    t = d.tris[d.nodes[i].offset]
    old = (t.mesh, t.v[0])
    t.mesh, t.v[0] = 0xFFFFFFFF, 1
    try:
        handle = C.c_void_p()
        assert emu.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle)) == _abi.PBRT_E_UNSUPPORTED
        assert b"nested instancing" in emu.pbrt_gpu_last_error()
    finally:
        t.mesh, t.v[0] = old


def test_null_pointers_and_empty_batches_at_every_entry_point():
    """Required arguments that are null, zero-length ray batches, part numbers outside [0, n_parts), a null handle in the device list: PBRT_E_INVALID (or a
    no-op for the destroy / zero-length cases), never a dereference."""
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu

    build_emu.build()
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "emu" / "misuse.py")], capture_output=True, text=True, timeout=300)
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and lines and lines[-1] == "done", (r.returncode, lines[-3:], r.stderr[-2000:])
    out = {" ".join(l.split()[:-1]): l.split()[-1] for l in lines[:-1] if not l.startswith(("destroy", "render null rect", "host_register", "bytes"))}
    assert out.pop("create") == "0" and out.pop("intersect n=0") == "0" and out.pop("intersect null arrays n=0") == "0"
    assert out and all(v == "-1" for v in out.values()), out
