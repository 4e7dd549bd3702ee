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
