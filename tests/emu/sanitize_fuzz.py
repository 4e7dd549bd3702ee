"""The randomised scenes of tests/test_emu_kernels.py rendered through an instrumented emulation build (thread engine): out-of-bounds accesses and undefined
behaviour in branches the fixed sanitizer scenes do not reach.  No oracle here -- parity is the test suite's business; this only asks the sanitizer.
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tests/emu/sanitize_fuzz.py asan 0 60
    LD_PRELOAD=$(gcc -print-file-name=libubsan.so) UBSAN_OPTIONS=print_stacktrace=1 python tests/emu/sanitize_fuzz.py usan 0 60"""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import test_emu_kernels as T  # noqa: E402
from rs_pbrt_b200 import GpuScene, _abi  # noqa: E402
from rs_pbrt_b200.host import PbrtError  # noqa: E402

E = _abi.bind(C.CDLL(str(ROOT / "tests" / "emu" / "_build" / ("librs_pbrt_b200_emu_%s.so" % sys.argv[1]))))
lo, hi = int(sys.argv[2]), int(sys.argv[3])
rays = [0]


def render_only(emu, oracle, h, rect=None, count_work=False, exact_weights=True):
    rp = h.params.contents
    if count_work:
        rp.flags |= _abi.RENDER_COUNT_WORK
    g = GpuScene(h.desc, 0, lib=E)
    try:
        _, st = g.render_samples(h.params, rect or list(rp.sample_bounds))
        g.render(h.params, rect=rect or list(rp.sample_bounds))
    finally:
        g.close()
    rays[0] += st["rays"]
    return st


T.check = render_only
for name in ("test_randomised_materials_and_settings", "test_randomised_scene_families", "test_randomised_textures", "test_randomised_triangle_soups_and_lights",
             "test_randomised_object_instances"):
    fn = getattr(T, name)
    rays[0] = 0
    for seed in range(lo, hi):
        try:
            fn(E, None, seed)
        except (PbrtError, AttributeError):  # a refused scene (too many sampler dimensions); None has no OracleScene for the refusal cross-check
            pass
    print(name, "seeds %d..%d" % (lo, hi), rays[0], flush=True)
print("done")
