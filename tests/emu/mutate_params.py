"""TEST INFRASTRUCTURE (tests/test_emu_abi_mutations.py): hostile PbrtRenderParams against the emulation library -- inverted / empty / huge bounds, zero
and huge sample counts, unknown sampler / integrator / strategy numbers, NaN / zero / negative / huge filter radii.  Every call returns (an error or a
film); none crashes, none runs away.    python tests/emu/mutate_params.py"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from rs_pbrt_b200 import _abi, scenes

emu = _abi.bind(C.CDLL(os.environ.get("RS_PBRT_EMU_LIB") or str(ROOT / "tests" / "emu" / "_build" / "librs_pbrt_b200_emu.so")))  # RS_PBRT_EMU_LIB: a sanitizer build
h = scenes.cornell_box(xres=6, yres=6, spp=1, materials="mixed", lights="delta")
handle = C.c_void_p()
assert emu.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle)) == 0
rp = h.params.contents
film = np.zeros((64, 64, 4), np.float32)  # larger than any accepted cropped window below


def render(tag):
    cb = list(rp.cropped_pixel_bounds)
    w, hgt = cb[2] - cb[0], cb[3] - cb[1]
    if w * hgt > 64 * 64 or w < 0 or hgt < 0:  # (the caller owns the film: a window larger than its buffer is the caller's bug, not the library's)
        buf = np.zeros((max(hgt, 0) if hgt < 4096 else 0, max(w, 0) if w < 4096 else 0, 4), np.float32) if 0 <= w < 4096 and 0 <= hgt < 4096 else None
    else:
        buf = film
    st = _abi.PbrtStats()
    rect = (C.c_int32 * 4)(*list(rp.sample_bounds))
    if buf is None:
        print("skipped", tag, "(no film of that size)", flush=True)
        return
    rc = emu.pbrt_gpu_render(handle, h.params, rect, buf.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st))
    print("returned", tag, rc, (emu.pbrt_gpu_last_error().decode()[:60] if rc else ""), flush=True)


def with_field(name, values, sub=None):
    for v in values:
        old = getattr(rp, name) if sub is None else getattr(rp, name)[sub]
        try:
            if sub is None:
                setattr(rp, name, v)
            else:
                getattr(rp, name)[sub] = v
        except (TypeError, OverflowError):
            continue
        render("%s%s=%r" % (name, "" if sub is None else "[%d]" % sub, v))
        if sub is None:
            setattr(rp, name, old)
        else:
            getattr(rp, name)[sub] = old


render("unmodified")
with_field("spp", [0, 3])  # (a huge count is a legitimate, long render)
with_field("max_depth", [0, 1 << 30, 0xffffffff])
with_field("rr_threshold", [float("nan"), -1.0, float("inf")])
with_field("light_strategy", [3, 99])
with_field("sampler", [2, 99])
with_field("integrator", [4, 99])
with_field("direct_strategy", [2, 99])
with_field("ao_samples", [0])
with_field("instancing", [2, 99])
with_field("flags", [0xffffffff])
with_field("max_sample_luminance", [float("nan"), 0.0, -1.0])
for k in range(2):
    with_field("filter_radius", [float("nan"), 0.0, -1.0, 1e9, float("inf")], sub=k)
for k in range(4):
    with_field("sample_bounds", [-(1 << 30), 1 << 30, 3], sub=k)
    with_field("cropped_pixel_bounds", [-(1 << 30), 1 << 30, 3], sub=k)
    with_field("pixel_bounds", [-(1 << 30), 1 << 30, 3], sub=k)
with_field("filter_table", [float("nan"), float("inf")], sub=0)
emu.pbrt_gpu_scene_destroy(handle)
print("done")
