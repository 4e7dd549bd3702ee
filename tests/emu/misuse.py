"""TEST INFRASTRUCTURE (tests/test_emu_abi_mutations.py): null pointers, empty batches and out-of-range part numbers at every entry point of the C ABI -- each call
returns a status (PBRT_E_INVALID where the argument is required), none dereferences a null pointer.    python tests/emu/misuse.py"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from rs_pbrt_b200 import _abi, scenes
emu = _abi.bind(C.CDLL(os.environ.get("RS_PBRT_EMU_LIB") or str(ROOT / "tests" / "emu" / "_build" / "librs_pbrt_b200_emu.so")))  # RS_PBRT_EMU_LIB: a sanitizer build
h = scenes.cornell_box(xres=4, yres=4, spp=1)
handle = C.c_void_p()
print("create null out", emu.pbrt_gpu_scene_create(h.desc, 0, None))
print("create", emu.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle)))
st = _abi.PbrtStats()
film = np.zeros((4, 4, 4), np.float32)
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
print("render null scene", emu.pbrt_gpu_render(None, h.params, None, fp(film), C.byref(st)), flush=True)
print("render null params", emu.pbrt_gpu_render(handle, None, None, fp(film), C.byref(st)), flush=True)
print("render null film", emu.pbrt_gpu_render(handle, h.params, None, None, C.byref(st)), flush=True)
print("render null stats", emu.pbrt_gpu_render(handle, h.params, None, fp(film), None), flush=True)
print("render null rect", emu.pbrt_gpu_render(handle, h.params, None, fp(film), C.byref(st)), st.rays, flush=True)
o = np.zeros((1, 3), np.float32); d = np.ones((1, 3), np.float32); tm = np.full(1, np.inf, np.float32)
prim = np.zeros(1, np.int32); t = np.zeros(1, np.float32); b = np.zeros((1, 3), np.float32); occ = np.zeros(1, np.uint8)
ip = prim.ctypes.data_as(C.POINTER(C.c_int32))
print("intersect n=0", emu.pbrt_gpu_intersect(handle, 0, fp(o), fp(d), fp(tm), ip, fp(t), fp(b), C.byref(st)), flush=True)
print("intersect null arrays n=0", emu.pbrt_gpu_intersect(handle, 0, None, None, None, None, None, None, None), flush=True)
print("intersect null o n=1", emu.pbrt_gpu_intersect(handle, 1, None, fp(d), fp(tm), ip, fp(t), fp(b), C.byref(st)), flush=True)
print("intersect null tmax n=1", emu.pbrt_gpu_intersect(handle, 1, fp(o), fp(d), None, ip, fp(t), fp(b), C.byref(st)), flush=True)
print("intersect_p null occ", emu.pbrt_gpu_intersect_p(handle, 1, fp(o), fp(d), fp(tm), None, C.byref(st)), flush=True)
print("intersect null scene", emu.pbrt_gpu_intersect(None, 1, fp(o), fp(d), fp(tm), ip, fp(t), fp(b), C.byref(st)), flush=True)
print("samples null", emu.pbrt_gpu_render_samples(handle, h.params, None, None, C.byref(st)), flush=True)
print("tiles n_parts=0", emu.pbrt_gpu_render_tiles_device(handle, h.params, 0, 0, film.ctypes.data, None, C.byref(st)), flush=True)
print("tiles part>=n", emu.pbrt_gpu_render_tiles_device(handle, h.params, 3, 2, film.ctypes.data, None, C.byref(st)), flush=True)
arr = (C.c_void_p * 2)(handle, None)
print("multi with null scene", emu.pbrt_gpu_render_multi(arr, 2, h.params, fp(film), C.byref(st)), flush=True)
print("multi n=0", emu.pbrt_gpu_render_multi(arr, 0, h.params, fp(film), C.byref(st)), flush=True)
print("bytes null", emu.pbrt_gpu_scene_bytes(None) if hasattr(emu, "pbrt_gpu_scene_bytes") else "n/a", flush=True)
emu.pbrt_gpu_scene_destroy(None); print("destroy null ok", flush=True)
emu.pbrt_gpu_scene_destroy(handle); print("destroy ok", flush=True)
print("host_register null", emu.pbrt_gpu_host_register(None, 0), emu.pbrt_gpu_host_unregister(None), flush=True)
print("done")
