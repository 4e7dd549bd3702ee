"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under rs_pbrt_b200/ imports this.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from rs_pbrt_b200 import _abi

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_build" / "liboracle.so"
_lib = None


def build():
    srcs = list((ROOT / "oracle").glob("*.hpp")) + list((ROOT / "oracle").glob("*.cpp")) + [ROOT / "include" / "pbrt_gpu.h"]
    if not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    return LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    try:
        build()
    except Exception:
        if not LIB.exists():
            raise
    L = C.CDLL(str(LIB))
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
    L.orc_last_error.restype = C.c_char_p
    L.orc_init.argtypes = [C.c_char_p]
    L.orc_scene_create.argtypes = [C.POINTER(_abi.PbrtSceneDesc)]
    L.orc_scene_create.restype = vp
    L.orc_scene_destroy.argtypes = [vp]
    L.orc_scene_destroy.restype = None
    L.orc_render.argtypes = [vp, C.POINTER(_abi.PbrtRenderParams), ip, fp, fp, C.c_int, C.POINTER(_abi.PbrtStats)]
    L.orc_intersect.argtypes = [vp, C.c_uint32, fp, fp, fp, ip, fp, fp, C.POINTER(_abi.PbrtStats)]
    L.orc_intersect_p.argtypes = [vp, C.c_uint32, fp, fp, fp, C.POINTER(C.c_uint8), C.POINTER(_abi.PbrtStats)]
    L.orc_interaction.argtypes = [vp, C.c_int32, fp, fp, fp]
    L.orc_bvh_build.argtypes = [fp, C.c_uint32, C.c_uint32, C.POINTER(_abi.PbrtBvhNode), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_gamma.argtypes = [C.c_int]
    L.orc_gamma.restype = C.c_float
    L.orc_next_float_up.argtypes = [C.c_float]
    L.orc_next_float_up.restype = C.c_float
    L.orc_next_float_down.argtypes = [C.c_float]
    L.orc_next_float_down.restype = C.c_float
    L.orc_offset_ray_origin.argtypes = [fp, fp, fp, fp, fp]
    L.orc_sobol_interval_to_index.argtypes = [C.c_uint32, C.c_uint64, C.c_int32, C.c_int32]
    L.orc_sobol_interval_to_index.restype = C.c_uint64
    L.orc_sobol_sample_float.argtypes = [C.c_int64, C.c_int, C.c_uint32]
    L.orc_sobol_sample_float.restype = C.c_float
    L.orc_radical_inverse.argtypes = [C.c_int, C.c_uint64]
    L.orc_radical_inverse.restype = C.c_float
    L.orc_sampler_dimension.argtypes = [C.POINTER(_abi.PbrtRenderParams), C.c_int32, C.c_int32, C.c_int64, C.c_int, C.POINTER(C.c_uint64)]
    L.orc_sampler_dimension.restype = C.c_float
    L.orc_halton_permutation.argtypes = [C.c_int, C.POINTER(C.c_uint16), C.c_int]
    L.orc_scrambled_radical_inverse.argtypes = [C.c_int, C.c_uint64, C.POINTER(C.c_uint16)]
    L.orc_scrambled_radical_inverse.restype = C.c_float
    L.orc_prime.restype = L.orc_prime_sum.restype = C.c_uint32
    L.orc_prime.argtypes = L.orc_prime_sum.argtypes = [C.c_int]
    L.orc_texture_lookup.argtypes = [C.POINTER(_abi.PbrtTexture), C.c_uint32, fp, fp, fp, fp]
    L.orc_texture_level.argtypes = [C.POINTER(_abi.PbrtTexture), C.c_uint32, fp]
    L.orc_camera_sample.argtypes = [vp, C.POINTER(_abi.PbrtRenderParams), C.c_int32, C.c_int32, C.c_int64, fp]
    L.orc_bsdf.argtypes = [C.POINTER(_abi.PbrtMaterial), fp, fp, fp, fp, fp, fp, C.c_int, fp]
    L.orc_bsdf_at.argtypes = [C.POINTER(_abi.PbrtMaterial), C.c_uint32, C.c_uint32, fp, fp, fp, fp, fp, fp, C.c_int, fp]
    L.orc_light_distribution.argtypes = [vp, C.c_int, fp, fp, fp]
    L.orc_light_distribution.restype = C.c_float
    L.orc_film_add_sample.argtypes = [C.POINTER(_abi.PbrtRenderParams), fp, fp, fp, C.c_float]
    if L.orc_init(str(ROOT / "data" / "sobol_tables.bin").encode()) != 0:
        raise RuntimeError(L.orc_last_error().decode())
    _lib = L
    return L


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class OracleScene:
    def __init__(self, desc):
        self.L = load()
        self.h = self.L.orc_scene_create(desc)
        if not self.h:
            raise RuntimeError(self.L.orc_last_error().decode())

    def close(self):
        if self.h:
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, params, rect=None, n_threads=8, want_samples=False):
        cb = params.contents.cropped_pixel_bounds
        film = np.zeros((cb[3] - cb[1], cb[2] - cb[0], 4), np.float32)
        r = np.ascontiguousarray(rect if rect is not None else list(params.contents.sample_bounds), np.int32)
        samples = None
        if want_samples:
            samples = np.zeros((r[3] - r[1], r[2] - r[0], params.contents.spp, 3), np.float32)
        st = _abi.PbrtStats()
        rc = self.L.orc_render(self.h, params, r.ctypes.data_as(C.POINTER(C.c_int32)), _fptr(film), _fptr(samples) if want_samples else None,
                               n_threads, C.byref(st))
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return film, samples, st.as_dict()

    def intersect(self, o, d, t_max=None):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        tm = np.ascontiguousarray(t_max, np.float32) if t_max is not None else np.full(n, np.inf, np.float32)
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        b = np.zeros((n, 3), np.float32)
        st = _abi.PbrtStats()
        self.L.orc_intersect(self.h, n, _fptr(o), _fptr(d), _fptr(tm), prim.ctypes.data_as(C.POINTER(C.c_int32)), _fptr(t), _fptr(b), C.byref(st))
        return prim, t, b, st.as_dict()

    def intersect_p(self, o, d, t_max=None):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        tm = np.ascontiguousarray(t_max, np.float32) if t_max is not None else np.full(n, np.inf, np.float32)
        occ = np.zeros(n, np.uint8)
        st = _abi.PbrtStats()
        self.L.orc_intersect_p(self.h, n, _fptr(o), _fptr(d), _fptr(tm), occ.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st))
        return occ, st.as_dict()

    def camera_sample(self, params, px, py, s):
        out = np.zeros(11, np.float32)
        self.L.orc_camera_sample(self.h, params, px, py, s, _fptr(out))
        return out

    def light_distribution(self, strategy, p, n_lights):
        func = np.zeros(max(n_lights, 1), np.float32)
        cdf = np.zeros(n_lights + 1, np.float32)
        pp = np.ascontiguousarray(p, np.float32)
        fi = self.L.orc_light_distribution(self.h, strategy, _fptr(pp), _fptr(func), _fptr(cdf))
        return func[:n_lights], cdf, fi


class OracleTexture:
    """MipMap<Spectrum> of an image as ImageTexture::new hands it to MipMap::new (texels: (h, w, 3), row 0 at t = 0)."""

    def __init__(self, texels, trilinear=False, max_anisotropy=8.0, wrap=0):
        self.L = load()
        self.texels = np.ascontiguousarray(texels, np.float32)
        t = _abi.PbrtTexture()
        t.res[0], t.res[1] = self.texels.shape[1], self.texels.shape[0]
        t.texels = _fptr(self.texels)
        t.channels = 3
        t.trilinear, t.max_anisotropy, t.wrap = int(trilinear), max_anisotropy, wrap
        t.su = t.sv = 1.0
        self.t = t

    def lookup(self, st, dst0=None, dst1=None):
        st = np.ascontiguousarray(st, np.float32).reshape(-1, 2)
        z = np.zeros_like(st)
        d0 = np.ascontiguousarray(dst0, np.float32).reshape(-1, 2) if dst0 is not None else z
        d1 = np.ascontiguousarray(dst1, np.float32).reshape(-1, 2) if dst1 is not None else z.copy()
        out = np.zeros((st.shape[0], 3), np.float32)
        assert self.L.orc_texture_lookup(C.byref(self.t), st.shape[0], _fptr(st), _fptr(d0), _fptr(d1), _fptr(out)) == 0
        return out

    def level(self, i):
        r = self.L.orc_texture_level(C.byref(self.t), i, None)
        if r < 0:
            return None
        us, vs = r & 0xffff, r >> 16
        out = np.zeros((vs, us, 3), np.float32)
        self.L.orc_texture_level(C.byref(self.t), i, _fptr(out))
        return out


def bvh_build(bounds, max_prims_in_node=4):
    L = load()
    b = np.ascontiguousarray(bounds, np.float32).reshape(-1, 6)
    n = b.shape[0]
    nodes = (_abi.PbrtBvhNode * max(2 * n, 1))()
    ordered = np.zeros(max(n, 1), np.uint32)
    nn = C.c_uint32(0)
    L.orc_bvh_build(_fptr(b), n, max_prims_in_node, nodes, C.byref(nn), ordered.ctypes.data_as(C.POINTER(C.c_uint32)))
    arr = np.frombuffer(nodes, dtype=np.uint8)[: nn.value * 32].reshape(nn.value, 32).copy()
    return arr, ordered[:n].copy()
