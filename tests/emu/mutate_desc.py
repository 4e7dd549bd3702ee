"""TEST INFRASTRUCTURE (tests/test_emu_abi_mutations.py runs it in a child process, so that a crash is a failed test and not a dead test run).
One field of a valid PbrtSceneDesc at a time is set to an out-of-range or otherwise hostile value -- node offsets, leaf ranges, vertex / mesh /
material / light / texture / instance indices, kinds and counts -- and handed to pbrt_gpu_scene_create of the emulation library: the description
must be rejected (PBRT_E_INVALID / PBRT_E_UNSUPPORTED), or, where the value happens to be harmless (an unused field, a smaller count), render to the end.
    python tests/emu/mutate_desc.py cornell|textured|landscape [seed]"""
import sys, ctypes as C, numpy as np, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from rs_pbrt_b200 import GpuScene, _abi, scenes
emu = _abi.bind(C.CDLL(os.environ.get("RS_PBRT_EMU_LIB") or str(ROOT / "tests" / "emu" / "_build" / "librs_pbrt_b200_emu.so")))  # RS_PBRT_EMU_LIB: a sanitizer build
which = sys.argv[1]
h = {"cornell": lambda: scenes.cornell_box(xres=6, yres=6, spp=1, materials="mix", lights="delta"),
     "textured": lambda: scenes.cornell_box(xres=6, yres=6, spp=1, textures="ewa+float+graph+bump", alpha="masks"),
     "landscape": lambda: scenes.landscape(xres=8, yres=4, spp=1, n_trees=12, grid=8, detail=4, instancing="fixed", n_prototypes=2)}[which]()
d = h.desc.contents
BIG = [0xffffffff, 0x7fffffff, 1 << 20]
def try_render(tag):
    handle = C.c_void_p()
    rc = emu.pbrt_gpu_scene_create(h.desc, 0, C.byref(handle))
    if rc != 0:
        print("rejected", tag, rc, emu.pbrt_gpu_last_error().decode()[:70], flush=True)
        return
    print("ACCEPTED", tag, flush=True)
    rp = h.params.contents
    film = np.zeros((rp.cropped_pixel_bounds[3] - rp.cropped_pixel_bounds[1], rp.cropped_pixel_bounds[2] - rp.cropped_pixel_bounds[0], 4), np.float32)
    st = _abi.PbrtStats()
    rect = (C.c_int32 * 4)(*list(rp.sample_bounds)); rc = emu.pbrt_gpu_render(handle, h.params, rect, film.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st))
    print("  rendered", tag, rc, flush=True)
    emu.pbrt_gpu_scene_destroy(handle)
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
def mutate_field(arr, n, field, values, sub=None, label=""):
    if n == 0: return
    for v in values:
        i = int(rng.integers(0, n))
        obj = arr[i]
        old = getattr(obj, field) if sub is None else getattr(obj, field)[sub]
        try:
            if sub is None: setattr(obj, field, v)
            else: getattr(obj, field)[sub] = v
        except (TypeError, OverflowError):
            continue
        try_render("%s[%d].%s%s=%s" % (label, i, field, "" if sub is None else "[%d]" % sub, v))
        if sub is None: setattr(obj, field, old)
        else: getattr(obj, field)[sub] = old
try_render("unmodified")
mutate_field(d.nodes, d.n_nodes, "offset", [-1, 0, 1, d.n_nodes, d.n_nodes + 5, 0x7fffffff, d.n_tris, d.n_tris + 1], label="nodes")
mutate_field(d.nodes, d.n_nodes, "n_prims", [1, 2, 16, 255, 65535], label="nodes")
mutate_field(d.nodes, d.n_nodes, "axis", [3, 200], label="nodes")
for k in range(3): mutate_field(d.tris, d.n_tris, "v", BIG, sub=k, label="tris")
mutate_field(d.tris, d.n_tris, "mesh", [d.n_meshes, 0x7fffffff, 0xffffffff], label="tris")
mutate_field(d.tris, d.n_tris, "material", [d.n_materials, 0x7fffffff, 0xfffffffe], label="tris")
mutate_field(d.tris, d.n_tris, "area_light", [d.n_lights, 0x7fffffff, -2, 0], label="tris")
mutate_field(d.meshes, d.n_meshes, "n_verts", [0, 1], label="meshes")
mutate_field(d.meshes, d.n_meshes, "alpha", [d.n_textures + 1, 0xffffffff], label="meshes")
mutate_field(d.meshes, d.n_meshes, "shadow_alpha", [d.n_textures + 1, 0xffffffff], label="meshes")
mutate_field(d.materials, d.n_materials, "kind", [9, 100, 0xffffffff, 8], label="materials")
mutate_field(d.materials, d.n_materials, "bump", [d.n_textures + 1, 0xffffffff], label="materials")
for g in (0, 1, 7): mutate_field(d.materials, d.n_materials, "tex", [d.n_textures + 1, 0xffffffff], sub=g, label="materials")
mutate_field(d.lights, d.n_lights, "kind", [5, 99, 0xffffffff], label="lights")
mutate_field(d.lights, d.n_lights, "tri", [d.n_tris, 0xffffffff], label="lights")
mutate_field(d.lights, d.n_lights, "n_samples", [0, 1 << 30], label="lights")
if d.n_instances:
    mutate_field(d.instances, d.n_instances, "root", [0, d.n_nodes, 0xffffffff, 1], label="instances")
if d.n_textures:
    mutate_field(d.textures, d.n_textures, "kind", [4, 99], label="textures")
    mutate_field(d.textures, d.n_textures, "channels", [0, 2, 4], label="textures")
    mutate_field(d.textures, d.n_textures, "wrap", [3, 99], label="textures")
    mutate_field(d.textures, d.n_textures, "mapping", [4, 99], label="textures")
    for k in range(3): mutate_field(d.textures, d.n_textures, "child", [d.n_textures + 1, 0xffffffff, d.n_textures], sub=k, label="textures")
    for k in range(2): mutate_field(d.textures, d.n_textures, "res", [0, 1 << 20], sub=k, label="textures")
# non-finite numbers where a renderer only ever puts finite ones: the camera, the world bound, instance matrices, vertex positions, light parameters --
# garbage in the film is the caller's to keep, a crash or a hang is not
NAN, INF = float("nan"), float("inf")
for v in (NAN, INF, 0.0):
    for k in (0, 5, 11, 15):
        old = d.camera.raster_to_camera[k]; d.camera.raster_to_camera[k] = v; try_render("camera.raster_to_camera[%d]=%r" % (k, v)); d.camera.raster_to_camera[k] = old
        old = d.camera.camera_to_world[k]; d.camera.camera_to_world[k] = v; try_render("camera.camera_to_world[%d]=%r" % (k, v)); d.camera.camera_to_world[k] = old
    for name in ("lens_radius", "focal_distance"):
        old = getattr(d.camera, name); setattr(d.camera, name, v if v == v else NAN); try_render("camera.%s=%r" % (name, v)); setattr(d.camera, name, old)
    for k in (0, 4):
        old = d.world_bound[k]; d.world_bound[k] = v; try_render("world_bound[%d]=%r" % (k, v)); d.world_bound[k] = old
    if d.n_instances:
        for k in (0, 3, 10):
            mutate_field(d.instances, d.n_instances, "m", [v], sub=k, label="instances")
            mutate_field(d.instances, d.n_instances, "m_inv", [v], sub=k, label="instances")
    for k in range(3):
        mutate_field(d.lights, d.n_lights, "L", [v], sub=k, label="lights")
        mutate_field(d.lights, d.n_lights, "p", [v], sub=k, label="lights")
    i = int(rng.integers(0, d.n_meshes)); j = int(rng.integers(0, 3 * d.meshes[i].n_verts))
    old = d.meshes[i].p[j]; d.meshes[i].p[j] = v; try_render("meshes[%d].p[%d]=%r" % (i, j, v)); d.meshes[i].p[j] = old
    mutate_field(d.nodes, d.n_nodes, "pmin", [v], sub=1, label="nodes")
    mutate_field(d.nodes, d.n_nodes, "pmax", [v], sub=2, label="nodes")
for name in ("n_nodes", "n_tris", "n_meshes", "n_materials", "n_lights", "n_instances", "n_textures"):
    old = getattr(d, name)
    for v in (0, old - 1 if old else 0, old + 1):
        if v == old: continue
        setattr(d, name, v); try_render("%s=%d (was %d)" % (name, v, old)); setattr(d, name, old)
print("done")
