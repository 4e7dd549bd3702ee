"""Scenes rendered through an instrumented emulation build (tools/emu_sanitize.sh): every kernel path once."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
import ctypes as C, numpy as np
from rs_pbrt_b200 import _abi, scenes, GpuScene
E = _abi.bind(C.CDLL(str(ROOT / 'tests' / 'emu' / '_build' / ('librs_pbrt_b200_emu_%s.so' % sys.argv[1]))))
def run(h, name):
    g = GpuScene(h.desc, 0, lib=E); gs, st = g.render_samples(h.params, list(h.params.contents.sample_bounds)); f,_ = g.render(h.params); g.close(); print(name, st["rays"], flush=True)
run(scenes.cornell_box(xres=10, yres=10, spp=2), "cornell")
run(scenes.cornell_box(xres=10, yres=10, spp=2, materials="mixed", lights="delta", strategy="spatial"), "mixed+delta")
run(scenes.sky_scene(xres=10, yres=10, spp=2, env="two"), "sky")
run(scenes.cornell_box(xres=10, yres=10, spp=3, sampler="halton"), "halton")
run(scenes.statue(n_side=40, xres=8, yres=8, spp=2), "statue")
run(scenes.cornell_box(xres=10, yres=10, spp=2, integrator=("ao", 5, True)), "ao")
run(scenes.conference(xres=12, yres=8, spp=2, n_chairs=3, detail=4, n_light_quads=8), "conference")
for mode in ("fixed", "reference"):
    run(scenes.landscape(xres=16, yres=10, spp=2, n_trees=40, grid=12, detail=6, instancing=mode), "landscape-" + mode)
run(scenes.cornell_box(xres=12, yres=12, spp=2, textures="ewa"), "textures-ewa")
run(scenes.cornell_box(xres=12, yres=12, spp=2, textures="trilinear+float", lensradius=6.0, focaldistance=900.0, sampler="halton"), "textures-trilinear")
run(scenes.cornell_box(xres=10, yres=10, spp=2, integrator=("direct", "all"), materials="mixed", lights="delta", lightsamples=2, maxdepth=3), "direct-all")
run(scenes.cornell_box(xres=10, yres=10, spp=2, integrator="whitted", materials="mixed", sampler="halton"), "whitted")
run(scenes.cornell_box(xres=10, yres=10, spp=2, integrator="whitted", textures="trilinear+float+graph+bump"), "whitted-textures")
run(scenes.cornell_box(xres=10, yres=10, spp=2, textures="ewa+float+graph+bump"), "textures-graph-bump")
run(scenes.landscape(xres=12, yres=8, spp=2, n_trees=30, grid=10, detail=6, instancing="fixed", integrator=("direct", "all"), maxdepth=3), "landscape-direct")
run(scenes.landscape(xres=12, yres=8, spp=2, n_trees=30, grid=10, detail=6, instancing="reference", integrator=("ao", 4, True)), "landscape-ao")
run(scenes.mapped_walls(12, 12, 2), "texture-mappings")
# round 2: alpha / shadow-alpha masks, TranslucentMaterial, MixMaterial (sc_opt), a mesh large enough for the wide two-box records, tile shares, the multi-device render
run(scenes.cornell_box(xres=10, yres=10, spp=2, alpha="masks", materials="mixed", lights="delta"), "alpha-masks")
run(scenes.cornell_box(xres=10, yres=10, spp=2, alpha="masks", integrator=("direct", "all"), lightsamples=2), "alpha-masks-direct")
run(scenes.cornell_box(xres=10, yres=10, spp=2, materials="translucent"), "translucent")
run(scenes.cornell_box(xres=10, yres=10, spp=2, materials="mix"), "mix")
run(scenes.cornell_box(xres=10, yres=10, spp=2, materials="mix", integrator="whitted"), "mix-whitted")
run(scenes.statue(n_side=90, xres=8, yres=8, spp=2), "statue-wide")
h = scenes.cornell_box(xres=40, yres=28, spp=2, materials="mixed")
gs = [GpuScene(h.desc, d, lib=E) for d in range(int(os.environ.get("PB_EMU_DEVICES", "1")))]
film = np.zeros((28, 40, 4), np.float32)
for k in range(3):
    gs[0].render_tiles_device(h.params, film.ctypes.data, k, 3)
from rs_pbrt_b200.host import render_multi
multi, sm = render_multi(gs, h.params)
print("tile-shares+multi", sm["rays"], flush=True)
for g in gs:
    g.close()
print("done")
