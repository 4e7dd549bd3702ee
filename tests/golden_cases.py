"""The scenes behind tests/golden/widened_16.npz and round2_16.npz (made by tests/make_golden.py): one list per file, for the generator and for the tests that read it."""
from rs_pbrt_b200 import scenes


def widened_cases():
    yield "textured", scenes.cornell_box(xres=16, yres=16, spp=4, textures="ewa+float+graph+bump")
    yield "landscape_fixed", scenes.landscape(xres=24, yres=14, spp=4, n_trees=40, grid=12, detail=6, instancing="fixed")
    yield "landscape_reference", scenes.landscape(xres=24, yres=14, spp=4, n_trees=40, grid=12, detail=6, instancing="reference")
    yield "direct_all", scenes.cornell_box(xres=16, yres=16, spp=4, integrator=("direct", "all"), materials="mixed", lights="delta", lightsamples=2)
    yield "whitted_textured", scenes.cornell_box(xres=16, yres=16, spp=4, integrator="whitted", textures="trilinear+bump")
    yield "ao", scenes.cornell_box(xres=16, yres=16, spp=4, integrator=("ao", 8, True))


def round2_cases():
    """Round 2's additions: alpha / shadow-alpha masks, TranslucentMaterial, MixMaterial (path and whitted), Halton over a mix."""
    yield "alpha_masks", scenes.cornell_box(xres=16, yres=16, spp=4, alpha="masks", materials="mixed", lights="delta")
    yield "translucent", scenes.cornell_box(xres=16, yres=16, spp=4, materials="translucent")
    yield "mix", scenes.cornell_box(xres=16, yres=16, spp=4, materials="mix")
    yield "mix_whitted", scenes.cornell_box(xres=16, yres=16, spp=4, materials="mix", integrator="whitted")
    yield "mix_halton_direct", scenes.cornell_box(xres=16, yres=16, spp=3, materials="mix", sampler="halton", integrator=("direct", "all"), lightsamples=2, lights="delta")
