"""GPU parity with the HaltonSampler (src/samplers/halton.rs): same kernels, other sample generator -- index of (pixel, sample)
through the base-2/3 strata, scrambled radical inverses with the PCG32-shuffled digit permutations, no power-of-two rounding of
pixelsamples."""
import numpy as np
import pytest

from rs_pbrt_b200 import GpuScene, PbrtError, scenes
from test_gpu_parity_materials import compare

pytestmark = pytest.mark.gpu


def test_halton_cornell(oracle):
    compare(scenes.cornell_box(xres=48, yres=48, spp=12, sampler="halton"), oracle)


def test_halton_mixed_materials_thin_lens(oracle):
    compare(scenes.cornell_box(xres=40, yres=40, spp=10, sampler="halton", materials="mixed", lensradius=8.0, focaldistance=900.0), oracle)


def test_halton_tile_repeat_beyond_128_pixels(oracle):
    """K_MAX_RESOLUTION = 128: pixels 128 apart share their Halton indices; base-3 stratum of 243 rows."""
    compare(scenes.cornell_box(xres=160, yres=140, spp=2, sampler="halton", maxdepth=3), oracle)


def test_halton_all_light_kinds(oracle):
    compare(scenes.sky_scene(xres=40, yres=40, spp=9, env="image", strategy="spatial", sampler="halton"), oracle)
    compare(scenes.cornell_box(xres=40, yres=40, spp=7, sampler="halton", lights="delta", strategy="power"), oracle)


def test_halton_sample_pixel_center(oracle):
    h = scenes.cornell_box(xres=32, yres=32, spp=4, sampler="halton", samplepixelcenter=True)
    compare(h, oracle)


def test_halton_index_range_is_checked():
    h = scenes.cornell_box(xres=200, yres=300, spp=1 << 18, sampler="halton")  # 2^18 * 128 * 243 >= 2^32
    g = GpuScene(h.desc, 0)
    with pytest.raises(PbrtError):
        g.render(h.params, rect=[0, 0, 2, 2])
    g.close()
