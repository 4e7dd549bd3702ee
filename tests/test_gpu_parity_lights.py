"""GPU parity for the delta lights (PointLight, SpotLight, DistantLight -- src/lights/{point,spot,distant}.rs): the is_delta_light
branch of estimate_direct (src/core/integrator.rs:470-480), power() of every kind in the power distribution and sample_li of every
kind in the spatial distribution, against the oracle on the same inputs."""
import numpy as np
import pytest

from rs_pbrt_b200 import GpuScene, scenes
from test_gpu_parity_materials import compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("strategy", ["uniform", "power", "spatial"])
def test_mixed_area_and_delta_lights(oracle, strategy):
    h = scenes.cornell_box(xres=48, yres=48, spp=16, lights="delta", strategy=strategy)
    compare(h, oracle)


@pytest.mark.parametrize("kind", ["point", "spot", "distant"])
def test_single_delta_light_no_emitters(oracle, kind):
    """One light => UniformLightDistribution whatever the strategy (integrator.rs create_light_sample_distribution); no BSDF-sampling
    strategy, no MIS rays: shadow rays only."""
    h = scenes.cornell_box(xres=40, yres=40, spp=8, lights=kind)
    g = GpuScene(h.desc, 0)
    _, st = g.render_samples(h.params, list(h.params.contents.sample_bounds))
    g.close()
    assert st["light_tri_tests"] == 0
    compare(h, oracle)


def test_delta_lights_with_specular_materials(oracle):
    h = scenes.cornell_box(xres=40, yres=40, spp=16, lights="delta", materials="mixed")
    compare(h, oracle)


# ---- InfiniteAreaLight (src/lights/infinite.rs) --------------------------------------------------------------------------------
@pytest.mark.parametrize("env,strategy", [("constant", "spatial"), ("image", "uniform"), ("image", "power"), ("image", "spatial"), ("two", "spatial")])
def test_infinite_light_scene(oracle, env, strategy):
    """Environment emission of escaping camera / specular rays (path.rs:267-275), sample_li through the Distribution2D, pdf_li,
    Le of escaping MIS rays (integrator.rs:560-562), power() via the MIP pyramid, all light kinds in one distribution."""
    h = scenes.sky_scene(xres=48, yres=48, spp=16, env=env, strategy=strategy)
    compare(h, oracle)


def test_infinite_light_only_empty_scene(oracle):
    """No geometry at all: every camera ray escapes."""
    from rs_pbrt_b200 import HostScene
    h = HostScene()
    h.light_infinite([1.0, 1.0, 1.0], texels=scenes.sky_map(32, 16), light_to_world=scenes.Y_UP)
    h.look_at([0, 0, 0], [0, 0.3, 1], [0, 1, 0])
    h.film(32, 32)
    h.camera(fov=70.0)
    h.sampler(4)
    h.integrator(maxdepth=3)
    h.world_end()
    gs, os_ = compare(h, oracle)
    assert gs.mean() > 0.05


def test_infinite_light_non_power_of_two_map(oracle):
    """MipMap::new resamples such a map to the next power of two (4-tap Lanczos, mipmap.rs:60-150) before anything else is
    derived from it; the library restates that on the host."""
    from rs_pbrt_b200 import HostScene, _abi
    rng = np.random.default_rng(2)
    tex = (rng.random((9, 20, 3)) ** 2 * 3).astype(np.float32)
    tex[1, 3] += 25.0
    h = HostScene()
    m = h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0])
    mir = h.material(_abi.MAT_MIRROR, [0.9, 0.9, 0.9])
    h.light_infinite([1, 1, 1], texels=tex, light_to_world=scenes.Y_UP)
    h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), np.array([[-4, 0, -4], [-4, 0, 4], [4, 0, 4], [4, 0, -4]], np.float32), material=m)
    h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), np.array([[-1, 0.5, -1], [-1, 1.5, 1], [1, 1.5, 1], [1, 0.5, -1]], np.float32), material=mir)
    h.look_at([0, 3, -6], [0, 0.5, 0], [0, 1, 0])
    h.film(40, 32)
    h.camera(fov=45.0)
    h.sampler(8)
    h.integrator(maxdepth=4, lightsamplestrategy="power")
    h.world_end()
    assert list(h.desc.contents.lights[0].env_res) == [20, 9]
    compare(h, oracle)
