"""AOIntegrator (src/integrators/ao.rs) in the oracle: closed forms.  The integrator adds
dot(wi, n) / (pdf * nsamples) for every unoccluded direction, so an unoccluded point gets pi (there is no 1/pi in it)."""
import math

import numpy as np
import pytest

from rs_pbrt_b200 import HostScene, _abi


def scene(build, nsamples=16, cossample=True, spp=4, res=12, sampler="sobol", eye=(0, 5, 0), look=(0, 0, 0), up=(0, 0, 1), fov=40.0):
    h = HostScene()
    m = h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0])
    build(h, m)
    h.look_at(list(eye), list(look), list(up))
    h.film(res, res)
    h.camera(fov=fov)
    h.sampler(spp, name=sampler)
    h.integrator_ao(nsamples=nsamples, cossample=cossample)
    h.world_end()
    rp = h.params.contents
    assert rp.integrator == _abi.INTEGRATOR_AO and rp.ao_samples == nsamples and rp.ao_cos_sample == int(cossample)
    return h


def quad(h, m, p, **kw):
    h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), np.array(p, np.float32), material=m, **kw)


def floor(h, m):
    quad(h, m, [[-50, 0, -50], [-50, 0, 50], [50, 0, 50], [50, 0, -50]])


@pytest.mark.parametrize("sampler", ["sobol", "halton"])
def test_open_plane_is_pi(oracle, sampler):
    h = scene(floor, sampler=sampler)
    _, samples, st = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=2)
    assert np.allclose(samples, math.pi, rtol=2e-6)  # cosine sampling: every direction weighs exactly pi / nsamples
    assert st["shadow_rays"] == 16 * st["camera_rays"] and st["closest_rays"] == st["camera_rays"]


def test_uniform_sampling_is_unbiased(oracle):
    h = scene(floor, nsamples=64, cossample=False, spp=16)
    _, samples, _ = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=4)
    assert samples.mean() == pytest.approx(math.pi, rel=0.01)
    assert samples.std() > 0  # 2 pi cos(theta) / n per direction: not constant


def test_closed_box_is_black_and_miss_is_black(oracle):
    def box(h, m):
        floor(h, m)
        quad(h, m, [[-50, 1, -50], [50, 1, -50], [50, 1, 50], [-50, 1, 50]])  # a lid one unit above the floor
        for a in (-2.0, 2.0):  # four walls
            quad(h, m, [[a, 0, -2], [a, 1, -2], [a, 1, 2], [a, 0, 2]])
            quad(h, m, [[-2, 0, a], [-2, 1, a], [2, 1, a], [2, 0, a]])
    h = scene(box, eye=(0, 0.5, 0), look=(0, 0, 0), up=(0, 0, 1))
    _, samples, st = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=2)
    assert not samples.any()
    h2 = scene(floor, eye=(0, 5, 0), look=(0, 10, 0))  # looking away: no hit, no rays but the camera rays
    _, s2, st2 = oracle.OracleScene(h2.desc).render(h2.params, want_samples=True, n_threads=2)
    assert not s2.any() and st2["shadow_rays"] == 0


def test_half_covered_sky_is_half_pi(oracle):
    """A ceiling over x > 0 only: a floor point under its edge sees half of the cosine-weighted hemisphere."""
    def build(h, m):
        floor(h, m)
        quad(h, m, [[0, 1, -500], [500, 1, -500], [500, 1, 500], [0, 1, 500]])
    # the camera sits below the ceiling's height, looking straight down at the edge line x = 0
    h = scene(build, nsamples=256, spp=16, res=8, eye=(0, 0.5, 0), look=(0, 0, 0), up=(0, 0, 1), fov=2.0)
    _, samples, _ = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=4)
    assert samples.mean() == pytest.approx(math.pi / 2, rel=0.03)


def test_one_direction_per_sample_is_all_or_nothing(oracle):
    """nsamples = 1 under the half ceiling: each camera sample is either fully open (pi) or fully blocked (0), about half each."""
    def build(h, m):
        floor(h, m)
        quad(h, m, [[0, 1, -500], [500, 1, -500], [500, 1, 500], [0, 1, 500]])
    h = scene(build, nsamples=1, spp=8, res=4, eye=(0, 0.5, 0), look=(0, 0, 0), up=(0, 0, 1), fov=2.0)
    _, samples, st = oracle.OracleScene(h.desc).render(h.params, want_samples=True, n_threads=1)
    v = samples[..., 0].ravel()
    assert np.all(np.isclose(v, 0.0, atol=1e-6) | np.isclose(v, math.pi, rtol=1e-5))
    assert 0.2 < (v > 1).mean() < 0.8
    assert st["shadow_rays"] == st["camera_rays"]
