"""HaltonSampler (src/samplers/halton.rs, the crate's default sampler) in the oracle and the host
mirror.  The reference holds no vectors for it, so the restatement is pinned by the properties that DEFINE the sampler: the
global index of (pixel, sample) must put the unscrambled base-2/base-3 radical inverses into that pixel's stratum (this checks
the extended-gcd inverses, the CRT offset and inverse_radical_inverse against plain radical inverses), every digit permutation is
a permutation, scrambling with the identity is the plain radical inverse, and the image converges to the Sobol' image."""
import ctypes as C

import numpy as np
import pytest

from rs_pbrt_b200 import _abi, scenes


def radical_inverse(base, a):
    """Exact (rational) radical inverse."""
    from fractions import Fraction
    r, f = Fraction(0), Fraction(1, base)
    while a:
        r += (a % base) * f
        a //= base
        f /= base
    return r


def params(xres, yres, spp, **kw):
    h = scenes.cornell_box(xres=xres, yres=yres, spp=spp, sampler="halton", **kw)
    return h


@pytest.mark.parametrize("res", [(24, 17), (150, 130), (128, 243), (300, 300)])
def test_halton_index_lands_in_the_pixel_stratum(oracle, res):
    h = params(res[0], res[1], 4)
    L = oracle.load()
    rp = h.params.contents
    assert rp.sampler == _abi.SAMPLER_HALTON and rp.spp == 4
    sb = list(rp.sample_bounds)
    scale = [1, 1]
    for i, base in enumerate((2, 3)):
        while scale[i] < min(sb[2 + i] - sb[i], 128):
            scale[i] *= base
    stride = scale[0] * scale[1]
    rng = np.random.default_rng(1)
    seen = set()
    for _ in range(200):
        px, py, s = int(rng.integers(sb[0], sb[2])), int(rng.integers(sb[1], sb[3])), int(rng.integers(0, 4))
        idx = C.c_uint64()
        u0 = L.orc_sampler_dimension(h.params, px, py, s, 0, C.byref(idx))
        u1 = L.orc_sampler_dimension(h.params, px, py, s, 1, None)
        i = idx.value
        assert i // stride == s  # sample s of a pixel is the s-th visit of the stratum grid
        # the unscrambled Halton point of index i, scaled to the stratum grid, falls into pixel (px, py) mod 128
        hx, hy = radical_inverse(2, i) * scale[0], radical_inverse(3, i) * scale[1]
        assert int(hx) == px % 128 % scale[0] and int(hy) == py % 128 % scale[1]
        # dims 0 / 1 are the offsets INSIDE the pixel: the remaining digits of the same radical inverses
        assert u0 == pytest.approx(float(hx - int(hx)), abs=2e-6) and u1 == pytest.approx(float(hy - int(hy)), abs=2e-6)
        assert 0.0 <= u0 < 1.0 and 0.0 <= u1 < 1.0
        seen.add((px % 128, py % 128, s, i))
    # the pattern repeats every K_MAX_RESOLUTION = 128 pixels; within a tile distinct (pixel, sample) have distinct indices
    assert len({v[3] for v in seen}) == len(seen)


def test_digit_permutations_and_scrambled_radical_inverse(oracle):
    L = oracle.load()
    buf = (C.c_uint16 * 8192)()
    primes = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29]
    for dim, p in enumerate(primes):
        assert L.orc_halton_permutation(dim, buf, 8192) == p
        perm = list(buf[:p])
        assert sorted(perm) == list(range(p))
    assert L.orc_halton_permutation(999, buf, 8192) == 7919  # PRIME_TABLE_SIZE = 1000, last prime 7919
    assert sorted(buf[:7919]) == list(range(7919))
    # identity permutation -> the plain radical inverse (the perm[0] tail term vanishes)
    for dim, p in list(enumerate(primes))[2:]:
        ident = (C.c_uint16 * p)(*range(p))
        for a in (0, 1, 5, 12345, 987654321):
            assert L.orc_scrambled_radical_inverse(dim, a, ident) == pytest.approx(float(radical_inverse(p, a)), abs=1e-6)
    # a real permutation: digits are mapped through it, and the infinite tail of zero digits contributes perm[0]/(p-1)
    L.orc_halton_permutation(2, buf, 8192)
    perm = list(buf[:5])
    a, x, f = 1234567, 0.0, 1.0 / 5
    while a:
        x += perm[a % 5] * f
        a //= 5
        f /= 5
    x += f * 5 * perm[0] / 4.0  # sum_{k>=K} perm[0] 5^-k
    pb = (C.c_uint16 * 5)(*perm)
    assert L.orc_scrambled_radical_inverse(2, 1234567, pb) == pytest.approx(x, rel=1e-6)


def test_halton_dimensions_are_uniform(oracle):
    L = oracle.load()
    h = params(32, 32, 64)
    for dim in (2, 3, 4, 5, 12, 40):
        v = np.array([L.orc_sampler_dimension(h.params, x, y, s, dim, None) for x in range(0, 32, 5) for y in range(0, 32, 5) for s in range(0, 64, 3)])
        assert v.min() >= 0.0 and v.max() < 1.0
        assert abs(v.mean() - 0.5) < 0.03 and abs(np.mean(v < 0.25) - 0.25) < 0.04


def test_halton_and_sobol_renders_agree(oracle):
    """Two unbiased estimators of the same image."""
    hs = scenes.cornell_box(xres=24, yres=24, spp=64)
    hh = scenes.cornell_box(xres=24, yres=24, spp=50, sampler="halton")  # any spp: no power-of-two rounding
    assert hh.params.contents.spp == 50
    fs, _, _ = oracle.OracleScene(hs.desc).render(hs.params, n_threads=8)
    fh, _, st = oracle.OracleScene(hh.desc).render(hh.params, n_threads=8)
    ims, imh = fs[..., :3] / fs[..., 3:], fh[..., :3] / fh[..., 3:]
    assert st["camera_rays"] == 24 * 24 * 50
    assert np.allclose(ims.mean((0, 1)), imh.mean((0, 1)), rtol=0.03)


def test_sample_pixel_center(oracle):
    L = oracle.load()
    h = params(16, 16, 4, samplepixelcenter=True)
    assert h.params.contents.sample_at_pixel_center == 1
    assert L.orc_sampler_dimension(h.params, 3, 4, 1, 0, None) == 0.5 and L.orc_sampler_dimension(h.params, 3, 4, 1, 1, None) == 0.5
    assert L.orc_sampler_dimension(h.params, 3, 4, 1, 2, None) != 0.5
