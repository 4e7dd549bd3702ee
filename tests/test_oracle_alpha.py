"""Alpha / shadow-alpha masks of triangle meshes (triangle.rs:313-330, 593-654) in the oracle, pinned by closed forms:
an alpha of exactly 0 removes the surface for every ray; a shadow-alpha of 0 removes it for shadow rays only (intersect_p tests
both masks, intersect the alpha mask alone); any non-zero value is opaque; the mask is looked up at the hit's uv."""
import numpy as np

import oracle_lib
from rs_pbrt_b200 import HostScene, _abi


def card_scene(alpha=None, shadow_alpha=None, res=16, spp=4, half=False):
    """A matte floor under a constant sky, a black card hovering above it and facing the camera from above."""
    h = HostScene()
    floor = h.material(_abi.MAT_MATTE, [0.5, 0.5, 0.5, 0.0])
    black = h.material(_abi.MAT_MATTE, [0.0, 0.0, 0.0, 0.0])
    h.light_infinite([1.0, 1.0, 1.0])
    P = np.array([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], np.float32)
    h.trianglemesh([0, 1, 2, 0, 2, 3], P, material=floor)
    C_ = np.array([[-2, 1, -2], [2, 1, -2], [2, 1, 2], [-2, 1, 2]], np.float32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    m = h.trianglemesh([0, 1, 2, 0, 2, 3], C_, UV=uv, material=black)
    tex = {}
    for name, v in (("alpha", alpha), ("shadow_alpha", shadow_alpha)):
        if v is None:
            continue
        if np.isscalar(v):
            tex[name] = h.texture_constant([float(v)], float_valued=True)
        else:
            img = np.repeat(np.asarray(v, np.float32)[..., None], 3, axis=2)
            tex[name] = h.texture_image(img, float_valued=True, wrap=_abi.WRAP_CLAMP)
    if tex:
        h.mesh_alpha(m, alpha=tex.get("alpha"), shadow_alpha=tex.get("shadow_alpha"))
    h.look_at([0, 6, 0.001], [0, 0, 0], [0, 0, 1])
    h.film(res, res)
    h.camera(fov=60.0)
    h.sampler(spp)
    h.integrator(maxdepth=1, lightsamplestrategy="uniform")
    h.world_end(n_threads=2)
    return h


def centre(h):
    film, _, st = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=2)
    img = film[..., :3] / film[..., 3:]
    n = img.shape[0]
    return float(img[n // 2 - 1:n // 2 + 1, n // 2 - 1:n // 2 + 1].mean()), st


def test_alpha_zero_removes_the_surface_and_non_zero_is_opaque():
    seen_card, st_card = centre(card_scene())                      # the black card fills the centre: radiance 0
    assert seen_card == 0.0
    open_floor, st_open = centre(card_scene(alpha=0.0))            # alpha 0: as if the card were not there
    assert 0.3 < open_floor < 0.55                                 # Kd * E[visible sky] = 0.5 * 1 minus noise; the card casts no shadow either
    faint, _ = centre(card_scene(alpha=0.25))                      # any non-zero alpha is fully opaque (== 0.0 test, triangle.rs:327)
    assert faint == 0.0
    # without the card at all the rays of an alpha-0 card scene are the same rays
    h = card_scene(alpha=0.0)
    film0, _, st0 = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=2)
    assert st0["rays"] == st_open["rays"]


def test_shadow_alpha_only_affects_shadow_rays():
    # Triangle::intersect ignores shadow_alpha_mask: the camera still sees the black card ...
    seen, _ = centre(card_scene(shadow_alpha=0.0))
    assert seen == 0.0
    # ... but intersect_p honours it: the floor under the card is lit as if nothing were above it.  Look at the floor beside the card,
    # where the card (2x2 at height 1) blocks a good part of the sky: with shadow_alpha = 0 the NEE shadow rays pass through it.
    def beside(h):
        film, _, _ = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=2)
        img = film[..., 0] / film[..., 3]
        return float(img[:, :2].mean())  # left edge columns: floor next to the card
    blocked = beside(card_scene(res=16, spp=64))
    passing = beside(card_scene(shadow_alpha=0.0, res=16, spp=64))
    nothing = beside(card_scene(alpha=0.0, res=16, spp=64))
    assert passing > blocked
    # with maxdepth 1 the only light transport to the floor is NEE + MIS: MIS rays are closest-hit rays (still blocked), so the result
    # lies between the fully blocked and the fully open scene
    assert blocked < passing <= nothing + 1e-6


def test_image_mask_is_looked_up_at_the_hits_uv():
    mask = np.zeros((2, 8), np.float32)
    mask[:, 4:] = 1.0  # bilinear lookups: exactly 0 (cut away) for u < 3.5 / 8, non-zero (opaque) beyond; both rows alike
    h = card_scene(alpha=mask, res=32, spp=4)
    film, _, _ = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=2)
    img = film[..., 0] / film[..., 3]
    # the card spans the centre of the image; one half of it must be black (opaque), the other show the lit floor
    row = img[16, 9:23]
    assert (row == 0.0).sum() >= 5 and (row > 0.1).sum() >= 4
    assert (row[:4] == 0.0).all() != (row[-4:] == 0.0).all()  # one side of the card is there, the other is not
