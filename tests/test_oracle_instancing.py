"""Object instancing (TransformedPrimitive, src/core/primitive.rs:198-272, ObjectBegin/End/Instance
src/core/api.rs:3001-3109) in the oracle and the host mirror, in both reporting modes (quirk Q7):

  FIXED      pbrt-v3's behaviour: an instance hit keeps its primitive.  Anchor: rendering an instanced object must equal rendering the
             same triangles placed in the world by hand (up to the rounding of the extra transform).
  REFERENCE  TransformedPrimitive::intersect as rs_pbrt has it: a non-identity instance loses its primitive (no material, no emission:
             the path walks through it) yet still blocks shadow rays; an identity instance overwrites the interaction and shortens the
             ray but reports no hit, so whatever lies behind it is culled."""
import numpy as np
import pytest

from rs_pbrt_b200 import HostScene, _abi


def quad(h, m, p, **kw):
    h.trianglemesh(np.array([0, 1, 2, 0, 2, 3], np.uint32), np.array(p, np.float32), material=m, **kw)


def box_mesh(lo, hi, M=None):
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    P = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], np.float64)
    if M is not None:
        P = (np.asarray(M, np.float64)[:3, :3] @ P.T).T + np.asarray(M, np.float64)[:3, 3]
    idx = np.array([0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7, 0, 1, 5, 0, 5, 4, 2, 3, 7, 2, 7, 6, 1, 2, 6, 1, 6, 5, 0, 4, 7, 0, 7, 3], np.uint32)
    return idx, P.astype(np.float32)


def scene(mode, transforms, instanced=True, wall=False, res=(24, 18), spp=16, sky=None):
    h = HostScene()
    grey = h.material(_abi.MAT_MATTE, [0.6, 0.6, 0.6, 0.0])
    red = h.material(_abi.MAT_MATTE, [0.7, 0.2, 0.2, 0.0])
    quad(h, grey, [[-6, 0, -6], [-6, 0, 6], [6, 0, 6], [6, 0, -6]])
    quad(h, grey, [[-1, 4, -1], [1, 4, -1], [1, 4, 1], [-1, 4, 1]], emit=[8, 8, 8])
    if wall:
        quad(h, grey, [[-6, 0, 3], [-6, 5, 3], [6, 5, 3], [6, 0, 3]])
    if sky is not None:
        h.light_infinite(sky)
    lo, hi = (-0.5, 0.0, -0.5), (0.5, 1.0, 0.5)
    if instanced:
        o = h.object_begin()
        h.trianglemesh(*box_mesh(lo, hi), material=red)
        h.object_end()
        for M in transforms:
            h.object_instance(o, M)
        h.instancing(mode)
    else:
        for M in transforms:
            h.trianglemesh(*box_mesh(lo, hi, M), material=red)
    h.look_at([0, 3, -7], [0, 0.5, 0], [0, 1, 0])
    h.film(*res)
    h.camera(fov=45.0)
    h.sampler(spp)
    h.integrator(maxdepth=3, lightsamplestrategy="uniform")
    h.world_end()
    return h


def translate(x, y, z):
    M = np.eye(4, dtype=np.float32)
    M[:3, 3] = [x, y, z]
    return M


def rot_scale(angle_deg, s, t):
    a = np.radians(angle_deg)
    M = np.eye(4, dtype=np.float64)
    M[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ np.diag(s)
    M[:3, 3] = t
    return M.astype(np.float32)


def image(oracle, h):
    f, _, st = oracle.OracleScene(h.desc).render(h.params, n_threads=8)
    return f[..., :3] / np.maximum(f[..., 3:], 1e-9), st


def test_host_mirror_layout(product_lib):
    h = scene("fixed", [translate(1.5, 0, 0.5), rot_scale(30, [1.5, 0.7, 1.0], [-2, 0, 1])])
    d = h.desc.contents
    assert d.n_instances == 2 and h.params.contents.instancing == _abi.INSTANCING_FIXED
    top = [d.tris[i] for i in range(d.n_tris)]
    markers = [t for t in top if t.mesh == _abi.MESH_INSTANCE]
    assert sorted(t.v[0] for t in markers) == [0, 1] and all(t.material == _abi.PBRT_NO_MATERIAL for t in markers)
    I = [d.instances[i] for i in range(2)]
    assert I[0].root == I[1].root and 0 < I[0].root < d.n_nodes  # one BVHAccel per object, shared by its instances
    assert not I[0].identity and np.allclose(np.array(list(I[0].m)).reshape(4, 4) @ np.array(list(I[0].m_inv)).reshape(4, 4), np.eye(4), atol=1e-6)
    # the object's tree: absolute offsets, leaves point at the 12 box triangles stored after the top-level primitives
    root = d.nodes[I[0].root]
    assert list(root.pmin) == [-0.5, 0.0, -0.5] and list(root.pmax) == [0.5, 1.0, 0.5]
    leaves = [d.nodes[i] for i in range(I[0].root, d.n_nodes) if d.nodes[i].n_prims > 0]
    covered = sorted(n.offset + k for n in leaves for k in range(n.n_prims))
    assert covered == list(range(d.n_tris - 12, d.n_tris)) and all(d.tris[i].mesh != _abi.MESH_INSTANCE for i in covered)
    # the top-level tree bounds the transformed boxes
    assert d.nodes[0].pmax[0] >= 2.0 and d.nodes[0].pmin[0] <= -2.5


@pytest.mark.parametrize("transforms", [[translate(1.5, 0, 0.5)], [rot_scale(30, [1.5, 0.7, 1.0], [-2, 0, 1]), translate(1.5, 0, 0.5), rot_scale(-50, [0.5, 2.0, 0.5], [0, 0, 2])]])
def test_fixed_mode_equals_hand_placed_geometry(oracle, transforms):
    a, sa = image(oracle, scene("fixed", transforms, instanced=True))
    b, sb = image(oracle, scene("fixed", transforms, instanced=False))
    assert (sa["camera_rays"], sa["closest_rays"]) == (sb["camera_rays"], sb["closest_rays"]) or abs(sa["rays"] - sb["rays"]) < 0.01 * sb["rays"]
    assert np.abs(a - b).mean() < 2e-3 * b.mean() + 1e-6  # a few samples may fall on the other side of an edge after the extra rounding
    assert (a[..., 0] > 1.5 * a[..., 1]).sum() > 5        # the red boxes are seen


def test_reference_mode_walks_through_instances_but_shadows_stay(oracle):
    T = [translate(1.5, 0, 0.5)]
    ref, sr = image(oracle, scene("reference", T, spp=64))
    fix, sf = image(oracle, scene("fixed", T, spp=64))
    # build the same scene without any box
    empty, _ = image(oracle, scene("fixed", [], instanced=False, spp=64))
    assert np.allclose(ref[..., 0], ref[..., 1], rtol=1e-5, atol=1e-6)          # the red box is never shaded ...
    assert (fix[..., 0] > 1.5 * fix[..., 1]).sum() > 5                          # ... unlike in the fixed mode
    assert sr["closest_rays"] > sf["closest_rays"]                              # pass-through rays
    darker = (ref.mean(-1) < 0.8 * empty.mean(-1)) & (empty.mean(-1) > 0.02)    # its shadow is still cast
    assert darker.sum() > 3


def test_reference_mode_identity_instance_culls_what_is_behind_it(oracle):
    """Identity transform: the inner hit shortens the ray and is then reported as a miss, so the wall behind the box is never
    found -- the camera ray leaves the scene (here: into a constant sky)."""
    sky = np.array([0.25, 0.5, 1.0], np.float32)
    I = [np.eye(4, dtype=np.float32)]
    ref, _ = image(oracle, scene("reference", I, wall=True, sky=sky, spp=4, res=(48, 36)))
    fix, _ = image(oracle, scene("fixed", I, wall=True, sky=sky, spp=4, res=(48, 36)))
    # pixels whose centre ray hits the box (fixed mode shows it red)
    m = fix[..., 0] > 1.3 * fix[..., 1]
    box_px = m.copy()  # interior pixels only: every sample of the pixel and of its neighbours sees the box
    box_px[1:, :] &= m[:-1, :]; box_px[:-1, :] &= m[1:, :]; box_px[:, 1:] &= m[:, :-1]; box_px[:, :-1] &= m[:, 1:]
    box_px[0, :] = box_px[-1, :] = False; box_px[:, 0] = box_px[:, -1] = False
    assert box_px.sum() >= 2
    assert np.allclose(ref[box_px], sky[None, :], rtol=1e-4)                    # pure sky where the box stands in front of the wall
    assert not np.allclose(fix[box_px], sky[None, :], rtol=0.2)
