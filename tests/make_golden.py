#!/usr/bin/env python3
"""Freeze oracle outputs as golden fixtures under tests/golden/ (T0/T1/T2 tiers of SURVEY.md section 8c).

The reference ships no golden vectors for this path and cannot be built here, so these vectors pin the
ORACLE (any later change to it shows up as a diff) and give the GPU box fixed inputs/outputs that do not
depend on /root/reference.  Re-run only when the oracle is deliberately changed.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle_lib as O  # noqa: E402
from rs_pbrt_b200 import scenes  # noqa: E402

G = ROOT / "tests" / "golden"
G.mkdir(parents=True, exist_ok=True)
L = O.load()

# --- Sobol' / radical inverse known answers
idx = np.array([0, 1, 2, 3, 5, 17, 255, 256, 4095, 65537, (1 << 27) + 12345, (1 << 31) + 7], np.int64)
dims = np.array([0, 1, 2, 3, 4, 5, 6, 7, 13, 44, 45, 100, 1023], np.int32)
sob = np.array([[L.orc_sobol_sample_float(int(a), int(d), 0) for d in dims] for a in idx], np.float32)
ival = np.array([[m, f, x, y, L.orc_sobol_interval_to_index(m, f, x, y)] for m in (1, 4, 9, 10, 11) for f in (0, 1, 7, 255)
                 for (x, y) in ((0, 0), (1, 0), (3, 5), ((1 << m) - 1, (1 << m) - 1))], np.uint64)
rad = np.array([[L.orc_radical_inverse(b, i) for b in range(5)] for i in range(128)], np.float32)
np.savez(G / "sobol_kat.npz", idx=idx, dims=dims, sobol=sob, interval=ival, radical=rad)

# --- Cornell box: rays, camera samples, per-sample radiance, film, light distributions
h = scenes.cornell_box(xres=32, yres=32, spp=8)
osc = O.OracleScene(h.desc)
rng = np.random.default_rng(11)
o = rng.uniform(1.0, 554.0, (4096, 3)).astype(np.float32)
d = rng.normal(size=(4096, 3))
d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
prim, t, b, st = osc.intersect(o, d)
occ, _ = osc.intersect_p(o, d * np.float32(250.0), np.full(4096, 1.0 - 1e-4, np.float32))
cam = np.array([osc.camera_sample(h.params, px, py, s) for (px, py, s) in ((0, 0, 0), (5, 7, 3), (31, 31, 7), (16, 2, 5))], np.float32)
film, samples, rst = osc.render(h.params, n_threads=4, want_samples=True)
pts = np.array([[10, 10, 10], [277, 540, 280], [500, 20, 500], [277, 277, 277]], np.float32)
ld = [osc.light_distribution(2, p, 2) for p in pts]
np.savez_compressed(G / "cornell_32x32x8.npz", o=o, d=d, prim=prim, t=t, b=b, occ=occ, nodes_visited=st["nodes_visited"], tris_tested=st["tris_tested"],
                    cam=cam, film=film, samples=samples, rays=rst["rays"], ld_pts=pts, ld_func=np.array([x[0] for x in ld]),
                    ld_cdf=np.array([x[1] for x in ld]), ld_int=np.array([x[2] for x in ld], np.float32))

# --- mixed materials (glass / metal / plastic) Cornell
h2 = scenes.cornell_box(xres=24, yres=24, spp=8, materials="mixed")
film2, samples2, rst2 = O.OracleScene(h2.desc).render(h2.params, n_threads=4, want_samples=True)
np.savez_compressed(G / "cornell_mixed_24x24x8.npz", film=film2, samples=samples2, rays=rst2["rays"])
# --- the widening of SURVEY.md section 8(f): textures + bump maps, object instances (both readings of quirk Q7), the sibling integrators
from golden_cases import widened_cases  # noqa: E402


if __name__ == "__main__":
    out = {}
    for name, hw in widened_cases():
        f, smp, st_ = O.OracleScene(hw.desc).render(hw.params, n_threads=4, want_samples=True)
        out[name + "_samples"] = smp
        out[name + "_rays"] = np.int64(st_["rays"])
    np.savez_compressed(G / "widened_16.npz", **out)
    from golden_cases import round2_cases

    out = {}
    for name, hw in round2_cases():
        f, smp, st_ = O.OracleScene(hw.desc).render(hw.params, n_threads=4, want_samples=True)
        out[name + "_samples"] = smp
        out[name + "_rays"] = np.int64(st_["rays"])
    np.savez_compressed(G / "round2_16.npz", **out)
    print("golden fixtures written to", G)
