"""Diagnostic: 40 end-to-end steps of the statue; scene_create phase stamps are printed only for calls slower than 100 ms (PB_TIMING=2)."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["PB_TIMING"] = "2"
import numpy as np
import bench
from rs_pbrt_b200 import GpuScene, pin_description, unpin_description
h = bench.make_scene("statue")
film = np.zeros((1024, 1024, 4), np.float32)
pinned = pin_description(h.desc)
ts = []
for i in range(40):
    t0 = time.perf_counter(); g = GpuScene(h.desc, 0); t1 = time.perf_counter(); film.fill(0.0); _, st = g.render(h.params, film=film); t2 = time.perf_counter(); g.close(); t3 = time.perf_counter()
    ts.append((t3 - t0) * 1e3)
    if ts[-1] > 260: print("SLOW step %d: create %.1f render call %.1f (device %.1f) destroy %.1f" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st["ms_total"], (t3 - t2) * 1e3), flush=True)
unpin_description(pinned)
ts.sort()
print("e2e step ms: median %.1f, p90 %.1f, max %.1f, mean %.1f" % (ts[20], ts[36], ts[-1], sum(ts) / len(ts)))
