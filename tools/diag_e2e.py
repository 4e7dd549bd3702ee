"""Diagnostic: where does bench.py's end-to-end step lose time against the bare plugin calls?  Runs the same step with / without the
NVML clock sampler thread, with / without torch holding a context, Cornell and statue."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import bench
from rs_pbrt_b200 import GpuScene, pin_description, unpin_description

def run(name, sampler_on, use_torch, resident_alive):
    h = bench.make_scene(name)
    rp = h.params.contents
    cb = list(rp.cropped_pixel_bounds)
    film = np.zeros((cb[3]-cb[1], cb[2]-cb[0], 4), np.float32)
    if use_torch:
        import torch
        t = torch.zeros((8,), device="cuda"); torch.cuda.synchronize()
    keep = GpuScene(h.desc, 0) if resident_alive else None
    pinned = pin_description(h.desc)
    s = bench.ClockSampler(0) if sampler_on else None
    if s: s.start()
    times = []
    for i in range(5):
        t0 = time.perf_counter()
        g = GpuScene(h.desc, 0); film.fill(0.0); _, st = g.render(h.params, film=film); g.close()
        times.append((time.perf_counter() - t0) * 1e3)
    if s:
        s.stop_flag = True; s.join(timeout=2)
    unpin_description(pinned)
    if keep: keep.close()
    print("%-8s sampler=%d torch=%d resident=%d: e2e step ms %s  (device %.1f)" % (name, sampler_on, use_torch, resident_alive, " ".join("%.1f" % x for x in times), st["ms_total"]), flush=True)

for name in ("cornell", "statue"):
    for cfg in ((0, 0, 0), (1, 0, 0), (0, 0, 1), (0, 1, 0), (1, 1, 1)):
        run(name, *cfg)
