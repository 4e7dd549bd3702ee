import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rs_pbrt_b200 import GpuScene, scenes
import os
for name, h in (("cornell", scenes.cornell_box(xres=64, yres=64, spp=4)), ("statue", scenes.statue(n_side=1468, xres=64, yres=64, spp=4, n_threads=os.cpu_count()))):
    for i in range(3):
        t = time.perf_counter(); g = GpuScene(h.desc, 0); dt = time.perf_counter() - t
        t = time.perf_counter(); g.close(); dc = time.perf_counter() - t
        print(name, "create %.1f ms destroy %.1f ms, %.1f MB" % (dt * 1e3, dc * 1e3, g.L.pbrt_gpu_scene_bytes(g.handle) / 1e6 if g.handle else 0))
