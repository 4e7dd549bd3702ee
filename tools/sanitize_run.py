"""Small renders of every scene type, meant to be run under compute-sanitizer (memcheck)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rs_pbrt_b200 import GpuScene, scenes  # noqa: E402

for name, h in (("cornell", scenes.cornell_box(xres=24, yres=24, spp=4)), ("mixed", scenes.cornell_box(xres=16, yres=16, spp=4, materials="mixed")),
                ("conference", scenes.conference(xres=32, yres=18, spp=2, n_chairs=4, detail=4, n_light_quads=4)),
                ("statue", scenes.statue(n_side=48, xres=24, yres=24, spp=2)),
                ("gauss", scenes.cornell_box(xres=16, yres=16, spp=2, filter="gaussian", xwidth=2.0, ywidth=2.0))):
    g = GpuScene(h.desc, 0)
    film, st = g.render(h.params)
    rng = np.random.default_rng(0)
    o = rng.uniform(-5, 560, (2000, 3)).astype(np.float32)
    d = rng.normal(size=(2000, 3)).astype(np.float32)
    g.intersect(o, d)
    g.intersect_p(o, d)
    g.close()
    print(name, "ok", st["rays"], float(film[..., :3].mean()))
