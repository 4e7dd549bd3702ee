#!/usr/bin/env python3
"""Render an exported scene with a real rs_pbrt build, where one exists, and hold it against this repository's render of the same scene.

    python tools/compare_with_rs_pbrt.py [--scene cornell|cornell-mixed|statue-small|cornell-translucent|cornell-mix|cornell-textured|landscape-small] [--rs-pbrt /path/to/rs_pbrt] [--device 0 | --oracle]

This image has no Rust toolchain (no cargo / rustc, no crates offline), so here the script stops after writing the .pbrt files and says so.
On a machine with `rs_pbrt` on PATH (cargo build --release in the reference repository) it runs
    rs_pbrt --integrator path -t <threads> --path <scene>.pbrt        (src/bin/rs_pbrt.rs:41-68; writes pbrt.png, film.rs:437-528)
parses the tiles-per-second progress line for a second CPU-baseline row, and compares the 8-bit image with
Film::write_image of our render (pbrt_host_write_image produces the same bytes from the same float film).  That comparison is what
would turn "parity unpinned" into a pinned oracle.
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell")
    ap.add_argument("--rs-pbrt", default=shutil.which("rs_pbrt") or "")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "pbrt_export"))
    ap.add_argument("--oracle", action="store_true", help="render our side with the CPU oracle instead of the GPU library")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    import numpy as np

    from rs_pbrt_b200 import pbrt_export, scenes

    makers = {
        "cornell": lambda: scenes.cornell_box(xres=400, yres=400, spp=64),  # BASELINE.json configs[0]
        "cornell-mixed": lambda: scenes.cornell_box(xres=200, yres=200, spp=32, materials="mixed", lights="delta"),
        "statue-small": lambda: scenes.statue(n_side=200, xres=256, yres=256, spp=16),
        # the widened rows: translucent / mix materials, image textures + bump maps + alpha masks (8-bit texels: the round trip is exact), instances
        "cornell-translucent": lambda: scenes.cornell_box(xres=200, yres=200, spp=32, materials="translucent"),
        "cornell-mix": lambda: scenes.cornell_box(xres=200, yres=200, spp=32, materials="mix"),
        "cornell-textured": lambda: scenes.cornell_box(xres=200, yres=200, spp=32, textures="ewa+float+graph+bump", alpha="masks", quantize_textures=True),
        "landscape-small": lambda: scenes.landscape(xres=320, yres=180, spp=16, n_trees=200, grid=48, detail=8, instancing="reference", n_prototypes=5, sky="constant"),
    }
    h = makers[args.scene]()
    out = Path(args.out)
    notes = pbrt_export.write(h, out / (args.scene + ".pbrt"))
    print("wrote", out / (args.scene + ".pbrt"), "notes:", notes)
    if not args.rs_pbrt:
        print("no rs_pbrt binary (this image has no Rust toolchain): nothing to compare against; the files above are the input for one")
        return 0
    t0 = time.perf_counter()
    r = subprocess.run([args.rs_pbrt, "--integrator", "path", "-t", str(os.cpu_count() or 1), "--path", str(out / (args.scene + ".pbrt"))], cwd=out, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    m = re.findall(r"([\d.]+)/s", r.stdout + r.stderr)
    print("rs_pbrt: rc %d, %.1f s wall, tiles/s %s" % (r.returncode, dt, m[-1] if m else "?"))
    if args.oracle:
        import oracle_lib

        film, _, _ = oracle_lib.OracleScene(h.desc).render(h.params, n_threads=os.cpu_count() or 1)
        h.film_clear(); h.film_add(film)
    else:
        h.render(device=args.device)
    h.write_image(out / (args.scene + "_ours.ppm"))
    ours = np.frombuffer((out / (args.scene + "_ours.ppm")).read_bytes().split(b"\n", 3)[3], np.uint8)
    import pbrt_reader

    theirs = pbrt_reader.read_png(out / "pbrt.png").reshape(-1)
    diff = np.abs(ours.astype(np.int32) - theirs.astype(np.int32))
    print("8-bit image: max |diff| %d, mean %.4f, pixels differing %.4f %%" % (diff.max(), diff.mean(), 100.0 * (diff > 0).mean()))
    return 0 if diff.max() <= 1 else 1


if __name__ == "__main__":
    sys.exit(main())
