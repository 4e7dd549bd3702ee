#!/usr/bin/env python3
"""Lean A/B harness for experiment builds of the library (run on the GPU box).

    python tools/exp_bench.py [--scenes statue,cornell,conference] [--libs default,variants/lib_x.so,...] [--env "A=1 B=2;C=3"] [--check]

Scenes are generated ONCE (the 4.3 M-triangle statue takes ~10 s to build) and every library variant renders them through the same
C ABI: one warm-up frame, one frame with two batches in flight (the throughput number) and one single-stream frame (per-kernel
times).  --check: every variant's film must equal the first variant's bit for bit (kernel experiments may not change results).
Variants are built here in the container (tools/build_variants.py) and travel to the box as .so files.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", default="statue,cornell")
    ap.add_argument("--libs", default="default")
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--parts", type=int, default=0, help="render the frame as N tile shares one after the other on this GPU (what N ranks would each do) and report the slowest")
    args = ap.parse_args()
    import numpy as np
    import torch

    import bench
    from rs_pbrt_b200 import GpuScene, _abi

    class A:
        small = False

    # "path@ENV=VAL+ENV2=VAL2": environment knobs the library reads once, at its first render; every entry gets its own copy of the
    # .so (dlopen would hand back the same handle, and with it the same already-initialised statics, for a path it has seen)
    import shutil
    import tempfile

    tmp = Path(tempfile.mkdtemp(prefix="pbvar"))
    libs = []
    for k, spec in enumerate(args.libs.split(",")):
        name, _, envs = spec.partition("@")
        path = ROOT / "rs_pbrt_b200" / "librs_pbrt_b200.so" if name == "default" else ROOT / name
        copy = tmp / ("v%d_%s" % (k, path.name))
        shutil.copy(path, copy)
        L = C.CDLL(str(copy))
        _abi.bind(L)
        env = dict(e.split("=", 1) for e in envs.split("+") if e)
        libs.append((spec, L, env))
    rows = []
    for sname in args.scenes.split(","):
        t0 = time.perf_counter()
        h = bench.make_scene(sname)
        rp = h.params.contents
        cb = list(rp.cropped_pixel_bounds)
        film = torch.zeros((cb[3] - cb[1], cb[2] - cb[0], 4), dtype=torch.float32, device="cuda")
        print("# scene %s built in %.1f s" % (sname, time.perf_counter() - t0), flush=True)
        ref = None
        for lname, L, env in libs:
            os.environ.update(env)
            g = GpuScene(h.desc, 0, lib=L)
            rp.flags = 0
            film.zero_()
            g.render_device(h.params, film.data_ptr())
            torch.cuda.synchronize()
            best = None
            for _ in range(args.frames):
                film.zero_()
                st = g.render_device(h.params, film.data_ptr())
                best = st if best is None or st["ms_total"] < best["ms_total"] else best
            out = film.cpu().numpy().copy()
            rp.flags = _abi.RENDER_SINGLE_STREAM
            film.zero_()
            ss = g.render_device(h.params, film.data_ptr())
            rp.flags = 0
            g.close()
            for k in env:
                os.environ.pop(k, None)
            same = None
            if args.check:
                if ref is None:
                    ref = out
                else:
                    same = bool(np.array_equal(out[..., 3], ref[..., 3]) and np.allclose(out, ref, rtol=1e-6, atol=1e-6))
            row = {"scene": sname, "lib": lname, "Mrays/s": round(best["rays"] / best["ms_total"] / 1e3, 1), "ms": round(best["ms_total"], 2),
                   "single_stream_ms": round(ss["ms_total"], 2), "k_trace": round(ss["ms_trace"], 2), "k_shade": round(ss["ms_shade"], 2),
                   "other": round(ss["ms_total"] - ss["ms_trace"] - ss["ms_shade"], 2), "launches": ss["kernel_launches"], "rays": best["rays"], "same_film": same}
            if args.parts > 1:
                os.environ.update(env)
                g = GpuScene(h.desc, 0, lib=L)
                per = []
                for part in range(args.parts):
                    g.render_tiles_device(h.params, film.data_ptr(), part, args.parts)  # warm
                    b = min((g.render_tiles_device(h.params, film.data_ptr(), part, args.parts) for _ in range(2)), key=lambda s: s["ms_total"])
                    per.append(round(b["ms_total"], 2))
                g.close()
                for k in env:
                    os.environ.pop(k, None)
                row.update({"parts": args.parts, "part_ms_max": max(per), "part_ms_mean": round(sum(per) / len(per), 2), "part_ms": per,
                            "parts_Mrays/s": round(best["rays"] / max(per) / 1e3, 1)})
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(r) for r in rows) + "\n")


if __name__ == "__main__":
    main()
