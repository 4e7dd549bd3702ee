#!/usr/bin/env python3
"""Build experiment variants of librs_pbrt_b200.so in parallel (here in the container; the .so files travel to the GPU box).

    python tools/build_variants.py name1:-DPB_X=1,-DPB_Y=2 name2:-DPB_Z=3 ...

Each variant is variants/lib_<name>.so: the same sources with extra nvcc flags.  tools/exp_bench.py A/B-tests them.
"""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rs_pbrt_b200 import _build  # noqa: E402


def main():
    specs = [a.split(":", 1) for a in sys.argv[1:]]
    _build.build()  # makes build/pbrt_host.o and build/sobol_blob.o
    out_dir = ROOT / "variants"
    out_dir.mkdir(exist_ok=True)
    nvcc = "/usr/local/cuda/bin/nvcc"

    def one(spec):
        name, flags = spec[0], [f for f in (spec[1].split(",") if len(spec) > 1 and spec[1] else []) if f]
        obj = ROOT / "build" / ("pbrt_gpu_var_%s.o" % name)
        out = out_dir / ("lib_%s.so" % name)
        subprocess.run([nvcc] + _build.NVCC_FLAGS + flags + ["-c", str(_build.CSRC / "pbrt_gpu.cu"), "-o", str(obj)], check=True)
        subprocess.run([nvcc, "-shared", "-o", str(out), str(obj), str(ROOT / "build" / "pbrt_host.o"), str(ROOT / "build" / "sobol_blob.o"), "-Xcompiler", "-pthread", "-lcudart"],
                       check=True)
        return str(out)

    with ThreadPoolExecutor(max_workers=4) as ex:
        for r in ex.map(one, specs):
            print("built", r)


if __name__ == "__main__":
    main()
