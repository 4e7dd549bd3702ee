"""Build librs_pbrt_b200.so (CUDA kernels + C ABI + C++ host mirror) in-tree for sm_100a.

Flags that matter for parity (DESIGN.md "Numerics"):
  -fmad=false           no FMA contraction on the device (rustc/LLVM never contracts)
  -ffp-contract=off     same for the host-side C++
  default -prec-div / -prec-sqrt / -ftz=false: IEEE division, sqrt and denormals
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "rs_pbrt_b200" / "csrc"
OUT = ROOT / "rs_pbrt_b200" / "librs_pbrt_b200.so"
BUILD = ROOT / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-pthread",
]


def _run(cmd):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True)


def _newer(target, sources):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build(force=False, verbose_ptxas=False, defines=(), out=None):
    """`defines` / `out`: experiment builds (tools/exp_variants.sh) with extra -D flags into another .so."""
    BUILD.mkdir(exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    blob_s = BUILD / "sobol_blob.S"
    blob_o = BUILD / "sobol_blob.o"
    table = ROOT / "data" / "sobol_tables.bin"
    blob_s.write_text(
        "    .section .rodata\n    .global pb_sobol_blob_start\n    .global pb_sobol_blob_end\n    .balign 256\n"
        "pb_sobol_blob_start:\n    .incbin \"%s\"\npb_sobol_blob_end:\n    .section .note.GNU-stack,\"\",@progbits\n" % table
    )
    sources = [CSRC / "pbrt_gpu.cu", CSRC / "pbrt_host.cpp"]
    deps = list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h")) + sources + [table, Path(__file__)]
    out = Path(out) if out else OUT
    if not force and not defines and not _newer(out, deps):
        return out
    _run(["gcc", "-c", blob_s, "-o", blob_o])
    gpu_o = BUILD / ("pbrt_gpu%s.o" % ("_" + out.stem if defines else ""))
    host_o = BUILD / "pbrt_host.o"
    flags = list(NVCC_FLAGS) + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-D" + d for d in defines]
    _run([nvcc] + flags + ["-c", CSRC / "pbrt_gpu.cu", "-o", gpu_o])
    _run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-Wall", "-c", CSRC / "pbrt_host.cpp", "-o", host_o])
    _run([nvcc, "-shared", "-o", out, gpu_o, host_o, blob_o, "-Xcompiler", "-pthread", "-lcudart"])
    return out


def build_oracle(force=False):
    """The oracle is test infrastructure; building the checker is not using it."""
    out = ROOT / "oracle" / "_build" / "liboracle.so"
    srcs = list((ROOT / "oracle").glob("*.hpp")) + list((ROOT / "oracle").glob("*.cpp")) + [ROOT / "include" / "pbrt_gpu.h"]
    if force or _newer(out, srcs):
        _run(["make", "-C", ROOT / "oracle", "-B" if force else "-s"])
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
    build_oracle()
