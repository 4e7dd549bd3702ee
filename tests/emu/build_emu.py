"""TEST INFRASTRUCTURE ONLY.  Builds tests/emu/_build/librs_pbrt_b200_emu.so: the product's CUDA sources compiled by g++ against the
stand-in cuda_runtime.h of this directory, so that the kernels' source runs on the CPU (see that header for what this can and
cannot show).  pbrt_gpu.cu is not edited: its `kernel<<<cfg>>>(args)` launches are rewritten on the fly into
`emu::launcher(kernel, cfg)(args)`.  The product never loads this library."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
EMU = Path(__file__).resolve().parent
OUT = EMU / "_build"
LIB = OUT / "librs_pbrt_b200_emu.so"
CSRC = ROOT / "rs_pbrt_b200" / "csrc"

LAUNCH = re.compile(r"([A-Za-z_][A-Za-z0-9_]*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>", re.S)


def transform(text):
    text, n = LAUNCH.subn(lambda m: "emu::launcher(%s, %s)" % (m.group(1), m.group(2)), text)
    if n < 20 or "<<<" in text:
        raise SystemExit("launch rewrite: %d launches rewritten, leftovers: %s" % (n, "<<<" in text))
    text = text.replace('#include "../../include/pbrt_gpu.h"', '#include "%s"' % (ROOT / "include" / "pbrt_gpu.h"))
    return text


def build(force=False, sanitize=None):
    """sanitize="address" / "thread" / "undefined": an instrumented copy (lib..._asan.so / _tsan.so / _usan.so) for tools/emu_sanitize.sh"""
    global LIB
    lib = OUT / ("librs_pbrt_b200_emu%s.so" % ("_" + sanitize[0] + "san" if sanitize else ""))
    OUT.mkdir(exist_ok=True)
    srcs = [CSRC / "pbrt_gpu.cu", CSRC / "pbrt_host.cpp", EMU / "emu_engine.cpp", EMU / "include" / "cuda_runtime.h", Path(__file__)] + list(CSRC.glob("*.cuh"))
    if not force and lib.exists() and all(s.stat().st_mtime <= lib.stat().st_mtime for s in srcs):
        return lib
    # (the sanitizers follow host threads, not hand-switched stacks: their builds run every CUDA thread on a host thread, and ThreadSanitizer needs that to see races)
    san = ["-fsanitize=" + sanitize, "-fno-omit-frame-pointer", "-DEMU_THREADS"] if sanitize else []
    if sanitize == "undefined":  # float -> int conversions out of range too: defined (saturating) on the device, undefined on the host
        san += ["-fsanitize=float-cast-overflow", "-fno-sanitize-recover=all"]
    tag = "_" + sanitize[0] + "san" if sanitize else ""
    gen = OUT / "pbrt_gpu_emu.cpp"
    gen.write_text(transform((CSRC / "pbrt_gpu.cu").read_text()))
    table = ROOT / "data" / "sobol_tables.bin"
    blob_s = OUT / "sobol_blob.S"
    blob_s.write_text("    .section .rodata\n    .global pb_sobol_blob_start\n    .global pb_sobol_blob_end\n    .balign 256\n"
                      "pb_sobol_blob_start:\n    .incbin \"%s\"\npb_sobol_blob_end:\n    .section .note.GNU-stack,\"\",@progbits\n" % table)
    cxx = ["g++"] + san + ["-std=c++20", "-O1", "-g", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-DPB_CACHE_HINTS=0", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-I", str(EMU / "include"), "-I", str(CSRC), "-I", str(ROOT / "include")]

    def run(cmd):
        print("+", " ".join(str(c) for c in cmd), flush=True)
        subprocess.run([str(c) for c in cmd], check=True)

    run(["gcc", "-c", blob_s, "-o", OUT / "sobol_blob.o"])
    run(cxx + ["-c", gen, "-o", OUT / ("pbrt_gpu_emu%s.o" % tag)])
    run(cxx + ["-c", EMU / "emu_engine.cpp", "-o", OUT / ("emu_engine%s.o" % tag)])
    run(["g++"] + san + ["-O2", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-c", CSRC / "pbrt_host.cpp", "-o", OUT / ("pbrt_host%s.o" % tag)])
    run(["g++"] + san + ["-shared", "-pthread", "-o", lib, OUT / ("pbrt_gpu_emu%s.o" % tag), OUT / ("emu_engine%s.o" % tag), OUT / ("pbrt_host%s.o" % tag), OUT / "sobol_blob.o"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, sanitize="address" if "--asan" in sys.argv else ("thread" if "--tsan" in sys.argv else ("undefined" if "--ubsan" in sys.argv else None))))
