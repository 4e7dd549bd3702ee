#!/usr/bin/env python3
"""Registers / stack / shared memory per kernel of an object or shared library (cuobjdump -res-usage), optionally against a second file:
    python tools/res_usage.py build/pbrt_gpu.o [build/pbrt_gpu_old.o]
Used before GPU time is spent on a change that should not touch the hot kernels' resource usage."""
import re
import subprocess
import sys


def usage(path):
    txt = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        if name and "REG:" in line:
            d = dict(kv.split(":") for kv in line.split() if ":" in kv)
            out[name] = (int(d["REG"]), int(d["STACK"]), int(d["SHARED"]))
            name = None
    return out


if __name__ == "__main__":
    a = usage(sys.argv[1])
    b = usage(sys.argv[2]) if len(sys.argv) > 2 else None
    for k in sorted(a):
        if b is None:
            print("%-110s reg %3d stack %5d smem %6d" % ((k[:110],) + a[k]))
        elif k not in b:
            print("%-110s reg %3d stack %5d smem %6d   (new)" % ((k[:110],) + a[k]))
        elif a[k] != b[k]:
            print("%-110s reg %3d stack %5d smem %6d   was reg %3d stack %5d smem %6d" % ((k[:110],) + a[k] + b[k]))
    if b is not None:
        print("%d kernels, %d unchanged" % (len(a), sum(1 for k in a if b.get(k) == a[k])))
