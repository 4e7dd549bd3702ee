#!/usr/bin/env python3
"""Map an ncu SASS source page back to CUDA source lines (needs the .so built with -lineinfo from the same sources).

usage: ncu_hotspots.py report.ncu-rep kernel_regex [launch_skip] [--so path] [--top N]
Prints instruction-weighted totals per innermost source line and per line of the kernel body (outermost inline frame).
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def disasm(so, kernel_re):
    tmp = Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(so)], cwd=tmp, capture_output=True)
    cubin = [p for p in tmp.glob("*.cubin") if "sm_100" in p.name][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout
    out, cur, chain, active = [], None, [], False
    for line in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", line)
        if m:
            active = re.search(kernel_re, m.group(1)) is not None
            chain = []
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            if not chain or chain[-1].get("closed"):
                chain = [{"frames": []}]
            chain[-1]["frames"].append((Path(m.group(1)).name, int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            frames = chain[-1]["frames"] if chain else []
            out.append((int(m.group(1), 16), m.group(2).strip(), list(frames)))
            if chain:
                chain[-1]["closed"] = True
    return out


def main():
    args = []
    it = iter(sys.argv[1:])
    for a in it:
        if a.startswith("--"):
            next(it, None)
        else:
            args.append(a)
    rep, kre = args[0], args[1]
    skip = int(args[2]) if len(args) > 2 else 0
    so = ROOT / "rs_pbrt_b200" / "librs_pbrt_b200.so"
    top = 25
    dis_re = kre
    for i, a in enumerate(sys.argv):
        if a == "--so":
            so = Path(sys.argv[i + 1])
        if a == "--top":
            top = int(sys.argv[i + 1])
        if a == "--dis-re":
            dis_re = sys.argv[i + 1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hi = his[min(skip, len(his) - 1)]
    hdr = rows[hi]
    data = []
    for r in rows[hi + 1:]:
        if not r or r[0] == "Kernel Name":
            break
        data.append(r)
    col = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_wait", "stall_short_sb",
                                      "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_lg", "stall_no_inst")}
    dis = disasm(so, dis_re)
    print("ncu instructions: %d, disassembly instructions: %d" % (len(data), len(dis)))
    n = min(len(data), len(dis))
    inner, outer = defaultdict(lambda: [0, 0, 0]), defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for i in range(n):
        r = data[i]
        try:
            smp, ins, tins = int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]])
        except ValueError:
            continue
        frames = dis[i][2]
        k_in = frames[0] if frames else ("?", 0)
        k_out = frames[-1] if frames else ("?", 0)
        for d, k in ((inner, k_in), (outer, k_out)):
            d[k][0] += smp; d[k][1] += ins; d[k][2] += tins
        tot[0] += smp; tot[1] += ins; tot[2] += tins
    print("total: samples %d, warp-instructions %d, avg active threads %.2f" % (tot[0], tot[1], tot[2] / max(tot[1], 1)))
    for name, d in (("innermost source line", inner), ("kernel-body line (outermost frame)", outer)):
        print("\n== by %s: samples%%  inst%%  avg-threads  file:line" % name)
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print("  %5.1f%%  %5.1f%%  %5.1f   %s:%d" % (100.0 * v[0] / max(tot[0], 1), 100.0 * v[1] / max(tot[1], 1), v[2] / max(v[1], 1), k[0], k[1]))


if __name__ == "__main__":
    main()
