#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON for profiles/.

usage: ncu_summary.py report.ncu-rep out.json [kernel-name-substring] [limiter: one sentence read off the stall picture, kept with the numbers]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_sample_buffer_full",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    # --print-units base: ncu otherwise auto-scales every VALUE on its own (one launch's dram__bytes_read in Gbyte, its
    # dram__bytes_write in Mbyte, the next launch's duration in us or ms) while the CSV carries a single unit row per column
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if filt and filt not in d.get("Kernel Name", ""):
            continue
        e = {"kernel": d.get("Kernel Name", "")[:80]}
        for k in KEYS:
            if k in d and d[k] != "":
                try:
                    e[k] = float(d[k])
                except ValueError:
                    e[k] = d[k]
        launches.append(e)
    u = {k: units[hdr.index(k)] for k in KEYS if k in hdr}
    res = {"report": rep.split("/")[-1], "units": u, "launches": launches}
    if launches:
        # base units are expected (byte, ns); if this ncu still scales a column, bring it back to base -- per column, which is
        # only right BECAUSE --print-units base makes the unit uniform down the column
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum"):
            f = scale[u.get(k, "byte")]
            if f != 1.0:
                for l in launches:
                    if k in l:
                        l[k] *= f
                u[k] = "byte" if "bytes" in k else "ns"
        tot = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in launches)
        res["dram_bytes_per_launch"] = tot / len(launches)
        ns = sum(l.get("gpu__time_duration.sum", 0) for l in launches)
        res["ns_per_launch"] = ns / len(launches)
        res["dram_gbs"] = tot / ns if ns else None  # bytes per ns == GB/s
        res["active_lanes_per_inst"] = sum(l.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) for l in launches) / len(launches)
    if len(sys.argv) > 4:
        res["limiter"] = sys.argv[4]
    json.dump(res, open(out, "w"), indent=1)
    print("wrote", out, len(launches), "launches")


if __name__ == "__main__":
    main()
