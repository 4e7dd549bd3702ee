#!/usr/bin/env python3
"""Per-kernel totals of ONE WHOLE FRAME from an ncu metrics CSV (run here or on the GPU box, no GPU needed).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,\\
smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,\\
sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s S -c <launches per frame> --csv --log-file frame.csv python bench.py ...
    python tools/ncu_frame_summary.py frame.csv out_dir workload [note]

The batches of a frame differ (tiles in Morton order: one batch sees the statue and the sky, the next the ground), so a capture of a few
launches is not the frame; -c <launches per frame> launches of the periodic bench loop are, wherever the window starts.  Writes
out_dir/ncu_k_trace_<workload>.json and ncu_k_shade_<workload>.json in the form bench.py::ncu_record reads (`dram_bytes_per_launch`,
`ns_per_launch`, `dram_gbs`, `active_lanes_per_inst`, `limiter`) plus the per-instantiation rows they were summed from."""
import collections
import csv
import json
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}


def main():
    path, out_dir, workload = sys.argv[1], sys.argv[2], sys.argv[3]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    idx = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= idx["Metric Value"]:
            continue
        e = launches.setdefault(r[idx["ID"]], {"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("pb::", "")})
        try:
            v = float(r[idx["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        e[r[idx["Metric Name"]]] = v * SCALE.get(r[idx["Metric Unit"]], 1.0)
    per = collections.OrderedDict()
    for e in launches.values():
        p = per.setdefault(e["kernel"], collections.Counter())
        ns = e.get("gpu__time_duration.sum", 0.0)
        inst = e.get("smsp__inst_executed.sum", 0.0)
        p["launches"] += 1
        p["ns"] += ns
        p["dram_bytes"] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
        p["warp_inst"] += inst
        p["thread_inst"] += inst * e.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
        p["issue_ns"] += ns * e.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
        p["occ_ns"] += ns * e.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
    total_ns = sum(p["ns"] for p in per.values())

    def fold(names):
        t = collections.Counter()
        for n in names:
            t.update(per[n])
        if not t["launches"]:
            return None
        return {"launches": int(t["launches"]), "ns_per_launch": t["ns"] / t["launches"], "dram_bytes_per_launch": t["dram_bytes"] / t["launches"],
                "dram_gbs": t["dram_bytes"] / t["ns"] if t["ns"] else None, "active_lanes_per_inst": t["thread_inst"] / t["warp_inst"] if t["warp_inst"] else None,
                "issue_active_pct": t["issue_ns"] / t["ns"] if t["ns"] else None, "warps_active_pct": t["occ_ns"] / t["ns"] if t["ns"] else None,
                "share_of_captured_time": t["ns"] / total_ns if total_ns else None, "warp_instructions": t["warp_inst"]}

    table = {n: fold([n]) for n in per}
    for fam in ("k_trace", "k_shade"):
        names = [n for n in per if n.startswith(fam)]
        f = fold(names)
        if not f:
            continue
        f.update({"kernel_family": fam, "source": path.split("/")[-1], "window": "%d consecutive launches (one frame of the periodic bench loop)" % len(launches),
                  "note": note, "instantiations": {n: table[n] for n in names}})
        lanes, issue, gbs = f["active_lanes_per_inst"], f["issue_active_pct"], f["dram_gbs"]
        f["limiter"] = ("%.1f of 32 lanes active per issued instruction, issue slots %.0f %% busy, %.0f GB/s of DRAM traffic measured "
                        "(time-weighted over the frame's launches, ncu --cache-control all: cold caches per launch)" % (lanes, issue, gbs))
        json.dump(f, open("%s/ncu_%s_%s.json" % (out_dir, fam, workload), "w"), indent=1)
        print(fam, json.dumps({k: f[k] for k in ("launches", "ns_per_launch", "dram_bytes_per_launch", "dram_gbs", "active_lanes_per_inst", "issue_active_pct", "warps_active_pct",
                                                  "share_of_captured_time")}))
    json.dump({"launches": len(launches), "total_ns": total_ns, "per_kernel": table}, open("%s/ncu_frame_%s.json" % (out_dir, workload), "w"), indent=1)


if __name__ == "__main__":
    main()
