#!/bin/bash
# Round 2, GPU call 16 (1 GPU): batches in flight (2 / 3 / 4 streams) with the faster k_shade; the device-decided loops polled late vs after every
# iteration (directlighting / whitted); where the conference frame's k_shade time goes, per instantiation (launch list).
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python tools/exp_bench.py --scenes statue,cornell,conference --libs "default,default@PB_STREAMS=3,default@PB_STREAMS=4" --check --out $o/c16_streams.jsonl > $o/c16_streams.log 2>&1; echo "streams: exit $?" | tee $o/c16_summary.txt
timeout 900 python tools/exp_bench.py --scenes cornell-direct,cornell-whitted --libs "default,default@PB_POLL_LAG=0" --check --out $o/c16_poll.jsonl > $o/c16_poll.log 2>&1; echo "poll: exit $?" >> $o/c16_summary.txt
export PB_STREAMS=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000 --csv --log-file $o/c16_launches_conference.csv python bench.py --workload conference --steps 1 --warmup 1 --no-cpu --no-extra > $o/c16_ncu.log 2>&1; echo "ncu launch list: exit $?" >> $o/c16_summary.txt
rm -f $o/c16_ncu.log
cat $o/c16_summary.txt
cut -c1-260 $o/c16_streams.jsonl $o/c16_poll.jsonl
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(open("gpurun_out/c16_launches_conference.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
for k, v in tot.most_common(): print("%-45s %5d %10.0f us %5.1f%%" % (k[:45], cnt[k], v, 100 * v / T))
PY
