#!/bin/bash
# Round 2, GPU call 6: full GPU suite with the full-size tile parity tests, device-side flatten + pinned inputs (e2e), bench lines of
# all four configs at N = 1, ncu captures of the final k_trace_wide / k_shade on the statue for profiles/.
mkdir -p gpurun_out
o=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > $o/c6_pytest.log 2>&1; echo "pytest -m gpu: exit $?" | tee $o/c6_summary.txt
tail -14 $o/c6_pytest.log >> $o/c6_summary.txt
PB_TIMING=1 timeout 300 python - > $o/c6_e2e_timing.txt 2>&1 <<'PY'
import os, time, numpy as np
from rs_pbrt_b200 import scenes, GpuScene, pin_description, unpin_description
h = scenes.statue(n_side=1468, xres=1024, yres=1024, spp=128, n_threads=os.cpu_count())
film = np.zeros((1024, 1024, 4), np.float32)
for pinned in (False, True):
    hd = pin_description(h.desc) if pinned else None
    for i in range(3):
        t0 = time.perf_counter(); g = GpuScene(h.desc, 0); t1 = time.perf_counter()
        film.fill(0.0); t2 = time.perf_counter()
        _, st = g.render(h.params, film=film); t3 = time.perf_counter()
        g.close(); t4 = time.perf_counter()
        print("STEP pinned=%s %d: create %.1f ms, fill %.1f, render call %.1f (device %.1f), destroy %.1f, total %.1f" % (pinned, i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, st["ms_total"], (t4 - t3) * 1e3, (t4 - t0) * 1e3), flush=True)
    if hd: unpin_description(hd)
PY
timeout 600 python bench.py --steps 5 --warmup 3 > $o/c6_bench_statue.json 2> $o/c6_bench_statue.err; echo "bench statue: exit $?" >> $o/c6_summary.txt
for w in conference landscape; do
  timeout 500 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > $o/c6_bench_$w.json 2> $o/c6_bench_$w.err; echo "bench $w: exit $?" >> $o/c6_summary.txt
done
export PB_STREAMS=1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:k_trace -s 97 -c 1 -o $o/c6_trace_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c6_ncu1.log 2>&1
$N -k regex:k_shade -s 97 -c 1 -o $o/c6_shade_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c6_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 500 --csv --log-file $o/c6_launches_statue.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c6_ncu3.log 2>&1
cat $o/c6_summary.txt; grep STEP $o/c6_e2e_timing.txt
