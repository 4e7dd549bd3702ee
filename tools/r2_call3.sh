#!/bin/bash
# Round 2, GPU call 3: wide-record traversal (k_trace_wide) against the reference-layout walk, staged node upload, e2e phase clock.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/c3_pytest.log 2>&1; echo "pytest -m gpu: exit $?" | tee $o/c3_summary.txt
tail -3 $o/c3_pytest.log >> $o/c3_summary.txt
timeout 900 python tools/exp_bench.py --scenes statue,cornell,conference --check --out $o/c3_exp.jsonl --libs default,default@PB_WIDE=0 > $o/c3_exp.log 2>&1
echo "exp_bench: exit $?" >> $o/c3_summary.txt
PB_TIMING=1 timeout 300 python - > $o/c3_e2e_timing.txt 2>&1 <<'PY'
import os, time, numpy as np
from rs_pbrt_b200 import scenes, GpuScene
h = scenes.statue(n_side=1468, xres=1024, yres=1024, spp=128, n_threads=os.cpu_count())
film = np.zeros((1024, 1024, 4), np.float32)
for i in range(4):
    t0 = time.perf_counter(); g = GpuScene(h.desc, 0); t1 = time.perf_counter()
    film.fill(0.0); t2 = time.perf_counter()
    _, st = g.render(h.params, film=film); t3 = time.perf_counter()
    g.close(); t4 = time.perf_counter()
    print("STEP %d: create %.1f ms, fill %.1f, render call %.1f (device %.1f), destroy %.1f, total %.1f" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, st["ms_total"], (t4 - t3) * 1e3, (t4 - t0) * 1e3), flush=True)
PY
timeout 600 python bench.py --steps 5 --warmup 3 > $o/c3_bench_statue.json 2> $o/c3_bench_statue.err; echo "bench statue: exit $?" >> $o/c3_summary.txt
export PB_STREAMS=1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 97 -c 1 -o $o/c3_trace_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c3_ncu1.log 2>&1
cat $o/c3_summary.txt; cat $o/c3_exp.jsonl; grep -v "dev 0" $o/c3_e2e_timing.txt | tail -40
