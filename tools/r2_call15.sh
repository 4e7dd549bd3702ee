#!/bin/bash
# Round 2, GPU call 15 (1 GPU): the suite with the translucent material; final 1-GPU bench lines; ncu --set full of one whole batch of the statue
# frame (6 x k_trace_wide, 12 x k_shade; PB_STREAMS=1), summarised on the box (the reports themselves are kept only while small), and a launch list.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/c15_pytest.log 2>&1; echo "pytest -m gpu: exit $?" | tee $o/c15_summary.txt
tail -2 $o/c15_pytest.log >> $o/c15_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $o/c15_bench_statue.json 2> $o/c15_bench_statue.err; echo "bench statue: exit $?" >> $o/c15_summary.txt
for w in cornell conference landscape-64; do
  timeout 500 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-extra > $o/c15_bench_$w.json 2> $o/c15_bench_$w.err; echo "bench $w: exit $?" >> $o/c15_summary.txt
done
export PB_STREAMS=1
N="ncu --set full --clock-control none"
timeout 600 $N -k regex:k_trace_wide -s 60 -c 6 -o $o/c15_trace_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c15_ncu1.log 2>&1; echo "ncu k_trace_wide: exit $?" >> $o/c15_summary.txt
timeout 600 $N -k regex:k_shade -s 120 -c 12 -o $o/c15_shade_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c15_ncu2.log 2>&1; echo "ncu k_shade: exit $?" >> $o/c15_summary.txt
python tools/ncu_summary.py $o/c15_trace_statue.ncu-rep $o/c15_ncu_k_trace_statue.json k_trace_wide >> $o/c15_summary.txt 2>&1
python tools/ncu_summary.py $o/c15_shade_statue.ncu-rep $o/c15_ncu_k_shade_statue.json k_shade >> $o/c15_summary.txt 2>&1
for f in $o/c15_trace_statue.ncu-rep $o/c15_shade_statue.ncu-rep; do
  if [ -f $f ] && [ $(stat -c %s $f) -gt 20000000 ]; then echo "$f: $(stat -c %s $f) bytes, summary kept, report dropped" >> $o/c15_summary.txt; rm -f $f; fi
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 500 --csv --log-file $o/c15_launches_statue.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c15_ncu3.log 2>&1; echo "ncu launch list: exit $?" >> $o/c15_summary.txt
rm -f $o/c15_ncu1.log $o/c15_ncu2.log $o/c15_ncu3.log
cat $o/c15_summary.txt
du -sh $o
for w in statue cornell conference landscape-64; do python - "$o/c15_bench_$w.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["kernel_ms_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
