#!/bin/bash
# Round 2, GPU call 14 (1 GPU): plastic-specialised k_shade (class 9) and the restored k_trace_wide bounds against the call-9 build; the GPU suite; bench lines; ncu --set full of one whole batch (6 x k_trace_wide + 12 x k_shade, PB_STREAMS=1) and a launch list.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/c14_pytest.log 2>&1; echo "pytest -m gpu: exit $?"; tail -2 $o/c14_pytest.log
timeout 900 python tools/exp_bench.py --scenes statue,cornell,conference,landscape-64 --libs "default,variants/lib_c9.so" --check --out $o/c14_exp.jsonl > $o/c14_exp.log 2>&1; echo "exp_bench: exit $?" | tee $o/c14_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $o/c14_bench_statue.json 2> $o/c14_bench_statue.err; echo "bench statue: exit $?" >> $o/c14_summary.txt
for w in cornell conference landscape-64; do
  timeout 500 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-extra > $o/c14_bench_$w.json 2> $o/c14_bench_$w.err; echo "bench $w: exit $?" >> $o/c14_summary.txt
done
export PB_STREAMS=1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_trace_wide|k_shade" -s 180 -c 18 -o $o/c14_batch_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c14_ncu1.log 2>&1; echo "ncu batch: exit $?" >> $o/c14_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 500 --csv --log-file $o/c14_launches_statue.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c14_ncu2.log 2>&1; echo "ncu launch list: exit $?" >> $o/c14_summary.txt
cat $o/c14_summary.txt
cut -c1-330 $o/c14_exp.jsonl
for w in statue cornell conference landscape-64; do python - "$o/c14_bench_$w.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["kernel_ms_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
