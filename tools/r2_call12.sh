#!/bin/bash
# Round 2, GPU call 12 (1 GPU): the suite on the build with the late-polled loops and the sparse light tables; scene_create outliers after the
# capability-check change; what N ranks would each do, one share after the other on this GPU, for tile runs / batch sizes; bench lines.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/c12_pytest.log 2>&1; echo "pytest -m gpu: exit $?" | tee $o/c12_summary.txt
tail -3 $o/c12_pytest.log >> $o/c12_summary.txt
timeout 300 python tools/diag_e2e2.py > $o/c12_diag_e2e2.txt 2>&1; echo "diag_e2e2: exit $?" >> $o/c12_summary.txt
V="default,default@PB_TILE_RUN=4,default@PB_TILE_RUN=16,default@PB_TILE_RUN=64,default@PB_BATCH_LOG2=21,default@PB_BATCH_LOG2=20,default@PB_TILE_RUN=16+PB_BATCH_LOG2=21"
timeout 900 python tools/exp_bench.py --scenes statue,cornell --parts 8 --libs "$V" --out $o/c12_parts8.jsonl > $o/c12_parts8.log 2>&1; echo "parts 8: exit $?" >> $o/c12_summary.txt
timeout 600 python tools/exp_bench.py --scenes statue --parts 4 --libs "default,default@PB_TILE_RUN=16,default@PB_BATCH_LOG2=21" --out $o/c12_parts4.jsonl > $o/c12_parts4.log 2>&1; echo "parts 4: exit $?" >> $o/c12_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $o/c12_bench_statue.json 2> $o/c12_bench_statue.err; echo "bench statue: exit $?" >> $o/c12_summary.txt
for w in cornell-direct cornell-whitted; do
  timeout 400 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > $o/c12_bench_$w.json 2> $o/c12_bench_$w.err; echo "bench $w: exit $?" >> $o/c12_summary.txt
done
cat $o/c12_summary.txt
cat $o/c12_parts8.jsonl $o/c12_parts4.jsonl | cut -c1-400
grep -E "slow|mean|median" $o/c12_diag_e2e2.txt | tail -12
for w in statue cornell-direct cornell-whitted; do python - "$o/c12_bench_$w.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
