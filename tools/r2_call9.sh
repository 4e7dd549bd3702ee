#!/bin/bash
# Round 2, GPU call 9: allocation-free scene re-creation (slow-step diagnostic again), bench lines statue / conference / landscape-64 at N = 1.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_full_configs.py > $o/c9_pytest.log 2>&1; echo "pytest -m gpu (without the full-size file): exit $?" | tee $o/c9_summary.txt
tail -3 $o/c9_pytest.log >> $o/c9_summary.txt
timeout 600 python tools/diag_e2e2.py > $o/c9_diag_e2e2.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $o/c9_bench_statue.json 2> $o/c9_bench_statue.err; echo "bench statue: exit $?" >> $o/c9_summary.txt
for w in conference landscape-64; do
  timeout 500 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > $o/c9_bench_$w.json 2> $o/c9_bench_$w.err; echo "bench $w: exit $?" >> $o/c9_summary.txt
done
cat $o/c9_summary.txt; grep -E "SLOW|e2e step" $o/c9_diag_e2e2.txt | tail -12
