#!/bin/bash
# Round 2, GPU call 7: instanced wide traversal + two-round light search A/B, e2e overhead diagnostic, fast parity subset.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_full_configs.py > $o/c7_pytest.log 2>&1; echo "pytest -m gpu (without the full-size file): exit $?" | tee $o/c7_summary.txt
tail -3 $o/c7_pytest.log >> $o/c7_summary.txt
timeout 900 python -m pytest tests/test_gpu_full_configs.py -q -m gpu -x -k "c5 or c4" > $o/c7_pytest_full.log 2>&1; echo "pytest C4/C5 tiles: exit $?" >> $o/c7_summary.txt
timeout 900 python tools/exp_bench.py --scenes landscape,conference,statue --check --out $o/c7_exp.jsonl --libs default,default@PB_WIDE=0 > $o/c7_exp.log 2>&1
echo "exp_bench: exit $?" >> $o/c7_summary.txt
timeout 600 python tools/diag_e2e.py > $o/c7_diag_e2e.txt 2>&1
cat $o/c7_summary.txt; cat $o/c7_exp.jsonl; grep "e2e step" $o/c7_diag_e2e.txt
