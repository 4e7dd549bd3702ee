#!/bin/bash
# Round 2, GPU call 10: look-ahead wide traversal, launch bounds and stack split A/B; slow-step diagnostic with exact-size buffer recycling.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_full_configs.py > $o/c10_pytest.log 2>&1; echo "pytest -m gpu (without the full-size file): exit $?" | tee $o/c10_summary.txt
PB_WIDE_SPEC=1 timeout 900 python -m pytest tests/test_gpu_parity_cornell.py tests/test_gpu_parity_materials.py tests/test_gpu_full_configs.py -q -m gpu -x -k "not c1 and not c5" > $o/c10_pytest_spec.log 2>&1; echo "pytest with PB_WIDE_SPEC=1: exit $?" >> $o/c10_summary.txt
tail -2 $o/c10_pytest_spec.log >> $o/c10_summary.txt
timeout 900 python tools/exp_bench.py --scenes statue,conference --check --out $o/c10_exp.jsonl \
  --libs default,default@PB_WIDE_SPEC=1,variants/lib_mb9.so,variants/lib_mb10.so,variants/lib_ws8.so,variants/lib_ws24.so,default@PB_WIDE_SPEC=1+PB_WIDE_WALK=4,default@PB_WIDE_SPEC=1+PB_WIDE_WALK=8 > $o/c10_exp.log 2>&1
echo "exp_bench: exit $?" >> $o/c10_summary.txt
timeout 600 python tools/diag_e2e2.py > $o/c10_diag_e2e2.txt 2>&1
cat $o/c10_summary.txt; cat $o/c10_exp.jsonl; grep -E "SLOW|e2e step" $o/c10_diag_e2e2.txt | tail -12
