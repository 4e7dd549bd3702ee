#!/bin/bash
# Round 2, GPU call 8: per-lobe-kind specialised shade kernels A/B (PB_SHADE_SPEC=1: Lambert only, 2: every single-lobe class), slow-step diagnostic.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_full_configs.py > $o/c8_pytest.log 2>&1; echo "pytest -m gpu (without the full-size file): exit $?" | tee $o/c8_summary.txt
tail -3 $o/c8_pytest.log >> $o/c8_summary.txt
timeout 900 python tools/exp_bench.py --scenes conference,cornell,statue --check --out $o/c8_exp.jsonl --libs default,default@PB_SHADE_SPEC=1 > $o/c8_exp.log 2>&1
echo "exp_bench: exit $?" >> $o/c8_summary.txt
timeout 600 python tools/diag_e2e2.py > $o/c8_diag_e2e2.txt 2>&1
cat $o/c8_summary.txt; cat $o/c8_exp.jsonl; cat $o/c8_diag_e2e2.txt | tail -40
