#!/bin/bash
# Round 2, GPU call 4: A/B of the wide traversal (this time with a working switch), its walk-step count, the interleaved state records,
# the SPEC kernel's launch bounds; bench line with the in-process NVML clock sampler.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/c4_pytest.log 2>&1; echo "pytest -m gpu: exit $?" | tee $o/c4_summary.txt
tail -3 $o/c4_pytest.log >> $o/c4_summary.txt
timeout 1200 python tools/exp_bench.py --scenes statue,cornell,conference --check --out $o/c4_exp.jsonl \
  --libs default,default@PB_WIDE=0,default@PB_STATE_AOS=0,variants/lib_wws6.so,variants/lib_wws10.so,variants/lib_spec5.so,variants/lib_spec6.so > $o/c4_exp.log 2>&1
echo "exp_bench: exit $?" >> $o/c4_summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $o/c4_bench_statue.json 2> $o/c4_bench_statue.err; echo "bench statue: exit $?" >> $o/c4_summary.txt
cat $o/c4_summary.txt; cat $o/c4_exp.jsonl
