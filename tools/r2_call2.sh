#!/bin/bash
# Round 2, GPU call 2: parity of the rewritten pieces (flat node walk, SPEC shade kernel, tile shares, multi-device entry point, staged
# scene upload), A/B of the kernel variants on all four scenes, scene_create phases on hardware, a bench line, fresh ncu captures.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/c2_pytest.log 2>&1; echo "pytest -m gpu: exit $?" | tee $o/c2_summary.txt
tail -3 $o/c2_pytest.log >> $o/c2_summary.txt
PB_TIMING=1 timeout 300 python tools/time_scene_create.py > $o/c2_scene_create.txt 2>&1
timeout 900 python tools/exp_bench.py --scenes statue,cornell,conference,landscape --check --out $o/c2_exp.jsonl \
  --libs default,variants/lib_flat0.so,variants/lib_ns32.so,variants/lib_ns16.so,default@PB_SHADE_SPEC=0 > $o/c2_exp.log 2>&1
echo "exp_bench: exit $?" >> $o/c2_summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $o/c2_bench_statue.json 2> $o/c2_bench_statue.err; echo "bench statue: exit $?" >> $o/c2_summary.txt
export PB_STREAMS=1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:k_trace -s 97 -c 1 -o $o/c2_trace_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c2_ncu1.log 2>&1
$N -k regex:k_shade -s 100 -c 1 -o $o/c2_shade_cornell python bench.py --workload cornell --steps 1 --warmup 3 --no-cpu > $o/c2_ncu2.log 2>&1
cat $o/c2_summary.txt; cat $o/c2_exp.jsonl; cat $o/c2_scene_create.txt | tail -30
