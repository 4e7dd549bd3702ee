#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp_raysort2.txt
: > $out
PB_RAY_SORT=1 timeout 120 python -m pytest tests/test_gpu_parity_cornell.py tests/test_gpu_parity_materials.py tests/test_gpu_parity_lights.py -m gpu -q -x 2>&1 | tail -1 >> $out
run() { PB_RAY_SORT=$1 timeout 100 python bench.py --workload $2 --no-cpu --steps $3 --warmup $4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sort=$1', '$2', round(d['value'],1), 'ms', round(d['ms_per_step'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms_per_step'].items() if k!='note'})" >> $out 2>&1; }
run 1 cornell 3 3
run 0 cornell 3 3
run 1 conference 1 2
run 1 statue 2 3
cat $out
