#!/bin/bash
# Run on the GPU box: ray-queue coherence order on / off for the three bench workloads (+ parity suite with it on).
mkdir -p gpurun_out
out=gpurun_out/exp_raysort.txt
: > $out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $out
for rs in 0 1; do
  export PB_RAY_SORT=$rs
  for w in cornell statue conference; do
    st=3; [ $w = conference ] && st=2
    timeout 300 python bench.py --workload $w --no-cpu --steps $st --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sort=$rs', '$w', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms_per_step'].items() if k!='note'})" >> $out 2>&1
  done
done
cat $out
