#!/bin/bash
# Which part of the coherence key pays?  (any-hit bit 0x1000, octant 0x0e00, origin cell 0x01ff)
mkdir -p gpurun_out
out=gpurun_out/exp_raykey.txt
: > $out
for m in 0x1000 0x0e00 0x1e00 0x01ff 0x1fff; do
  export PB_RAY_SORT=1 PB_RAY_KEY_MASK=$m
  for w in cornell conference; do
    timeout 300 python bench.py --workload $w --no-cpu --steps 1 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mask=$m', '$w', round(d['value'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms_per_step'].items() if k!='note'})" >> $out 2>&1
  done
done
cat $out
