#!/bin/bash
# Run on the GPU box: parity subset + bench of every experiment build in variants/ (built by tools/build_variants.py).
mkdir -p gpurun_out
out=gpurun_out/exp_variants.txt
: > $out
for lib in rs_pbrt_b200/librs_pbrt_b200.so variants/lib_*.so; do
  name=$(basename $lib .so)
  export RS_PBRT_B200_LIB=$PWD/$lib
  echo "=== $name" >> $out
  timeout 200 python -m pytest tests/test_gpu_parity_cornell.py tests/test_gpu_parity_materials.py -m gpu -q -x 2>&1 | tail -1 >> $out
  timeout 200 python bench.py --no-cpu --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cornell', round(d['value'],1), 'ms', round(d['ms_per_step'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms_per_step'].items() if k!='note'})" >> $out 2>&1
  timeout 300 python bench.py --workload conference --no-cpu --steps 1 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('conference', round(d['value'],1), 'ms', round(d['ms_per_step'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms_per_step'].items() if k!='note'})" >> $out 2>&1
done
cat $out
