for v in 3 4 5 6 8; do
  for w in cornell statue; do
    PB_SHADE_MINB=$v python bench.py --workload $w --steps 1 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('minb=$v $w', round(d['value']), round(d['ms_per_step'],1), {k:round(x,1) for k,x in d['kernel_ms_per_step'].items()})"
  done
done
