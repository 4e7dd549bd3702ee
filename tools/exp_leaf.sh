for v in $1; do
  sed -i "s/\"-fmad=false\",/\"-fmad=false\", \"-DPB_LEAF_MIN=$v\",/" rs_pbrt_b200/_build.py
  python rs_pbrt_b200/_build.py --force > /dev/null 2>&1
  sed -i "s/ \"-DPB_LEAF_MIN=$v\",//" rs_pbrt_b200/_build.py
  for w in cornell statue; do
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('leaf_min=$v $w', round(d['value']), round(d['ms_per_step'],1), {k[:7]:round(x,1) for k,x in d['kernel_ms_per_step'].items()})"
  done
done
python rs_pbrt_b200/_build.py --force > /dev/null 2>&1
