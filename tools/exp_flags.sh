# usage: bash tools/exp_flags.sh "<flags variant 1>" "<flags variant 2>" ...   e.g. "-DPB_CACHE_HINTS=0 -DPB_SMEM_STACK_ENTRIES=24"
for v in "$@"; do
  flags=$(python - "$v" <<'PY'
import sys
print(", ".join('"%s"' % f for f in sys.argv[1].split()))
PY
)
  sed -i "s/\"-fmad=false\",/\"-fmad=false\", $flags,/" rs_pbrt_b200/_build.py
  python rs_pbrt_b200/_build.py --force > /dev/null 2>&1
  sed -i "s/ $flags,//" rs_pbrt_b200/_build.py
  for w in ${WORKLOADS:-statue cornell}; do
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$v] $w', round(d['value']), round(d['ms_per_step'],1), {k[:7]:(round(x,1) if isinstance(x,float) else x) for k,x in d['kernel_ms_per_step'].items() if k!='note'})"
  done
done
python rs_pbrt_b200/_build.py --force > /dev/null 2>&1
