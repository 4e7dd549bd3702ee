#!/bin/bash
# Round 2, GPU call 19 (8 GPUs): the final build on 8 / 4 / 2 devices -- statue scaling curve, conference on 4, the config-shaped landscape on 8 --
# and pbrt_gpu_render_multi over all 8 devices from one process.
mkdir -p gpurun_out
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $o/c19_pytest.log 2>&1; echo "pytest multi (8 devices): exit $?" | tee $o/c19_summary.txt
tail -2 $o/c19_pytest.log >> $o/c19_summary.txt
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > $o/c19_bench_statue_n8.json 2> $o/c19_bench_statue_n8.err; echo "statue N=8: exit $?" >> $o/c19_summary.txt
timeout 900 $T --nproc-per-node 4 --master-port 29632 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-inproc --no-extra > $o/c19_bench_statue_n4.json 2> $o/c19_bench_statue_n4.err; echo "statue N=4: exit $?" >> $o/c19_summary.txt
timeout 900 $T --nproc-per-node 2 --master-port 29633 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-inproc --no-extra > $o/c19_bench_statue_n2.json 2> $o/c19_bench_statue_n2.err; echo "statue N=2: exit $?" >> $o/c19_summary.txt
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-extra > $o/c19_bench_statue_n1.json 2> $o/c19_bench_statue_n1.err; echo "statue N=1: exit $?" >> $o/c19_summary.txt
timeout 900 $T --nproc-per-node 4 --master-port 29634 bench.py --gpus 4 --workload conference --steps 3 --warmup 3 --no-cpu > $o/c19_bench_conference_n4.json 2> $o/c19_bench_conference_n4.err; echo "conference N=4: exit $?" >> $o/c19_summary.txt
timeout 900 $T --nproc-per-node 8 --master-port 29635 bench.py --gpus 8 --workload landscape --steps 2 --warmup 2 --no-cpu > $o/c19_bench_landscape_n8.json 2> $o/c19_bench_landscape_n8.err; echo "landscape N=8: exit $?" >> $o/c19_summary.txt
cat $o/c19_summary.txt
for f in statue_n8 statue_n4 statue_n2 statue_n1 conference_n4 landscape_n8; do python - "$o/c19_bench_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    x = d.get("extra", {})
    print(sys.argv[1], "N=%d value %.0f e2e %.0f ms/step %.1f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]), x.get("render_multi"), (x.get("cornell") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
