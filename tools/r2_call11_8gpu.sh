#!/bin/bash
# Round 2, GPU call 11 (8 GPUs): the configs BASELINE.json quotes on several GPUs -- statue at N = 8 / 4 (scaling curve with N = 1, 2 of earlier calls),
# conference on 4, the config-shaped landscape (3000 instances of 20 prototypes, 1024 spp) on 8 -- each with the CPU arm next to it, and
# pbrt_gpu_render_multi over all 8 devices from one process.
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi topo -m > $o/c11_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $o/c11_pytest.log 2>&1; echo "pytest multi (8 devices): exit $?" | tee $o/c11_summary.txt
tail -2 $o/c11_pytest.log >> $o/c11_summary.txt
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > $o/c11_bench_statue_n8.json 2> $o/c11_bench_statue_n8.err; echo "statue N=8: exit $?" >> $o/c11_summary.txt
timeout 900 $T --nproc-per-node 4 --master-port 29622 bench.py --gpus 4 --steps 5 --warmup 3 --no-inproc > $o/c11_bench_statue_n4.json 2> $o/c11_bench_statue_n4.err; echo "statue N=4: exit $?" >> $o/c11_summary.txt
timeout 900 $T --nproc-per-node 4 --master-port 29623 bench.py --gpus 4 --workload conference --steps 3 --warmup 3 > $o/c11_bench_conference_n4.json 2> $o/c11_bench_conference_n4.err; echo "conference N=4: exit $?" >> $o/c11_summary.txt
timeout 900 $T --nproc-per-node 8 --master-port 29624 bench.py --gpus 8 --workload landscape --steps 2 --warmup 3 > $o/c11_bench_landscape_n8.json 2> $o/c11_bench_landscape_n8.err; echo "landscape N=8: exit $?" >> $o/c11_summary.txt
timeout 400 python bench.py --impl reference --workload conference --steps 1 --warmup 0 > $o/c11_ref_conference.json 2> $o/c11_ref_conference.err; echo "reference conference: exit $?" >> $o/c11_summary.txt
timeout 400 python bench.py --impl reference --workload landscape --steps 1 --warmup 0 > $o/c11_ref_landscape.json 2> $o/c11_ref_landscape.err; echo "reference landscape: exit $?" >> $o/c11_summary.txt
cat $o/c11_summary.txt
for f in statue_n8 statue_n4 conference_n4 landscape_n8; do python - "$o/c11_bench_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "N=%d value %.0f e2e %.0f ms/step %.1f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("extra", {}).get("render_multi"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
