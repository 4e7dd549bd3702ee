#!/bin/bash
# Round 2, GPU call 5 (2 GPUs): the multi-GPU pieces on hardware -- pbrt_gpu_render_multi with NVLink peer access, the per-process tile
# shares + NCCL reduce through bench.py (N = 2), and the walk-step / dense-L follow-ups at N = 1.
mkdir -p gpurun_out
o=gpurun_out
nvidia-smi topo -m > $o/c5_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity_cornell.py -q -m gpu -x > $o/c5_pytest.log 2>&1; echo "pytest multi: exit $?" | tee $o/c5_summary.txt
tail -3 $o/c5_pytest.log >> $o/c5_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > $o/c5_bench_statue_n2.json 2> $o/c5_bench_statue_n2.err
echo "bench N=2: exit $?" >> $o/c5_summary.txt
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > $o/c5_bench_statue_n1.json 2> $o/c5_bench_statue_n1.err; echo "bench N=1: exit $?" >> $o/c5_summary.txt
timeout 600 python tools/exp_bench.py --scenes statue,cornell,conference --out $o/c5_exp.jsonl --libs default,default@PB_WIDE_WALK=8,default@PB_WIDE_WALK=16 > $o/c5_exp.log 2>&1
cat $o/c5_summary.txt; cat $o/c5_exp.jsonl; tail -5 $o/c5_bench_statue_n2.err
