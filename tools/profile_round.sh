#!/bin/bash
# Run on the GPU box: ncu captures of the two hot kernels (one batch in flight, like bench.py's roofline pass) + launch list.
# usage: tools/profile_round.sh <tag>
tag=${1:-rXX}
mkdir -p gpurun_out
export PB_STREAMS=1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:k_shade -s 100 -c 1 -o gpurun_out/prof_shade_cornell_$tag python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.1.log 2>&1
$N -k regex:k_trace -s 100 -c 1 -o gpurun_out/prof_trace_cornell_$tag python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.2.log 2>&1
$N -k regex:k_trace -s 97 -c 1 -o gpurun_out/prof_trace_statue_$tag python bench.py --workload statue --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.3.log 2>&1
$N -k regex:k_shade -s 60 -c 1 -o gpurun_out/prof_shade_conference_$tag python bench.py --workload conference --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$tag.4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/launches_cornell_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.5.log 2>&1
ls -la gpurun_out/*$tag*
