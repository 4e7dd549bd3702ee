#!/bin/bash
# Round 2, GPU call 1: the whole -m gpu suite (the former xfail file is strict now), bench lines of every BASELINE.json config on
# one GPU, what the never-timed kernels cost, the fixed cost of a render call, and ncu captures (with source counters) of k_trace
# and k_shade on the 4.31 M-triangle scene.   gpurun --timeout 1500 -- 'bash tools/r2_call1.sh'
mkdir -p gpurun_out
o=gpurun_out
{ nvidia-smi -L; nproc; free -g | head -2; } > $o/c1_box.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > $o/c1_pytest.log 2>&1; echo "pytest -m gpu: exit $?" | tee $o/c1_summary.txt
tail -3 $o/c1_pytest.log >> $o/c1_summary.txt
# bench lines (the default is the statue now, with extra.cornell)
timeout 600 python bench.py --steps 5 --warmup 3 > $o/c1_bench_statue.json 2> $o/c1_bench_statue.err; echo "bench statue: exit $?" >> $o/c1_summary.txt
for w in conference landscape cornell-textured cornell-direct cornell-whitted cornell-ao; do
  timeout 400 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > $o/c1_bench_$w.json 2> $o/c1_bench_$w.err; echo "bench $w: exit $?" >> $o/c1_summary.txt
done
# A/B: refill thresholds, ray-prep, ray-sort modes -- scenes built once
timeout 600 python tools/exp_bench.py --scenes statue,cornell,conference --check --out $o/c1_exp.jsonl \
  --libs default,default@PB_RAY_PREP=1,variants/lib_refill4.so,variants/lib_refill8.so,variants/lib_refill16.so,default@PB_RAY_SORT=2,default@PB_RAY_SORT=1 > $o/c1_exp.log 2>&1
echo "exp_bench: exit $?" >> $o/c1_summary.txt
# fixed cost of one render call: Cornell, whole frame and a 1/8 band, host-side phase clock
PB_TIMING=1 timeout 200 python - > $o/c1_timing.txt 2>&1 <<'PY'
import time, torch
from rs_pbrt_b200 import scenes, GpuScene
h = scenes.cornell_box(xres=1024, yres=1024, spp=256)
g = GpuScene(h.desc, 0)
film = torch.zeros((1024, 1024, 4), dtype=torch.float32, device="cuda")
for rect in ([0, 0, 1024, 1024], [0, 0, 1024, 128], [0, 0, 1024, 128], [0, 448, 1024, 576]):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = g.render_device(h.params, film.data_ptr(), rect=rect)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print("rect", rect, "wall %.2f ms, ms_total %.2f, trace %.2f shade %.2f, rays %d, launches %d" % (dt, st["ms_total"], st["ms_trace"], st["ms_shade"], st["rays"], st["kernel_launches"]), flush=True)
PY
# ncu: one secondary-ray k_trace launch and one k_shade launch of the statue frame, full set with source counters; launch list
export PB_STREAMS=1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:k_trace -s 97 -c 1 -o $o/c1_trace_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c1_ncu1.log 2>&1
$N -k regex:k_shade -s 97 -c 1 -o $o/c1_shade_statue python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c1_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 500 --csv --log-file $o/c1_launches_statue.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c1_ncu3.log 2>&1
ls -la $o >> $o/c1_summary.txt
cat $o/c1_summary.txt
