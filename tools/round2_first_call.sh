#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, in the order that tells the most
# per minute.  Run on the GPU box:   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/ (scratch); copy what is to be judged into profiles/.
mkdir -p gpurun_out
o=gpurun_out
# 1. the verified parity suite (must stay green with the instancing / texture changes in the shared kernels), then the unverified tests
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_zz_gpu_unverified.py > $o/r2_pytest_verified.log 2>&1; echo "verified suite: exit $?" | tee $o/r2_summary.txt
timeout 600 python -m pytest tests/test_zz_gpu_unverified.py -q -rxX > $o/r2_pytest_unverified.log 2>&1; echo "unverified suite: exit $?" | tee -a $o/r2_summary.txt
grep -E "XPASS|XFAIL|passed|failed|xpassed|xfailed" $o/r2_pytest_unverified.log | tail -40 >> $o/r2_summary.txt
# 2. bench lines: the three measured workloads (fdiv0, k_shade's texture bit and INST template went in unmeasured), then landscape
for w in cornell statue conference landscape cornell-textured cornell-direct cornell-whitted cornell-ao; do
  timeout 400 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > $o/r2_bench_$w.json 2> $o/r2_bench_$w.err; echo "bench $w: exit $?" >> $o/r2_summary.txt
done
# 3. the two-level ray-order scatter against OFF and mode 1
for m in 0 1 2; do
  PB_RAY_SORT=$m timeout 300 python bench.py --workload cornell --steps 3 --warmup 3 --no-cpu > $o/r2_raysort${m}_cornell.json 2>/dev/null
  PB_RAY_SORT=$m timeout 300 python bench.py --workload conference --steps 2 --warmup 2 --no-cpu > $o/r2_raysort${m}_conference.json 2>/dev/null
done
# 4. textured Cornell: cost of k_texture (per-kernel times of one frame)
timeout 300 python - > $o/r2_textured_cornell.txt 2>&1 <<'PY'
import time
from rs_pbrt_b200 import scenes, GpuScene
for tex in (None, "ewa", "trilinear"):
    h = scenes.cornell_box(xres=1024, yres=1024, spp=64, textures=tex)
    g = GpuScene(h.desc, 0)
    g.render(h.params)
    t0 = time.perf_counter(); _, st = g.render(h.params); dt = time.perf_counter() - t0
    print(tex, "ms_total %.1f trace %.1f shade %.1f rays %d Mrays/s %.0f launches %d" % (st["ms_total"], st["ms_trace"], st["ms_shade"], st["rays"], st["rays"] / st["ms_total"] / 1e3, st["kernel_launches"]), "wall %.1f ms" % (dt * 1e3))
    g.close()
PY
# 4b. the sibling integrators on the Cornell box (per-frame times; none of them has run on hardware yet)
timeout 300 python - > $o/r2_sibling_integrators.txt 2>&1 <<'PY'
from rs_pbrt_b200 import scenes, GpuScene
for integ, tex in (("path", None), (("ao", 16, True), None), (("direct", "all"), None), (("direct", "one"), None), ("whitted", None), ("path", "ewa+float+graph+bump"),
                   ("whitted", "ewa+bump")):
    h = scenes.cornell_box(xres=512, yres=512, spp=16, materials="mixed", integrator=integ, textures=tex)
    g = GpuScene(h.desc, 0)
    g.render(h.params)
    _, st = g.render(h.params)
    print(integ, tex, "ms_total %.1f trace %.1f shade %.1f rays %d Mrays/s %.0f launches %d" % (st["ms_total"], st["ms_trace"], st["ms_shade"], st["rays"], st["rays"] / st["ms_total"] / 1e3, st["kernel_launches"]))
    g.close()
PY
# 5. launch list of one textured frame and of one landscape frame (who costs what)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2_launches_landscape.csv python bench.py --workload landscape --small --steps 1 --warmup 1 --no-cpu > $o/r2_ncu_landscape.log 2>&1
cat $o/r2_summary.txt
