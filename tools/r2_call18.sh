#!/bin/bash
# Round 2, GPU call 18 (1 GPU): the suite and the 1-GPU bench lines of every workload on the final build (2^24 camera samples per batch);
# 2^23 / 2^25 beside it; one whole statue frame under ncu with DRAM / lane / issue metrics per launch.
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/c18_pytest.log 2>&1; echo "pytest -m gpu: exit $?" | tee $o/c18_summary.txt
tail -2 $o/c18_pytest.log >> $o/c18_summary.txt
timeout 900 python tools/exp_bench.py --scenes statue,cornell --libs "default,default@PB_BATCH_LOG2=25,default@PB_BATCH_LOG2=23" --check --parts 8 --out $o/c18_batch.jsonl > $o/c18_batch.log 2>&1; echo "batch size: exit $?" >> $o/c18_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $o/c18_bench_statue.json 2> $o/c18_bench_statue.err; echo "bench statue: exit $?" >> $o/c18_summary.txt
for w in cornell conference landscape-64 cornell-direct cornell-whitted cornell-ao cornell-textured; do
  timeout 500 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-extra > $o/c18_bench_$w.json 2> $o/c18_bench_$w.err; echo "bench $w: exit $?" >> $o/c18_summary.txt
done
export PB_STREAMS=1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none -s 400 -c 304 --csv --log-file $o/c18_frame_statue.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c18_ncu.log 2>&1; echo "ncu frame: exit $?" >> $o/c18_summary.txt
python tools/ncu_frame_summary.py $o/c18_frame_statue.csv $o statue "statue frame (304 launches), round-2 final kernels, 2^24 camera samples per batch, PB_STREAMS=1" >> $o/c18_summary.txt 2>&1
gzip -f $o/c18_frame_statue.csv
rm -f $o/c18_ncu.log
cat $o/c18_summary.txt
cut -c1-420 $o/c18_batch.jsonl
for w in statue cornell conference landscape-64 cornell-direct cornell-whitted cornell-ao cornell-textured; do python - "$o/c18_bench_$w.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]), d["kernel_ms_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
