#!/bin/bash
# Round 2, GPU call 21 (1 GPU): textured frames on ONE stream (call 20: 1 565.6 ms single-stream vs 1 832.0 with two batches in flight) -- the new default
# against two streams and against smaller batches; the textured GPU tests on this build; the bench line of the textured workload.
mkdir -p gpurun_out
o=gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity_siblings.py tests/test_gpu_parity_materials.py -q -m gpu -x -k "texture or bump or mix or alpha or two_batches" > $o/c21_pytest.log 2>&1; echo "pytest textured: exit $?" | tee $o/c21_summary.txt
tail -2 $o/c21_pytest.log >> $o/c21_summary.txt
V="default,default@PB_STREAMS=2,default@PB_BATCH_LOG2=21,default@PB_BATCH_LOG2=20"
timeout 150 python tools/exp_bench.py --scenes cornell-textured --libs "$V" --out $o/c21_textured.jsonl > $o/c21_textured.log 2>&1; echo "textured: exit $?" >> $o/c21_summary.txt
timeout 150 python bench.py --workload cornell-textured --steps 3 --warmup 3 --no-cpu --no-extra > $o/c21_bench_cornell-textured.json 2> $o/c21_bench_cornell-textured.err; echo "bench textured: exit $?" >> $o/c21_summary.txt
cat $o/c21_summary.txt
cut -c1-400 $o/c21_textured.jsonl
cut -c1-300 $o/c21_bench_cornell-textured.json
