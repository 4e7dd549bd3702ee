#!/bin/bash
# Round 2, GPU call 20 (1 GPU): the textured Cornell frame went 1 388 -> 1 979 ms between call 1 and call 18 -- batch size (2^22 -> 2^24 camera
# samples: 8 GB of per-hit material records per context), two batches in flight, or the kernels?  Plus the GPU test added after call 18.
mkdir -p gpurun_out
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_materials.py -q -m gpu -x -k "two_batches or translucent" > $o/c20_pytest.log 2>&1; echo "pytest: exit $?" | tee $o/c20_summary.txt
tail -2 $o/c20_pytest.log >> $o/c20_summary.txt
V="default,default@PB_STREAMS=2,default@PB_BATCH_LOG2=22,default@PB_BATCH_LOG2=22+PB_STREAMS=2,default@PB_BATCH_LOG2=23,variants/lib_c9.so"
timeout 900 python tools/exp_bench.py --scenes cornell-textured --libs "$V" --out $o/c20_textured.jsonl > $o/c20_textured.log 2>&1; echo "textured: exit $?" >> $o/c20_summary.txt
timeout 600 python tools/exp_bench.py --scenes cornell,statue,conference,landscape-64 --libs "default,default@PB_STREAMS=2" --out $o/c20_streams.jsonl > $o/c20_streams.log 2>&1; echo "streams: exit $?" >> $o/c20_summary.txt
cat $o/c20_summary.txt
cut -c1-300 $o/c20_textured.jsonl $o/c20_streams.jsonl
