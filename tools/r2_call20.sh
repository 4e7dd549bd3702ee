#!/bin/bash
# Round 2, GPU call 20 (1 GPU, the last 8 GPU-minutes of the round): the build with MixMaterial (DLobe::has_sc) on hardware --
# the new parity tests first, then the whole -m gpu suite without the full-size file, then the textured Cornell frame at 2^22 (the new default for textured scenes) / 2^23 /
# 2^24 camera samples per batch (it went 1 388 -> 1 979 ms between call 1 and call 18: 8 GB of per-hit material records per context at
# 2^24), then a statue frame as a regression check of the resident rate.  Every step writes its result at once.
mkdir -p gpurun_out
o=gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity_materials.py -q -m gpu -x -k "mix_material" > $o/c20_pytest_mix.log 2>&1; echo "pytest mix: exit $?" | tee $o/c20_summary.txt
tail -2 $o/c20_pytest_mix.log >> $o/c20_summary.txt
timeout 240 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_full_configs.py --deselect tests/test_gpu_multi.py > $o/c20_pytest.log 2>&1; echo "pytest -m gpu (without full-size / multi files): exit $?" >> $o/c20_summary.txt
tail -2 $o/c20_pytest.log >> $o/c20_summary.txt
V="default,default@PB_BATCH_LOG2=23,default@PB_BATCH_LOG2=24"
timeout 150 python tools/exp_bench.py --scenes cornell-textured --libs "$V" --out $o/c20_textured.jsonl > $o/c20_textured.log 2>&1; echo "textured: exit $?" >> $o/c20_summary.txt
timeout 150 python tools/exp_bench.py --scenes statue --libs "default" --out $o/c20_statue.jsonl > $o/c20_statue.log 2>&1; echo "statue: exit $?" >> $o/c20_summary.txt
cat $o/c20_summary.txt
cut -c1-400 $o/c20_textured.jsonl $o/c20_statue.jsonl
