#!/bin/bash
# Round 2, GPU call 17 (1 GPU): camera samples per batch (2^22 default vs 2^23 / 2^24: fewer latency-bound tail iterations per frame); one whole
# frame of the statue under ncu with DRAM / lane / issue metrics per launch (frame-level roofline record).
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python tools/exp_bench.py --scenes statue,cornell,conference --libs "default,default@PB_BATCH_LOG2=23,default@PB_BATCH_LOG2=24" --check --parts 8 --out $o/c17_batch.jsonl > $o/c17_batch.log 2>&1; echo "batch size: exit $?" | tee $o/c17_summary.txt
export PB_STREAMS=1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none -s 1300 -c 1216 --csv --log-file $o/c17_frame_statue.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $o/c17_ncu.log 2>&1; echo "ncu frame: exit $?" >> $o/c17_summary.txt
python tools/ncu_frame_summary.py $o/c17_frame_statue.csv $o statue "statue frame, round-2 final kernels, PB_STREAMS=1" >> $o/c17_summary.txt 2>&1
gzip -f $o/c17_frame_statue.csv
rm -f $o/c17_ncu.log
cat $o/c17_summary.txt
cut -c1-420 $o/c17_batch.jsonl
