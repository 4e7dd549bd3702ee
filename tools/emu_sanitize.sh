#!/bin/bash
# AddressSanitizer, UndefinedBehaviorSanitizer and ThreadSanitizer over the kernels' source, run on host threads by the emulation harness (tests/emu):
# out-of-bounds accesses and intra-block data races of every kernel path, without a GPU.  (Blocks run one at a time, so races
# BETWEEN blocks are invisible to it.)   usage: tools/emu_sanitize.sh > profiles/rNN_emu_sanitizers.txt
cd "$(dirname "$0")/.."
export PB_EMU_DEVICES=3  # the multi-device render at the end of tests/emu/sanitize_scenes.py
python tests/emu/build_emu.py --asan > /dev/null 2>&1 || exit 1
python tests/emu/build_emu.py --tsan > /dev/null 2>&1 || exit 1
python tests/emu/build_emu.py --ubsan > /dev/null 2>&1 || exit 1
for mode in 0 1 2; do
  echo "== AddressSanitizer, PB_RAY_SORT=$mode"
  PB_RAY_SORT=$mode LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tests/emu/sanitize_scenes.py asan 2>&1 | grep -E "ERROR|SUMMARY|done|^[a-z+-]+ [0-9]+$"
done
echo "== UndefinedBehaviorSanitizer (+ float-cast-overflow, no recovery), PB_RAY_SORT=0"
LD_PRELOAD=$(gcc -print-file-name=libubsan.so) UBSAN_OPTIONS=print_stacktrace=1 python tests/emu/sanitize_scenes.py usan 2>&1 | grep -E "runtime error|done|^[a-z+-]+ [0-9]+$"
echo "== ThreadSanitizer, PB_RAY_SORT=0"
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="report_signal_unsafe=0 halt_on_error=0" python tests/emu/sanitize_scenes.py tsan > /tmp/emu_tsan.log 2>&1
grep -E "^[a-z+-]+ [0-9]+$|done|ThreadSanitizer: reported" /tmp/emu_tsan.log
grep -A3 "WARNING: ThreadSanitizer" /tmp/emu_tsan.log | grep "#0" | sed 's/(librs.*//' | sort | uniq -c
