#!/bin/bash
# Round 2, GPU call 22 (1 GPU, the round's last GPU-minutes): the final build through the WHOLE -m gpu suite (full-size configs and the multi-device
# file included) and smoke(), as the driver will run them at round end.
mkdir -p gpurun_out
o=gpurun_out
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $o/c22_smoke.log 2>&1; echo "smoke: exit $?" | tee $o/c22_summary.txt
tail -1 $o/c22_smoke.log >> $o/c22_summary.txt
timeout 170 python -m pytest tests -q -m gpu -x --durations=12 > $o/c22_pytest.log 2>&1; echo "pytest -m gpu: exit $?" >> $o/c22_summary.txt
tail -18 $o/c22_pytest.log >> $o/c22_summary.txt
cat $o/c22_summary.txt
