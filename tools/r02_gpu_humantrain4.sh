#!/bin/bash
# fourth pass: the two corrected tests; if green, the whole GPU suite and smoke()
mkdir -p gpurun_out
T=${1:-ht4}
timeout 150 python -m pytest tests/test_gpu_human_train.py tests/test_gpu_dropin.py -q -m gpu > gpurun_out/r02_${T}_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -8 gpurun_out/r02_${T}_tests.log | cut -c1-260
grep -n "^E  " gpurun_out/r02_${T}_tests.log | head -12 | cut -c1-300
if [ $rc -eq 0 ]; then
  timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r02_${T}_suite.log 2>&1; echo "suite rc=$?"
  tail -5 gpurun_out/r02_${T}_suite.log | cut -c1-260
  timeout 150 python __graft_entry__.py smoke > gpurun_out/r02_${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_${T}_smoke.log | cut -c1-400
fi
