#!/bin/bash
mkdir -p gpurun_out
T=${1:-k3}
timeout 120 ./tools/tmem_bench > gpurun_out/r02_tmem_bench.log 2>&1; echo "tmem rc=$?"
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_render.py -x -q -m gpu > gpurun_out/r02_${T}_tests.log 2>&1; echo "tests rc=$?"
timeout 300 python tools/tc_check.py tc > gpurun_out/r02_${T}_tc_check.log 2>&1; echo "tc_check rc=$?"
timeout 300 python tools/tc_trace.py inference > gpurun_out/r02_${T}_trace.log 2>&1; echo "trace rc=$?"
NEUMAN_TC_RANGE=0 timeout 300 python tools/tc_trace.py inference > gpurun_out/r02_${T}_trace_norange.log 2>&1; echo "trace rc=$?"
NEUMAN_TC_RANGE=0 timeout 300 python tools/tc_check.py tc > gpurun_out/r02_${T}_tc_check_norange.log 2>&1
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r02_${T}_tests2.log 2>&1; echo "tests2 rc=$?"
cat gpurun_out/r02_tmem_bench.log
tail -4 gpurun_out/r02_${T}_tests.log; tail -3 gpurun_out/r02_${T}_tc_check.log; tail -3 gpurun_out/r02_${T}_tc_check_norange.log; grep -A1 "step period\|commit -> leader\|leader epilogue: accumulator seen" gpurun_out/r02_${T}_trace.log | cut -c1-360; echo NORANGE; grep -A1 "step period\|leader epilogue: accumulator seen" gpurun_out/r02_${T}_trace_norange.log | cut -c1-360; tail -5 gpurun_out/r02_${T}_tests2.log
