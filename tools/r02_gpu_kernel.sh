#!/bin/bash
# kernel iteration: MLP tests, numerics check + throughput, timeline, bench
mkdir -p gpurun_out
T=${1:-k1}
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_render.py -x -q -m gpu > gpurun_out/r02_${T}_tests.log 2>&1; echo "tests rc=$?"
timeout 300 python tools/tc_check.py tc > gpurun_out/r02_${T}_tc_check.log 2>&1; echo "tc_check rc=$?"
timeout 300 python tools/tc_trace.py inference > gpurun_out/r02_${T}_trace.log 2>&1; echo "trace rc=$?"
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/r02_${T}_tests2.log 2>&1; echo "tests2 rc=$?"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_${T}_bench.json 2> gpurun_out/r02_${T}_bench.err; echo "bench rc=$?"
tail -15 gpurun_out/r02_${T}_tests.log; tail -8 gpurun_out/r02_${T}_tc_check.log; grep -A1 "step period\|commit -> leader\|seen -> drained" gpurun_out/r02_${T}_trace.log | cut -c1-330; tail -5 gpurun_out/r02_${T}_tests2.log; cat gpurun_out/r02_${T}_bench.json | cut -c1-1500
