#!/bin/bash
# quick gate after a kernel change: MLP + render + train tests, numerics/throughput, timeline -- every step tightly bounded
mkdir -p gpurun_out
T=${1:-q1}
timeout 180 python -m pytest tests/test_gpu_mlp.py -x -q -m gpu > gpurun_out/r02_${T}_mlp.log 2>&1; echo "mlp rc=$?"
tail -3 gpurun_out/r02_${T}_mlp.log
timeout 240 python -m pytest tests/test_gpu_render.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r02_${T}_tests.log 2>&1; echo "render+train rc=$?"
tail -3 gpurun_out/r02_${T}_tests.log
timeout 120 python tools/tc_check.py tc > gpurun_out/r02_${T}_tc_check.log 2>&1; echo "tc_check rc=$?"
tail -3 gpurun_out/r02_${T}_tc_check.log
timeout 120 python tools/tc_trace.py inference > gpurun_out/r02_${T}_trace.log 2>&1; echo "trace rc=$?"
grep -A1 "step period\|leader epilogue: accumulator seen" gpurun_out/r02_${T}_trace.log | cut -c1-360
NEUMAN_TC_RANGE=0 timeout 120 python tools/tc_trace.py inference > gpurun_out/r02_${T}_trace_norange.log 2>&1
grep -A1 "step period" gpurun_out/r02_${T}_trace_norange.log | cut -c1-360
timeout 300 python bench.py --no-configs --no-cpu-baseline > gpurun_out/r02_${T}_bench.json 2> gpurun_out/r02_${T}_bench.err; echo "bench rc=$?"
cut -c1-1800 gpurun_out/r02_${T}_bench.json; tail -3 gpurun_out/r02_${T}_bench.err
if ! grep -q '"value"' gpurun_out/r02_${T}_bench.json; then
  CUDA_LAUNCH_BLOCKING=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_${T}_bench_blocking.json 2> gpurun_out/r02_${T}_bench_blocking.err; echo "bench blocking rc=$?"
  tail -4 gpurun_out/r02_${T}_bench_blocking.err
fi
