#!/bin/bash
# fifth pass: split signed-distance query + device-side bounds, the loss_func drop-in test; timings again
mkdir -p gpurun_out
T=${1:-ht5}
timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_human_train.py tests/test_gpu_dropin.py tests/test_gpu_render.py -q -m gpu > gpurun_out/r02_${T}_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02_${T}_tests.log | cut -c1-260
grep -n "^E  " gpurun_out/r02_${T}_tests.log | head -14 | cut -c1-300
timeout 90 python tools/human_train_bench.py > gpurun_out/r02_${T}_bench.json 2> gpurun_out/r02_${T}_bench.err; echo "human bench rc=$?"
cat gpurun_out/r02_${T}_bench.json; tail -3 gpurun_out/r02_${T}_bench.err | cut -c1-300
