#!/bin/bash
# after the human-trainer kernels: their parity tests, the files they touch, then the stage timings -- every step bounded
mkdir -p gpurun_out
T=${1:-ht1}
timeout 200 python -m pytest tests/test_gpu_human_train.py -q -m gpu > gpurun_out/r02_${T}_human_train.log 2>&1; echo "human_train rc=$?"
tail -25 gpurun_out/r02_${T}_human_train.log | cut -c1-300
timeout 280 python -m pytest tests/test_gpu_stages.py tests/test_gpu_dropin.py -q -m gpu > gpurun_out/r02_${T}_stages_dropin.log 2>&1; echo "stages+dropin rc=$?"
tail -25 gpurun_out/r02_${T}_stages_dropin.log | cut -c1-300
timeout 200 python -m pytest tests/test_gpu_train.py -q -m gpu -k "batch or train_step" > gpurun_out/r02_${T}_train.log 2>&1; echo "train rc=$?"
tail -12 gpurun_out/r02_${T}_train.log | cut -c1-300
timeout 120 python tools/human_train_bench.py > gpurun_out/r02_${T}_bench.json 2> gpurun_out/r02_${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/r02_${T}_bench.json; tail -5 gpurun_out/r02_${T}_bench.err | cut -c1-300
