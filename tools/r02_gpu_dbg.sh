#!/bin/bash
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "empty_and_scale" > gpurun_out/r02_dbg_train.log 2>&1; echo "train rc=$?"
grep -B2 -A14 "Invalid\|Error\|trap\|Illegal" gpurun_out/r02_dbg_train.log | head -80
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_dbg_bench.log 2>&1; echo "bench rc=$?"
grep -B2 -A14 "Invalid\|Error\|trap\|Illegal" gpurun_out/r02_dbg_bench.log | head -80
