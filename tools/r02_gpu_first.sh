#!/bin/bash
# round-2 baseline capture: tests, timeline, ncu source-level capture of the MLP kernel on the bench workload, sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
lscpu | head -25 >> gpurun_out/r02_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests0.log 2>&1; echo "tests rc=$?"
timeout 300 python tools/tc_trace.py inference > gpurun_out/r02_trace_before.log 2>&1; echo "trace rc=$?"
timeout 300 python tools/tc_check.py tc > gpurun_out/r02_tc_check_before.log 2>&1; echo "tc_check rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 120 -c 2 -f -o gpurun_out/r02_mlp_before python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_before.log 2>&1; echo "ncu rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mlp.py -x -q -m gpu > gpurun_out/r02_memcheck_mlp.log 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mlp.py -x -q -m gpu > gpurun_out/r02_racecheck_mlp.log 2>&1; echo "racecheck rc=$?"
timeout 300 python bench.py > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_tests0.log; tail -12 gpurun_out/r02_trace_before.log; tail -3 gpurun_out/r02_memcheck_mlp.log; tail -3 gpurun_out/r02_racecheck_mlp.log; cat gpurun_out/r02_bench0.json
