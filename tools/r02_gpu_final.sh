#!/bin/bash
# final single-GPU records: smoke, the full bench line (CPU arm, configs, train step), reference arm, sanitizer passes
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log | cut -c1-400
timeout 420 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_final_bench.err
cut -c1-600 gpurun_out/r02_final_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference arm rc=$?"
cut -c1-900 gpurun_out/r02_final_bench_reference.json
timeout 240 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_train.py -x -q -m gpu -k "not additivity" > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_memcheck.log | cut -c1-200
timeout 300 compute-sanitizer --tool racecheck --print-limit 400 python -m pytest tests/test_gpu_mlp.py -x -q -m gpu > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_racecheck.log | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_train.py -q -m gpu > gpurun_out/r02_final_tests_mlp_train.log 2>&1; echo "mlp+train rc=$?"; tail -2 gpurun_out/r02_final_tests_mlp_train.log
