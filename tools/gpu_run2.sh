#!/bin/bash
# correctness + bench + profiles
mkdir -p gpurun_out
NEUMAN_TC_PAIR=2 timeout 300 python tools/tc_check.py tc > gpurun_out/diag_tc2.log 2>&1; echo "tc2 rc=$?" >> gpurun_out/diag_tc2.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
if [ "$1" == "prof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 8 -c 1 -o gpurun_out/prof_mlp python tools/tc_check.py tc > gpurun_out/ncu_mlp.log 2>&1
fi
tail -n 4 gpurun_out/diag_tc2.log; tail -n 3 gpurun_out/t_gpu.log; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
