#!/bin/bash
mkdir -p gpurun_out
NEUMAN_DEBUG=1 timeout 900 python tools/human_bench.py cfg3 cfg4 cfg5 --check > gpurun_out/human.json 2> gpurun_out/human.err; echo "rc=$?" >> gpurun_out/human.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_cfg5.csv python tools/human_bench.py cfg5 > gpurun_out/human_ncu.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
cat gpurun_out/human.json; tail -n 8 gpurun_out/human.err; tail -n 3 gpurun_out/t_gpu.log
