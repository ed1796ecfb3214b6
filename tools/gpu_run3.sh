#!/bin/bash
mkdir -p gpurun_out
NEUMAN_TC_PAIR=2 timeout 300 python tools/tc_check.py tc > gpurun_out/diag_tc2.log 2>&1; echo "tc2 rc=$?" >> gpurun_out/diag_tc2.log
NEUMAN_TC_PAIR=1 timeout 300 python tools/tc_check.py tc > gpurun_out/diag_tc1.log 2>&1; echo "tc1 rc=$?" >> gpurun_out/diag_tc1.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 8 -c 1 -o gpurun_out/prof_mlp python tools/tc_check.py tc > gpurun_out/ncu_mlp.log 2>&1
tail -n 4 gpurun_out/diag_tc2.log; tail -n 4 gpurun_out/diag_tc1.log; tail -n 3 gpurun_out/t_gpu.log
