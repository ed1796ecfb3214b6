#!/bin/bash
# first GPU session: diagnostics in isolation (each under its own timeout), then the test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/tc_check.py simt 1048576 > gpurun_out/diag_simt.log 2>&1; echo "simt rc=$?" >> gpurun_out/diag_simt.log
NEUMAN_TC_PAIR=1 timeout 300 python tools/tc_check.py tc > gpurun_out/diag_tc1.log 2>&1; echo "tc1 rc=$?" >> gpurun_out/diag_tc1.log
NEUMAN_TC_PAIR=2 timeout 300 python tools/tc_check.py tc > gpurun_out/diag_tc2.log 2>&1; echo "tc2 rc=$?" >> gpurun_out/diag_tc2.log
timeout 600 python -m pytest tests/test_gpu_stages.py -q --timeout 200 -p no:cacheprovider > gpurun_out/t_stages.log 2>&1
NEUMAN_MLP_MODE=simt timeout 900 python -m pytest tests/test_gpu_render.py -q --timeout 400 -p no:cacheprovider > gpurun_out/t_render_simt.log 2>&1
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_render.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t_tc.log 2>&1
tail -5 gpurun_out/diag_simt.log gpurun_out/diag_tc1.log gpurun_out/diag_tc2.log
tail -3 gpurun_out/t_stages.log gpurun_out/t_render_simt.log gpurun_out/t_tc.log
