#!/bin/bash
# third pass: the two corrected tests, then the ncu launch list of one human-trainer step
mkdir -p gpurun_out
T=${1:-ht3}
timeout 150 python -m pytest tests/test_gpu_human_train.py tests/test_gpu_dropin.py -q -m gpu > gpurun_out/r02_${T}_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r02_${T}_tests.log | cut -c1-260
timeout 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_human_step_launches.csv python tools/human_train_step_once.py > gpurun_out/r02_${T}_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02_${T}_ncu.log | cut -c1-200
python tools/human_train_step_once.py --summarize gpurun_out/r02_human_step_launches.csv | head -40
