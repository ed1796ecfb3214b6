#!/bin/bash
# full validation of the round-2 state: every GPU test file (each bounded), the bench line, ncu launch list + full capture, sanitizer
mkdir -p gpurun_out
T=${1:-f2}
for f in test_gpu_stages test_gpu_fullsize test_gpu_dropin test_oracle_golden; do
  timeout 420 python -m pytest tests/$f.py -q -m "gpu or not gpu" > gpurun_out/r02_${T}_$f.log 2>&1; echo "$f rc=$?"
  tail -4 gpurun_out/r02_${T}_$f.log | cut -c1-400
done
timeout 420 python bench.py > gpurun_out/r02_${T}_bench.json 2> gpurun_out/r02_${T}_bench.err; echo "bench rc=$?"
cut -c1-7000 gpurun_out/r02_${T}_bench.json; tail -3 gpurun_out/r02_${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_${T}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 120 -c 2 -f -o gpurun_out/r02_mlp_after python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
