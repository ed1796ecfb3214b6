#!/bin/bash
# full validation of the round-2 state: every GPU test, the bench line (+ range-off comparison), ncu launch list + full capture
mkdir -p gpurun_out
T=${1:-f1}
timeout 1500 python -m pytest tests -m gpu -q -k "${2:-not zzz}" > gpurun_out/r02_${T}_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py > gpurun_out/r02_${T}_bench.json 2> gpurun_out/r02_${T}_bench.err; echo "bench rc=$?"
NEUMAN_TC_RANGE=0 timeout 300 python bench.py --no-configs --no-cpu-baseline > gpurun_out/r02_${T}_bench_norange.json 2> gpurun_out/r02_${T}_bench_norange.err; echo "bench norange rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_${T}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 120 -c 2 -f -o gpurun_out/r02_mlp_after python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -25 gpurun_out/r02_${T}_tests.log | cut -c1-300; cat gpurun_out/r02_${T}_bench.json | cut -c1-6000; tail -3 gpurun_out/r02_${T}_bench.err; cat gpurun_out/r02_${T}_bench_norange.json | cut -c1-900
