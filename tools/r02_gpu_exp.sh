#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/tmem_bench > gpurun_out/r02_tmem_bench.log 2>&1; echo "tmem rc=$?"
for e in 1 2; do
NEUMAN_TC_EXP=$e NEUMAN_TC_RANGE=0 timeout 300 python tools/tc_trace.py inference > gpurun_out/r02_exp${e}_trace.log 2>&1; echo "trace rc=$?"
done
cat gpurun_out/r02_tmem_bench.log
for e in 1 2; do echo EXP $e; grep -A1 "step period\|leader epilogue: accumulator seen\|commit -> leader" gpurun_out/r02_exp${e}_trace.log | cut -c1-360; done
