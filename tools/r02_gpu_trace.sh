#!/bin/bash
# timeline + throughput of the MLP kernel under the env settings given as arguments (VAR=value ...)
mkdir -p gpurun_out
T=$1; shift
for kv in "$@"; do export "$kv"; done
timeout 300 python tools/tc_trace.py inference > gpurun_out/r02_${T}_trace.log 2>&1; echo "trace rc=$?"
timeout 300 python tools/tc_check.py tc > gpurun_out/r02_${T}_tc_check.log 2>&1; echo "tc_check rc=$?"
tail -3 gpurun_out/r02_${T}_tc_check.log; grep -A1 "step period\|commit -> leader\|leader epilogue: accumulator seen" gpurun_out/r02_${T}_trace.log | cut -c1-330
