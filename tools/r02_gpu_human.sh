#!/bin/bash
# after a change to the warp / hybrid drivers: stage + render + full-size parity, then cfg3-5 timing per packet shape
mkdir -p gpurun_out
T=${1:-h1}
timeout 240 python -m pytest tests/test_gpu_stages.py tests/test_gpu_render.py -x -q -m gpu > gpurun_out/r02_${T}_tests.log 2>&1; echo "stages+render rc=$?"
tail -3 gpurun_out/r02_${T}_tests.log
timeout 240 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/r02_${T}_fullsize.log 2>&1; echo "fullsize+dropin rc=$?"
tail -3 gpurun_out/r02_${T}_fullsize.log
for lg in 2 0 1 3; do
  NEUMAN_WARP_PACKET=$lg timeout 120 python tools/human_bench.py cfg3 cfg4 cfg5 > gpurun_out/r02_${T}_human_lg$lg.json 2> gpurun_out/r02_${T}_human_lg$lg.err; echo "human lg=$lg rc=$?"
  python - <<PY
import json
try:
    r = json.load(open("gpurun_out/r02_${T}_human_lg$lg.json"))
    print({k: (round(v["ms"], 2), round(v["mlp_ms"], 2), round(1 - v["mlp_ms"] / v["ms"], 4)) for k, v in r.items()})
except Exception as e:
    print("no result", e)
PY
done
