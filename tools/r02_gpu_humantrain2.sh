#!/bin/bash
# second pass after the human-trainer kernels: changed tests first, stage timings, the full bench line, then the rest of the suite
mkdir -p gpurun_out
T=${1:-ht2}
timeout 240 python -m pytest tests/test_gpu_human_train.py tests/test_gpu_dropin.py tests/test_gpu_train.py tests/test_gpu_stages.py -q -m gpu > gpurun_out/r02_${T}_changed.log 2>&1; echo "changed tests rc=$?"
tail -30 gpurun_out/r02_${T}_changed.log | cut -c1-260
timeout 90 python tools/human_train_bench.py > gpurun_out/r02_${T}_bench.json 2> gpurun_out/r02_${T}_bench.err; echo "human bench rc=$?"
cat gpurun_out/r02_${T}_bench.json; tail -4 gpurun_out/r02_${T}_bench.err | cut -c1-300
timeout 330 python bench.py > gpurun_out/r02_${T}_benchline.json 2> gpurun_out/r02_${T}_benchline.err; echo "bench.py rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_${T}_benchline.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"])
    print("train_step", d.get("train_step")); print("human_train_step", d.get("human_train_step"))
    print("cpu_baseline", {k: d["cpu_baseline"][k] for k in ("value", "cores", "kind")} if "cpu_baseline" in d else None)
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r02_${T}_benchline.err | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_render.py tests/test_gpu_fullsize.py -q -m gpu > gpurun_out/r02_${T}_rest.log 2>&1; echo "rest of the suite rc=$?"
tail -6 gpurun_out/r02_${T}_rest.log | cut -c1-260
