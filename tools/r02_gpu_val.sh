#!/bin/bash
# validation of the state before the multi-GPU runs: human-path tests, fullsize + dropin parity, bench with configs, cfg5 launch list
mkdir -p gpurun_out
T=${1:-v1}
for f in test_gpu_stages test_gpu_render test_gpu_fullsize test_gpu_dropin; do
  timeout 300 python -m pytest tests/$f.py -q -m "gpu" > gpurun_out/r02_${T}_$f.log 2>&1; echo "$f rc=$?"
  tail -4 gpurun_out/r02_${T}_$f.log | cut -c1-300
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_${T}_bench.json 2> gpurun_out/r02_${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02_'+__import__('sys').argv[1] if False else 'gpurun_out/r02_%s_bench.json' % __import__('os').environ.get('TT','v1')))
print({k:j[k] for k in ('value','ms_per_step')}, j['roofline']['frac'], j['e2e']['value'])
for k,v in j['configs'].items():
    print(k, round(v['ms_per_frame'],2), round(v['Mrays_s'],3), round(v['non_mlp_share'],3), round(v['mlp_frac_of_peak'],3))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_cfg5_launches.csv python tools/human_bench.py cfg5 > gpurun_out/r02_${T}_cfg5_ncu.log 2>&1; echo "cfg5 ncu rc=$?"
