#!/bin/bash
# bench.py on N GPUs of one box (ray-sharded frame, one all_gather), as the driver launches it
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -$N
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r02_bench_${N}gpu.json; tail -5 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
j=[json.loads(l) for l in open('gpurun_out/r02_bench_${N}gpu.json') if l.startswith('{')][-1]
print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'mlp_share', j['roofline']['mlp_share_of_step'])
for k,v in j['configs'].items():
    print(k, round(v['ms_per_frame'],2), round(v['Mrays_s'],3), 'per-rank render ms', [round(x,1) for x in v['per_rank_render_ms']], 'hits', v['per_rank_hit_rays'])
PY
