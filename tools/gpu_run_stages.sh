#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k 'regex:k_(raygen|near_far|ray_to_samples|importance|raw2outputs|merge|warp_points|warp_dirs|bvh_refit|bvh_hierarchy|move_rows|compact_hits)' -c 26 -o gpurun_out/prof_stages python tools/human_bench.py cfg4 > gpurun_out/ncu_stages.log 2>&1
tail -2 gpurun_out/ncu_stages.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
