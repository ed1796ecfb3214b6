#!/bin/bash
# final gate of the session: the whole GPU suite, then smoke()
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/r02_final2_suite.log 2>&1; echo "suite rc=$?"
tail -4 gpurun_out/r02_final2_suite.log | cut -c1-200
grep -n "^E  \|^FAILED" gpurun_out/r02_final2_suite.log | head -10 | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02_final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_final2_smoke.log | cut -c1-300
