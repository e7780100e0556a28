#!/bin/bash
# round-2 batch 19 (1 GPU): final sanity of the committed tree (GPU test lane, smoke) + the missing conus12 ncu capture
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/b19_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b19_gpu_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/b19_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/b19_smoke.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:amt_pipe -s 2 -c 1 -f -o gpurun_out/r2_pipe_conus12 $B --steps 1 --warmup 3 --workload conus12 > gpurun_out/b19_ncu_conus12.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:amt_pipe -s 4 -c 1 -f -o gpurun_out/r2_pipe_patch8 $B --steps 1 --warmup 3 --workload patch8 > gpurun_out/b19_ncu_patch8.log 2>&1
