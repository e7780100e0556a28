#!/bin/bash
# round-2 batch 15 (1 GPU): what the driver runs at round end: GPU test lane, smoke(), the default bench, the reference arm
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/b15_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b15_gpu_tests.log
timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/b15_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/b15_smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/b15_bench_ref.json 2> gpurun_out/b15_bench_ref.err
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b15_bench.json 2> gpurun_out/b15_bench.err
echo "bench rc=$?" >> gpurun_out/b15_bench.err
