#!/bin/bash
# round-2 batch 3 (1 GPU): fused-comm tests again (kernels preloaded), full suite with the mixed-tail kernel,
# tail-split A/B on the small shapes
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_comm.py -q -m gpu --timeout 300 > gpurun_out/b3_fused.log 2>&1
echo "fused rc=$?" >> gpurun_out/b3_fused.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_fused_comm.py > gpurun_out/b3_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b3_gpu_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
for wl in patch8 conus12 tiny; do
  for tail in 0 1 8 16 24 32 48; do
    echo "== $wl tail=$tail" >> gpurun_out/b3_sweep.log
    WRFB200_PIPE_TAIL=$tail timeout 120 $B --workload $wl 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b3_sweep.log 2>&1
  done
done
