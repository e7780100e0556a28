#!/bin/bash
# round-2 batch 12 (1 GPU): shorter strip blocks: parity + timing
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/b12_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b12_gpu_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b12_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b12_ab.log 2>&1
}
for rep in 1 2; do
for wl in conus3 patch8 patch4; do
  run "$wl strip=1 rep$rep" timeout 300 $B --workload $wl
  run "$wl strip=0 rep$rep" WRFB200_PIPE_STRIP=0 timeout 300 $B --workload $wl
done
done
