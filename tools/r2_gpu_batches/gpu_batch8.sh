#!/bin/bash
# round-2 batch 8 (1 GPU): launch-shape policy for a rank's patch at 8 / 4 GPUs: strips x tail split
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b8_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b8_ab.log 2>&1
}
for wl in patch8 patch4; do
for st in 0 2; do
  for tail in 0 1 5 7 9 15 23; do
    run "$wl strip=$st tail=$tail" WRFB200_PIPE_STRIP=$st WRFB200_PIPE_TAIL=$tail timeout 200 $B --workload $wl
  done
done
done
for cfg in 12 13; do
  run "patch8 strip=2 cfg=$cfg" WRFB200_PIPE_STRIP=2 WRFB200_PIPE_CFG=$cfg timeout 200 $B --workload patch8
done
