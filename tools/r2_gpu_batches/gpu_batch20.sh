#!/bin/bash
# round-2 batch 20 (1 GPU): one code body for all tile blocks on small launches (instruction-cache pressure) A/B
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b20_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b20_ab.log 2>&1
}
for rep in 1 2; do
for wl in conus12 patch8 patch4 tiny; do
  run "$wl two-bodies rep$rep" timeout 200 $B --workload $wl
  run "$wl one-body rep$rep" WRFB200_PIPE_ONE_BODY=1 timeout 200 $B --workload $wl
done
done
run "conus12 one-body tail=2" WRFB200_PIPE_ONE_BODY=1 WRFB200_PIPE_TAIL=2 timeout 200 $B --workload conus12
run "patch8 one-body tail=0" WRFB200_PIPE_ONE_BODY=1 WRFB200_PIPE_TAIL=0 timeout 200 $B --workload patch8
run "conus3 one-body" WRFB200_PIPE_ONE_BODY=1 timeout 200 $B --workload conus3
WRFB200_PIPE_ONE_BODY=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden or ragged or tutorial or conus12" --timeout 500 > gpurun_out/b20_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b20_tests.log
