#!/bin/bash
# round-2 batch 5 (1 GPU): hoisted-reciprocal division (self-test + full parity suite), A/B vs the plain build,
# deep-column 2-blocks/SM variant, signal-kernel protocol (ranks sharing the device)
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/b5_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b5_gpu_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { # label env...
  echo "== $1" >> gpurun_out/b5_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b5_ab.log 2>&1
}
for rep in 1 2; do
for wl in conus3 patch8 conus12 deep120 weak2048; do
  run "$wl hoist rep$rep" timeout 300 $B --workload $wl
  run "$wl nohoist rep$rep" WRFB200_LIB=$PWD/wrf_model_cuda_sample_b200/libwrfb200_nohoist.so timeout 300 $B --workload $wl
done
done
run "deep120 cfg14 (old auto)" WRFB200_PIPE_CFG=14 timeout 300 $B --workload deep120
run "deep120 cfg62" WRFB200_PIPE_CFG=62 timeout 300 $B --workload deep120
run "deep120 cfg12" WRFB200_PIPE_CFG=12 timeout 300 $B --workload deep120
