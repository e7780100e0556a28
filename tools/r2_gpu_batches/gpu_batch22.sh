#!/bin/bash
# round-2 batch 22 (1 GPU): phase-3 source order (horizontal fluxes before the first use of the stream registers) A/B + parity
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 500 > gpurun_out/b22_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b22_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b22_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b22_ab.log 2>&1
}
L=$PWD/wrf_model_cuda_sample_b200
for rep in 1 2 3; do
for wl in conus3 patch8; do
  run "$wl flux-first rep$rep" timeout 300 $B --workload $wl
  run "$wl old-order rep$rep" WRFB200_LIB=$L/libwrfb200_oldorder.so timeout 300 $B --workload $wl
done
done
for wl in conus12 deep120 weak2048; do
  run "$wl flux-first" timeout 300 $B --workload $wl
  run "$wl old-order" WRFB200_LIB=$L/libwrfb200_oldorder.so timeout 300 $B --workload $wl
done
