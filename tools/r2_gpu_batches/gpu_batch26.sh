#!/bin/bash
# round-2 batch 26 (1 GPU): phase 3 with ONE [136x1x3] t_1 box per level (3 TMA operations per job instead of 5): parity + A/B
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
L=$PWD/wrf_model_cuda_sample_b200
WRFB200_LIB=$L/libwrfb200_box3.so timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 300 > gpurun_out/b26_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b26_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b26_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b26_ab.log 2>&1
}
for rep in 1 2; do
for wl in conus3 patch8; do
  run "$wl five-copies rep$rep" timeout 200 $B --workload $wl
  run "$wl box3 rep$rep" WRFB200_LIB=$L/libwrfb200_box3.so timeout 200 $B --workload $wl
done
done
for wl in conus12 deep120; do
  run "$wl five-copies" timeout 200 $B --workload $wl
  run "$wl box3" WRFB200_LIB=$L/libwrfb200_box3.so timeout 200 $B --workload $wl
done
