#!/bin/bash
# round-2 batch 11 (1 GPU): L2 prefetch of the first phase-3 operands before the scan: A/B (0 / 4 / 16 levels)
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b11_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b11_ab.log 2>&1
}
L=$PWD/wrf_model_cuda_sample_b200
for rep in 1 2; do
for wl in conus3 patch8 patch4 conus12 deep120; do
  run "$wl pf0 rep$rep" timeout 300 $B --workload $wl
  run "$wl pf4 rep$rep" WRFB200_LIB=$L/libwrfb200_pf4.so timeout 300 $B --workload $wl
  run "$wl pf16 rep$rep" WRFB200_LIB=$L/libwrfb200_pf16.so timeout 300 $B --workload $wl
done
done
run "weak2048 pf0" timeout 300 $B --workload weak2048
run "weak2048 pf16" WRFB200_LIB=$L/libwrfb200_pf16.so timeout 300 $B --workload weak2048
WRFB200_LIB=$L/libwrfb200_pf16.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fused_comm.py -q -m gpu --timeout 600 > gpurun_out/b11_tests_pf16.log 2>&1
echo "rc=$?" >> gpurun_out/b11_tests_pf16.log
timeout 900 python -m pytest tests/test_fused_comm.py tests/test_c_harness.py -q -m gpu --timeout 600 > gpurun_out/b11_tests_fused.log 2>&1
echo "rc=$?" >> gpurun_out/b11_tests_fused.log
