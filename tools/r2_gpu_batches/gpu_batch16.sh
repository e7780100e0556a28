#!/bin/bash
# round-2 batch 16 (1 GPU): per-row named barriers around the scan: parity + A/B against the block-wide barrier
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_fused_comm.py -q -m gpu --timeout 600 > gpurun_out/b16_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b16_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b16_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b16_ab.log 2>&1
}
L=$PWD/wrf_model_cuda_sample_b200
for rep in 1 2 3; do
for wl in conus3 patch8 conus12 deep120; do
  run "$wl rowbar rep$rep" timeout 300 $B --workload $wl
  run "$wl blockbar rep$rep" WRFB200_LIB=$L/libwrfb200_norb.so timeout 300 $B --workload $wl
done
done
run "weak2048 rowbar" timeout 300 $B --workload weak2048
run "weak2048 blockbar" WRFB200_LIB=$L/libwrfb200_norb.so timeout 300 $B --workload weak2048
