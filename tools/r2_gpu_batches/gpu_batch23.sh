#!/bin/bash
# round-2 batch 23 (1 GPU): code-size diet (three instead of five copies of the phase-3 level body; scan unroll 2) A/B
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b23_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b23_ab.log 2>&1
}
L=$PWD/wrf_model_cuda_sample_b200
for wl in conus12 patch8 conus3 tiny; do
  run "$wl baseline" timeout 200 $B --workload $wl
  run "$wl compact" WRFB200_LIB=$L/libwrfb200_compact.so timeout 200 $B --workload $wl
  run "$wl compact+scan2" WRFB200_LIB=$L/libwrfb200_compact2.so timeout 200 $B --workload $wl
done
WRFB200_LIB=$L/libwrfb200_compact.so timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden or ragged or tutorial or nz120" --timeout 300 > gpurun_out/b23_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b23_tests.log
