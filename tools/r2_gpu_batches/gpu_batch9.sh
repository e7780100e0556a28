#!/bin/bash
# round-2 batch 9 (2 GPUs): k-parallel strip blocks: parity suite, shape sweep, 2-rank runs
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/b9_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b9_gpu_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b9_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'], d['roofline']['kernel'])
" >> gpurun_out/b9_ab.log 2>&1
}
for wl in conus3 patch8 patch4; do
  for st in 0 1; do
    for tail in 0 1; do
      run "$wl strip=$st tail=$tail" WRFB200_PIPE_STRIP=$st WRFB200_PIPE_TAIL=$tail timeout 300 $B --workload $wl
    done
  done
done
run "patch8 strip=1 tail=15" WRFB200_PIPE_STRIP=1 WRFB200_PIPE_TAIL=15 timeout 300 $B --workload patch8
run "patch8 strip=1 tail=7" WRFB200_PIPE_STRIP=1 WRFB200_PIPE_TAIL=7 timeout 300 $B --workload patch8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for pg in 1x2 2x1; do
  timeout 300 $TR --master-port 29511 tools/multi_gpu_check.py --pgrid $pg > gpurun_out/b9_check_fused_$pg.log 2>&1
  echo "rc=$?" >> gpurun_out/b9_check_fused_$pg.log
  timeout 300 $TR --master-port 29513 tools/multi_gpu_check.py --pgrid $pg --shape 1800x266x50 --steps 6 > gpurun_out/b9_check_fused_big_$pg.log 2>&1
  echo "rc=$?" >> gpurun_out/b9_check_fused_big_$pg.log
done
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-e2e > gpurun_out/b9_bench_n2.json 2> gpurun_out/b9_bench_n2.err
echo "rc=$?" >> gpurun_out/b9_bench_n2.err
