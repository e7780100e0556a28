#!/bin/bash
# round-2 batch 7 (2 GPUs): epoch/index flag protocol (no per-step signal kernel), batched column body, strips
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/b7_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b7_gpu_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b7_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'], d['roofline']['kernel'])
" >> gpurun_out/b7_ab.log 2>&1
}
for wl in conus3 patch8; do
  for st in 0 1 2; do
    run "$wl strip=$st" WRFB200_PIPE_STRIP=$st timeout 300 $B --workload $wl
  done
done
run "conus3 column kernel" timeout 300 $B --workload conus3 --kernel column
run "conus12 strip=2" WRFB200_PIPE_STRIP=2 timeout 300 $B --workload conus12
run "conus12 strip=0" WRFB200_PIPE_STRIP=0 timeout 300 $B --workload conus12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for pg in 1x2 2x1; do
  timeout 300 $TR --master-port 29511 tools/multi_gpu_check.py --pgrid $pg > gpurun_out/b7_check_fused_$pg.log 2>&1
  echo "rc=$?" >> gpurun_out/b7_check_fused_$pg.log
  WRFB200_PIPE_STRIP=2 timeout 300 $TR --master-port 29513 tools/multi_gpu_check.py --pgrid $pg --shape 1800x266x50 --steps 6 > gpurun_out/b7_check_fused_big_$pg.log 2>&1
  echo "rc=$?" >> gpurun_out/b7_check_fused_big_$pg.log
done
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-e2e > gpurun_out/b7_bench_n2.json 2> gpurun_out/b7_bench_n2.err
echo "rc=$?" >> gpurun_out/b7_bench_n2.err
