#!/bin/bash
# round-2 batch 6 (2 GPUs): strip-folded launch + in-kernel v push + signal kernel: parity suite, A/B, 2-rank runs
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/b6_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b6_gpu_tests.log
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras --steps 20 --warmup 3"
run() { echo "== $1" >> gpurun_out/b6_ab.log; shift
  env "$@" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b6_ab.log 2>&1
}
for rep in 1 2; do
for wl in conus3 patch8; do
  run "$wl strip rep$rep" timeout 300 $B --workload $wl
  run "$wl nostrip rep$rep" WRFB200_PIPE_STRIP=0 timeout 300 $B --workload $wl
done
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for pg in 1x2 2x1; do
  timeout 300 $TR --master-port 29511 tools/multi_gpu_check.py --pgrid $pg > gpurun_out/b6_check_fused_$pg.log 2>&1
  echo "rc=$?" >> gpurun_out/b6_check_fused_$pg.log
done
timeout 300 $TR --master-port 29513 tools/multi_gpu_check.py --pgrid 1x2 --shape 1800x266x50 --steps 6 > gpurun_out/b6_check_fused_big.log 2>&1
echo "rc=$?" >> gpurun_out/b6_check_fused_big.log
timeout 300 $TR --master-port 29516 tools/multi_gpu_check.py --pgrid 2x1 --shape 1800x266x50 --steps 6 > gpurun_out/b6_check_fused_big_2x1.log 2>&1
echo "rc=$?" >> gpurun_out/b6_check_fused_big_2x1.log
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-e2e > gpurun_out/b6_bench_n2.json 2> gpurun_out/b6_bench_n2.err
echo "rc=$?" >> gpurun_out/b6_bench_n2.err
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --workload patch8 --no-extras --no-e2e > gpurun_out/b6_bench_n2_patch.json 2> gpurun_out/b6_bench_n2_patch.err
echo "rc=$?" >> gpurun_out/b6_bench_n2_patch.err
