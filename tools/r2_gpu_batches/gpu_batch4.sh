#!/bin/bash
# round-2 batch 4 (2 GPUs): fused halo exchange between two PROCESSES on two devices (CUDA IPC over NVLink)
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/b4_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for pg in 1x2 2x1; do
  timeout 300 $TR --master-port 29511 tools/multi_gpu_check.py --pgrid $pg > gpurun_out/b4_check_fused_$pg.log 2>&1
  echo "rc=$?" >> gpurun_out/b4_check_fused_$pg.log
done
timeout 300 $TR --master-port 29512 tools/multi_gpu_check.py --pgrid 1x2 --mode nccl > gpurun_out/b4_check_nccl_1x2.log 2>&1
echo "rc=$?" >> gpurun_out/b4_check_nccl_1x2.log
timeout 300 $TR --master-port 29513 tools/multi_gpu_check.py --pgrid 1x2 --shape 1800x266x50 --steps 6 > gpurun_out/b4_check_fused_big.log 2>&1
echo "rc=$?" >> gpurun_out/b4_check_fused_big.log
timeout 600 python -m pytest tests/test_c_harness.py tests/test_multi_gpu.py -q -m gpu --timeout 500 > gpurun_out/b4_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b4_tests.log
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b4_bench_n2.json 2> gpurun_out/b4_bench_n2.err
echo "rc=$?" >> gpurun_out/b4_bench_n2.err
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --pgrid 2x1 --no-extras --no-e2e > gpurun_out/b4_bench_n2_2x1.json 2> gpurun_out/b4_bench_n2_2x1.err
echo "rc=$?" >> gpurun_out/b4_bench_n2_2x1.err
