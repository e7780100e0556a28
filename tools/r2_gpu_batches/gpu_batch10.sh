#!/bin/bash
# round-2 batch 10 (8 GPUs): parity at 8 ranks (1x8, 2x4; BASELINE size), the driver's bench command at N=8, 2x4, N=4
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/b10_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tools/multi_gpu_check.py --pgrid 1x8 > gpurun_out/b10_check_1x8.log 2>&1
echo "rc=$?" >> gpurun_out/b10_check_1x8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29512 tools/multi_gpu_check.py --pgrid 2x4 > gpurun_out/b10_check_2x4.log 2>&1
echo "rc=$?" >> gpurun_out/b10_check_2x4.log
timeout 300 $TR --nproc-per-node 8 --master-port 29518 tools/multi_gpu_check.py --pgrid 4x2 --shape 1800x266x50 --steps 3 > gpurun_out/b10_check_4x2_wide.log 2>&1
echo "rc=$?" >> gpurun_out/b10_check_4x2_wide.log
timeout 600 $TR --nproc-per-node 8 --master-port 29513 tools/multi_gpu_check.py --pgrid 1x8 --shape 1800x1060x50 --steps 6 > gpurun_out/b10_check_1x8_conus3.log 2>&1
echo "rc=$?" >> gpurun_out/b10_check_1x8_conus3.log
nvidia-smi nvlink -gt d -i 0 > gpurun_out/b10_nvlink_before.txt 2>&1
timeout 900 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/b10_bench_n8.json 2> gpurun_out/b10_bench_n8.err
echo "rc=$?" >> gpurun_out/b10_bench_n8.err
nvidia-smi nvlink -gt d -i 0 > gpurun_out/b10_nvlink_after.txt 2>&1
timeout 600 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 5 --pgrid 2x4 --no-extras --no-e2e > gpurun_out/b10_bench_n8_2x4.json 2> gpurun_out/b10_bench_n8_2x4.err
echo "rc=$?" >> gpurun_out/b10_bench_n8_2x4.err
timeout 600 $TR --nproc-per-node 4 --master-port 29516 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras --no-e2e > gpurun_out/b10_bench_n4.json 2> gpurun_out/b10_bench_n4.err
echo "rc=$?" >> gpurun_out/b10_bench_n4.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu --no-ref-cuda > gpurun_out/b10_bench_n1.json 2> gpurun_out/b10_bench_n1.err
timeout 600 $TR --nproc-per-node 8 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-e2e --no-graph > gpurun_out/b10_bench_n8_eager.json 2> gpurun_out/b10_bench_n8_eager.err
