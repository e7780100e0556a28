#!/bin/bash
# round-2 batch 17 (8 GPUs): the driver's command at N = 8 and N = 2, every key of the bench line
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/b17_bench_n8.json 2> gpurun_out/b17_bench_n8.err
echo "rc=$?" >> gpurun_out/b17_bench_n8.err
timeout 120 $TR --nproc-per-node 8 --master-port 29515 bench.py --impl reference --gpus 8 --steps 5 --warmup 2 > gpurun_out/b17_bench_ref_n8.json 2> gpurun_out/b17_bench_ref_n8.err
timeout 600 $TR --nproc-per-node 2 --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b17_bench_n2.json 2> gpurun_out/b17_bench_n2.err
echo "rc=$?" >> gpurun_out/b17_bench_n2.err
