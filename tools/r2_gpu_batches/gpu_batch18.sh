#!/bin/bash
# round-2 batch 18 (8 GPUs): 2-D process grids with the final build
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
BA="--steps 20 --warmup 5 --no-extras --no-e2e"
timeout 300 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 $BA --pgrid 2x4 > gpurun_out/b18_n8_2x4.json 2> gpurun_out/b18_n8_2x4.err
timeout 300 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 $BA > gpurun_out/b18_n8_1x8.json 2> gpurun_out/b18_n8_1x8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29516 bench.py --gpus 4 $BA --pgrid 2x2 > gpurun_out/b18_n4_2x2.json 2> gpurun_out/b18_n4_2x2.err
