#!/bin/bash
# round-2 batch 14 (8 GPUs): N=8 strong scaling with the north-row-second dispatch order (A/B), N=1 and N=4 on the same box
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
BA="--steps 20 --warmup 5 --no-extras --no-e2e"
timeout 300 python bench.py $BA --no-cpu --no-ref-cuda > gpurun_out/b14_n1.json 2> gpurun_out/b14_n1.err
timeout 400 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 $BA > gpurun_out/b14_n8_a.json 2> gpurun_out/b14_n8_a.err
WRFB200_NORTH_SECOND=0 timeout 400 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 $BA > gpurun_out/b14_n8_nosecond.json 2> gpurun_out/b14_n8_nosecond.err
timeout 400 $TR --nproc-per-node 8 --master-port 29516 bench.py --gpus 8 $BA > gpurun_out/b14_n8_b.json 2> gpurun_out/b14_n8_b.err
WRFB200_PIPE_STRIP=0 timeout 400 $TR --nproc-per-node 8 --master-port 29517 bench.py --gpus 8 $BA > gpurun_out/b14_n8_nostrip.json 2> gpurun_out/b14_n8_nostrip.err
timeout 400 $TR --nproc-per-node 4 --master-port 29518 bench.py --gpus 4 $BA > gpurun_out/b14_n4.json 2> gpurun_out/b14_n4.err
timeout 400 $TR --nproc-per-node 2 --master-port 29519 bench.py --gpus 2 $BA > gpurun_out/b14_n2.json 2> gpurun_out/b14_n2.err
timeout 300 python bench.py $BA --no-cpu --no-ref-cuda > gpurun_out/b14_n1_b.json 2> gpurun_out/b14_n1_b.err
