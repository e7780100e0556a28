#!/bin/bash
# round-2 batch 2 (1 GPU): new fused-comm tests (ranks share the device), the full GPU suite, the new bench line
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/b2_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_fused_comm.py -q -m gpu -x --timeout 300 > gpurun_out/b2_fused.log 2>&1
echo "fused rc=$?" >> gpurun_out/b2_fused.log
timeout 600 python -m pytest tests/test_c_harness.py -q -m gpu --timeout 400 > gpurun_out/b2_charness.log 2>&1
echo "charness rc=$?" >> gpurun_out/b2_charness.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_fused_comm.py --deselect tests/test_c_harness.py > gpurun_out/b2_gpu_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/b2_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/b2_bench.json 2> gpurun_out/b2_bench.err
echo "bench rc=$?" >> gpurun_out/b2_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/b2_bench_ref.json 2> gpurun_out/b2_bench_ref.err
