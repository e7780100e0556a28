#!/bin/bash
# round-2 batch 1: tile-config sweep on the off-design shapes + ncu source capture of the product kernel
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --steps 20 --warmup 3"
for wl in patch8 conus12 deep120; do
  for cfg in 0 12 13 14 22 23 24; do
    echo "== $wl cfg=$cfg" >> gpurun_out/b1_sweep.log
    WRFB200_PIPE_CFG=$cfg timeout 120 $B --workload $wl 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b1_sweep.log 2>&1
  done
done
echo "== conus3 cfg sweep" >> gpurun_out/b1_sweep.log
for cfg in 0 12 13 23; do
  echo "== conus3 cfg=$cfg" >> gpurun_out/b1_sweep.log
  WRFB200_PIPE_CFG=$cfg timeout 120 $B --workload conus3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])
" >> gpurun_out/b1_sweep.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:amt_pipe -s 4 -c 1 -f -o gpurun_out/b1_pipe_conus3 \
  python bench.py --no-e2e --no-cpu --no-ref-cuda --steps 1 --warmup 3 --workload conus3 > gpurun_out/b1_ncu.log 2>&1
nvidia-smi topo -m > gpurun_out/b1_topo.txt 2>&1
ls -la gpurun_out >> gpurun_out/b1_sweep.log
