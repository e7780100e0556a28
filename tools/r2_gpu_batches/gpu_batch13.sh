#!/bin/bash
# round-2 batch 13 (1 GPU): ncu evidence of the final kernels + compute-sanitizer
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-ref-cuda --no-extras"
# launch list of the bench command (serialised, cold-cache: compare shares)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_conus3.csv $B --steps 2 --warmup 3 > gpurun_out/b13_launches.log 2>&1
for wl in conus3 conus12 deep120; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:amt_pipe -s 4 -c 1 -f -o gpurun_out/r2_pipe_$wl $B --steps 1 --warmup 3 --workload $wl > gpurun_out/b13_ncu_$wl.log 2>&1
done
# sanitizer: the single-GPU kernels (smoke's first part) and the fused two-rank loop; a flag wait gives up after 2 s
# in case the tool serialises the two ranks' kernels
export WRFB200_FLAG_TIMEOUT_MS=2000
timeout 300 compute-sanitizer --tool memcheck python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_sanitizer_memcheck_smoke.txt 2>&1
timeout 300 compute-sanitizer --tool racecheck python -c "
import __graft_entry__ as g
g.smoke_fused = lambda: None     # racecheck forces blocking launches: two ranks on one device cannot overlap
g.smoke()" > gpurun_out/r2_sanitizer_racecheck_smoke.txt 2>&1
