"""Race hunt: run the deep-column case many times through one kernel and count mismatching values
against the oracle (a correct kernel is bit-identical every time)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wrf_model_cuda_sample_b200 as wrf  # noqa: E402
from oracle import loader  # noqa: E402
from tests import cases  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kernel = {"pipe": wrf.KERNEL_PIPE, "tile": wrf.KERNEL_TILE, "column": wrf.KERNEL_COLUMN}[sys.argv[2] if len(sys.argv) > 2 else "pipe"]
total_bad = 0
for (nx, ny, nz, variant, scal) in ((200, 96, 120, "periodic_specified", cases.SCALARS_3KM),
                                    (425, 300, 35, "specified", cases.SCALARS_12KM),
                                    (900, 200, 50, "specified", cases.SCALARS_3KM)):
    g = cases.grid(nx, ny, nz, halo=5, variant=variant)
    fin = wrf.synth_fields(g, seed=4)
    want = cases.copy_fields(fin)
    loader.oracle_c(want, g, scal, tiles=16)
    bad_runs = bad_vals = 0
    with wrf.Patch(g) as p:
        p.set_scalars(*scal)
        p.set_kernel(kernel)
        for r in range(reps):
            got = cases.copy_fields(fin)
            p.upload(got)
            p.step()
            p.download(got)
            nb = sum(int(np.count_nonzero(cases.bits(got[n]) != cases.bits(want[n]))) for n in cases.OUTPUTS)
            bad_runs += nb > 0
            bad_vals += nb
    print(f"{nx}x{ny}x{nz} {variant}: {bad_runs}/{reps} runs differ, {bad_vals} values", flush=True)
    total_bad += bad_vals
print("RACE_FREE" if total_bad == 0 else "MISMATCHES", total_bad)
