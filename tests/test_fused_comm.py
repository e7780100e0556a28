"""The multi-GPU path of csrc/comm.cu through the C ABI: 2-D patches whose one-cell halo exchange is fused
into the kernels over peer-mapped memory.  Here all ranks are handles of ONE process on device 0 (the
library then shares plain device pointers instead of CUDA IPC handles), so the test runs on a single-GPU
box; the flag protocol, the peer stores and the kernels are exactly those of the multi-process runs
(tools/multi_gpu_check.py under torchrun, tests/c_comm_harness.c with fork()).
Bar: every rank's patch BIT-IDENTICAL to the single-domain oracle loop, halos poisoned beforehand."""
import numpy as np
import pytest

import wrf_model_cuda_sample_b200 as wrf
from wrf_model_cuda_sample_b200 import parallel
from tests import cases

pytestmark = pytest.mark.gpu

C_UV = 0.25


def run_ranks(G, px, py, nsteps, standin=True, graph=True, seed=99, scalars=cases.SCALARS_3KM, halo=3,
              stepwise=False, repeats=1):
    decomp = parallel.Decomposition(G, px, py, halo=halo)
    world = px * py
    whole = wrf.synth_fields(G, seed=seed)
    ranks = []
    for r in range(world):
        pg = decomp.patch_grid(r)
        f = cases.carve_patch(whole, G, pg)
        cases.poison_neighbour_halos(f, decomp, r, pg, (parallel.CONSTANT_HALOS, parallel.STEP_HALOS, parallel.OUTPUT_HALOS))
        p = wrf.Patch(pg, device=0)
        p.set_scalars(*scalars)
        p.upload(f)
        ranks.append((p, pg, f))
    try:
        infos = [p.comm_init(px, py, r, *decomp.patch_extents(r)) for r, (p, _, _) in enumerate(ranks)]
        for p, _, _ in ranks:
            p.comm_connect(infos)
        for p, _, _ in ranks:                       # asynchronous: the neighbour barrier inside spins on the
            p.comm_push_constants()                 # device until the other ranks' kernels have been enqueued
        for _ in range(repeats):
            if stepwise:
                for s in range(nsteps):
                    for p, _, _ in ranks:
                        p.comm_push_uv()
                        p.comm_step()
                        if standin and s + 1 < nsteps:
                            p.comm_standin_advance_uv(C_UV)
            else:
                for p, _, _ in ranks:
                    p.comm_loop(nsteps, standin=standin, c=C_UV, graph=graph)
        for r, (p, pg, f) in enumerate(ranks):
            timeouts, steps = p.comm_status()
            assert timeouts == 0, f"rank {r}: {timeouts} halo waits timed out"
            assert steps == nsteps * repeats
            p.download(f, names=cases.OUTPUTS + ("u", "v"))
    finally:
        for p, _, _ in ranks:
            p.close()
    return decomp, whole, ranks


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("px,py", [(1, 2), (2, 1), (2, 2), (1, 4)])
def test_fused_exchange_matches_single_domain_oracle(px, py, graph):
    G = cases.grid(300, 160, 20, halo=5, variant="specified")
    nsteps = 4
    decomp, whole, ranks = run_ranks(G, px, py, nsteps, standin=True, graph=graph)
    want = cases.oracle_loop(G, whole, cases.SCALARS_3KM, nsteps, c=C_UV)
    for r, (_, pg, f) in enumerate(ranks):
        bad = cases.patch_mismatches(f, want, G, pg, decomp.patch_extents(r))
        assert not bad, f"{px}x{py} rank {r}: {bad}"


@pytest.mark.parametrize("variant", ["open", "periodic_specified"])
def test_fused_exchange_flag_variants_stepwise(variant):
    """Edge clamps on the boundary ranks only (global ids..jde), driven step by step through the C ABI."""
    G = cases.grid(280, 90, 12, halo=4, variant=variant)
    nsteps = 3
    decomp, whole, ranks = run_ranks(G, 2, 2, nsteps, standin=True, stepwise=True, seed=7, halo=2)
    want = cases.oracle_loop(G, whole, cases.SCALARS_3KM, nsteps, c=C_UV)
    for r, (_, pg, f) in enumerate(ranks):
        bad = cases.patch_mismatches(f, want, G, pg, decomp.patch_extents(r))
        assert not bad, f"{variant} rank {r}: {bad}"


def test_fused_loop_replays_keep_epochs_consistent():
    """bench.py replays the captured n-step loop many times: the epoch flags must stay in step."""
    G = cases.grid(260, 64, 10, halo=5, variant="specified")
    decomp, whole, ranks = run_ranks(G, 1, 2, 3, standin=False, graph=True, repeats=5)
    want = cases.oracle_loop(G, whole, cases.SCALARS_3KM, 15)
    for r, (_, pg, f) in enumerate(ranks):
        bad = cases.patch_mismatches(f, want, G, pg, decomp.patch_extents(r), names=cases.OUTPUTS)
        assert not bad, f"rank {r}: {bad}"


def test_comm_requires_owned_mirrors_and_connect():
    import torch
    g = cases.grid(64, 32, 6, halo=2)
    with wrf.Patch(g, device=0) as p:
        with pytest.raises(wrf.WrfB200Error):
            p.comm_step()                                       # no comm_init
        p.comm_init(1, 1, 0, 1, 64, 1, 32)
        with pytest.raises(wrf.WrfB200Error):
            p.comm_step()                                       # not connected
    with wrf.Patch(g, device=0, allocate=False) as p:
        with pytest.raises(wrf.WrfB200Error):
            p.comm_init(1, 1, 0, 1, 64, 1, 32)                  # caller-owned buffers cannot be IPC-exported
