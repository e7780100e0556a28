"""Multi-GPU parity check (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/multi_gpu_check.py [--pgrid PXxPY] [--shape NXxNYxNZ] [--steps S]

Every rank runs the CUDA path on its patch (tile = patch, domain = global), exchanges the one-cell ring with
NCCL send/recv through HaloExchanger, overlaps the exchange with the interior tile, and applies the
advance_uv stand-in between steps.  Each rank then checks its patch BIT FOR BIT against the single-domain
oracle loop.  Exit code 0 = every rank identical.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import wrf_model_cuda_sample_b200 as wrf  # noqa: E402
from wrf_model_cuda_sample_b200 import parallel  # noqa: E402
from tests import cases  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pgrid", default="")
    ap.add_argument("--shape", default="300x160x20")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--variant", default="specified")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nx, ny, nz = (int(x) for x in args.shape.lower().split("x"))
    G = cases.grid(nx, ny, nz, halo=5, variant=args.variant)
    px, py = (int(x) for x in args.pgrid.lower().split("x")) if args.pgrid else parallel.choose_process_grid(world, nx, ny)
    assert px * py == world
    decomp = parallel.Decomposition(G, px, py, halo=3)
    pg, ext = decomp.patch_grid(rank), decomp.patch_extents(rank)
    c_uv = 0.25

    whole = wrf.synth_fields(G, seed=99)
    J = slice(pg.jms - G.jms, pg.jme - G.jms + 1); I = slice(pg.ims - G.ims, pg.ime - G.ims + 1)
    f = {n: np.ascontiguousarray(whole[n][J, :, I] if n in wrf.FIELDS_3D else
                                 whole[n][J, I] if n in wrf.FIELDS_2D else whole[n]) for n in wrf.FIELDS}
    # poison every halo that a neighbour must fill, so a missing exchange cannot go unnoticed
    ips, ipe, jps, jpe = ext
    for halos in (parallel.CONSTANT_HALOS, parallel.STEP_HALOS):
        for field, sides in halos:
            a = f[field]
            for side in sides:
                if decomp.neighbour(rank, side) is None:
                    continue
                if side == wrf.EAST: a[..., ipe + 1 - pg.ims:] = 12345.0
                if side == wrf.WEST: a[..., :ips - pg.ims] = 12345.0
                if side == wrf.NORTH: a[jpe + 1 - pg.jms:] = 12345.0
                if side == wrf.SOUTH: a[:jps - pg.jms] = 12345.0

    main_stream = torch.cuda.current_stream()
    comm = torch.cuda.Stream(device=dev)
    with wrf.Patch(pg, device=local) as patch:
        patch.set_stream(main_stream.cuda_stream)
        patch.set_scalars(*cases.SCALARS_3KM)
        patch.upload(f)
        halo = parallel.GpuPatchHalo(patch, decomp, rank, dev)
        ex = parallel.HaloExchanger(decomp, rank, halo.pack, halo.recv_buffer, halo.unpack)
        ex.exchange(parallel.CONSTANT_HALOS)
        interior, strips = decomp.interior_and_boundary_tiles(rank)
        ubox, vbox = cases.standin_boxes(G, *ext)
        for s in range(args.steps):
            comm.wait_stream(main_stream)
            with torch.cuda.stream(comm):
                patch.set_stream(comm.cuda_stream)
                ex.finish(ex.start(parallel.STEP_HALOS))
                for t in strips:                     # halo-dependent strips right behind the unpack, same stream
                    patch.step(pg.with_tile(*t))
            patch.set_stream(main_stream.cuda_stream)
            if interior:
                patch.step(pg.with_tile(*interior))  # concurrent with exchange + strips
            main_stream.wait_stream(comm)
            ex.exchange(parallel.OUTPUT_HALOS)
            if s + 1 < args.steps:
                patch.standin_advance_uv("u", c_uv, *ubox)
                patch.standin_advance_uv("v", c_uv, *vbox)
        patch.download(f, names=cases.OUTPUTS + ("u", "v"))

    want = cases.oracle_loop(G, whole, cases.SCALARS_3KM, args.steps, c=c_uv)
    Jp = slice(jps - pg.jms, jpe - pg.jms + 1); Ip = slice(ips - pg.ims, ipe - pg.ims + 1)
    Jg = slice(jps - G.jms, jpe - G.jms + 1); Ig = slice(ips - G.ims, ipe - G.ims + 1)
    bad = 0
    for n in cases.OUTPUTS + ("u", "v"):
        got = f[n][Jp, :, Ip] if f[n].ndim == 3 else f[n][Jp, Ip]
        ref = want[n][Jg, :, Ig] if want[n].ndim == 3 else want[n][Jg, Ig]
        nb = int(np.count_nonzero(cases.bits(got) != cases.bits(ref)))
        if nb:
            print(f"rank {rank}: field {n}: {nb} of {got.size} values differ", flush=True)
        bad += nb
    t = torch.tensor([bad], device=dev, dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        print(f"multi_gpu_check {px}x{py} on {nx}x{ny}x{nz}, {args.steps} steps: "
              f"{'BIT-IDENTICAL to the single-domain oracle' if t.item() == 0 else str(t.item()) + ' MISMATCHES'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
