"""The N>1 path on CPU: world_size 2 and 4 over gloo.  Each rank owns a patch of the global domain, calls
the operator with ITS tile and the GLOBAL domain extents, and exchanges the one-cell ring with its
neighbours through the product's HaloExchanger (the same plan NCCL executes on the GPUs).  The compute on
each rank is the ORACLE (this is a CPU test; the CUDA path is exercised by tools/multi_gpu_check.py and
test_multi_gpu.py) -- what is under test is the decomposition, the exchange plan and the tile-call contract.
Result must equal the single-domain oracle bit for bit after a multi-step loop with the advance_uv stand-in.
"""
import os
import socket
import tempfile

import numpy as np
import pytest

from tests import cases

NSTEPS = 3
C_UV = 0.25
POISON = np.float32(12345.0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class NumpyPatchHalo:
    """CPU stand-in for GpuPatchHalo: pack / recv_buffer / unpack on numpy patch arrays."""

    def __init__(self, fields, grid, ext):
        self.f, self.g, self.ext = fields, grid, ext

    def _box(self, side, width, inside):
        from wrf_model_cuda_sample_b200 import EAST, NORTH, SOUTH, WEST
        ips, ipe, jps, jpe = self.ext
        if side == WEST:
            i0 = ips if inside else ips - width; return i0, i0 + width - 1, jps, jpe
        if side == EAST:
            i0 = ipe - width + 1 if inside else ipe + 1; return i0, i0 + width - 1, jps, jpe
        if side == SOUTH:
            j0 = jps if inside else jps - width; return ips, ipe, j0, j0 + width - 1
        j0 = jpe - width + 1 if inside else jpe + 1; return ips, ipe, j0, j0 + width - 1

    def _view(self, field, box):
        i0, i1, j0, j1 = box
        g = self.g
        a = self.f[field]
        J = slice(j0 - g.jms, j1 - g.jms + 1); I = slice(i0 - g.ims, i1 - g.ims + 1)
        return a[J, :, I] if a.ndim == 3 else a[J, I]

    def pack(self, field, side, width):
        import torch
        return torch.from_numpy(np.ascontiguousarray(self._view(field, self._box(side, width, True))).reshape(-1).copy())

    def recv_buffer(self, field, side, width):
        import torch
        return torch.empty(self._view(field, self._box(side, width, False)).size, dtype=torch.float32)

    def unpack(self, field, side, width, buf):
        v = self._view(field, self._box(side, width, False))
        v[...] = buf.numpy().reshape(v.shape)


def _poison(halo, decomp, rank, halos):
    """Overwrite every halo that has a neighbour: the exchange must restore it."""
    for field, sides in halos:
        for side in sides:
            if decomp.neighbour(rank, side) is not None:
                halo._view(field, halo._box(side, 1, False))[...] = POISON


def _worker(rank, world, px, py, port, outdir, shape, variant):
    import torch.distributed as dist
    import wrf_model_cuda_sample_b200 as wrf
    from wrf_model_cuda_sample_b200 import parallel
    from oracle import loader
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny, nz = shape
        G = cases.grid(nx, ny, nz, halo=5, variant=variant)
        decomp = parallel.Decomposition(G, px, py, halo=2)
        pg = decomp.patch_grid(rank)
        ext = decomp.patch_extents(rank)
        whole = wrf.synth_fields(G, seed=77)
        J = slice(pg.jms - G.jms, pg.jme - G.jms + 1); I = slice(pg.ims - G.ims, pg.ime - G.ims + 1)
        f = {n: np.ascontiguousarray(whole[n][J, :, I] if n in wrf.FIELDS_3D else
                                     whole[n][J, I] if n in wrf.FIELDS_2D else whole[n]) for n in wrf.FIELDS}
        halo = NumpyPatchHalo(f, pg, ext)
        ex = parallel.HaloExchanger(decomp, rank, halo.pack, halo.recv_buffer, halo.unpack)

        _poison(halo, decomp, rank, parallel.CONSTANT_HALOS)
        ex.exchange(parallel.CONSTANT_HALOS)                       # once per RK sub-step
        ubox, vbox = cases.standin_boxes(G, *ext)
        for s in range(NSTEPS):
            _poison(halo, decomp, rank, parallel.STEP_HALOS)
            ex.exchange(parallel.STEP_HALOS)                       # u east, v north: read by this step
            loader.oracle_c(f, pg, cases.SCALARS_3KM)              # tile = patch, domain = global
            _poison(halo, decomp, rank, parallel.OUTPUT_HALOS)
            ex.exchange(parallel.OUTPUT_HALOS)                     # mudf/mu/muts west+south: read by advance_uv
            if s + 1 < NSTEPS:
                cases.standin_advance_uv_numpy(f, pg, C_UV, ubox, vbox)
        ips, ipe, jps, jpe = ext
        Jp = slice(jps - pg.jms, jpe - pg.jms + 1); Ip = slice(ips - pg.ims, ipe - pg.ims + 1)
        out = {n: (f[n][Jp, :, Ip] if f[n].ndim == 3 else f[n][Jp, Ip]) for n in cases.OUTPUTS + ("u", "v")}
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), ext=np.array(ext), **out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("px,py,variant", [(1, 2, "specified"), (2, 1, "specified"), (2, 2, "periodic_specified"),
                                            (1, 4, "open")])
def test_decomposed_loop_equals_single_domain(px, py, variant):
    import torch.multiprocessing as mp
    import wrf_model_cuda_sample_b200 as wrf
    shape = (61, 43, 9)
    world = px * py
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(world, px, py, _free_port(), outdir, shape, variant), nprocs=world, join=True)
        G = cases.grid(*shape, halo=5, variant=variant)
        want = cases.oracle_loop(G, wrf.synth_fields(G, seed=77), cases.SCALARS_3KM, NSTEPS, c=C_UV)
        covered = np.zeros(G.shape2, dtype=bool)
        for r in range(world):
            z = np.load(os.path.join(outdir, f"rank{r}.npz"))
            ips, ipe, jps, jpe = (int(x) for x in z["ext"])
            J = slice(jps - G.jms, jpe - G.jms + 1); I = slice(ips - G.ims, ipe - G.ims + 1)
            covered[J, I] = True
            for n in cases.OUTPUTS + ("u", "v"):
                ref = want[n][J, :, I] if want[n].ndim == 3 else want[n][J, I]
                assert np.array_equal(cases.bits(z[n]), cases.bits(ref)), f"rank {r} field {n} differs"
        dom = np.zeros(G.shape2, dtype=bool)
        dom[1 - G.jms:G.jde - G.jms + 1, 1 - G.ims:G.ide - G.ims + 1] = True
        assert np.array_equal(covered, dom), "patches must tile the domain exactly"


def test_decomposition_geometry():
    from wrf_model_cuda_sample_b200 import EAST, NORTH, SOUTH, WEST, parallel
    G = cases.grid(1800, 1060, 50)
    assert parallel.choose_process_grid(8, 1800, 1060) == (1, 8)
    d = parallel.Decomposition(G, 2, 4, halo=5)
    seen = set()
    for r in range(8):
        ips, ipe, jps, jpe = d.patch_extents(r)
        assert 1 <= ips <= ipe <= 1800 and 1 <= jps <= jpe <= 1060
        seen.add((ips, ipe, jps, jpe))
        pg = d.patch_grid(r)
        assert (pg.ids, pg.ide, pg.jds, pg.jde) == (1, 1800, 1, 1060) and (pg.its, pg.ite) == (ips, ipe)
        interior, strips = d.interior_and_boundary_tiles(r)
        cells = (interior[1] - interior[0] + 1) * (interior[3] - interior[2] + 1) + \
            sum((s[1] - s[0] + 1) * (s[3] - s[2] + 1) for s in strips)
        assert cells == (ipe - ips + 1) * (jpe - jps + 1)          # interior + strips tile the patch
    assert len(seen) == 8
    assert d.neighbour(0, WEST) is None and d.neighbour(0, EAST) == 1 and d.neighbour(0, NORTH) == 2
    assert d.neighbour(7, EAST) is None and d.neighbour(7, SOUTH) == 5 and d.neighbour(7, NORTH) is None
    # every receive has exactly one matching send on the peer, in the same order
    for halos in (parallel.STEP_HALOS, parallel.CONSTANT_HALOS, parallel.OUTPUT_HALOS):
        plans = {r: parallel.HaloExchanger(d, r, None, None, None).plan(halos) for r in range(8)}
        for r, plan in plans.items():
            for peer in range(8):
                mine = [(f, s) for kind, f, s, p in plan if kind == "recv" and p == peer]
                theirs = [(f, parallel.OPPOSITE[s]) for kind, f, s, p in plans[peer] if kind == "send" and p == r]
                assert mine == theirs


def test_split_range_and_tiles_properties():
    """Property checks (hypothesis): patches tile the domain, interior + strips tile the patch."""
    from hypothesis import given, settings, strategies as st
    from wrf_model_cuda_sample_b200 import parallel

    @settings(max_examples=200, deadline=None)
    @given(st.integers(8, 400), st.integers(8, 300), st.integers(1, 4), st.integers(1, 4))
    def check(nx, ny, px, py):
        if nx < px or ny < py:
            return
        G = cases.grid(nx, ny, 6)
        d = parallel.Decomposition(G, px, py, halo=1)
        cover = np.zeros((ny + 1, nx + 1), dtype=np.int32)
        for r in range(d.world):
            ips, ipe, jps, jpe = d.patch_extents(r)
            assert ips <= ipe and jps <= jpe
            cover[jps:jpe + 1, ips:ipe + 1] += 1
            interior, strips = d.interior_and_boundary_tiles(r)
            tiles = ([interior] if interior else []) + strips
            inner = np.zeros_like(cover)
            for (a, b, c, e) in tiles:
                inner[c:e + 1, a:b + 1] += 1
            assert np.array_equal(inner[jps:jpe + 1, ips:ipe + 1], np.ones((jpe - jps + 1, ipe - ips + 1), np.int32))
            assert inner.sum() == (ipe - ips + 1) * (jpe - jps + 1)
            # a tile that reads a per-step halo is never the interior one
            if interior and d.neighbour(r, wrf_EAST) is not None:
                assert interior[1] < ipe
            if interior and d.neighbour(r, wrf_NORTH) is not None:
                assert interior[3] < jpe
        assert np.array_equal(cover[1:, 1:], np.ones((ny, nx), np.int32))

    from wrf_model_cuda_sample_b200 import EAST as wrf_EAST, NORTH as wrf_NORTH
    check()


def _gather_worker(rank, world, port, outdir):
    import torch.distributed as dist
    from wrf_model_cuda_sample_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = bytes([rank]) * 2048                       # the size of a wrfb200_comm_init info blob
        got = parallel.torch_allgather_bytes()(blob)
        ok = len(got) == world and all(got[r] == bytes([r]) * 2048 for r in range(world))
        open(os.path.join(outdir, f"ok{rank}"), "w").write("1" if ok else "0")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_bootstrap_allgather_keeps_rank_order():
    """The fused multi-GPU path needs one thing from the host's transport: the info blobs of all ranks, in rank
    order (parallel.connect_fused).  The torch.distributed helper used by bench.py, over gloo, world size 3."""
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_gather_worker, args=(3, _free_port(), outdir), nprocs=3, join=True)
        assert [open(os.path.join(outdir, f"ok{r}")).read() for r in range(3)] == ["1", "1", "1"]
