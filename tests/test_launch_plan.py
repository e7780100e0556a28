"""The launch geometry of the TMA kernel (tile columns, remainder strips, 2-row / 1-row block rows), checked on the CPU
through the host-only wrfb200_pipe_plan: every computed column and row is covered exactly once, the shared-memory
budget holds, and the shapes of the BASELINE configs are the ones the profiles were taken with."""
import ctypes as C

import numpy as np
import pytest

import wrf_model_cuda_sample_b200 as wrf
from tests import cases

TI, STRIP_W, STRIP_ROWS, SLOTS = 128, 16, 2, 2 * 148
KEYS = ("cfg", "tj", "stages", "nbx", "nby2", "nby1", "strip_blocks", "strip_i0", "grid", "smem")


def plan(g, slots=SLOTS):
    out = (C.c_longlong * 10)()
    d = g.domain()
    rc = wrf.lib().wrfb200_pipe_plan(C.byref(d), g.its, g.ite, g.jts, g.jte, g.kts, g.kte, slots, out)
    assert rc == 0, wrf.lib().wrfb200_last_error()
    return dict(zip(KEYS, list(out)))


def check_cover(g, p):
    i0, i1, j0, j1, k0, k1 = g.bounds()
    mi0, mi1 = i0 - g.ims, i1 - g.ims                     # memory indices, as the kernels see them
    nj = j1 - j0 + 1
    origin = mi0 & ~31
    cols = np.zeros(mi1 + 2 * TI, dtype=int)
    for bx in range(p["nbx"]):                            # tile columns
        lo, hi = origin + bx * TI, origin + bx * TI + TI - 1
        cols[max(lo, mi0):min(hi, mi1) + 1] += 1
    if p["strip_blocks"]:
        assert p["strip_i0"] == origin + p["nbx"] * TI and mi1 - p["strip_i0"] + 1 <= STRIP_W
        cols[p["strip_i0"]:mi1 + 1] += 1
        assert p["strip_blocks"] == -(-nj // STRIP_ROWS)  # strips sweep all computed rows
    assert (cols[mi0:mi1 + 1] == 1).all() and cols[:mi0].sum() == 0 and cols[mi1 + 1:].sum() == 0
    # rows: 2-row block rows first, then 1-row block rows
    assert 2 * p["nby2"] + p["nby1"] >= nj and 2 * (p["nby2"] - 1) + p["nby1"] < nj if p["nby1"] == 0 else \
        2 * p["nby2"] + p["nby1"] == nj
    assert p["grid"] == p["strip_blocks"] + p["nbx"] * (p["nby2"] + p["nby1"])
    # shared memory: two resident blocks per SM for the two-block configurations, one otherwise
    per_sm = 227 * 1024
    blocks = 2 if p["cfg"] in (22, 13, 12, 62) else 1
    assert blocks * (p["smem"] + 1024) <= per_sm


def test_baseline_shapes_match_the_profiled_launches():
    g = wrf.Grid.from_shape(1800, 1060, 50, halo=5)
    p = plan(g)
    assert (p["cfg"], p["nbx"], p["nby2"], p["nby1"], p["strip_blocks"], p["grid"]) == (22, 14, 529, 0, 529, 7935)  # profiles/r2_launches_conus3.csv
    p = plan(wrf.Grid.from_shape(1800, 133, 50, halo=5))
    assert p["cfg"] == 22 and p["nbx"] == 14 and p["nby1"] >= 15 and p["grid"] == 1087     # r2_ncu_details_pipe_patch8.txt
    p = plan(wrf.Grid.from_shape(425, 300, 35, halo=5))
    assert (p["cfg"], p["nbx"], p["nby2"], p["nby1"], p["strip_blocks"], p["grid"]) == (22, 4, 149, 0, 0, 596)      # r2_ncu_details_pipe_conus12.txt
    p = plan(wrf.Grid.from_shape(512, 512, 120, halo=5))
    assert p["cfg"] == 62 and p["tj"] == 1 and 2 * (p["smem"] + 1024) <= 227 * 1024          # two blocks per SM for nk = 119
    p = plan(wrf.Grid.from_shape(74, 61, 28, halo=5))
    assert p["nby2"] == 0 and p["nby1"] == 58                                                # less than one wave: all 1-row


@pytest.mark.parametrize("variant", sorted(cases.FLAG_VARIANTS))
def test_every_column_and_row_is_covered_exactly_once(variant):
    rs = np.random.RandomState(hash(variant) % 1000)
    for _ in range(150):
        nx, ny, nz = int(rs.randint(5, 2100)), int(rs.randint(5, 400)), int(rs.randint(2, 130))
        halo = int(rs.randint(1, 8))
        g = cases.grid(nx, ny, nz, halo=halo, variant=variant)
        i0, i1, j0, j1, _, _ = g.bounds()
        if i0 > i1 or j0 > j1:
            continue
        # a random sub-tile, as WRF calls the routine per tile
        if rs.rand() < 0.5:
            its = int(rs.randint(1, nx + 1)); ite = int(rs.randint(its, nx + 1))
            jts = int(rs.randint(1, ny + 1)); jte = int(rs.randint(jts, ny + 1))
            g = g.with_tile(its, ite, jts, jte)
            i0, i1, j0, j1, _, _ = g.bounds()
            if i0 > i1 or j0 > j1:
                continue
        if nz - 1 > 400:
            continue
        out = (C.c_longlong * 10)()
        d = g.domain()
        rc = wrf.lib().wrfb200_pipe_plan(C.byref(d), g.its, g.ite, g.jts, g.jte, g.kts, g.kte,
                                         int(rs.choice([SLOTS, 2 * 132, 64])), out)
        if rc != 0:
            continue
        check_cover(g, dict(zip(KEYS, list(out))))


def test_plan_rejects_empty_and_unsupported_calls():
    g = cases.grid(3, 3, 4, halo=1, variant="specified")           # empty index sets
    out = (C.c_longlong * 10)()
    d = g.domain()
    assert wrf.lib().wrfb200_pipe_plan(C.byref(d), g.its, g.ite, g.jts, g.jte, g.kts, g.kte, 0, out) == 1
    g = cases.grid(40, 30, 10, halo=2)
    d = g.domain()
    assert wrf.lib().wrfb200_pipe_plan(C.byref(d), g.its, g.ite, g.jts, g.jte, 2, g.kte, 0, out) == 2
    assert wrf.lib().wrfb200_pipe_plan(None, 1, 2, 1, 2, 1, 2, 0, out) == 1
