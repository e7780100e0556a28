"""Shared test cases: grids, inputs and bit-level comparison helpers."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import wrf_model_cuda_sample_b200 as wrf  # noqa: E402
from wrf_model_cuda_sample_b200 import Grid  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
OUTPUTS = ("ww", "t", "t_ave", "mu", "muave", "muts", "mudf")
SCALARS_12KM = (np.float32(1.0 / 12000.0), np.float32(1.0 / 12000.0), np.float32(12.0), np.float32(0.1))
SCALARS_3KM = (np.float32(1.0 / 3000.0), np.float32(1.0 / 3000.0), np.float32(3.0), np.float32(0.1))

# (periodic_x, specified, nested): the three rows of the index-set table + nested alias
FLAG_VARIANTS = {
    "specified": (False, True, False),
    "periodic_specified": (True, True, False),
    "open": (False, False, False),
    "nested": (False, False, True),
    "periodic_open": (True, False, False),
}


def grid(nx, ny, nz, halo=5, variant="specified") -> Grid:
    px, sp, ne = FLAG_VARIANTS[variant]
    return Grid.from_shape(nx, ny, nz, halo=halo, periodic_x=px, specified=sp, nested=ne)


def copy_fields(f):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in f.items() if isinstance(v, np.ndarray)}


def random_fields(g: Grid, seed: int, adversarial: bool = False):
    """numpy-only inputs (independent of the product library), atmosphere-like magnitudes, or with
    ``adversarial`` uniform +-1 everywhere (cancellation stress; map factors kept away from 0)."""
    rs = np.random.RandomState(seed)
    f = {}

    def u(shape, lo, hi):
        return rs.uniform(lo, hi, size=shape).astype(np.float32)

    s3, s2, s1 = g.shape3, g.shape2, g.shape1
    if adversarial:
        for n in wrf.FIELDS_3D:
            f[n] = u(s3, -1, 1)
        for n in wrf.FIELDS_2D:
            f[n] = u(s2, -1, 1)
        for n in ("msfuy", "msfty", "msftx", "msfvx_inv"):
            f[n] = np.where(np.abs(f[n]) < 0.25, np.float32(0.5), f[n]).astype(np.float32)
        for n in wrf.FIELDS_1D:
            f[n] = u(s1, -1, 1)
        return f
    f["u"] = u(s3, -2e3, 2e3); f["v"] = u(s3, -2e3, 2e3)
    f["u_1"] = u(s3, -30, 30); f["v_1"] = u(s3, -30, 30)
    lev = np.linspace(0, 150, s3[1], dtype=np.float32)[None, :, None]
    f["t_1"] = (lev + u(s3, -2, 2)).astype(np.float32)
    f["t"] = u(s3, -50, 50); f["ft"] = u(s3, -5, 5)
    f["ww"] = u(s3, -0.5, 0.5); f["ww_1"] = u(s3, -0.5, 0.5)
    f["t_ave"] = u(s3, 3900, 4100)
    f["mut"] = u(s2, 9.0e4, 9.8e4)
    f["muu"] = (f["mut"] + u(s2, -50, 50)).astype(np.float32)
    f["muv"] = (f["mut"] + u(s2, -50, 50)).astype(np.float32)
    f["mu"] = u(s2, -300, 300); f["mu_tend"] = u(s2, -0.5, 0.5)
    for n in ("msfuy", "msftx", "msfty", "msfvx_inv"):
        f[n] = u(s2, 0.9, 1.1)
    f["muave"] = u(s2, 900, 1100); f["muts"] = u(s2, 1900, 2100); f["mudf"] = u(s2, 2900, 3100)
    nk = s1[0]
    f["dnw"] = (-(1.0 / nk) * rs.uniform(0.5, 1.5, size=nk)).astype(np.float32)
    f["rdnw"] = (np.float32(1.0) / f["dnw"]).astype(np.float32)
    f["fnm"] = u(s1, 0.4, 0.6); f["fnp"] = (np.float32(1.0) - f["fnm"]).astype(np.float32)
    return f


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_bit_equal(got, want, names=OUTPUTS, what=""):
    for n in names:
        g_, w_ = bits(got[n]), bits(want[n])
        if not np.array_equal(g_, w_):
            bad = np.argwhere(g_ != w_)
            first = tuple(bad[0])
            raise AssertionError(
                f"{what}{n}: {len(bad)} of {g_.size} values differ; first at [j,k,i]={first}: "
                f"got {got[n][first]!r} want {want[n][first]!r}")


def assert_inputs_untouched(after, before):
    for n in wrf.FIELDS:
        if n not in OUTPUTS:
            assert np.array_equal(bits(after[n]), bits(before[n])), f"input field {n} was modified"


def assert_outside_untouched(after, before, g: Grid):
    """Cells outside i_start..i_end x j_start..j_end (and level kte of 3-D fields) keep their bytes."""
    i0, i1, j0, j1, k0, k1 = g.bounds()
    for n in OUTPUTS:
        a, b = bits(after[n]), bits(before[n])
        mask = np.ones(a.shape, dtype=bool)
        if a.ndim == 3:
            mask[j0 - g.jms:j1 - g.jms + 1, k0 - g.kms:k1 - g.kms + 1, i0 - g.ims:i1 - g.ims + 1] = False
        else:
            mask[j0 - g.jms:j1 - g.jms + 1, i0 - g.ims:i1 - g.ims + 1] = False
        assert np.array_equal(a[mask], b[mask]), f"{n}: cells outside the computed range were written"


def standin_boxes(g: Grid, ips, ipe, jps, jpe):
    """Index boxes (Fortran-numbered, inclusive) of the stand-in advance_uv update restricted to a patch:
    u over i_start+1..i_end, v over j_start+1..j_end of the GLOBAL computed range."""
    gi0, gi1, gj0, gj1, _, _ = Grid(g.ids, g.ide, g.jds, g.jde, g.kde, g.ims, g.ime, g.jms, g.jme, g.kms, g.kme,
                                    g.ids, g.ide, g.jds, g.jde, g.kts, g.kte, g.periodic_x, g.specified,
                                    g.nested).bounds()
    ubox = (max(ips, gi0 + 1), min(ipe, gi1), max(jps, gj0), min(jpe, gj1))
    vbox = (max(ips, gi0), min(ipe, gi1), max(jps, gj0 + 1), min(jpe, gj1))
    return ubox, vbox


def standin_advance_uv_numpy(f, g: Grid, c, ubox, vbox):
    """u += c*(mudf(i)-mudf(i-1)), v += c*(mudf(j)-mudf(j-1)) in float32, on arrays with g's memory extents."""
    c = np.float32(c)
    i0, i1, j0, j1 = ubox
    if i0 <= i1 and j0 <= j1:
        I = slice(i0 - g.ims, i1 - g.ims + 1); Im = slice(i0 - 1 - g.ims, i1 - g.ims)
        J = slice(j0 - g.jms, j1 - g.jms + 1)
        du = c * (f["mudf"][J, I] - f["mudf"][J, Im])
        f["u"][J, :, I] = f["u"][J, :, I] + du[:, None, :]
    i0, i1, j0, j1 = vbox
    if i0 <= i1 and j0 <= j1:
        I = slice(i0 - g.ims, i1 - g.ims + 1)
        J = slice(j0 - g.jms, j1 - g.jms + 1); Jm = slice(j0 - 1 - g.jms, j1 - g.jms)
        dv = c * (f["mudf"][J, I] - f["mudf"][Jm, I])
        f["v"][J, :, I] = f["v"][J, :, I] + dv[:, None, :]


def oracle_loop(g: Grid, fin, scalars, nsteps, c=None):
    """nsteps of the oracle on the whole domain, with the stand-in for advance_uv between steps."""
    from oracle import loader
    f = copy_fields(fin)
    ubox, vbox = standin_boxes(g, g.ids, g.ide, g.jds, g.jde)
    for s in range(nsteps):
        loader.oracle_c(f, g, scalars)
        if c is not None and s + 1 < nsteps:
            standin_advance_uv_numpy(f, g, c, ubox, vbox)
    return f


# ---------------------------------------------------------------- multi-rank helpers (patch carving)
POISON = np.float32(12345.0)


def carve_patch(whole, G: Grid, pg: Grid):
    """The rank's share (patch + halo memory) of the single-domain fields ``whole``."""
    J = slice(pg.jms - G.jms, pg.jme - G.jms + 1)
    I = slice(pg.ims - G.ims, pg.ime - G.ims + 1)
    return {n: np.ascontiguousarray(whole[n][J, :, I] if n in wrf.FIELDS_3D else
                                    whole[n][J, I] if n in wrf.FIELDS_2D else whole[n]) for n in wrf.FIELDS}


def poison_neighbour_halos(f, decomp, rank, pg: Grid, halo_sets):
    """Overwrite every halo cell a neighbour must fill, so a missing exchange cannot go unnoticed."""
    ips, ipe, jps, jpe = decomp.patch_extents(rank)
    for halos in halo_sets:
        for field, sides in halos:
            a = f[field]
            for side in sides:
                if decomp.neighbour(rank, side) is None:
                    continue
                if side == wrf.EAST: a[..., ipe + 1 - pg.ims:] = POISON
                if side == wrf.WEST: a[..., :ips - pg.ims] = POISON
                if side == wrf.NORTH: a[jpe + 1 - pg.jms:] = POISON
                if side == wrf.SOUTH: a[:jps - pg.jms] = POISON


def patch_mismatches(f, want, G: Grid, pg: Grid, ext, names=OUTPUTS + ("u", "v")):
    """Number of values of the rank's patch (no halo) that differ bit-wise from the single-domain result."""
    ips, ipe, jps, jpe = ext
    Jp = slice(jps - pg.jms, jpe - pg.jms + 1); Ip = slice(ips - pg.ims, ipe - pg.ims + 1)
    Jg = slice(jps - G.jms, jpe - G.jms + 1); Ig = slice(ips - G.ims, ipe - G.ims + 1)
    bad = {}
    for n in names:
        got = f[n][Jp, :, Ip] if f[n].ndim == 3 else f[n][Jp, Ip]
        ref = want[n][Jg, :, Ig] if want[n].ndim == 3 else want[n][Jg, Ig]
        nb = int(np.count_nonzero(bits(got) != bits(ref)))
        if nb:
            bad[n] = nb
    return bad
