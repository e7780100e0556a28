"""oracle/synth_np.py -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

numpy port of the deterministic atmosphere-like input generator (``wrfb200_synth_field``,
wrf_model_cuda_sample_b200/csrc/host_utils.cu): same counter-based RNG, same formulas, same float32
rounding points.  It exists so that ``bench.py --impl reference`` can build the reference arm's inputs
WITHOUT loading the product library (the reference arm runs the reference's own C on the host cores;
nothing of ours may sit on that path).  The reference itself reads its inputs from un-shipped ``.bin`` dumps
(/root/reference/advance_mu_t_driver.f90:36-167); these fields have the same roles and magnitudes.

``tests/test_oracle_pinned.py`` checks this port against the C generator (bit-equal up to libm's last-ulp
differences in sin/cos, which almost never survive the rounding to float32).

Arrays: numpy float32, C order [j,k,i] / [j,i] / [k]; ``grid`` is any object with the Fortran index members
ids..kme and periodic_x/specified/nested (e.g. wrf_model_cuda_sample_b200.Grid -- importing that class does
not load libwrfb200.so).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np

FIELDS_3D = ("ww", "ww_1", "u", "u_1", "v", "v_1", "t", "t_1", "t_ave", "ft")
FIELDS_2D = ("mu", "mut", "muave", "muts", "muu", "muv", "mudf", "mu_tend",
             "msfuy", "msfvx_inv", "msftx", "msfty")
FIELDS_1D = ("dnw", "fnm", "fnp", "rdnw")
FIELDS = FIELDS_3D + FIELDS_2D + FIELDS_1D
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}

_M = (1 << 64) - 1
_TWO_PI = 6.283185307179586
F = np.float32


def _mix_scalar(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return z ^ (z >> 31)


def _mix(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _sym(seed: int, field: int, gi, gk, gj) -> np.ndarray:
    """2*unit-1 with unit = counter-based uniform in [0,1): a pure function of (seed, field, global i,k,j)."""
    key0 = np.uint64(_mix_scalar(seed ^ _mix_scalar((field + 0x51ED27) & _M)))
    ij = (gi.astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)) | \
         ((gj.astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)) << np.uint64(32))
    key = _mix(key0 ^ ij)
    key = _mix(key ^ (gk.astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)))
    unit = (key >> np.uint64(40)).astype(np.float32) * F(1.0 / 16777216.0)
    return F(2.0) * unit - F(1.0)


def bounds(g, its=None, ite=None, jts=None, jte=None):
    """Index sets of module_small_step_em.f90:91-106 for the tile (default: the grid's own tile)."""
    its = g.its if its is None else its
    ite = g.ite if ite is None else ite
    jts = g.jts if jts is None else jts
    jte = g.jte if jte is None else jte
    i0, i1 = its, min(ite, g.ide - 1)
    j0, j1 = jts, min(jte, g.jde - 1)
    spec = g.specified or g.nested
    if not g.periodic_x and spec:
        i0, i1 = max(its, g.ids + 1), min(ite, g.ide - 2)
    if spec:
        j0, j1 = max(jts, g.jds + 1), min(jte, g.jde - 2)
    return i0, i1, j0, j1, g.kts, g.kte - 1


def updated_points(g):
    i0, i1, j0, j1, k0, k1 = bounds(g)
    n2 = max(0, i1 - i0 + 1) * max(0, j1 - j0 + 1)
    return n2 * max(0, k1 - k0 + 1), n2


def _norm(g, gi, gj):
    nx = float(g.ide - g.ids) if g.ide > g.ids else 1.0
    ny = float(g.jde - g.jds) if g.jde > g.jds else 1.0
    return (gi - g.ids) / nx, (gj - g.jds) / ny


def _mut(g, gi, gj):
    x, y = _norm(g, gi.astype(np.float64), gj.astype(np.float64))
    return (94000.0 + 3000.0 * np.sin(_TWO_PI * 2.0 * x) * np.cos(_TWO_PI * 1.5 * y)
            + 800.0 * np.sin(_TWO_PI * 7.0 * x + 1.0) * np.sin(_TWO_PI * 5.0 * y)).astype(np.float32)


def _value2d(name, seed, g, gi, gj):
    fid = FIELD_ID[name]
    x, y = _norm(g, gi.astype(np.float64), gj.astype(np.float64))
    zero = np.zeros_like(gi)
    if name == "mut":
        return _mut(g, gi, gj)
    if name == "muu":
        return F(0.5) * (_mut(g, gi, gj) + _mut(g, gi - 1, gj))
    if name == "muv":
        return F(0.5) * (_mut(g, gi, gj) + _mut(g, gi, gj - 1))
    if name == "mu":
        return F(300.0) * _sym(seed, fid, gi, zero, gj)
    if name == "mu_tend":
        return F(0.5) * _sym(seed, fid, gi, zero, gj)
    if name == "msftx":
        return (1.0 + 0.1 * np.sin(_TWO_PI * y) * np.cos(0.5 * _TWO_PI * x)).astype(np.float32)
    if name == "msfty":
        return (1.0 + 0.1 * np.cos(_TWO_PI * 0.7 * y + 0.3) + 0.0 * x).astype(np.float32)
    if name == "msfuy":
        return (1.0 + 0.1 * np.cos(_TWO_PI * 0.7 * y + 0.3) + 0.01 * np.sin(_TWO_PI * 3.0 * x)).astype(np.float32)
    if name == "msfvx_inv":
        return (1.0 / (1.0 + 0.1 * np.sin(_TWO_PI * y - 0.2) * np.cos(0.5 * _TWO_PI * x))).astype(np.float32)
    base = {"muave": 1000.0, "muts": 2000.0, "mudf": 3000.0}[name]
    return F(base) + F(100.0) * _sym(seed, fid, gi, zero, gj)


def _value3d(name, seed, g, gi, gk, gj):
    fid = FIELD_ID[name]
    r = _sym(seed, fid, gi, gk, gj)
    if name in ("t", "ft", "ww_1", "ww", "t_ave"):
        if name == "t":
            return F(50.0) * r
        if name == "ft":
            return F(5.0) * r
        if name == "ww_1":
            return np.where(gk <= 1, F(0.0), F(0.5) * r).astype(np.float32)
        if name == "ww":
            return np.where((gk <= 1) | (gk >= g.kde), F(0.0), F(0.5) * r).astype(np.float32)
        return F(4000.0) + F(100.0) * r
    x, y = _norm(g, gi.astype(np.float64), gj.astype(np.float64))
    z = (gk.astype(np.float64) - 1.0) / float(g.kde - 1) if g.kde > 1 else 0.0 * gk
    if name == "u_1":
        return (10.0 + 20.0 * z * np.sin(_TWO_PI * (x + y))).astype(np.float32) + F(2.0) * r
    if name == "v_1":
        return (-5.0 + 20.0 * z * np.cos(_TWO_PI * (x - y))).astype(np.float32) + F(2.0) * r
    if name == "u":
        return (500.0 * np.sin(_TWO_PI * 3.0 * x) * np.cos(_TWO_PI * 2.0 * y) + 0.0 * z).astype(np.float32) + F(1500.0) * r
    if name == "v":
        return (500.0 * np.cos(_TWO_PI * 2.0 * x) * np.sin(_TWO_PI * 3.0 * y) + 0.0 * z).astype(np.float32) + F(1500.0) * r
    if name == "t_1":
        return (150.0 * z + 0.0 * x).astype(np.float32) + F(2.0) * r
    raise KeyError(name)


def _znw(g, k):
    s = (k - 1) / float(g.kde - 1) if g.kde > 1 else 0.0
    a = 2.5
    return (np.exp(-a * s) - np.exp(-a)) / (1.0 - np.exp(-a))


def _dnw(g, k):
    kk = min(max(k, 1), g.kde - 1)
    kk = max(kk, 1)
    return F(_znw(g, kk + 1) - _znw(g, kk))


def _value1d(name, g, k):
    if k < 1 or k > g.kde:
        return F(0.0)
    dnw = _dnw(g, k)
    if name == "dnw":
        return dnw
    if name == "rdnw":
        return F(1.0) / dnw
    if k < 2:
        return F(0.0)
    dnwm = _dnw(g, k - 1)
    dn = F(0.5) * (dnw + dnwm)
    return F(0.5) * dnwm / dn if name == "fnp" else F(0.5) * dnw / dn


def synth_field(name: str, g, seed: int = 20240617, threads: int = 0) -> np.ndarray:
    ni, nj, nk = g.ime - g.ims + 1, g.jme - g.jms + 1, g.kme - g.kms + 1
    if name in FIELDS_1D:
        return np.array([_value1d(name, g, g.kms + k) for k in range(nk)], dtype=np.float32)
    gi1 = np.arange(g.ims, g.ime + 1, dtype=np.int64)
    threads = threads or min(32, os.cpu_count() or 1)
    if name in FIELDS_2D:
        gj, gi = np.meshgrid(np.arange(g.jms, g.jme + 1, dtype=np.int64), gi1, indexing="ij")
        return np.ascontiguousarray(_value2d(name, seed, g, gi, gj), dtype=np.float32)
    out = np.empty((nj, nk, ni), dtype=np.float32)
    gk1 = np.arange(g.kms, g.kme + 1, dtype=np.int64)

    def rows(j0, j1):
        gj, gk, gi = np.meshgrid(np.arange(g.jms + j0, g.jms + j1, dtype=np.int64), gk1, gi1, indexing="ij")
        out[j0:j1] = _value3d(name, seed, g, gi, gk, gj)

    step = 8
    chunks = [(a, min(nj, a + step)) for a in range(0, nj, step)]
    if threads > 1 and len(chunks) > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda c: rows(*c), chunks))
    else:
        for c in chunks:
            rows(*c)
    return out


def synth_fields(g, seed: int = 20240617, names=FIELDS):
    return {n: synth_field(n, g, seed) for n in names}
