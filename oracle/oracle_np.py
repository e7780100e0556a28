"""oracle/oracle_np.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

A second, independent CPU restatement of WRF's ``advance_mu_t`` in numpy float32,
vectorised over (i, j) with an explicit *sequential* k loop so that the column
sum ``dmdt`` and the ``ww`` prefix keep the Fortran summation order.

Follows /root/reference/module_small_step_em.f90:91-106 (index sets),
:112-174 (mass / omega), :208-250 (theta).  Lines :175-189 (debug file dumps)
are not reproduced.  Every binary operation is a float32 numpy ufunc call, i.e.
one IEEE-754 binary32 rounding per operation, in the order the Fortran states.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import this module.  Arrays use numpy C order ``[j, k, i]`` which is byte-identical
to the Fortran ``(i, k, j)`` layout (i fastest).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def bounds(periodic_x, specified, nested, ids, ide, jds, jde, its, ite, jts, jte, kts, kte):
    """Index sets, module_small_step_em.f90:91-106 (Fortran numbering, inclusive)."""
    i_start, i_end = its, min(ite, ide - 1)
    j_start, j_end = jts, min(jte, jde - 1)
    if not periodic_x and (specified or nested):
        i_start, i_end = max(its, ids + 1), min(ite, ide - 2)
    if specified or nested:
        j_start, j_end = max(jts, jds + 1), min(jte, jde - 2)
    return i_start, i_end, j_start, j_end, kts, kte - 1


def advance_mu_t(ww, ww_1, u, u_1, v, v_1, mu, mut, muave, muts, muu, muv, mudf,
                 t, t_1, t_ave, ft, mu_tend, rdx, rdy, dts, epssm,
                 dnw, fnm, fnp, rdnw, msfuy, msfvx_inv, msftx, msfty,
                 periodic_x, specified, nested,
                 ids, ide, jds, jde, kde, ims, ime, jms, jme, kms, kme,
                 its, ite, jts, jte, kts, kte):
    """In-place update of ww, mu, muave, muts, mudf, t, t_ave (float32 ``[j,k,i]`` / ``[j,i]``)."""
    assert kts == 1 and kms <= 1 and kde == kte
    for a in (ww, ww_1, u, u_1, v, v_1, t, t_1, t_ave, ft):
        assert a.dtype == np.float32 and a.shape == (jme - jms + 1, kme - kms + 1, ime - ims + 1)
    i_start, i_end, j_start, j_end, k_start, k_end = bounds(
        periodic_x, specified, nested, ids, ide, jds, jde, its, ite, jts, jte, kts, kte)
    if i_start > i_end or j_start > j_end:
        return
    rdx, rdy, dts, epssm = F(rdx), F(rdy), F(dts), F(epssm)

    # numpy slices over the compute range and its one-cell shifts
    I = slice(i_start - ims, i_end - ims + 1)
    Ip = slice(i_start - ims + 1, i_end - ims + 2)
    Im = slice(i_start - ims - 1, i_end - ims)
    J = slice(j_start - jms, j_end - jms + 1)
    Jp = slice(j_start - jms + 1, j_end - jms + 2)
    Jm = slice(j_start - jms - 1, j_end - jms)

    def K(k):  # Fortran level -> memory index
        return k - kms

    cof = msftx[J, I] * msfty[J, I]
    nlev = k_end - k_start + 1
    dvdxi = np.empty((nlev,) + cof.shape, dtype=np.float32)
    dmdt = np.zeros_like(cof)                                              # :114-116
    for k in range(k_start, k_end + 1):                                     # :140-149
        vn = v[Jp, K(k), I] + (muv[Jp, I] * v_1[Jp, K(k), I]) * msfvx_inv[Jp, I]
        vs = v[J, K(k), I] + (muv[J, I] * v_1[J, K(k), I]) * msfvx_inv[J, I]
        ue = u[J, K(k), Ip] + (muu[J, Ip] * u_1[J, K(k), Ip]) / msfuy[J, Ip]
        uw = u[J, K(k), I] + (muu[J, I] * u_1[J, K(k), I]) / msfuy[J, I]
        dv = cof * ((rdy * (vn - vs)) + (rdx * (ue - uw)))
        dvdxi[k - k_start] = dv
        dmdt = dmdt + dnw[K(k)] * dv

    mu_old = mu[J, I].copy()                                                # :151-157
    tend = dmdt + mu_tend[J, I]
    mu_new = mu_old + dts * tend
    mu[J, I] = mu_new
    mudf[J, I] = tend
    muts[J, I] = mut[J, I] + mu_new
    muave[J, I] = F(0.5) * (((F(1.0) + epssm) * mu_new) + ((F(1.0) - epssm) * mu_old))

    for k in range(2, k_end + 1):                                           # :159-163
        inner = (dmdt + dvdxi[k - 1 - k_start]) + mu_tend[J, I]
        ww[J, K(k), I] = ww[J, K(k - 1), I] - (dnw[K(k - 1)] * inner) / msfty[J, I]
    for k in range(1, k_end + 1):                                           # :168-172
        ww[J, K(k), I] = ww[J, K(k), I] - ww_1[J, K(k), I]

    for k in range(1, k_end + 1):                                           # :208-215
        t_ave[J, K(k), I] = t[J, K(k), I]
        t[J, K(k), I] = t[J, K(k), I] + (msfty[J, I] * dts) * ft[J, K(k), I]

    wdtn = np.zeros((kde + 2,) + cof.shape, dtype=np.float32)               # wdtn(1)=wdtn(kde)=0, :219-222
    for k in range(2, k_end + 1):                                           # :224-229
        wdtn[k] = ww[J, K(k), I] * ((fnm[K(k)] * t_1[J, K(k), I]) + (fnp[K(k)] * t_1[J, K(k - 1), I]))
    hrdy = F(0.5) * rdy
    hrdx = F(0.5) * rdx
    for k in range(1, k_end + 1):                                           # :234-248
        tc = t_1[J, K(k), I]
        fy = hrdy * ((v[Jp, K(k), I] * (t_1[Jp, K(k), I] + tc)) - (v[J, K(k), I] * (tc + t_1[Jm, K(k), I])))
        fx = hrdx * ((u[J, K(k), Ip] * (t_1[J, K(k), Ip] + tc)) - (u[J, K(k), I] * (tc + t_1[J, K(k), Im])))
        fz = rdnw[K(k)] * (wdtn[k + 1] - wdtn[k])
        t[J, K(k), I] = t[J, K(k), I] - (dts * msfty[J, I]) * ((msftx[J, I] * (fy + fx)) + fz)
