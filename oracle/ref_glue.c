/*
 * oracle/ref_glue.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Glue that lets tests and bench.py call the reference's OWN, UNMODIFIED C
 * translation of advance_mu_t (/root/reference/advance_mu_t.c:17-239) through a
 * flat C ABI.  oracle/Makefile compiles that reference source where it lies
 * (never copied into this repository) together with this file into
 * oracle/_ref/libref_advance_mu_t.so.
 *
 * The reference source needs exactly one external symbol, `alloc_mem`
 * (declared /root/reference/advance_mu_t.h:25; the reference defines it in its
 * driver, advance_mu_t_driver.c:292-300, which cannot be built: it needs the
 * un-shipped ulps.h and /data2 input files).  We supply it here.
 */
#include <stdio.h>
#include <stdlib.h>
#include "advance_mu_t.h"   /* the reference's header, found via -I/root/reference */

void *alloc_mem(int size)
{
    void *p = malloc((size_t)size);
    if (!p) { fprintf(stderr, "ref_glue: alloc_mem(%d) failed\n", size); abort(); }
    return p;
}

/* One call of the reference routine.  `kds` is an extra argument of the C
 * translation (advance_mu_t.h:10-23) that cancels algebraically
 * (advance_mu_t.c:38,45,52); the Fortran contract has kds == kms == 1. */
void ref_advance_mu_t(
    float *ww, float *ww_1, float *u, float *u_1, float *v, float *v_1,
    float *mu, float *mut, float *muave, float *muts, float *muu, float *muv, float *mudf,
    float *t, float *t_1, float *t_ave, float *ft, float *mu_tend,
    float rdx, float rdy, float dts, float epssm,
    float *dnw, float *fnm, float *fnp, float *rdnw,
    float *msfuy, float *msfvx_inv, float *msftx, float *msfty,
    int periodic_x, int specified, int nested,
    int ids, int ide, int jds, int jde, int kde,
    int ims, int ime, int jms, int jme, int kms, int kme,
    int its, int ite, int jts, int jte, int kts, int kte)
{
    config_flags cfg;
    cfg.nested = nested;
    cfg.periodic_x = periodic_x;
    cfg.specified = specified;
    advance_mu_t(ww, ww_1, u, u_1, v, v_1, mu, mut, muave, muts, muu, muv, mudf,
                 t, t_1, t_ave, ft, mu_tend, rdx, rdy, dts, epssm,
                 dnw, fnm, fnp, rdnw, msfuy, msfvx_inv, msftx, msfty, cfg,
                 ids, ide, jds, jde, /*kds=*/kms, kde,
                 ims, ime, jms, jme, kms, kme,
                 its, ite, jts, jte, kts, kte);
}

/* The reference routine called once per j-tile from an OpenMP loop: WRF's own
 * tiling scheme, as sketched (commented out) in the reference Fortran driver,
 * /root/reference/advance_mu_t_driver.f90:175-205.  The routine is re-entrant
 * per tile, so this is "the reference with all the host threads it can use". */
void ref_advance_mu_t_tiled(
    float *ww, float *ww_1, float *u, float *u_1, float *v, float *v_1,
    float *mu, float *mut, float *muave, float *muts, float *muu, float *muv, float *mudf,
    float *t, float *t_1, float *t_ave, float *ft, float *mu_tend,
    float rdx, float rdy, float dts, float epssm,
    float *dnw, float *fnm, float *fnp, float *rdnw,
    float *msfuy, float *msfvx_inv, float *msftx, float *msfty,
    int periodic_x, int specified, int nested,
    int ids, int ide, int jds, int jde, int kde,
    int ims, int ime, int jms, int jme, int kms, int kme,
    int its, int ite, int jts, int jte, int kts, int kte,
    int num_tiles)
{
    const int nj = jte - jts + 1;
    if (num_tiles < 1) num_tiles = 1;
    if (num_tiles > nj) num_tiles = nj > 0 ? nj : 1;
#pragma omp parallel for schedule(static)
    for (int tile = 0; tile < num_tiles; ++tile) {
        const int j0 = jts + (int)(((long)nj * tile) / num_tiles);
        const int j1 = jts + (int)(((long)nj * (tile + 1)) / num_tiles) - 1;
        if (j1 < j0) continue;
        ref_advance_mu_t(ww, ww_1, u, u_1, v, v_1, mu, mut, muave, muts, muu, muv, mudf,
                         t, t_1, t_ave, ft, mu_tend, rdx, rdy, dts, epssm,
                         dnw, fnm, fnp, rdnw, msfuy, msfvx_inv, msftx, msfty,
                         periodic_x, specified, nested,
                         ids, ide, jds, jde, kde, ims, ime, jms, jme, kms, kme,
                         its, ite, j0, j1, kts, kte);
    }
}
