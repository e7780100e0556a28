/*
 * oracle/advance_mu_t_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A CPU restatement of WRF's acoustic small-step `advance_mu_t` in plain C,
 * following the reference Fortran subroutine
 *     /root/reference/module_small_step_em.f90:7-252
 * (index sets :91-106, mass/omega stage :112-174, theta stage :208-250; the
 * debugging file dumps at :175-189 are a reference-only side effect and are
 * NOT reproduced).  It is the checker that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg compare the CUDA path against.  Nothing in
 * the product package may import, link or execute it.
 *
 * Parity pin: this file is checked bit-for-bit against the reference's own C
 * translation (/root/reference/advance_mu_t.c, compiled in place into
 * oracle/_ref/ by oracle/Makefile) by tests/test_oracle_pinned.py, and against
 * the committed outputs of that reference in tests/golden/.
 *
 * Arithmetic contract: IEEE-754 binary32, every operation rounded once, in the
 * order the Fortran source states (left to right, parentheses honoured).  Build
 * with -ffp-contract=off and without -ffast-math (see oracle/Makefile) so that
 * gcc neither fuses multiply-adds nor reassociates.  Every product/sum below is
 * written with explicit parentheses that spell out the Fortran association.
 *
 * All indices are Fortran-numbered, exactly as the subroutine receives them.
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK            0
#define ORACLE_ERR_ARGS      1
#define ORACLE_ERR_NOMEM     2

typedef struct {
    int ims, jms, kms;
    size_t ni, nk;         /* memory extents in i and k */
} oracle_shape;

/* (i,k,j) with i fastest, then k, then j: module_small_step_em.f90:30-44 */
static inline size_t at3(const oracle_shape *s, int i, int k, int j) {
    return ((size_t)(j - s->jms) * s->nk + (size_t)(k - s->kms)) * s->ni + (size_t)(i - s->ims);
}
/* (i,j): module_small_step_em.f90:46-59 */
static inline size_t at2(const oracle_shape *s, int i, int j) {
    return (size_t)(j - s->jms) * s->ni + (size_t)(i - s->ims);
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* Index sets, module_small_step_em.f90:91-106. */
void oracle_advance_mu_t_bounds(int periodic_x, int specified, int nested,
                                int ids, int ide, int jds, int jde,
                                int its, int ite, int jts, int jte, int kts, int kte,
                                int *i_start, int *i_end, int *j_start, int *j_end,
                                int *k_start, int *k_end)
{
    int is = its, ie = imin(ite, ide - 1);           /* :91-92 */
    int js = jts, je = imin(jte, jde - 1);           /* :93-94 */
    if (!periodic_x) {                               /* :97 */
        if (specified || nested) {                   /* :98 */
            is = imax(its, ids + 1);                 /* :99 */
            ie = imin(ite, ide - 2);                 /* :100 */
        }
    }
    if (specified || nested) {                       /* :103 */
        js = imax(jts, jds + 1);                     /* :104 */
        je = imin(jte, jde - 2);                     /* :105 */
    }
    *i_start = is; *i_end = ie; *j_start = js; *j_end = je;
    *k_start = kts; *k_end = kte - 1;                /* :95-96 */
}

int oracle_advance_mu_t(
    float *ww, const float *ww_1, const float *u, const float *u_1,
    const float *v, const float *v_1,
    float *mu, const float *mut, float *muave, float *muts,
    const float *muu, const float *muv, float *mudf,
    float *t, const float *t_1, float *t_ave, const float *ft, const float *mu_tend,
    float rdx, float rdy, float dts, float epssm,
    const float *dnw, const float *fnm, const float *fnp, const float *rdnw,
    const float *msfuy, const float *msfvx_inv, const float *msftx, const float *msfty,
    int periodic_x, int specified, int nested,
    int ids, int ide, int jds, int jde, int kde,
    int ims, int ime, int jms, int jme, int kms, int kme,
    int its, int ite, int jts, int jte, int kts, int kte)
{
    /* The Fortran uses the literals k=1 and k=2 (:159,168,209,220,224,234) and
     * indexes its (kts:kte) scratch at kde (:221): it is only meaningful for
     * kms<=kts==1 and kde==kte. */
    if (kts != 1 || kms > 1 || kde != kte || kme < kte) return ORACLE_ERR_ARGS;
    if (ime < ims || jme < jms) return ORACLE_ERR_ARGS;

    oracle_shape s;
    s.ims = ims; s.jms = jms; s.kms = kms;
    s.ni = (size_t)(ime - ims + 1);
    s.nk = (size_t)(kme - kms + 1);

    int i_start, i_end, j_start, j_end, k_start, k_end;
    oracle_advance_mu_t_bounds(periodic_x, specified, nested, ids, ide, jds, jde,
                               its, ite, jts, jte, kts, kte,
                               &i_start, &i_end, &j_start, &j_end, &k_start, &k_end);
    if (i_start > i_end || j_start > j_end) return ORACLE_OK;   /* empty tile */
    /* one-cell ring of memory is read around the compute range (:143-146, :241-245) */
    if (i_start - 1 < ims || i_end + 1 > ime || j_start - 1 < jms || j_end + 1 > jme)
        return ORACLE_ERR_ARGS;

    /* Local arrays "from the stack (note tile size)", :74-75.
     * dvdxi(i,k), wdtn(i,k) for i in its:ite, k in kts:kte; dmdt(i). */
    const size_t ti = (size_t)(ite - its + 1);
    const size_t tk = (size_t)(kte - kts + 1);
    float *dvdxi = (float *)malloc(ti * tk * sizeof(float));
    float *wdtn  = (float *)malloc(ti * tk * sizeof(float));
    float *dmdt  = (float *)malloc(ti * sizeof(float));
    if (!dvdxi || !wdtn || !dmdt) { free(dvdxi); free(wdtn); free(dmdt); return ORACLE_ERR_NOMEM; }
#define LOC(i, k) ((size_t)((k) - kts) * ti + (size_t)((i) - its))

    /* ---- CALCULATION OF WW (dETA/dt), :112-174 ---- */
    for (int j = j_start; j <= j_end; ++j) {
        for (int i = i_start; i <= i_end; ++i) dmdt[i - its] = 0.0f;          /* :114-116 */

        for (int k = k_start; k <= k_end; ++k) {                              /* :140 */
            for (int i = i_start; i <= i_end; ++i) {                          /* :141 */
                /* :142-146 */
                const float cof = msftx[at2(&s, i, j)] * msfty[at2(&s, i, j)];
                const float vn = v[at3(&s, i, k, j + 1)]
                               + (muv[at2(&s, i, j + 1)] * v_1[at3(&s, i, k, j + 1)]) * msfvx_inv[at2(&s, i, j + 1)];
                const float vs = v[at3(&s, i, k, j)]
                               + (muv[at2(&s, i, j)] * v_1[at3(&s, i, k, j)]) * msfvx_inv[at2(&s, i, j)];
                const float ue = u[at3(&s, i + 1, k, j)]
                               + (muu[at2(&s, i + 1, j)] * u_1[at3(&s, i + 1, k, j)]) / msfuy[at2(&s, i + 1, j)];
                const float uw = u[at3(&s, i, k, j)]
                               + (muu[at2(&s, i, j)] * u_1[at3(&s, i, k, j)]) / msfuy[at2(&s, i, j)];
                const float dv = cof * ((rdy * (vn - vs)) + (rdx * (ue - uw)));
                dvdxi[LOC(i, k)] = dv;
                dmdt[i - its] = dmdt[i - its] + (dnw[k - kms] * dv);          /* :147 */
            }
        }

        for (int i = i_start; i <= i_end; ++i) {                              /* :151-157 */
            const size_t c = at2(&s, i, j);
            const float mu_old = mu[c];                                       /* :152 (held in muave) */
            const float tend = dmdt[i - its] + mu_tend[c];
            const float mu_new = mu_old + (dts * tend);                       /* :153 */
            mu[c] = mu_new;
            mudf[c] = tend;                                                   /* :154 */
            muts[c] = mut[c] + mu_new;                                        /* :155 */
            muave[c] = 0.5f * (((1.0f + epssm) * mu_new) + ((1.0f - epssm) * mu_old));  /* :156 */
        }

        for (int k = 2; k <= k_end; ++k) {                                    /* :159 */
            for (int i = i_start; i <= i_end; ++i) {
                const size_t c = at2(&s, i, j);
                /* :161  ww(k)=ww(k-1)-dnw(k-1)*(dmdt+dvdxi(k-1)+mu_tend)/msfty */
                const float inner = (dmdt[i - its] + dvdxi[LOC(i, k - 1)]) + mu_tend[c];
                ww[at3(&s, i, k, j)] = ww[at3(&s, i, k - 1, j)] - ((dnw[k - 1 - kms] * inner) / msfty[c]);
            }
        }

        for (int k = 1; k <= k_end; ++k) {                                    /* :168-172 */
            for (int i = i_start; i <= i_end; ++i) {
                const size_t c3 = at3(&s, i, k, j);
                ww[c3] = ww[c3] - ww_1[c3];                                   /* :170 */
            }
        }
    }

    /* ---- CALCULATION OF THETA, :208-250 ---- */
    for (int j = j_start; j <= j_end; ++j) {                                  /* :208-215 */
        for (int k = 1; k <= k_end; ++k) {
            for (int i = i_start; i <= i_end; ++i) {
                const size_t c3 = at3(&s, i, k, j);
                t_ave[c3] = t[c3];                                            /* :211 */
                t[c3] = t[c3] + ((msfty[at2(&s, i, j)] * dts) * ft[c3]);      /* :212 */
            }
        }
    }

    for (int j = j_start; j <= j_end; ++j) {                                  /* :217 */
        for (int i = i_start; i <= i_end; ++i) {                              /* :219-222 */
            wdtn[LOC(i, 1)] = 0.0f;
            wdtn[LOC(i, kde)] = 0.0f;
        }
        for (int k = 2; k <= k_end; ++k) {                                    /* :224-229 */
            for (int i = i_start; i <= i_end; ++i) {
                wdtn[LOC(i, k)] = ww[at3(&s, i, k, j)]
                                * ((fnm[k - kms] * t_1[at3(&s, i, k, j)]) + (fnp[k - kms] * t_1[at3(&s, i, k - 1, j)]));
            }
        }
        for (int k = 1; k <= k_end; ++k) {                                    /* :234-248 */
            for (int i = i_start; i <= i_end; ++i) {
                const size_t c2 = at2(&s, i, j);
                const size_t c3 = at3(&s, i, k, j);
                const float tc = t_1[c3];
                const float fy = (0.5f * rdy) * ((v[at3(&s, i, k, j + 1)] * (t_1[at3(&s, i, k, j + 1)] + tc))
                                               - (v[c3] * (tc + t_1[at3(&s, i, k, j - 1)])));      /* :240-242 */
                const float fx = (0.5f * rdx) * ((u[at3(&s, i + 1, k, j)] * (t_1[at3(&s, i + 1, k, j)] + tc))
                                               - (u[c3] * (tc + t_1[at3(&s, i - 1, k, j)])));      /* :243-245 */
                const float fz = rdnw[k - kms] * (wdtn[LOC(i, k + 1)] - wdtn[LOC(i, k)]);          /* :246 */
                t[c3] = t[c3] - ((dts * msfty[c2]) * ((msftx[c2] * (fy + fx)) + fz));              /* :237-246 */
            }
        }
    }
#undef LOC
    free(dvdxi); free(wdtn); free(dmdt);
    return ORACLE_OK;
}

/*
 * The same routine driven over j-tiles, one call per tile, mirroring the
 * (commented-out) OpenMP tile loop of the reference Fortran driver,
 * /root/reference/advance_mu_t_driver.f90:175-205.  Each tile call is an
 * independent legal call of the subroutine (its:ite x jts:jte sub-range), so the
 * result is bit-identical to the single call.  Used as the all-cores CPU port.
 */
int oracle_advance_mu_t_tiled(
    float *ww, const float *ww_1, const float *u, const float *u_1,
    const float *v, const float *v_1,
    float *mu, const float *mut, float *muave, float *muts,
    const float *muu, const float *muv, float *mudf,
    float *t, const float *t_1, float *t_ave, const float *ft, const float *mu_tend,
    float rdx, float rdy, float dts, float epssm,
    const float *dnw, const float *fnm, const float *fnp, const float *rdnw,
    const float *msfuy, const float *msfvx_inv, const float *msftx, const float *msfty,
    int periodic_x, int specified, int nested,
    int ids, int ide, int jds, int jde, int kde,
    int ims, int ime, int jms, int jme, int kms, int kme,
    int its, int ite, int jts, int jte, int kts, int kte,
    int num_tiles)
{
    if (num_tiles < 1) num_tiles = 1;
    const int nj = jte - jts + 1;
    if (num_tiles > nj) num_tiles = nj > 0 ? nj : 1;
    int status = ORACLE_OK;
#pragma omp parallel for schedule(static) reduction(max : status)
    for (int tile = 0; tile < num_tiles; ++tile) {
        const int j0 = jts + (int)(((long)nj * tile) / num_tiles);
        const int j1 = jts + (int)(((long)nj * (tile + 1)) / num_tiles) - 1;
        if (j1 < j0) continue;
        int rc = oracle_advance_mu_t(ww, ww_1, u, u_1, v, v_1, mu, mut, muave, muts, muu, muv, mudf,
                                     t, t_1, t_ave, ft, mu_tend, rdx, rdy, dts, epssm,
                                     dnw, fnm, fnp, rdnw, msfuy, msfvx_inv, msftx, msfty,
                                     periodic_x, specified, nested,
                                     ids, ide, jds, jde, kde, ims, ime, jms, jme, kms, kme,
                                     its, ite, j0, j1, kts, kte);
        if (rc > status) status = rc;
    }
    return status;
}
