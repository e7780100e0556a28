/*
 * oracle/ref_kernel_glue.cu -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.
 *
 * Launches the reference's OWN, UNMODIFIED CUDA-C kernel
 * (/root/reference/advance_mu_t_kernel.cu:16-198, compiled in place for sm_100a
 * with the reference's -fmad=false, see oracle/Makefile) on caller-provided
 * device buffers, so that bench.py can time "the repo's CUDA-C version" on a
 * B200 beside ours and tests can cross-check it.  The reference *host* wrapper
 * (advance_mu_t_no_async.cu) is not used: it hard-codes `#define GPUs 3`
 * (:12), allocates and copies every field on every call (:178-306) and exits
 * the process on error (:22-32).
 *
 * Launch geometry and index normalisation restate what that wrapper does:
 * grid (idim/64+1, jdim), block 64 (advance_mu_t_no_async.cu:11,54-55); indices
 * shifted to 0-based memory offsets (:57-79); scratch dvdxi/wdtn are full 3-D
 * global arrays and dmdt a 2-D one (:175-176, :221-222, :236).
 */
#include <cuda_runtime.h>
#include "config_flags.h"        /* reference headers, via -I/root/reference */
#include "advance_mu_t_cu.h"

extern "C" int ref_cuda_kernel_launch(
    float *ww, float *ww_1, float *u, float *u_1, float *v, float *v_1,
    float *mu, float *mut, float *muave, float *muts, float *muu, float *muv, float *mudf,
    float *t, float *t_1, float *t_ave, float *ft, float *mu_tend,
    float rdx, float rdy, float dts, float epssm,
    float *dnw, float *fnm, float *fnp, float *rdnw,
    float *msfuy, float *msfvx_inv, float *msftx, float *msfty,
    float *wdtn_scratch, float *dvdxi_scratch, float *dmdt_scratch,
    int periodic_x, int specified, int nested,
    int ids, int ide, int jds, int jde, int kde,
    int ims, int ime, int jms, int jme, int kms, int kme,
    int its, int ite, int jts, int jte, int kts, int kte,
    void *stream)
{
    const int idim = ime - ims + 1;
    const int kdim = kme - kms + 1;
    const int jdim = jme - jms + 1;
    const int kds = kms;

    config_flags cfg = {};
    cfg.periodic_x = periodic_x;
    cfg.specified = specified;
    cfg.nested = nested;

    /* 0-based memory offsets, as advance_mu_t_no_async.cu:57-79 */
    const int m_ids = ids - ims, m_ide = ide - ims;
    const int m_jds = jds - jms, m_jde = jde - jms;
    const int m_kds = kds - kms, m_kde = kde - kms;
    const int m_its = its - ims, m_ite = ite - ims;
    const int m_jts = jts - jms, m_jte = jte - jms;
    const int m_kts = 0, m_kte = kte - kts;

    dim3 grid(idim / 64 + 1, jdim, 1);
    dim3 block(64, 1, 1);
    advance_mu_t_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        ww, ww_1, u, u_1, v, v_1, mu, mut, muave, muts, muu, muv, mudf, t, t_1, t_ave, ft, mu_tend,
        rdx, rdy, dts, epssm, dnw, fnm, fnp, rdnw, msfuy, msfvx_inv, msftx, msfty,
        wdtn_scratch, dvdxi_scratch, dmdt_scratch, cfg,
        m_ids, m_ide, m_jds, m_jde, m_kds, m_kde, idim, jdim, kdim,
        m_its, m_ite, m_jts, m_jte, m_kts, m_kte);
    return (int)cudaGetLastError();
}
