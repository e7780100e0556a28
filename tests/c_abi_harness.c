/* tests/c_abi_harness.c -- calls libwrfb200.so exactly the way the Fortran shim
 * (wrf_model_cuda_sample_b200/fortran/module_small_step_em.F90) does: arrays by reference in the Fortran
 * order, scalars by value, Fortran-numbered indices.  No Fortran compiler exists in the build image, so this
 * C program is what enforces the shim's contract ("Hard parts").
 *
 *   gcc tests/c_abi_harness.c -Iinclude -Lwrf_model_cuda_sample_b200 -lwrfb200 -o /tmp/c_abi_harness
 *   /tmp/c_abi_harness            -> prints a checksum of every output field (needs a GPU), or, without a
 *                                    GPU, the library's "no CPU fallback" error and exit code 3.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "wrfb200.h"

int main(void)
{
    const int ids = 1, ide = 40, jds = 1, jde = 30, kde = 12;
    const int ims = -2, ime = 43, jms = -2, jme = 33, kms = 1, kme = 12;
    wrfb200_domain dom = {ids, ide, jds, jde, kde, ims, ime, jms, jme, kms, kme, 0, 1, 0};
    const size_t n3 = (size_t)(ime - ims + 1) * (kme - kms + 1) * (jme - jms + 1);
    const size_t n2 = (size_t)(ime - ims + 1) * (jme - jms + 1);
    const size_t n1 = (size_t)(kme - kms + 1);
    float *f[WRFB200_NUM_FIELDS];
    for (int i = 0; i < WRFB200_NUM_FIELDS; ++i) {
        size_t n = i < WRFB200_NUM_3D ? n3 : i < WRFB200_NUM_3D + WRFB200_NUM_2D ? n2 : n1;
        f[i] = (float *)malloc(n * sizeof(float));
        if (wrfb200_synth_field(i, 20240617ull, &dom, 12000.0f, f[i]) != WRFB200_OK) {
            fprintf(stderr, "synth: %s\n", wrfb200_last_error());
            return 2;
        }
    }
    int rc = wrfb200_advance_mu_t(
        f[WRFB200_WW], f[WRFB200_WW_1], f[WRFB200_U], f[WRFB200_U_1], f[WRFB200_V], f[WRFB200_V_1],
        f[WRFB200_MU], f[WRFB200_MUT], f[WRFB200_MUAVE], f[WRFB200_MUTS], f[WRFB200_MUU], f[WRFB200_MUV],
        f[WRFB200_MUDF], f[WRFB200_T], f[WRFB200_T_1], f[WRFB200_T_AVE], f[WRFB200_FT], f[WRFB200_MU_TEND],
        1.0f / 12000.0f, 1.0f / 12000.0f, 12.0f, 0.1f,
        f[WRFB200_DNW], f[WRFB200_FNM], f[WRFB200_FNP], f[WRFB200_RDNW],
        f[WRFB200_MSFUY], f[WRFB200_MSFVX_INV], f[WRFB200_MSFTX], f[WRFB200_MSFTY],
        0, 1, 0, ids, ide, jds, jde, kde, ims, ime, jms, jme, kms, kme,
        ids, ide, jds, jde, 1, kde);
    if (rc != WRFB200_OK) {
        fprintf(stderr, "wrfb200_advance_mu_t: status %d: %s\n", rc, wrfb200_last_error());
        return rc;
    }
    const int outs[] = {WRFB200_WW, WRFB200_T, WRFB200_T_AVE, WRFB200_MU, WRFB200_MUAVE, WRFB200_MUTS, WRFB200_MUDF};
    for (unsigned q = 0; q < sizeof(outs) / sizeof(outs[0]); ++q) {
        const size_t n = outs[q] < WRFB200_NUM_3D ? n3 : n2;
        unsigned long long h = 1469598103934665603ull;           /* FNV-1a over the raw bytes */
        const unsigned char *b = (const unsigned char *)f[outs[q]];
        for (size_t x = 0; x < n * 4; ++x) { h ^= b[x]; h *= 1099511628211ull; }
        printf("field %d fnv1a %016llx\n", outs[q], h);
    }
    return 0;
}
