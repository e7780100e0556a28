// amt_column.cu -- advance_mu_t, one thread per (i,j) column.  Works on ANY layout (no alignment
// or pitch requirement), so it serves caller-owned device arrays in the dense Fortran layout and is
// the simple cross-check for the tile kernel.  sm_100a.
//
// Per column ("Exact semantics"; module_small_step_em.f90 lines cited inline):
//   sweep 1 (k ascending): dvdxi(k) -> shared memory, dmdt += dnw(k)*dvdxi(k)
//   2-D update of mu, mudf, muts, muave
//   sweep 2 (k ascending): ww prefix, ww -= ww_1, t_ave, t update with the vertical flux wdtn
// The reference kernel keeps dvdxi/wdtn as full 3-D global arrays and dmdt as a 2-D global array
// (advance_mu_t_kernel.cu:86,112,117,164-171) and sweeps k six times; here dvdxi lives in shared
// memory ([k][thread], conflict-free), wdtn and dmdt in registers, and k is swept twice.
#include "amt_params.h"

namespace {

constexpr int kColThreads = 128;

__global__ void __launch_bounds__(kColThreads)
amt_column_kernel(const AmtParams p, const int nbx)
{
    extern __shared__ float s_dvdxi[];                  // [nk][kColThreads]
    const int tid = threadIdx.x;
    const int bx = blockIdx.x % nbx;
    const int by = blockIdx.x / nbx;
    const int i = p.i0 + bx * kColThreads + tid;
    const int j = p.j0 + by;
    if (i > p.i1) return;                               // no block-wide barrier is used below

    const long long c2 = (long long)j * p.pitch2 + i;
    const float msftx = p.msftx[c2];
    const float msfty = p.msfty[c2];
    const float cof = f_mul(msftx, msfty);              // :142 msftx*msfty*( ... )
    const float muv_s = p.muv[c2],           muv_n = p.muv[c2 + p.pitch2];
    const float mvi_s = p.msfvx_inv[c2],     mvi_n = p.msfvx_inv[c2 + p.pitch2];
    const float muu_w = p.muu[c2],           muu_e = p.muu[c2 + 1];
    const float mfu_w = p.msfuy[c2],         mfu_e = p.msfuy[c2 + 1];
    const float mu_tend = p.mu_tend[c2];

    const long long base = (long long)j * p.jstride + (long long)p.k0 * p.pitch + i;
    const float *dnw = p.dnw + p.k0;

    // ---- sweep 1: :140-149 ----
    float dmdt = 0.0f;                                  // :115
#pragma unroll 4
    for (int k = 0; k < p.nk; ++k) {
        const long long o = base + (long long)k * p.pitch;
        const float vn = f_add(p.v[o + p.jstride], f_mul(f_mul(muv_n, p.v_1[o + p.jstride]), mvi_n));   // :143
        const float vs = f_add(p.v[o],             f_mul(f_mul(muv_s, p.v_1[o]), mvi_s));               // :144
        const float ue = f_add(p.u[o + 1],         f_div(f_mul(muu_e, p.u_1[o + 1]), mfu_e));           // :145
        const float uw = f_add(p.u[o],             f_div(f_mul(muu_w, p.u_1[o]), mfu_w));               // :146
        const float dv = f_mul(cof, f_add(f_mul(p.rdy, f_sub(vn, vs)), f_mul(p.rdx, f_sub(ue, uw))));
        s_dvdxi[k * kColThreads + tid] = dv;
        dmdt = f_add(dmdt, f_mul(dnw[k], dv));          // :147
    }

    // ---- 2-D update: :151-157 ----
    const float mu_old = p.mu[c2];
    const float tend = f_add(dmdt, mu_tend);
    const float mu_new = f_add(mu_old, f_mul(p.dts, tend));                     // :153
    p.mu[c2] = mu_new;
    p.mudf[c2] = tend;                                                          // :154
    p.muts[c2] = f_add(p.mut[c2], mu_new);                                      // :155
    p.muave[c2] = f_mul(0.5f, f_add(f_mul(f_add(1.0f, p.epssm), mu_new),
                                    f_mul(f_sub(1.0f, p.epssm), mu_old)));      // :156

    // ---- sweep 2: ww prefix :159-172, theta :208-248, fused with a one-level look-ahead ----
    const float *fnm = p.fnm + p.k0, *fnp = p.fnp + p.k0, *rdnw = p.rdnw + p.k0;
    const float dts_msfty = f_mul(p.dts, msfty);        // :237 dts*msfty  (== msfty*dts of :212)
    const float hrdy = f_mul(0.5f, p.rdy);              // :240 .5*rdy
    const float hrdx = f_mul(0.5f, p.rdx);              // :243 .5*rdx

    float w_raw = p.ww[base];                           // ww(i,1,j): input value, never re-integrated (:159 starts at k=2)
    float w_fin = f_sub(w_raw, p.ww_1[base]);           // :170 at k=1
    float wdtn_k = 0.0f;                                // :220 wdtn(i,1)=0
    float t1_c = p.t_1[base];                           // t_1(i,k,j)
    for (int k = 0; k < p.nk; ++k) {
        const long long o = base + (long long)k * p.pitch;
        // level k+1 of the prefix, its final value and the flux through the top face of level k
        float w_raw_n = 0.0f, w_fin_n = 0.0f, wdtn_n = 0.0f, t1_n = 0.0f;       // :221 wdtn(i,kde)=0
        if (k + 1 < p.nk) {
            const float inner = f_add(f_add(dmdt, s_dvdxi[k * kColThreads + tid]), mu_tend);
            w_raw_n = f_sub(w_raw, f_div(f_mul(dnw[k], inner), msfty));         // :161
            w_fin_n = f_sub(w_raw_n, p.ww_1[o + p.pitch]);                      // :170
            t1_n = p.t_1[o + p.pitch];
            wdtn_n = f_mul(w_fin_n, f_add(f_mul(fnm[k + 1], t1_n), f_mul(fnp[k + 1], t1_c)));   // :227
        }
        const float t_old = p.t[o];
        const float t_mid = f_add(t_old, f_mul(dts_msfty, p.ft[o]));            // :212
        const float fy = f_mul(hrdy, f_sub(f_mul(p.v[o + p.jstride], f_add(p.t_1[o + p.jstride], t1_c)),
                                           f_mul(p.v[o], f_add(t1_c, p.t_1[o - p.jstride]))));  // :240-242
        const float fx = f_mul(hrdx, f_sub(f_mul(p.u[o + 1], f_add(p.t_1[o + 1], t1_c)),
                                           f_mul(p.u[o], f_add(t1_c, p.t_1[o - 1]))));          // :243-245
        const float fz = f_mul(rdnw[k], f_sub(wdtn_n, wdtn_k));                                  // :246
        p.ww[o] = w_fin;
        p.t_ave[o] = t_old;                                                                      // :211
        p.t[o] = f_sub(t_mid, f_mul(dts_msfty, f_add(f_mul(msftx, f_add(fy, fx)), fz)));         // :237
        w_raw = w_raw_n; w_fin = w_fin_n; wdtn_k = wdtn_n; t1_c = t1_n;
    }
}

}  // namespace

cudaError_t amt_launch_column(const AmtParams &p, cudaStream_t stream)
{
    const int ni = p.i1 - p.i0 + 1, nj = p.j1 - p.j0 + 1;
    if (ni <= 0 || nj <= 0 || p.nk <= 0) return cudaSuccess;
    const int nbx = (ni + kColThreads - 1) / kColThreads;
    const size_t smem = (size_t)p.nk * kColThreads * sizeof(float);
    if (smem > (size_t)kMaxDynSmemOptIn) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {                   // opt in once per device; the limit is only ever raised
        static bool raised[64] = {};
        cudaError_t e = amt_raise_smem_limit(amt_column_kernel, raised);
        if (e != cudaSuccess) return e;
    }
    (void)cudaGetLastError();   // a launch status must not inherit a stale error of some earlier, unrelated call
    amt_column_kernel<<<(unsigned)((long long)nbx * nj), kColThreads, smem, stream>>>(p, nbx);
    return cudaGetLastError();
}
