// amt_params.h -- kernel argument block and IEEE helpers shared by the advance_mu_t kernels.
//
// Semantics follow the reference Fortran, /root/reference/module_small_step_em.f90:91-250
// (restated in SURVEY.md section 8).  All indices in here are 0-based MEMORY indices
// (Fortran index minus ims / kms / jms); the host layer (capi.cu) does that shift, like the
// reference wrapper does at advance_mu_t_no_async.cu:57-79.
#pragma once
#include <cuda.h>            // CUtensorMap (types only; the driver is reached through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>

struct AmtParams {
    // 3-D fields, (i fastest, k, j): element (i,k,j) at  j*jstride + k*pitch + i
    float *ww;
    const float *ww_1, *u, *u_1, *v, *v_1;
    float *t;
    const float *t_1;
    float *t_ave;
    const float *ft;
    // 2-D fields: element (i,j) at  j*pitch2 + i
    float *mu;
    const float *mut;
    float *muave, *muts;
    const float *muu, *muv;
    float *mudf;
    const float *mu_tend, *msfuy, *msfvx_inv, *msftx, *msfty;
    // 1-D fields, index 0 = level kms
    const float *dnw, *fnm, *fnp, *rdnw;
    float rdx, rdy, dts, epssm;
    long long pitch;    // 3-D row stride in floats
    long long jstride;  // 3-D plane stride in floats (= pitch * kdim)
    long long pitch2;   // 2-D row stride in floats
    int i0, i1;         // computed i range (i_start..i_end), inclusive, memory index
    int j0, j1;         // computed j range
    int k0;             // memory index of level kts
    int nk;             // number of computed levels: k_start..k_end = kts..kte-1
    int kdim, jdim;     // memory extents in k and j (jstride == pitch * kdim)
};

// Tensor maps of the TMA-staged kernel (amt_pipe.cu): boxes are [columns x 1 level x rows].
struct AmtTmaMaps {
    alignas(64) CUtensorMap u132;       // u    [132 x 1 x 1]
    alignas(64) CUtensorMap u1_132;     // u_1  [132 x 1 x 1]
    alignas(64) CUtensorMap v_2rows;    // v    [128 x 1 x 2]  rows j, j+1
    alignas(64) CUtensorMap v1_2rows;   // v_1  [128 x 1 x 2]
    alignas(64) CUtensorMap t1_136;     // t_1  [136 x 1 x 1]  both ring columns
    alignas(64) CUtensorMap t1_128;     // t_1  [128 x 1 x 1]
    const void *key_ptrs[5];            // what the maps were built for
    long long key_dims[3];
    int valid;
};

// Round-to-nearest single operations that the compiler may never contract into an FMA,
// whatever -fmad says: this is what keeps the result bit-identical to the Fortran's real*4
// arithmetic (the reference gets the same effect with `-fmad=false`, Makefile:12).
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }

// Launchers (defined in the kernel translation units).
cudaError_t amt_launch_column(const AmtParams &p, cudaStream_t stream);
cudaError_t amt_launch_tile(const AmtParams &p, cudaStream_t stream);
bool amt_tile_supported(const AmtParams &p);
bool amt_pipe_supported(const AmtParams &p);
bool amt_build_tma_maps(const AmtParams &p, AmtTmaMaps *maps);                    // cached: rebuilds only on change
cudaError_t amt_launch_pipe(const AmtParams &p, const AmtTmaMaps &maps, cudaStream_t stream, int cfg);   // cfg 0 = automatic
