/*
 * wrfb200.h -- C ABI of the B200-native `advance_mu_t` (WRF acoustic small step).
 *
 * This is the drop-in boundary: plain pointers, ints and floats, no C++ or torch
 * types.  Everything is exported from libwrfb200.so (built from
 * wrf_model_cuda_sample_b200/csrc/ by __graft_entry__.build()).
 *
 * What each entry point replaces in the reference (lydia-schiff/wrf-model-cuda-sample):
 *
 *   wrfb200_advance_mu_t            the operator itself:
 *                                     Fortran  module_small_step_em.f90:7-18   (48 dummy arguments)
 *                                     C        advance_mu_t.h:10-23
 *                                     CUDA     advance_mu_t_cu.h:3-17 / advance_mu_t_no_async.cu:35-48
 *   wrfb200_create/destroy/upload/
 *   download/step/...               the CUDA host layer advance_mu_t_no_async.cu:35-424
 *                                   (per-call cudaMalloc :178-244, H2D :245-306, launch :329-353,
 *                                   sync :354-357, D2H :366-390, free :392-423), split so that state
 *                                   stays device-resident across the acoustic sub-steps
 *   wrfb200_pack_halo/unpack_halo   the host re-upload of overlapping j-slabs that stands in for a halo
 *                                   exchange, advance_mu_t_no_async.cu:87-162, :276-298
 *   wrfb200_compare                 the error report of common.cu:68-164 (`compare`), same metric set
 *   wrfb200_last_error              HANDLE_ERROR's print+exit, advance_mu_t_no_async.cu:22-32
 *                                   (a library never calls exit(); it returns a status)
 *
 * Conventions
 *   - All grid indices are Fortran-numbered and inclusive, exactly as the Fortran subroutine receives
 *     them (ids..kte).  There is no `kds` argument (the Fortran has none; in the C translation it cancels,
 *     advance_mu_t.c:38,45,52).
 *   - 3-D fields are (ims:ime, kms:kme, jms:jme) with i fastest, then k, then j
 *     (module_small_step_em.f90:30-44); 2-D fields are (ims:ime, jms:jme) (:46-59); 1-D are (kms:kme) (:61-64).
 *   - Preconditions taken from the reference code itself: kts == 1 and kms <= 1 (literal k=1 / k=2 loops,
 *     module_small_step_em.f90:159,168,209,220,224,234), kde == kte (scratch indexed at kde, :221),
 *     and a one-cell ring of memory around the computed range (ims <= i_start-1, ime >= i_end+1, same in j).
 *   - Every function returns a wrfb200_status (0 = ok).  After a non-zero status,
 *     wrfb200_last_error() gives a human-readable message (thread-local).
 *   - No CPU fallback exists: without a CUDA device every compute entry point fails with WRFB200_ERR_CUDA.
 */
#ifndef WRFB200_H
#define WRFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WRFB200_VERSION 100

typedef enum wrfb200_status {
    WRFB200_OK = 0,
    WRFB200_ERR_INVALID_ARG = 1,   /* bad pointer / index set / field id */
    WRFB200_ERR_UNSUPPORTED = 2,   /* violates a precondition listed above */
    WRFB200_ERR_CUDA = 3,          /* CUDA runtime error (message has the detail) */
    WRFB200_ERR_NOMEM = 4,
    WRFB200_ERR_STATE = 5          /* e.g. step before upload, field not bound */
} wrfb200_status;

/* Field ids, in the order of the Fortran argument list within each rank. */
typedef enum wrfb200_field {
    /* 3-D (i,k,j) */
    WRFB200_WW = 0, WRFB200_WW_1, WRFB200_U, WRFB200_U_1, WRFB200_V, WRFB200_V_1,
    WRFB200_T, WRFB200_T_1, WRFB200_T_AVE, WRFB200_FT,
    /* 2-D (i,j) */
    WRFB200_MU, WRFB200_MUT, WRFB200_MUAVE, WRFB200_MUTS, WRFB200_MUU, WRFB200_MUV,
    WRFB200_MUDF, WRFB200_MU_TEND, WRFB200_MSFUY, WRFB200_MSFVX_INV, WRFB200_MSFTX, WRFB200_MSFTY,
    /* 1-D (k) */
    WRFB200_DNW, WRFB200_FNM, WRFB200_FNP, WRFB200_RDNW,
    WRFB200_NUM_FIELDS
} wrfb200_field;
#define WRFB200_NUM_3D 10
#define WRFB200_NUM_2D 12
#define WRFB200_NUM_1D 4

/* Domain / memory description of one patch (one rank).  d = domain (global), m = memory (patch + halo). */
typedef struct wrfb200_domain {
    int ids, ide, jds, jde, kde;
    int ims, ime, jms, jme, kms, kme;
    int periodic_x, specified, nested;   /* the three grid_config_rec_type members the routine reads,
                                            module_configure.f90:434,436,447 */
} wrfb200_domain;

/* Kernel selection (testing / benchmarking; AUTO is the product default). */
typedef enum wrfb200_kernel {
    WRFB200_KERNEL_AUTO = 0,
    WRFB200_KERNEL_COLUMN = 1,   /* one thread per (i,j) column, any layout */
    WRFB200_KERNEL_TILE = 2,     /* k-parallel float4 tile kernel (register staged), needs 16-byte aligned rows */
    WRFB200_KERNEL_PIPE = 3      /* tile kernel with per-warp TMA bulk-copy pipelines (the AUTO choice when the
                                    layout allows it) */
} wrfb200_kernel;

/* Halo sides of a patch. */
typedef enum wrfb200_side { WRFB200_WEST = 0, WRFB200_EAST = 1, WRFB200_SOUTH = 2, WRFB200_NORTH = 3 } wrfb200_side;

typedef struct wrfb200_handle wrfb200_handle;

/* ------------------------------------------------------------------------------------------------
 * 1. The operator, with the reference Fortran subroutine's argument list
 *    (module_small_step_em.f90:7-18; config_flags expanded to its three used members).
 *
 *    Pointers may be HOST pointers (compat mode: upload -> step -> download of the written ranges ->
 *    sync; device mirrors are cached per thread and re-used by later calls with the same extents) or
 *    DEVICE pointers (launched in place on the current wrfb200 stream, returns without host sync).
 *    All array pointers of one call must be of the same kind.  Writes exactly the cells the Fortran
 *    writes: ww,t,t_ave for k<=kte-1 and mu,muave,muts,mudf, inside i_start..i_end x j_start..j_end only.
 * ---------------------------------------------------------------------------------------------- */
int wrfb200_advance_mu_t(
    float *ww, const float *ww_1, const float *u, const float *u_1, const float *v, const float *v_1,
    float *mu, const float *mut, float *muave, float *muts, const float *muu, const float *muv,
    float *mudf, float *t, const float *t_1, float *t_ave, const float *ft, const float *mu_tend,
    float rdx, float rdy, float dts, float epssm,
    const float *dnw, const float *fnm, const float *fnp, const float *rdnw,
    const float *msfuy, const float *msfvx_inv, const float *msftx, const float *msfty,
    int periodic_x, int specified, int nested,
    int ids, int ide, int jds, int jde, int kde,
    int ims, int ime, int jms, int jme, int kms, int kme,
    int its, int ite, int jts, int jte, int kts, int kte);

/* Same call applied `nsteps` times back to back on device-resident state (the acoustic loop of one
 * RK3 sub-step with u,v held fixed): one upload, nsteps launches, one download.  Host pointers only. */
int wrfb200_advance_mu_t_loop(
    float *ww, const float *ww_1, const float *u, const float *u_1, const float *v, const float *v_1,
    float *mu, const float *mut, float *muave, float *muts, const float *muu, const float *muv,
    float *mudf, float *t, const float *t_1, float *t_ave, const float *ft, const float *mu_tend,
    float rdx, float rdy, float dts, float epssm,
    const float *dnw, const float *fnm, const float *fnp, const float *rdnw,
    const float *msfuy, const float *msfvx_inv, const float *msftx, const float *msfty,
    int periodic_x, int specified, int nested,
    int ids, int ide, int jds, int jde, int kde,
    int ims, int ime, int jms, int jme, int kms, int kme,
    int its, int ite, int jts, int jte, int kts, int kte,
    int nsteps);

/* Stream used by the device-pointer form of wrfb200_advance_mu_t (a cudaStream_t; NULL = default stream)
 * and kernel selection for it. Thread-local. */
int wrfb200_set_default_stream(void *cuda_stream);
int wrfb200_set_default_kernel(int kernel /* wrfb200_kernel */);

/* Release the per-thread device mirrors cached by the host-pointer form (and every host registration). */
int wrfb200_release_cache(void);

/* Acoustic-loop residency for HOST-pointer callers of wrfb200_advance_mu_t (thread-local).  Between begin and
 * end, the first call uploads everything; every later call with the SAME arrays, index sets and scalars
 * uploads only u and v -- what advance_uv changed -- runs the step on the device-resident state and
 * downloads the seven outputs.  Contract: between those calls the caller changes nothing but u and v (ww, t,
 * mu on the device are this routine's own previous outputs; the rest are constants of the RK sub-step).
 * This is the reference's per-call H2D of all 26 arrays (advance_mu_t_no_async.cu:245-306) reduced to the
 * two that change. */
int wrfb200_acoustic_loop_begin(void);
int wrfb200_acoustic_loop_end(void);

/* Page-locking of the caller's host arrays (process-wide; default off, or WRFB200_PIN_HOST=1).  When on,
 * every host array of a compat call is cudaHostRegister'ed the first time it is seen and stays registered
 * until wrfb200_release_cache / wrfb200_host_unregister_all: copies then run at the pinned link rate.  The
 * caller must keep registered arrays allocated until then.  wrfb200_host_register pins one array explicitly
 * (the role of cudaHostAlloc in advance_mu_t_driver.cu:97-167). */
int wrfb200_set_host_pinning(int enable);
int wrfb200_host_register(const float *host, size_t bytes);
int wrfb200_host_unregister_all(void);

/* Kernel the most recent wrfb200_advance_mu_t call of this thread ran (a wrfb200_kernel id; AUTO resolved). */
int wrfb200_default_last_kernel(int *kernel);

/* ------------------------------------------------------------------------------------------------
 * 2. Device-resident state (replaces the per-call malloc/H2D/D2H/free of advance_mu_t_no_async.cu).
 * ---------------------------------------------------------------------------------------------- */

/* Create a patch on CUDA device `device` (-1 = current).  If `allocate` is non-zero all 26 fields get
 * device mirrors with rows padded to a multiple of 32 floats; otherwise fields must be bound with
 * wrfb200_bind_device before use. */
int wrfb200_create(wrfb200_handle **out, const wrfb200_domain *dom, int device, int allocate);
int wrfb200_destroy(wrfb200_handle *h);

int wrfb200_set_stream(wrfb200_handle *h, void *cuda_stream);
int wrfb200_set_scalars(wrfb200_handle *h, float rdx, float rdy, float dts, float epssm);
int wrfb200_set_kernel(wrfb200_handle *h, int kernel /* wrfb200_kernel */);

/* Adopt a caller-owned device buffer for `field`.  pitch = row stride in floats (>= ime-ims+1);
 * for 1-D fields pitch is ignored. */
int wrfb200_bind_device(wrfb200_handle *h, int field, float *device_ptr, long pitch);
/* Query the device buffer of `field` (owned or bound). */
int wrfb200_device_ptr(wrfb200_handle *h, int field, float **device_ptr, long *pitch);

/* Async copies on the handle's stream between DENSE host arrays (Fortran extents ims:ime etc.) and the
 * device mirrors.  Host memory should be pinned for the copies to be truly asynchronous. */
int wrfb200_upload(wrfb200_handle *h, int field, const float *host);
int wrfb200_download(wrfb200_handle *h, int field, float *host);
/* Only the sub-range [i0..i1] x [k0..k1] x [j0..j1] (Fortran-numbered, inclusive; k ignored for 2-D). */
int wrfb200_upload_range(wrfb200_handle *h, int field, const float *host, int i0, int i1, int k0, int k1, int j0, int j1);
int wrfb200_download_range(wrfb200_handle *h, int field, float *host, int i0, int i1, int k0, int k1, int j0, int j1);

/* Resident-state verbs: the arrays a host-resident caller moves at the three cadences of the acoustic loop
 * (once per RK sub-step: constants and the state the loop starts from; every small step: u, v up and the
 * outputs down).  Dense host arrays, async on the handle's stream, NULL skips an array.
 * wrfb200_download_outputs copies exactly the cells the routine writes for the given tile. */
int wrfb200_upload_constants(
    wrfb200_handle *h, const float *ww_1, const float *u_1, const float *v_1, const float *t_1, const float *ft,
    const float *mut, const float *muu, const float *muv, const float *mu_tend,
    const float *msfuy, const float *msfvx_inv, const float *msftx, const float *msfty,
    const float *dnw, const float *fnm, const float *fnp, const float *rdnw);
int wrfb200_upload_state(wrfb200_handle *h, const float *ww, const float *t, const float *mu);
int wrfb200_set_uv(wrfb200_handle *h, const float *u, const float *v);
int wrfb200_download_outputs(wrfb200_handle *h, int its, int ite, int jts, int jte, int kts, int kte,
                             float *ww, float *t, float *t_ave, float *mu, float *muave, float *muts, float *mudf);

/* One advance_mu_t over the tile its:ite x jts:jte (kts:kte as the Fortran), asynchronous on the stream. */
int wrfb200_step(wrfb200_handle *h, int its, int ite, int jts, int jte, int kts, int kte);
/* `nsteps` back-to-back steps replayed from a CUDA graph captured on first use for this tile. */
int wrfb200_step_graph(wrfb200_handle *h, int its, int ite, int jts, int jte, int kts, int kte, int nsteps);
int wrfb200_sync(wrfb200_handle *h);

/* Number of kernels this library has launched on behalf of `h` since creation (bench accounting). */
int wrfb200_launch_count(wrfb200_handle *h, long *count);
/* Kernel of the most recent launch on `h` (a wrfb200_kernel id; AUTO resolved to what actually ran). */
int wrfb200_last_kernel(wrfb200_handle *h, int *kernel);

/* ------------------------------------------------------------------------------------------------
 * 3. Halo support for the 2-D (i,j) decomposition (one patch per rank).
 *    pack:   copy the `width` cells just INSIDE the patch edge `side` of `field` into a dense buffer
 *    unpack: copy a dense buffer into the `width` halo cells just OUTSIDE that edge.
 *    Buffer layout is [j][k][i] over the packed sub-box (i fastest); ips..ipe/jps..jpe is the patch.
 *    Buffer sizes: WEST/EAST width*nk*(jpe-jps+1) floats, SOUTH/NORTH (ipe-ips+1)*nk*width floats
 *    (nk = kme-kms+1 for 3-D fields, 1 for 2-D).
 * ---------------------------------------------------------------------------------------------- */
int wrfb200_pack_halo(wrfb200_handle *h, int field, int side, int width,
                      int ips, int ipe, int jps, int jpe, float *device_buf);
int wrfb200_unpack_halo(wrfb200_handle *h, int field, int side, int width,
                        int ips, int ipe, int jps, int jpe, const float *device_buf);

/* Deterministic stand-in for advance_uv (which in WRF updates u,v between two advance_mu_t calls but is
 * not part of the reference): field = WRFB200_U: u(i,k,j) += c*(mudf(i,j)-mudf(i-1,j));
 * field = WRFB200_V: v(i,k,j) += c*(mudf(i,j)-mudf(i,j-1)); over i0..i1 x all memory levels x j0..j1.
 * It makes the multi-step loop and the halo exchange load-bearing in tests. */
int wrfb200_standin_advance_uv(wrfb200_handle *h, int field, float c, int i0, int i1, int j0, int j1);

/* ------------------------------------------------------------------------------------------------
 * 3b. Multi-GPU behind the boundary: one rank per GPU of one NVLink box, 2-D (i,j) patches, the one-cell
 *     halo exchange fused into the kernels over peer-mapped memory (csrc/comm.cu).  Replaces the j-slab
 *     plan + per-call host re-upload + per-device launch loop of advance_mu_t_no_async.cu:87-162,
 *     :276-298, :329-357.  The handle must own its mirrors (wrfb200_create(..., allocate=1)) with
 *     memory = patch + halo (>= 1 cell) and the GLOBAL ids..jde, so the boundary clamps of
 *     module_small_step_em.f90:91-106 fire on edge ranks only.  Rank r sits at (r % px, r / px).
 *
 *     Bootstrap (transport-agnostic, like an ncclUniqueId): every rank calls wrfb200_comm_init, the caller
 *     all-gathers the WRFB200_COMM_INFO_BYTES blobs in rank order (MPI_Allgather, torch.distributed, a
 *     pipe ...), every rank calls wrfb200_comm_connect.  Ranks may be processes (CUDA IPC) or several
 *     handles of one process (peer access).
 *
 *     Per RK sub-step, after uploading:  wrfb200_comm_push_constants  (u_1, muu, msfuy, v_1, muv,
 *     msfvx_inv, t_1 edges into the neighbours' halos, bracketed by a stream-ordered neighbour barrier).
 *     Per acoustic step:  [caller's advance_uv]  ->  wrfb200_comm_push_uv  ->  wrfb200_comm_step.
 *     wrfb200_comm_step stores this patch's south row of v into the south neighbour's north halo (its
 *     south-row blocks do that before they start), waits -- in the blocks that own the east column / north row
 *     only -- for its own east / north halo of u / v, and stores mu, muts, mudf of its east column / north row
 *     into the east / north neighbour's west / south halo for that neighbour's next advance_uv; a kernel that
 *     reads those halos must be preceded by wrfb200_comm_wait_outputs on the same stream.  Everything is
 *     asynchronous on the handle's stream; there is no host synchronisation and no NCCL call inside the loop.
 *     All ranks must issue the same sequence of steps (the flags are step counters).  A halo wait gives up
 *     after 5 s (WRFB200_FLAG_TIMEOUT_MS) and is reported by wrfb200_comm_status.
 * ---------------------------------------------------------------------------------------------- */
#define WRFB200_COMM_INFO_BYTES 2048
int wrfb200_comm_info_bytes(void);
int wrfb200_comm_init(wrfb200_handle *h, int px, int py, int rank,
                      int ips, int ipe, int jps, int jpe, void *info_out /* WRFB200_COMM_INFO_BYTES */);
int wrfb200_comm_connect(wrfb200_handle *h, const void *all_infos /* nranks blobs, rank order */, int nranks);
int wrfb200_comm_barrier(wrfb200_handle *h);          /* stream-ordered barrier with the neighbours */
int wrfb200_comm_push_constants(wrfb200_handle *h);
int wrfb200_comm_push_uv(wrfb200_handle *h);          /* u west column -> west neighbour (px > 1); the v south row is
                                                         pushed by wrfb200_comm_step itself */
int wrfb200_comm_wait_outputs(wrfb200_handle *h);     /* west / south halos of mu, muts, mudf have arrived */
int wrfb200_comm_step(wrfb200_handle *h);             /* advance_mu_t over the patch, fused with its exchange */
/* stand-in advance_uv (see wrfb200_standin_advance_uv) over this patch's share of the global update boxes,
 * preceded by wrfb200_comm_wait_outputs */
int wrfb200_comm_standin_advance_uv(wrfb200_handle *h, float c);
/* nsteps x (push_uv, step[, stand-in between steps]) -- the acoustic loop of one RK sub-step; with
 * use_graph the sequence is captured once and replayed from a CUDA graph. */
int wrfb200_comm_loop(wrfb200_handle *h, int nsteps, int standin, float c, int use_graph);
/* Synchronises the stream; flag_timeouts != 0 means some halo wait gave up (results are then invalid);
 * steps_done = advance_mu_t launches completed since wrfb200_comm_init. */
int wrfb200_comm_status(wrfb200_handle *h, int *flag_timeouts, long *steps_done);

/* ------------------------------------------------------------------------------------------------
 * 4. Harness utilities (host side; usable without a GPU).
 * ---------------------------------------------------------------------------------------------- */

/* Index sets of module_small_step_em.f90:91-106. */
int wrfb200_bounds(int periodic_x, int specified, int nested,
                   int ids, int ide, int jds, int jde,
                   int its, int ite, int jts, int jte, int kts, int kte,
                   int *i_start, int *i_end, int *j_start, int *j_end, int *k_start, int *k_end);

/* Deterministic atmosphere-like synthetic field (counter-based: the value of a cell depends only on
 * (seed, field, GLOBAL i,k,j, domain extents), so every decomposition generates the same global field).
 * Fills the dense host array of `field` for the patch described by `dom`.  dx_m (grid spacing in metres) is
 * accepted for interface stability and currently unused: the field magnitudes are grid-spacing independent;
 * rdx, rdy and dts are the caller's scalars. */
int wrfb200_synth_field(int field, uint64_t seed, const wrfb200_domain *dom, float dx_m, float *host_out);

/* The reference's comparison metrics (common.cu:68-164) over n floats: number of bit-equal values,
 * max relative error (|a-b|/max(|a|,|b|), or max(|a|,|b|) when either is zero, :117-120), max absolute
 * error, max ulp distance (:51-66), rmse.  Returns WRFB200_ERR_INVALID_ARG if either side holds a NaN
 * (the reference aborts, :108-115). */
typedef struct wrfb200_compare_result {
    long n, n_equal, n_different;
    float max_rel, max_abs, rmse;
    long max_ulp;
} wrfb200_compare_result;
int wrfb200_compare(const float *a, const float *b, long n, wrfb200_compare_result *out);

/* Host-only description of the launch the TMA kernel makes for the tile its:ite x jts:jte of `dom` on a device with
 * `resident_blocks` resident blocks (0: a B200, 2 x 148): plan10 = {configuration TJ*10+STAGES, TJ, STAGES, tile
 * columns, block rows of 2-row tiles, block rows of 1-row tiles, remainder-strip blocks, first strip column (memory
 * index), grid size, dynamic shared memory in bytes}.  Needs no GPU; the launch geometry is unit-tested with it. */
int wrfb200_pipe_plan(const wrfb200_domain *dom, int its, int ite, int jts, int jte, int kts, int kte,
                      int resident_blocks, long long *plan10);

/* Device self-test of the kernel's division by a loop-invariant divisor (reciprocal hoisted out of the level
 * loop, csrc/amt_pipe.cu) against IEEE division: all 2^23 divisor mantissas at three exponents, each with
 * `dividends_per_divisor` dividends.  *mismatches must come back 0. */
int wrfb200_selftest_division(long long *mismatches, long long *checked, int dividends_per_divisor);

const char *wrfb200_last_error(void);
int wrfb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WRFB200_H */
