!-----------------------------------------------------------------------------------------------------------
! module_small_step_em_resident.F90 -- ISO_C_BINDING interfaces of the DEVICE-RESIDENT and MULTI-GPU entry
! points of libwrfb200.so (include/wrfb200.h sections 1, 2 and 3b), for a Fortran host that wants more than
! the one-call drop-in of module_small_step_em.F90:
!
!   (A) acoustic-loop residency with the UNCHANGED 48-argument call (smallest source change in solve_em):
!           CALL wrfb200_loop_begin()
!           DO iteration = 1, number_of_small_timesteps
!              CALL advance_uv( ... )                       ! host code, changes u, v only
!              CALL advance_mu_t( ww, ww1, u_2, ... )       ! module_small_step_em.F90 -> wrfb200_advance_mu_t
!           END DO
!           CALL wrfb200_loop_end()
!       The first call of the loop uploads all 26 arrays, later calls only u and v; every call downloads the
!       seven outputs.  Replaces the reference's H2D copy of every array on every call
!       (advance_mu_t_no_async.cu:245-306).
!
!   (B) an explicit patch handle with the three cadences spelled out (wrfb200_create, wrfb200_upload_constants
!       + wrfb200_upload_state once per RK sub-step, then per small step wrfb200_set_uv / wrfb200_step /
!       wrfb200_download_outputs) -- what replaces the allocate/copy/launch/copy/free body of
!       advance_mu_t_no_async.cu:178-423.
!
!   (C) one MPI rank per GPU: wrfb200_comm_init -> MPI_Allgather of the WRFB200_COMM_INFO_BYTES blobs ->
!       wrfb200_comm_connect, then per RK sub-step wrfb200_comm_push_constants and per small step
!       wrfb200_comm_push_uv + wrfb200_comm_step (or wrfb200_comm_loop for the whole loop).  Replaces the
!       j-slab plan and per-device launch loop of advance_mu_t_no_async.cu:87-162, :329-357.  Sketch:
!
!           TYPE(c_ptr) :: h
!           CHARACTER(KIND=c_char), TARGET :: mine(WRFB200_COMM_INFO_BYTES), everyone(WRFB200_COMM_INFO_BYTES*nproc)
!           st = wrfb200_create( h, dom, MOD(myrank, gpus_per_node), 1_c_int )
!           ...uploads...
!           st = wrfb200_comm_init( h, px, py, myrank, ips, ipe, jps, jpe, C_LOC(mine) )
!           CALL MPI_Allgather( mine, WRFB200_COMM_INFO_BYTES, MPI_BYTE, everyone, WRFB200_COMM_INFO_BYTES, MPI_BYTE, comm, ierr )
!           st = wrfb200_comm_connect( h, C_LOC(everyone), nproc )
!           st = wrfb200_comm_push_constants( h )
!           st = wrfb200_comm_loop( h, number_of_small_timesteps, 0_c_int, 0.0_c_float, 1_c_int )
!
! Source only: no Fortran compiler exists in the build image (INTEGRATION.md); tests/c_abi_harness.c and
! tests/c_comm_harness.c drive exactly these entry points from C, and tests/test_cabi_host.py checks that every
! NAME= below is an exported symbol of the built library.
!-----------------------------------------------------------------------------------------------------------
MODULE module_small_step_em_resident

USE, INTRINSIC :: iso_c_binding, ONLY : c_float, c_int, c_long, c_ptr, c_size_t

IMPLICIT NONE
PUBLIC

INTEGER, PARAMETER :: WRFB200_COMM_INFO_BYTES = 2048

! struct wrfb200_domain (include/wrfb200.h)
TYPE, BIND(C) :: wrfb200_domain
   INTEGER(c_int) :: ids, ide, jds, jde, kde
   INTEGER(c_int) :: ims, ime, jms, jme, kms, kme
   INTEGER(c_int) :: periodic_x, specified, nested
END TYPE wrfb200_domain

INTERFACE
   ! ---- (A) acoustic-loop residency of the 48-argument call ----
   FUNCTION wrfb200_acoustic_loop_begin() BIND(C, NAME="wrfb200_acoustic_loop_begin") RESULT(status)
      IMPORT :: c_int
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_acoustic_loop_end() BIND(C, NAME="wrfb200_acoustic_loop_end") RESULT(status)
      IMPORT :: c_int
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_set_host_pinning(enable) BIND(C, NAME="wrfb200_set_host_pinning") RESULT(status)
      IMPORT :: c_int
      INTEGER(c_int), VALUE :: enable
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_host_register(host, bytes) BIND(C, NAME="wrfb200_host_register") RESULT(status)
      IMPORT :: c_int, c_float, c_size_t
      REAL(c_float), DIMENSION(*), INTENT(IN) :: host
      INTEGER(c_size_t), VALUE :: bytes
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_release_cache() BIND(C, NAME="wrfb200_release_cache") RESULT(status)
      IMPORT :: c_int
      INTEGER(c_int) :: status
   END FUNCTION

   ! ---- (B) explicit device-resident patch ----
   FUNCTION wrfb200_create(handle, dom, device, allocate) BIND(C, NAME="wrfb200_create") RESULT(status)
      IMPORT :: c_int, c_ptr, wrfb200_domain
      TYPE(c_ptr), INTENT(OUT) :: handle
      TYPE(wrfb200_domain), INTENT(IN) :: dom
      INTEGER(c_int), VALUE :: device, allocate
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_destroy(handle) BIND(C, NAME="wrfb200_destroy") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_set_scalars(handle, rdx, rdy, dts, epssm) BIND(C, NAME="wrfb200_set_scalars") RESULT(status)
      IMPORT :: c_int, c_ptr, c_float
      TYPE(c_ptr), VALUE :: handle
      REAL(c_float), VALUE :: rdx, rdy, dts, epssm
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_upload_constants(handle, ww_1, u_1, v_1, t_1, ft, mut, muu, muv, mu_tend,        &
                                     msfuy, msfvx_inv, msftx, msfty, dnw, fnm, fnp, rdnw)           &
            BIND(C, NAME="wrfb200_upload_constants") RESULT(status)
      IMPORT :: c_int, c_ptr, c_float
      TYPE(c_ptr), VALUE :: handle
      REAL(c_float), DIMENSION(*), INTENT(IN) :: ww_1, u_1, v_1, t_1, ft, mut, muu, muv, mu_tend
      REAL(c_float), DIMENSION(*), INTENT(IN) :: msfuy, msfvx_inv, msftx, msfty, dnw, fnm, fnp, rdnw
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_upload_state(handle, ww, t, mu) BIND(C, NAME="wrfb200_upload_state") RESULT(status)
      IMPORT :: c_int, c_ptr, c_float
      TYPE(c_ptr), VALUE :: handle
      REAL(c_float), DIMENSION(*), INTENT(IN) :: ww, t, mu
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_set_uv(handle, u, v) BIND(C, NAME="wrfb200_set_uv") RESULT(status)
      IMPORT :: c_int, c_ptr, c_float
      TYPE(c_ptr), VALUE :: handle
      REAL(c_float), DIMENSION(*), INTENT(IN) :: u, v
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_step(handle, its, ite, jts, jte, kts, kte) BIND(C, NAME="wrfb200_step") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int), VALUE :: its, ite, jts, jte, kts, kte
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_step_graph(handle, its, ite, jts, jte, kts, kte, nsteps)                          &
            BIND(C, NAME="wrfb200_step_graph") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int), VALUE :: its, ite, jts, jte, kts, kte, nsteps
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_download_outputs(handle, its, ite, jts, jte, kts, kte,                            &
                                     ww, t, t_ave, mu, muave, muts, mudf)                            &
            BIND(C, NAME="wrfb200_download_outputs") RESULT(status)
      IMPORT :: c_int, c_ptr, c_float
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int), VALUE :: its, ite, jts, jte, kts, kte
      REAL(c_float), DIMENSION(*), INTENT(INOUT) :: ww, t, t_ave, mu, muave, muts, mudf
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_sync(handle) BIND(C, NAME="wrfb200_sync") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int) :: status
   END FUNCTION

   ! ---- (C) one rank per GPU: halo exchange fused into the kernels over peer-mapped memory ----
   FUNCTION wrfb200_comm_init(handle, px, py, rank, ips, ipe, jps, jpe, info_out)                     &
            BIND(C, NAME="wrfb200_comm_init") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int), VALUE :: px, py, rank, ips, ipe, jps, jpe
      TYPE(c_ptr), VALUE :: info_out                      ! WRFB200_COMM_INFO_BYTES bytes
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_comm_connect(handle, all_infos, nranks) BIND(C, NAME="wrfb200_comm_connect") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      TYPE(c_ptr), VALUE :: all_infos                     ! nranks blobs in rank order (MPI_Allgather)
      INTEGER(c_int), VALUE :: nranks
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_comm_barrier(handle) BIND(C, NAME="wrfb200_comm_barrier") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_comm_push_constants(handle) BIND(C, NAME="wrfb200_comm_push_constants") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_comm_push_uv(handle) BIND(C, NAME="wrfb200_comm_push_uv") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_comm_wait_outputs(handle) BIND(C, NAME="wrfb200_comm_wait_outputs") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_comm_step(handle) BIND(C, NAME="wrfb200_comm_step") RESULT(status)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_comm_loop(handle, nsteps, standin, c, use_graph) BIND(C, NAME="wrfb200_comm_loop") RESULT(status)
      IMPORT :: c_int, c_ptr, c_float
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int), VALUE :: nsteps, standin
      REAL(c_float), VALUE :: c
      INTEGER(c_int), VALUE :: use_graph
      INTEGER(c_int) :: status
   END FUNCTION
   FUNCTION wrfb200_comm_status(handle, flag_timeouts, steps_done) BIND(C, NAME="wrfb200_comm_status") RESULT(status)
      IMPORT :: c_int, c_ptr, c_long
      TYPE(c_ptr), VALUE :: handle
      INTEGER(c_int), INTENT(OUT) :: flag_timeouts
      INTEGER(c_long), INTENT(OUT) :: steps_done
      INTEGER(c_int) :: status
   END FUNCTION
END INTERFACE

CONTAINS

SUBROUTINE wrfb200_loop_begin()
   INTEGER(c_int) :: st
   st = wrfb200_acoustic_loop_begin()
END SUBROUTINE wrfb200_loop_begin

SUBROUTINE wrfb200_loop_end()
   INTEGER(c_int) :: st
   st = wrfb200_acoustic_loop_end()
END SUBROUTINE wrfb200_loop_end

END MODULE module_small_step_em_resident
