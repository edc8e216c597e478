!*************************************************************************************!
! lsf_b200_mod.f90 -- ISO_C_BINDING interface to liblsf_b200.so (include/lsf_b200.h)
!
! Drop-in replacements for the grid hot path of musheen/LevelSetFortran:
!
!   reference                                            this module
!   ---------------------------------------------------  ------------------------------
!   inline sign-search loop       set3d.f90:196-268      CALL signSearch_b200(...)
!   CALL reinit(...)              subs.f90:717-931       CALL reinit_b200(...)      (same argument list)
!   CALL narrowBand(...)          subs.f90:178-207       CALL narrowBand_b200(...)  (same argument list)
!   inline min/max DO n loop      set3d.f90:394-462      CALL minMaxFlow_b200(...)
!   "Advect Nodes" block          set3d.f90:465-501      CALL advectNodes_b200(...)
!
! Build (the reference is compiled with -fdefault-real-8, Makefile:4, so REAL == REAL(c_double)
! and the default INTEGER == INTEGER(c_int); the module states the kinds explicitly so it is
! correct with or without that flag as long as the caller's arrays are 8-byte reals):
!
!   gfortran -O3 -fdefault-real-8 -c fortran/lsf_b200_mod.f90 subs.f90 set3d.f90
!   gfortran -o set3d.exec lsf_b200_mod.o subs.o set3d.o -L<repo>/levelsetfortran_b200 -llsf_b200 \
!            -Wl,-rpath,<repo>/levelsetfortran_b200
!
! NOTE: no Fortran compiler exists in the image this library was developed in; this file is kept
! small and conservative (F2003 only) and has been reviewed but not compiled there (its statements are restricted to the subset the
! translator oracle/f90_to_c.py understands wherever that was possible).  The same C
! symbols are exercised by the ctypes mirror levelsetfortran_b200/set_subs.py in the test-suite.
!*************************************************************************************!
MODULE lsf_b200

USE, INTRINSIC :: ISO_C_BINDING
IMPLICIT NONE
PRIVATE

PUBLIC :: lsf_init, lsf_finalize, lsf_set_arith, lsf_set_sched, lsf_set_minmax_algo, lsf_set_precision
PUBLIC :: lsf_slab_range, lsf_sgrid_create, lsf_sgrid_create_f32, lsf_sgrid_ipc_handle, lsf_sgrid_attach, lsf_grid_destroy
PUBLIC :: lsf_sgrid_sync_ghosts, lsf_grid_create, lsf_grid_create_f32, lsf_grid_fill, lsf_grid_upload, lsf_grid_download
PUBLIC :: lsf_grid_download_phiN, lsf_grid_sign_init, lsf_grid_reinit, lsf_grid_narrowband, lsf_grid_minmax
PUBLIC :: lsf_grid_reinit_rk3, lsf_grid_advect_nodes, lsf_grid_checksum, lsf_host_register, lsf_host_unregister
PUBLIC :: signSearch_b200, reinit_b200, reinit_nograd_b200, narrowBand_b200, minMaxFlow_b200, advectNodes_b200
PUBLIC :: gridReinit_b200, gridMinMaxFlow_b200
PUBLIC :: stlRead_b200, writeVti_b200, writeS3d_b200
PUBLIC :: LSF_OK, LSF_NAN, LSF_ARITH_FAST, LSF_ARITH_EXACT, LSF_ARITH_AUTO, LSF_PREC_F64, LSF_PREC_F32

INTEGER(c_int), PARAMETER :: LSF_OK = 0, LSF_NAN = 1
INTEGER(c_int), PARAMETER :: LSF_ARITH_FAST = 0, LSF_ARITH_EXACT = 1, LSF_ARITH_AUTO = 2
INTEGER(c_int), PARAMETER :: LSF_PREC_F64 = 0, LSF_PREC_F32 = 1

INTERFACE

   FUNCTION lsf_init(device) BIND(C, NAME='lsf_init') RESULT(rc)
      IMPORT :: c_int
      INTEGER(c_int), VALUE :: device
      INTEGER(c_int) :: rc
   END FUNCTION lsf_init

   FUNCTION lsf_finalize() BIND(C, NAME='lsf_finalize') RESULT(rc)
      IMPORT :: c_int
      INTEGER(c_int) :: rc
   END FUNCTION lsf_finalize

   FUNCTION lsf_set_arith(arith) BIND(C, NAME='lsf_set_arith') RESULT(rc)
      IMPORT :: c_int
      INTEGER(c_int), VALUE :: arith
      INTEGER(c_int) :: rc
   END FUNCTION lsf_set_arith

   FUNCTION lsf_set_sched(sched) BIND(C, NAME='lsf_set_sched') RESULT(rc)
      IMPORT :: c_int
      INTEGER(c_int), VALUE :: sched
      INTEGER(c_int) :: rc
   END FUNCTION lsf_set_sched

   FUNCTION lsf_set_minmax_algo(algo) BIND(C, NAME='lsf_set_minmax_algo') RESULT(rc)
      IMPORT :: c_int
      INTEGER(c_int), VALUE :: algo          ! 0 = active list (default), 1 = whole-grid march
      INTEGER(c_int) :: rc
   END FUNCTION lsf_set_minmax_algo

   ! optional fp32 mode of reinit_b200 (device fields and WENO5 arithmetic in single precision; the REAL(8)
   ! host arrays are unchanged; phi within 1e-4 relative of the fp64 path; gradPhi/gradPhiMag not written)
   FUNCTION lsf_set_precision(prec) BIND(C, NAME='lsf_set_precision') RESULT(rc)
      IMPORT :: c_int
      INTEGER(c_int), VALUE :: prec          ! LSF_PREC_F64 (default) / LSF_PREC_F32
      INTEGER(c_int) :: rc
   END FUNCTION lsf_set_precision

   ! ---- z-slab sharding, one MPI rank per GPU (include/lsf_b200.h; calling sequence: INTEGRATION.md) ----
   FUNCTION lsf_slab_range(nz,nranks,rank,k0,k1) BIND(C, NAME='lsf_slab_range') RESULT(rc)
      IMPORT :: c_int
      INTEGER(c_int), VALUE :: nz,nranks,rank
      INTEGER(c_int) :: k0,k1                ! this rank owns planes k0 .. k1-1 of phi(0:nx,0:ny,0:nz)
      INTEGER(c_int) :: rc
   END FUNCTION lsf_slab_range

   FUNCTION lsf_sgrid_create(g,nx,ny,nz,rank,nranks) BIND(C, NAME='lsf_sgrid_create') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr) :: g                        ! lsf_grid**
      INTEGER(c_int), VALUE :: nx,ny,nz,rank,nranks
      INTEGER(c_int) :: rc
   END FUNCTION lsf_sgrid_create

   FUNCTION lsf_sgrid_create_f32(g,nx,ny,nz,rank,nranks) BIND(C, NAME='lsf_sgrid_create_f32') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr) :: g                        ! lsf_grid**, the slab in the optional fp32 mode
      INTEGER(c_int), VALUE :: nx,ny,nz,rank,nranks
      INTEGER(c_int) :: rc
   END FUNCTION lsf_sgrid_create_f32

   FUNCTION lsf_sgrid_ipc_handle(g,handle) BIND(C, NAME='lsf_sgrid_ipc_handle') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_char
      TYPE(c_ptr), VALUE :: g
      CHARACTER(KIND=c_char) :: handle(64)    ! LSF_IPC_HANDLE_BYTES; MPI_Allgather these as MPI_BYTE
      INTEGER(c_int) :: rc
   END FUNCTION lsf_sgrid_ipc_handle

   FUNCTION lsf_sgrid_attach(g,handles) BIND(C, NAME='lsf_sgrid_attach') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_char
      TYPE(c_ptr), VALUE :: g
      CHARACTER(KIND=c_char) :: handles(*)    ! nranks*64 bytes, rank order
      INTEGER(c_int) :: rc
   END FUNCTION lsf_sgrid_attach

   FUNCTION lsf_grid_destroy(g) BIND(C, NAME='lsf_grid_destroy') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: g
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_destroy

   FUNCTION lsf_sgrid_sync_ghosts(g) BIND(C, NAME='lsf_sgrid_sync_ghosts') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: g
      INTEGER(c_int) :: rc
   END FUNCTION lsf_sgrid_sync_ghosts

   ! ---- device-resident grids (whole grid: lsf_grid_create; z-slab: lsf_sgrid_create): the same calls drive both.
   ! On a z-slab the host arrays hold the rank's OWNED planes k0..k1-1, i.e. phi(0:nx,0:ny,k0:k1-1). ----
   FUNCTION lsf_grid_create(g,nx,ny,nz) BIND(C, NAME='lsf_grid_create') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr) :: g                        ! lsf_grid**
      INTEGER(c_int), VALUE :: nx,ny,nz
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_create

   FUNCTION lsf_grid_create_f32(g,nx,ny,nz) BIND(C, NAME='lsf_grid_create_f32') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr) :: g
      INTEGER(c_int), VALUE :: nx,ny,nz
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_create_f32

   FUNCTION lsf_grid_fill(g,value) BIND(C, NAME='lsf_grid_fill') RESULT(rc)       ! phi = value (set3d.f90:161)
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: g
      REAL(c_double), VALUE :: value
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_fill

   FUNCTION lsf_grid_upload(g,phi) BIND(C, NAME='lsf_grid_upload') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: g
      REAL(c_double), INTENT(IN) :: phi(*)
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_upload

   FUNCTION lsf_grid_download(g,phi) BIND(C, NAME='lsf_grid_download') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: g
      REAL(c_double) :: phi(*)
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_download

   FUNCTION lsf_grid_download_phiN(g,phiN) BIND(C, NAME='lsf_grid_download_phiN') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: g
      REAL(c_double) :: phiN(*)
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_download_phiN

   FUNCTION lsf_grid_sign_init(g,xLo,dx,surfX,nSurfNode,surfElem,nSurfElem,im,ip,jm,jp,km,kp) &
                               BIND(C, NAME='lsf_grid_sign_init') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double, c_int32_t
      TYPE(c_ptr), VALUE :: g
      REAL(c_double), INTENT(IN) :: xLo(3)
      REAL(c_double), VALUE :: dx
      REAL(c_double), INTENT(IN) :: surfX(*)
      INTEGER(c_int), VALUE :: nSurfNode
      INTEGER(c_int32_t), INTENT(IN) :: surfElem(*)
      INTEGER(c_int), VALUE :: nSurfElem,im,ip,jm,jp,km,kp      ! GLOBAL sub-box (set3d.f90:180-186) on every rank
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_sign_init

   FUNCTION lsf_grid_reinit(g,iter,dx,h,tol,n_exit,rms_hist) BIND(C, NAME='lsf_grid_reinit') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: g
      INTEGER(c_int), VALUE :: iter
      REAL(c_double), VALUE :: dx,h,tol        ! tol = 1.E-5 in the reference (subs.f90:915)
      INTEGER(c_int) :: n_exit
      REAL(c_double) :: rms_hist(*)            ! 0:iter
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_reinit

   ! K2' throughput mode: Jacobi WENO5 + TVD-RK3 -- NOT the reference's Gauss-Seidel algorithm (include/lsf_b200.h)
   FUNCTION lsf_grid_reinit_rk3(g,steps,dx,dt,tol,n_exit,rms_hist) BIND(C, NAME='lsf_grid_reinit_rk3') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: g
      INTEGER(c_int), VALUE :: steps
      REAL(c_double), VALUE :: dx,dt,tol
      INTEGER(c_int) :: n_exit
      REAL(c_double) :: rms_hist(*)
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_reinit_rk3

   FUNCTION lsf_grid_narrowband(g,dx,phiNB,phiSB) BIND(C, NAME='lsf_grid_narrowband') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double, c_int32_t
      TYPE(c_ptr), VALUE :: g
      REAL(c_double), VALUE :: dx
      INTEGER(c_int32_t) :: phiNB(*), phiSB(*)
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_narrowband

   FUNCTION lsf_grid_minmax(g,iter,dx,h1,tol,n_exit,rms_hist) BIND(C, NAME='lsf_grid_minmax') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: g
      INTEGER(c_int), VALUE :: iter
      REAL(c_double), VALUE :: dx,h1,tol       ! tol = 1.E-7 in the reference (set3d.f90:448)
      INTEGER(c_int) :: n_exit
      REAL(c_double) :: rms_hist(*)            ! 1:iter
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_minmax

   FUNCTION lsf_grid_advect_nodes(g,xLo,dx,surfXX,nSurfNode,phiSurf,gradPhiSurf,iter,n_moves) &
                                  BIND(C, NAME='lsf_grid_advect_nodes') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_double, c_long_long
      TYPE(c_ptr), VALUE :: g
      REAL(c_double), INTENT(IN) :: xLo(3)
      REAL(c_double), VALUE :: dx
      REAL(c_double) :: surfXX(*), phiSurf(*), gradPhiSurf(*)
      INTEGER(c_int), VALUE :: nSurfNode, iter
      INTEGER(c_long_long) :: n_moves
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_advect_nodes

   ! digest(1) = sum of bits(phi(q))*(2q+1) mod 2**64, digest(2) = xor of the bit patterns, over the OWNED points;
   ! the ranks' digests add / xor up to the single-GPU digest (MPI_Allreduce with MPI_SUM on INTEGER(8) wraps the same way)
   FUNCTION lsf_grid_checksum(g,digest) BIND(C, NAME='lsf_grid_checksum') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_int64_t
      TYPE(c_ptr), VALUE :: g
      INTEGER(c_int64_t) :: digest(2)
      INTEGER(c_int) :: rc
   END FUNCTION lsf_grid_checksum

   ! Page-lock a host array for the lifetime of its allocation: the host-buffer entry points (reinit_b200, ...) then
   ! move it at PCIe speed instead of through the driver's pageable staging path.  Call right after ALLOCATE, and
   ! lsf_host_unregister before DEALLOCATE.  nbytes = 8*SIZE(phi).
   FUNCTION lsf_host_register(ptr,nbytes) BIND(C, NAME='lsf_host_register') RESULT(rc)
      IMPORT :: c_int, c_ptr, c_size_t
      TYPE(c_ptr), VALUE :: ptr
      INTEGER(c_size_t), VALUE :: nbytes
      INTEGER(c_int) :: rc
   END FUNCTION lsf_host_register

   FUNCTION lsf_host_unregister(ptr) BIND(C, NAME='lsf_host_unregister') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: ptr
      INTEGER(c_int) :: rc
   END FUNCTION lsf_host_unregister

   ! ---- host-side pieces either side of the path (no device work): .vti / .s3d writers, stlRead de-duplication ----
   FUNCTION c_lsf_write_vti(path,phi,nx,ny,nz,xLo,dx) BIND(C, NAME='lsf_write_vti') RESULT(rc)
      IMPORT :: c_int, c_double, c_char
      CHARACTER(KIND=c_char), INTENT(IN) :: path(*)          ! NUL-terminated
      REAL(c_double), INTENT(IN) :: phi(*)
      INTEGER(c_int), VALUE :: nx,ny,nz
      REAL(c_double), INTENT(IN) :: xLo(3)
      REAL(c_double), VALUE :: dx
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_write_vti

   FUNCTION c_lsf_write_s3d(path,nSurfElem,nSurfNode,nBndElem,nBndComp,surfOrder,surfElem,surfElemTag,surfXX,bndNormal) &
                            BIND(C, NAME='lsf_write_s3d') RESULT(rc)
      IMPORT :: c_int, c_double, c_char, c_int32_t
      CHARACTER(KIND=c_char), INTENT(IN) :: path(*)
      INTEGER(c_int), VALUE :: nSurfElem,nSurfNode,nBndElem,nBndComp
      INTEGER(c_int32_t), INTENT(IN) :: surfOrder(*), surfElem(*), surfElemTag(*)
      REAL(c_double), INTENT(IN) :: surfXX(*), bndNormal(*)
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_write_s3d

   FUNCTION c_lsf_stl_count(path,ntri) BIND(C, NAME='lsf_stl_count') RESULT(rc)
      IMPORT :: c_int, c_char
      CHARACTER(KIND=c_char), INTENT(IN) :: path(*)
      INTEGER(c_int) :: ntri
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_stl_count

   FUNCTION c_lsf_stl_read_triangles(path,ntri,tri) BIND(C, NAME='lsf_stl_read_triangles') RESULT(rc)
      IMPORT :: c_int, c_char, c_float
      CHARACTER(KIND=c_char), INTENT(IN) :: path(*)
      INTEGER(c_int), VALUE :: ntri
      REAL(c_float) :: tri(*)                                 ! triangles(3,ntri*3) of subs.f90:38
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_stl_read_triangles

   FUNCTION c_lsf_stl_dedup(tri,ntri,nodes,surfElem,nSurfNode) BIND(C, NAME='lsf_stl_dedup') RESULT(rc)
      IMPORT :: c_int, c_float, c_int32_t
      REAL(c_float), INTENT(IN) :: tri(*)
      INTEGER(c_int), VALUE :: ntri
      REAL(c_float) :: nodes(*)                               ! nodesT(3,k)
      INTEGER(c_int32_t) :: surfElem(*)                       ! (ntri,3), 1-based
      INTEGER(c_int) :: nSurfNode
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_stl_dedup

   FUNCTION lsf_last_error() BIND(C, NAME='lsf_last_error') RESULT(msg)
      IMPORT :: c_ptr
      TYPE(c_ptr) :: msg
   END FUNCTION lsf_last_error

   FUNCTION c_lsf_sign_init(phi,nx,ny,nz,xLo,dx,surfX,nSurfNode,surfElem,nSurfElem, &
                            im,ip,jm,jp,km,kp) BIND(C, NAME='lsf_sign_init') RESULT(rc)
      IMPORT :: c_int, c_double, c_int32_t
      REAL(c_double) :: phi(*)
      INTEGER(c_int), VALUE :: nx,ny,nz
      REAL(c_double), INTENT(IN) :: xLo(3)
      REAL(c_double), VALUE :: dx
      REAL(c_double), INTENT(IN) :: surfX(*)
      INTEGER(c_int), VALUE :: nSurfNode
      INTEGER(c_int32_t), INTENT(IN) :: surfElem(*)
      INTEGER(c_int), VALUE :: nSurfElem,im,ip,jm,jp,km,kp
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_sign_init

   FUNCTION c_lsf_reinit(phi,gradPhi,gradPhiMag,nx,ny,nz,iter,dx,h,n_exit,rms_hist) &
                         BIND(C, NAME='lsf_reinit') RESULT(rc)
      IMPORT :: c_int, c_double, c_ptr
      REAL(c_double) :: phi(*)
      TYPE(c_ptr), VALUE :: gradPhi, gradPhiMag       ! C_NULL_PTR: not wanted (both are dead downstream, set3d.f90:372-375)
      INTEGER(c_int), VALUE :: nx,ny,nz,iter
      REAL(c_double), VALUE :: dx,h
      INTEGER(c_int) :: n_exit
      REAL(c_double) :: rms_hist(*)
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_reinit

   FUNCTION c_lsf_narrowband(nx,ny,nz,dx,phi,phiNB,phiSB) BIND(C, NAME='lsf_narrowband') RESULT(rc)
      IMPORT :: c_int, c_double, c_int32_t
      INTEGER(c_int), VALUE :: nx,ny,nz
      REAL(c_double), VALUE :: dx
      REAL(c_double), INTENT(IN) :: phi(*)
      INTEGER(c_int32_t) :: phiNB(*), phiSB(*)
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_narrowband

   FUNCTION c_lsf_minmax(phi,phiN,phiNB,phiSB,nx,ny,nz,iter,dx,h1,tol,n_exit,rms_hist) &
                         BIND(C, NAME='lsf_minmax') RESULT(rc)
      IMPORT :: c_int, c_double, c_int32_t
      REAL(c_double) :: phi(*), phiN(*)
      INTEGER(c_int32_t) :: phiNB(*), phiSB(*)
      INTEGER(c_int), VALUE :: nx,ny,nz,iter
      REAL(c_double), VALUE :: dx,h1,tol
      INTEGER(c_int) :: n_exit
      REAL(c_double) :: rms_hist(*)
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_minmax

   FUNCTION c_lsf_advect_nodes(phi,phiSB,nx,ny,nz,xLo,dx,surfXX,nSurfNode,phiSurf,gradPhiSurf,iter,n_moves) &
                               BIND(C, NAME='lsf_advect_nodes') RESULT(rc)
      IMPORT :: c_int, c_double, c_int32_t, c_long_long
      REAL(c_double), INTENT(IN) :: phi(*)
      INTEGER(c_int32_t), INTENT(IN) :: phiSB(*)
      INTEGER(c_int), VALUE :: nx,ny,nz
      REAL(c_double), INTENT(IN) :: xLo(3)
      REAL(c_double), VALUE :: dx
      REAL(c_double) :: surfXX(*), phiSurf(*), gradPhiSurf(*)
      INTEGER(c_int), VALUE :: nSurfNode, iter
      INTEGER(c_long_long) :: n_moves
      INTEGER(c_int) :: rc
   END FUNCTION c_lsf_advect_nodes

END INTERFACE

CONTAINS

!*************************************************************************************!
! library failure (rc < 0): print the library's message and stop, like the reference's
! other fatal paths (set3d.f90:79-80)
!*************************************************************************************!
SUBROUTINE lsf_fail(where, rc)
   CHARACTER(LEN=*), INTENT(IN) :: where
   INTEGER(c_int), INTENT(IN) :: rc
   TYPE(c_ptr) :: p
   CHARACTER(KIND=c_char), POINTER :: s(:)
   INTEGER :: q
   p = lsf_last_error()
   PRINT*, " liblsf_b200 failure in ", where, " code ", rc
   IF (C_ASSOCIATED(p)) THEN
      CALL C_F_POINTER(p, s, (/512/))
      DO q = 1,512
         IF (s(q) == C_NULL_CHAR) EXIT
         WRITE(*,'(A)',ADVANCE='NO') s(q)
      END DO
      PRINT*,
   END IF
   STOP 2
END SUBROUTINE lsf_fail

!*************************************************************************************!
! Inside/outside sign search: replaces the loop nest set3d.f90:196-268 (centroids,
! nearest-centroid search, triple product, phiSign with gM = 1).  phi keeps its fill
! value (1., set3d.f90:161) outside the sub-box im..ip, jm..jp, km..kp.
!*************************************************************************************!
SUBROUTINE signSearch_b200(phi,nx,ny,nz,xLo,dx,surfX,nSurfNode,surfElem,nSurfElem,im,ip,jm,jp,km,kp)
   INTEGER, INTENT(IN) :: nx,ny,nz,im,ip,jm,jp,km,kp
   INTEGER(c_int32_t), INTENT(IN) :: nSurfNode,nSurfElem
   REAL(c_double), INTENT(IN) :: xLo(3),dx
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(INOUT) :: phi
   REAL(c_double), DIMENSION(nSurfNode,3), INTENT(IN) :: surfX
   INTEGER(c_int32_t), DIMENSION(nSurfElem,3), INTENT(IN) :: surfElem
   INTEGER(c_int) :: rc
   rc = c_lsf_sign_init(phi,INT(nx,c_int),INT(ny,c_int),INT(nz,c_int),xLo,dx,surfX,INT(nSurfNode,c_int), &
                        surfElem,INT(nSurfElem,c_int),INT(im,c_int),INT(ip,c_int),INT(jm,c_int), &
                        INT(jp,c_int),INT(km,c_int),INT(kp,c_int))
   IF (rc < 0) CALL lsf_fail("signSearch_b200", rc)
END SUBROUTINE signSearch_b200

!*************************************************************************************!
! SUBROUTINE reinit(phi,gradPhi,gradPhiMag,nx,ny,nz,iter,dx,h), subs.f90:717-931.
! Same argument list; prints the reference's per-iteration lines (subs.f90:916,923)
! from the returned RMS history and STOPs on a NaN RMS like subs.f90:926.
!*************************************************************************************!
SUBROUTINE reinit_b200(phi,gradPhi,gradPhiMag,nx,ny,nz,iter,dx,h)
   INTEGER, INTENT(IN) :: nx,ny,nz,iter
   REAL(c_double), INTENT(IN) :: dx,h
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(INOUT) :: phi
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(INOUT), TARGET :: gradPhiMag
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz,3), INTENT(INOUT), TARGET :: gradPhi
   CALL reinit_core_b200(phi,C_LOC(gradPhi),C_LOC(gradPhiMag),nx,ny,nz,iter,dx,h)
END SUBROUTINE reinit_b200

!*************************************************************************************!
! reinit without the gradPhi / gradPhiMag outputs.  Both are dead at the reference's two
! call sites (zeroed at set3d.f90:372-375 after the first call, never read after the second,
! :582), and shipping them costs 4 x 8 B per point over PCIe in each direction plus a replay
! of the last sweep; a driver that does not need them calls this instead of reinit_b200.
!*************************************************************************************!
SUBROUTINE reinit_nograd_b200(phi,nx,ny,nz,iter,dx,h)
   INTEGER, INTENT(IN) :: nx,ny,nz,iter
   REAL(c_double), INTENT(IN) :: dx,h
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(INOUT) :: phi
   CALL reinit_core_b200(phi,C_NULL_PTR,C_NULL_PTR,nx,ny,nz,iter,dx,h)
END SUBROUTINE reinit_nograd_b200

SUBROUTINE reinit_core_b200(phi,pGrad,pGradMag,nx,ny,nz,iter,dx,h)
   INTEGER, INTENT(IN) :: nx,ny,nz,iter
   REAL(c_double), INTENT(IN) :: dx,h
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(INOUT) :: phi
   TYPE(c_ptr), INTENT(IN) :: pGrad,pGradMag
   REAL(c_double), ALLOCATABLE :: hist(:)
   INTEGER(c_int) :: rc, n_exit
   ALLOCATE(hist(0:iter))
   rc = c_lsf_reinit(phi,pGrad,pGradMag,INT(nx,c_int),INT(ny,c_int),INT(nz,c_int),INT(iter,c_int), &
                     dx,h,n_exit,hist)
   IF (rc < 0) CALL lsf_fail("reinit_b200", rc)
   CALL print_reinit_history(hist,iter,n_exit,rc)
   DEALLOCATE(hist)
   IF (rc == LSF_NAN) STOP
   PRINT*,
END SUBROUTINE reinit_core_b200

! the reference's per-sweep lines (subs.f90:916,923) from the returned RMS history
SUBROUTINE print_reinit_history(hist,iter,n_exit,rc)
   INTEGER, INTENT(IN) :: iter
   REAL(c_double), INTENT(IN) :: hist(0:iter)
   INTEGER(c_int), INTENT(IN) :: n_exit, rc
   INTEGER :: n
   DO n = 0,n_exit
      IF (n == n_exit .AND. rc == LSF_OK .AND. hist(n) < 1.E-5) THEN
         PRINT*, " Distance function time integration has reached steady state "
      ELSE
         PRINT*, " Iteration: ",n," ", " RMS Error: ",hist(n)
      END IF
   END DO
END SUBROUTINE print_reinit_history

!*************************************************************************************!
! reinit (subs.f90:717-931) on a device-resident grid -- whole (lsf_grid_create) or z-slab
! (lsf_sgrid_create, one MPI rank per GPU; every rank makes the same call and gets the same
! n_exit / history).  iprint /= 0: this rank prints the reference's lines (rank 0 in an MPI run).
!*************************************************************************************!
SUBROUTINE gridReinit_b200(g,iter,dx,h,iprint)
   TYPE(c_ptr), INTENT(IN) :: g
   INTEGER, INTENT(IN) :: iter,iprint
   REAL(c_double), INTENT(IN) :: dx,h
   REAL(c_double), ALLOCATABLE :: hist(:)
   INTEGER(c_int) :: rc, n_exit
   ALLOCATE(hist(0:iter))
   rc = lsf_grid_reinit(g,INT(iter,c_int),dx,h,1.E-5_c_double,n_exit,hist)
   IF (rc < 0) CALL lsf_fail("gridReinit_b200", rc)
   IF (iprint /= 0) CALL print_reinit_history(hist,iter,n_exit,rc)
   DEALLOCATE(hist)
   IF (rc == LSF_NAN) STOP
   IF (iprint /= 0) PRINT*,
END SUBROUTINE gridReinit_b200

!*************************************************************************************!
! the min/max flow loop (set3d.f90:394-462) on a device-resident grid; nDone = loop index at exit
!*************************************************************************************!
SUBROUTINE gridMinMaxFlow_b200(g,iter,dx,h1,tol,nDone,iprint)
   TYPE(c_ptr), INTENT(IN) :: g
   INTEGER, INTENT(IN) :: iter,iprint
   REAL(c_double), INTENT(IN) :: dx,h1,tol
   INTEGER, INTENT(OUT) :: nDone
   REAL(c_double), ALLOCATABLE :: hist(:)
   INTEGER(c_int) :: rc, n_exit
   INTEGER :: n
   ALLOCATE(hist(MAX(iter,1)))
   rc = lsf_grid_minmax(g,INT(iter,c_int),dx,h1,tol,n_exit,hist)
   IF (rc < 0) CALL lsf_fail("gridMinMaxFlow_b200", rc)
   nDone = n_exit
   IF (iprint /= 0) THEN
      DO n = 1,n_exit
         IF (n == n_exit .AND. rc == LSF_OK .AND. hist(n) < tol) THEN
            PRINT*, " Min/max time integration has reached steady state "
         ELSE
            PRINT*, " Iteration: ",n," ", " RMS Error: ",hist(n)
         END IF
      END DO
   END IF
   DEALLOCATE(hist)
   IF (rc == LSF_NAN) STOP
END SUBROUTINE gridMinMaxFlow_b200

!*************************************************************************************!
! SUBROUTINE narrowBand(nx,ny,nz,dx,phi,phiNB,phiSB), subs.f90:178-207
!*************************************************************************************!
SUBROUTINE narrowBand_b200(nx,ny,nz,dx,phi,phiNB,phiSB)
   INTEGER, INTENT(IN) :: nx,ny,nz
   REAL(c_double), INTENT(IN) :: dx
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(IN) :: phi
   INTEGER(c_int32_t), DIMENSION(0:nx,0:ny,0:nz), INTENT(INOUT) :: phiNB,phiSB
   INTEGER(c_int) :: rc
   rc = c_lsf_narrowband(INT(nx,c_int),INT(ny,c_int),INT(nz,c_int),dx,phi,phiNB,phiSB)
   IF (rc < 0) CALL lsf_fail("narrowBand_b200", rc)
END SUBROUTINE narrowBand_b200

!*************************************************************************************!
! The min/max flow time loop, set3d.f90:394-462 (DO n = 1,iter ... END DO), including the
! per-iteration narrowBand (:460), the RMS / steady-state EXIT (:435-451) and the NaN STOP
! (:458).  tol = 1.E-7 in the reference.  On return phi, phiN, phiNB, phiSB hold what the
! reference loop leaves in them; nDone = the loop index at which it left.
!*************************************************************************************!
SUBROUTINE minMaxFlow_b200(phi,phiN,phiNB,phiSB,nx,ny,nz,iter,dx,h1,tol,nDone)
   INTEGER, INTENT(IN) :: nx,ny,nz,iter
   REAL(c_double), INTENT(IN) :: dx,h1,tol
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(INOUT) :: phi,phiN
   INTEGER(c_int32_t), DIMENSION(0:nx,0:ny,0:nz), INTENT(INOUT) :: phiNB,phiSB
   INTEGER, INTENT(OUT) :: nDone
   REAL(c_double), ALLOCATABLE :: hist(:)
   INTEGER(c_int) :: rc, n_exit
   INTEGER :: n
   ALLOCATE(hist(MAX(iter,1)))
   rc = c_lsf_minmax(phi,phiN,phiNB,phiSB,INT(nx,c_int),INT(ny,c_int),INT(nz,c_int),INT(iter,c_int), &
                     dx,h1,tol,n_exit,hist)
   IF (rc < 0) CALL lsf_fail("minMaxFlow_b200", rc)
   nDone = n_exit
   DO n = 1,n_exit
      IF (n == n_exit .AND. rc == LSF_OK .AND. hist(n) < tol) THEN
         PRINT*, " Min/max time integration has reached steady state "
      ELSE
         PRINT*, " Iteration: ",n," ", " RMS Error: ",hist(n)
      END IF
   END DO
   DEALLOCATE(hist)
   IF (rc == LSF_NAN) STOP
END SUBROUTINE minMaxFlow_b200

!*************************************************************************************!
! The "Advect Nodes" block, set3d.f90:465-501: firstDeriv(order 8) on the stencil band
! (:469-478), surfXX = surfX (:485), setPhiSurf (:487) and the node loop (:489-501, iter =
! 1000 in the reference).  The reference re-interpolates ALL nodes after every single node
! move (O(iter*nSurfNode**2) trilinear interpolations -- the dominant cost of a run on
! cube40.stl); the library moves every node independently and returns bit-identical
! surfXX, phiSurf and gradPhiSurf.  gradPhi itself is not returned: the reference overwrites
! it at :527-535 (firstDeriv order 2 over the whole grid) before any further use.
!*************************************************************************************!
SUBROUTINE advectNodes_b200(xLo,nx,ny,nz,dx,phi,phiSB,nSurfNode,surfX,surfXX,phiSurf,gradPhiSurf,iter)
   INTEGER, INTENT(IN) :: nx,ny,nz,iter
   INTEGER(c_int32_t), INTENT(IN) :: nSurfNode
   REAL(c_double), INTENT(IN) :: xLo(3),dx
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(IN) :: phi
   INTEGER(c_int32_t), DIMENSION(0:nx,0:ny,0:nz), INTENT(IN) :: phiSB
   REAL(c_double), DIMENSION(nSurfNode,3), INTENT(IN) :: surfX
   REAL(c_double), DIMENSION(nSurfNode,3), INTENT(OUT) :: surfXX,gradPhiSurf
   REAL(c_double), DIMENSION(nSurfNode), INTENT(OUT) :: phiSurf
   INTEGER(c_int) :: rc
   INTEGER(c_long_long) :: n_moves
   surfXX = surfX                                                     ! set3d.f90:485
   rc = c_lsf_advect_nodes(phi,phiSB,INT(nx,c_int),INT(ny,c_int),INT(nz,c_int),xLo,dx,surfXX, &
                           INT(nSurfNode,c_int),phiSurf,gradPhiSurf,INT(iter,c_int),n_moves)
   IF (rc < 0) CALL lsf_fail("advectNodes_b200", rc)
END SUBROUTINE advectNodes_b200

!*************************************************************************************!
! SUBROUTINE stlRead(...), subs.f90:17-121, same argument list and results; the vertex
! de-duplication (:68-93, O(ntri*nSurfNode) in the reference) runs hash-based in the library
! with the reference's own match predicate and numbering.  (nBndComp is set BEFORE bndNormal is
! allocated here; the reference allocates with nBndComp still undefined, :112-117.)
!*************************************************************************************!
SUBROUTINE stlRead_b200(surfX,nSurfNode,surfElem,filename,nSurfElem,surfElemTag,surfOrder,nBndComp,nBndElem,bndNormal)
   CHARACTER, INTENT(IN) :: filename*80
   INTEGER(c_int32_t) :: nSurfNode,nSurfElem
   INTEGER(c_int32_t),ALLOCATABLE,DIMENSION(:,:),INTENT(OUT) :: surfElem
   REAL(c_double),ALLOCATABLE,DIMENSION(:,:),INTENT(OUT) :: surfX
   INTEGER,ALLOCATABLE,DIMENSION(:),INTENT(OUT) :: surfElemTag,surfOrder
   REAL(c_double),ALLOCATABLE,DIMENSION(:,:),INTENT(OUT) :: bndNormal
   INTEGER,INTENT(OUT) :: nBndComp,nBndElem
   REAL(c_float),ALLOCATABLE,DIMENSION(:,:) :: triangles,nodesT
   CHARACTER(KIND=c_char) :: cpath(81)
   INTEGER(c_int) :: rc,ntri,nn
   INTEGER :: k
   PRINT*,
   PRINT*, " Reading in .stl Mesh "
   PRINT*,
   CALL c_path(filename,cpath)
   rc = c_lsf_stl_count(cpath,ntri)
   IF (rc < 0) CALL lsf_fail("stlRead_b200", rc)
   ALLOCATE(triangles(3,MAX(ntri*3,1)))
   ALLOCATE(nodesT(3,MAX(ntri*3,1)))
   ALLOCATE(surfElem(ntri,3))
   rc = c_lsf_stl_read_triangles(cpath,ntri,triangles)
   IF (rc < 0) CALL lsf_fail("stlRead_b200", rc)
   rc = c_lsf_stl_dedup(triangles,ntri,nodesT,surfElem,nn)
   IF (rc < 0) CALL lsf_fail("stlRead_b200", rc)
   nSurfElem = ntri
   nSurfNode = nn
   ALLOCATE(surfX(nSurfNode,3))
   DO k = 1,nSurfNode                                                 ! subs.f90:99-103
      surfX(k,1) = nodesT(1,k)
      surfX(k,2) = nodesT(2,k)
      surfX(k,3) = nodesT(3,k)
   END DO
   DEALLOCATE(nodesT)
   DEALLOCATE(triangles)
   nBndElem = 0
   nBndComp = 1
   ALLOCATE(surfOrder(nSurfElem))
   ALLOCATE(surfElemTag(nSurfElem))
   ALLOCATE(bndNormal(nBndComp,3))
   surfOrder = 1
   surfElemTag = 0
   bndNormal = 0.
END SUBROUTINE stlRead_b200

!*************************************************************************************!
! The ParaView writer of set3d.f90:320-351 / :539-569 (byte-identical file, including the
! 4-byte nbytePhi = (nx+1)**3*24 of :330): CALL writeVti_b200('signedDistanceFunction.vti',...)
!*************************************************************************************!
SUBROUTINE writeVti_b200(fname,phi,nx,ny,nz,xLo,dx)
   CHARACTER(LEN=*), INTENT(IN) :: fname
   INTEGER, INTENT(IN) :: nx,ny,nz
   REAL(c_double), DIMENSION(0:nx,0:ny,0:nz), INTENT(IN) :: phi
   REAL(c_double), INTENT(IN) :: xLo(3),dx
   CHARACTER(KIND=c_char) :: cpath(LEN(fname)+1)
   INTEGER(c_int) :: rc
   CALL c_path(fname,cpath)
   rc = c_lsf_write_vti(cpath,phi,INT(nx,c_int),INT(ny,c_int),INT(nz,c_int),xLo,dx)
   IF (rc < 0) CALL lsf_fail("writeVti_b200", rc)
END SUBROUTINE writeVti_b200

!*************************************************************************************!
! The .s3d writer of set3d.f90:600-614; surfElem is passed 0-based, i.e. after :590-594
!*************************************************************************************!
SUBROUTINE writeS3d_b200(meshname,nSurfElem,nSurfNode,nBndElem,nBndComp,surfOrder,surfElem,surfElemTag,surfXX,bndNormal)
   CHARACTER(LEN=*), INTENT(IN) :: meshname
   INTEGER(c_int32_t), INTENT(IN) :: nSurfElem,nSurfNode
   INTEGER, INTENT(IN) :: nBndElem,nBndComp
   INTEGER(c_int32_t), INTENT(IN) :: surfOrder(nSurfElem),surfElem(nSurfElem,3),surfElemTag(nSurfElem)
   REAL(c_double), INTENT(IN) :: surfXX(nSurfNode,3),bndNormal(nBndComp,3)
   CHARACTER(KIND=c_char) :: cpath(LEN(meshname)+1)
   INTEGER(c_int) :: rc
   CALL c_path(meshname,cpath)
   rc = c_lsf_write_s3d(cpath,INT(nSurfElem,c_int),INT(nSurfNode,c_int),INT(nBndElem,c_int),INT(nBndComp,c_int), &
                        surfOrder,surfElem,surfElemTag,surfXX,bndNormal)
   IF (rc < 0) CALL lsf_fail("writeS3d_b200", rc)
END SUBROUTINE writeS3d_b200

! blank-padded Fortran name -> NUL-terminated C string
SUBROUTINE c_path(fname,cpath)
   CHARACTER(LEN=*), INTENT(IN) :: fname
   CHARACTER(KIND=c_char), INTENT(OUT) :: cpath(*)
   INTEGER :: q,n
   n = LEN_TRIM(fname)
   DO q = 1,n
      cpath(q) = fname(q:q)
   END DO
   cpath(n+1) = C_NULL_CHAR
END SUBROUTINE c_path

END MODULE lsf_b200
