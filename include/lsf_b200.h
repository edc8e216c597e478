/*
 * lsf_b200.h -- C ABI of the B200-native grid hot path for LevelSetFortran.
 *
 * The reference (musheen/LevelSetFortran) has no FFI layer: its boundary is the
 * set of `set_subs` module procedures and two inline loops that `set3d.f90`
 * runs over phi(0:nx,0:ny,0:nz).  Per-cell routines (weno, phiSign,
 * secondDeriv, minMax) are the wrong granularity for a device, so this ABI
 * exports the enclosing whole-grid loops.  Each entry point cites the reference
 * code it replaces.  The Fortran ISO_C_BINDING interface for these symbols is in
 * fortran/lsf_b200_mod.f90 and INTEGRATION.md shows the calls to substitute in
 * set3d.f90.
 *
 * Conventions
 *  - All grid arrays are the reference's own: REAL(8)/INTEGER(4), Fortran
 *    column-major, extents (0:nx,0:ny,0:nz), i fastest, contiguous, owned by the
 *    caller.  Host entry points copy in/out; the library owns all device memory.
 *  - Return value: 0 = ok, 1 = a NaN RMS was produced (the reference STOPs at
 *    subs.f90:926 / set3d.f90:458; the caller decides), < 0 = failure
 *    (LSF_ERR_*); lsf_last_error() gives the text.  The library never aborts.
 *  - Calls are synchronous and must come from one host thread per process
 *    (one process per GPU).  There is no CPU fallback: without a CUDA device
 *    every compute entry point returns LSF_ERR_CUDA.
 */
#ifndef LSF_B200_H
#define LSF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSF_OK 0
#define LSF_NAN 1
#define LSF_ERR_CUDA (-1)
#define LSF_ERR_ARG (-2)
#define LSF_ERR_BAND_ON_BOUNDARY (-3) /* a narrow-band cell lies on the grid boundary: the
                                         reference would read phi(-1,..) (set3d.f90:402-403) */
#define LSF_ERR_TIMEOUT (-4)          /* sharded grid: a neighbouring rank stopped answering */
#define LSF_ERR_NODE_OFF_GRID (-5)    /* surface-node projection: a node left the grid (out-of-bounds read in the reference) */
#define LSF_IPC_HANDLE_BYTES 64       /* size of the opaque per-rank handle of lsf_sgrid_ipc_handle */

/* arithmetic of the WENO5 cell update */
#define LSF_ARITH_FAST 0  /* FMA + reciprocal-reduced form; <= 1e-10 of the reference on well-conditioned data */
#define LSF_ARITH_EXACT 1 /* the reference's operation order, IEEE div/sqrt, no FMA: bit-identical */
#define LSF_ARITH_AUTO 2  /* (default) FAST with an on-the-fly conditioning guard; if any cell update is
                             ill-conditioned (flat extremum of phi next to the interface -- where the reference
                             itself produces its 0/0 NaN, subs.f90:169) the call restarts in EXACT */
/* schedule of the in-place Gauss-Seidel sweeps (both are exact re-orderings) */
#define LSF_SCHED_MARCH 0 /* skewed x-marching column tiles, one launch per sweep (default) */
#define LSF_SCHED_PLANE 1 /* one launch per global hyperplane; simple cross-check path */

/* algorithm of the min/max flow iteration (all are exact re-orderings of set3d.f90:399-431, bit-identical) */
#define LSF_MINMAX_LIST 0  /* (default) active list of the cells that can still change, order-free speculative update */
#define LSF_MINMAX_MARCH 1 /* whole-grid skewed march, one fused kernel per iteration; cross-check path */

/* storage / arithmetic precision of the device fields (host arrays are REAL(8) in either mode) */
#define LSF_PREC_F64 0 /* (default) the reference's REAL(8): phi within 1e-10 of the reference, bit-exact in EXACT arithmetic */
#define LSF_PREC_F32 1 /* optional single-precision mode: 12 B per cell update, WENO5 sweep on the FP32 pipe.  Contract per stage:
                          sign search: the fp64 sign field rounded to float (signs and zeros exact); reinit:
                          max|phi - phi_ref| <= 1e-4 * max|phi_ref| at equal sweep counts; min/max flow: exactly
                          float(reference flow applied to the float field) -- the flow is a discontinuous map (band
                          membership, min-or-max switch), so float rounding of its INPUT flips individual cells, which then
                          drift by h1*L per iteration: no uniform bound against the fp64 pipeline exists for that stage */

typedef struct lsf_grid lsf_grid; /* a device-resident phi(0:nx,0:ny,0:nz) plus work arrays */

/* ---- library ---------------------------------------------------------------------------- */
int lsf_init(int device);          /* device < 0: use $LOCAL_RANK, else 0 */
int lsf_finalize(void);
const char *lsf_last_error(void);
int lsf_set_arith(int arith);      /* LSF_ARITH_*  */
int lsf_last_arith(void);          /* arithmetic the most recent lsf_*reinit call finished in (FAST or EXACT) */
int lsf_set_sched(int sched);      /* LSF_SCHED_*  */
int lsf_set_minmax_algo(int algo); /* LSF_MINMAX_* (used with LSF_SCHED_MARCH) */
int lsf_set_overlap(int on);       /* 1: reinit (fp64, single GPU, march schedule) runs its sweeps in overlapped batches of 8 --
                                      one launch per batch, CTAs start the next sweep's tiles while the previous one drains
                                      (same results; opt-in until measured on the target, env LSF_SWEEP_OVERLAP=1) */
int lsf_set_precision(int prec);   /* LSF_PREC_*: precision lsf_reinit (host-buffer entry point) runs in; env LSF_PRECISION=f32 */
long long lsf_last_minmax_active(void); /* LSF_MINMAX_LIST: cells on the active list of the most recent min/max call (this rank) */
/* Timing of the kernels of the most recent lsf_*reinit / lsf_*minmax / lsf_*sign_init call,
 * CUDA events on the library's own stream: total ms, number of kernel launches. */
int lsf_last_timing(double *kernel_ms, int *n_launches);
/* With profiling on, every sweep kernel launch of lsf_*reinit is bracketed by its own event pair;
 * lsf_last_sweep_timing returns their summed duration and count for the most recent call. */
int lsf_set_profile(int on);
int lsf_last_sweep_timing(double *sweep_ms, int *n_sweeps);

/* ---- host-buffer entry points (the drop-in boundary) ------------------------------------ */

/* Optional: page-lock a caller-owned host array (e.g. the driver's ALLOCATEd phi, set3d.f90:160) for the lifetime of
 * its allocation, so that the entry points below move it at PCIe speed instead of through the driver's pageable
 * staging path.  Explicit because only the caller knows when the array is freed: unregister BEFORE DEALLOCATE. */
int lsf_host_register(void *ptr, size_t nbytes);
int lsf_host_unregister(void *ptr);

/* Inside/outside sign search, replaces the inline loop set3d.f90:196-268 (+ phiSign,
 * subs.f90:169).  phi must already hold the caller's fill value (reference: 1., set3d.f90:161);
 * only points of the sub-box [im..ip]x[jm..jp]x[km..kp] are overwritten.
 * surfX: REAL(8) (nSurfNode,3); surfElem: INTEGER(4) (nSurfElem,3), 1-based (subs.f90:84-88). */
int lsf_sign_init(double *phi, int nx, int ny, int nz, const double xLo[3], double dx,
                  const double *surfX, int nSurfNode, const int32_t *surfElem, int nSurfElem,
                  int im, int ip, int jm, int jp, int km, int kp);

/* SUBROUTINE reinit(phi,gradPhi,gradPhiMag,nx,ny,nz,iter,dx,h), subs.f90:717-931.
 * gradPhi (0:nx,0:ny,0:nz,3) / gradPhiMag may be NULL (both are dead downstream,
 * set3d.f90:372-375); when given they receive the last sweep's weno outputs (subs.f90:696-703).
 * n_exit: loop index n at which the routine left; rms_hist[0..iter]: phiErr per sweep
 * (the values the reference prints at subs.f90:923).  Exit tolerance 1.E-5 (subs.f90:915).
 * phiErr is a sum over all points: the library adds per-tile partial sums in a fixed order (deterministic run to run), the
 * reference accumulates serially over i, j, k -- rms_hist agrees to ~1e-15 relative, so even where phi is bit-identical
 * (LSF_ARITH_EXACT) an RMS that lands within that distance of the tolerance can leave the loop one sweep apart from the
 * reference (never observed on the reference's inputs: the RMS changes by ~1e-3 relative per sweep there). */
int lsf_reinit(double *phi, double *gradPhi, double *gradPhiMag, int nx, int ny, int nz,
               int iter, double dx, double h, int *n_exit, double *rms_hist);

/* SUBROUTINE narrowBand(nx,ny,nz,dx,phi,phiNB,phiSB), subs.f90:178-207. */
int lsf_narrowband(int nx, int ny, int nz, double dx, const double *phi,
                   int32_t *phiNB, int32_t *phiSB);

/* The min/max flow time loop, set3d.f90:394-462 (secondDeriv subs.f90:370-407, minMax
 * subs.f90:413-483, RMS/exit :435-451, narrowBand :460).  On entry phiNB/phiSB hold
 * narrowBand(phi) (set3d.f90:360) and phiN = phi (:377); on exit all five arrays hold what
 * the reference loop leaves in them.  tol is 1.E-7 in the reference (:448).
 * rms_hist[0..iter-1]: phiErr of iteration n = 1..iter. */
int lsf_minmax(double *phi, double *phiN, int32_t *phiNB, int32_t *phiSB,
               int nx, int ny, int nz, int iter, double dx, double h1, double tol,
               int *n_exit, double *rms_hist);

/* Surface-node projection ("Advect Nodes"), set3d.f90:465-501: gradPhi = firstDeriv(order 8) (subs.f90:311-347,
 * including the jp1 typo of :346) on the stencil band phiSB == 1 and 0 elsewhere (set3d.f90:372), setPhiSurf
 * (subs.f90:1057-1170), then `iter` (reference: 1000) passes in which every node with phiSurf > 1E-13 moves by
 * phiSurf*gradPhiSurf and is re-interpolated.  surfXX (nSurfNode,3): surfX on entry (set3d.f90:485), the moved
 * nodes on exit; phiSurf (nSurfNode) and gradPhiSurf (nSurfNode,3) as the reference leaves them.  Bit-identical to
 * the reference loop (which re-interpolates ALL nodes after every move: O(iter*nSurfNode^2)).
 * n_moves (may be NULL): node moves executed.  LSF_ERR_NODE_OFF_GRID / LSF_ERR_BAND_ON_BOUNDARY where the reference
 * would read out of bounds. */
int lsf_advect_nodes(const double *phi, const int32_t *phiSB, int nx, int ny, int nz, const double xLo[3], double dx,
                     double *surfXX, int nSurfNode, double *phiSurf, double *gradPhiSurf, int iter, long long *n_moves);

/* ---- host-side pieces either side of the path (no device work) ---------------------------- */

/* The reference's ParaView writer, byte for byte (set3d.f90:320-351 signedDistanceFunction.vti, :539-569
 * smoothedDistanceFunction.vti), including the wrong 4-byte length field nbytePhi = (nx+1)**3*24 (:330). */
int lsf_write_vti(const char *path, const double *phi, int nx, int ny, int nz, const double xLo[3], double dx);
/* The .s3d mesh file of set3d.f90:604-614 (list-directed records, gfortran spacing).  surfElem (nSurfElem,3) is
 * 0-based here, as the reference makes it at :590-594 before writing; surfXX (nSurfNode,3); bndNormal (nBndComp,3). */
int lsf_write_s3d(const char *path, int nSurfElem, int nSurfNode, int nBndElem, int nBndComp, const int32_t *surfOrder,
                  const int32_t *surfElem, const int32_t *surfElemTag, const double *surfXX, const double *bndNormal);
/* stlRead (subs.f90:17-121) without its O(ntri * nSurfNode) search: count, read the raw REAL*4 triangles
 * (9*ntri floats, vertex-major = the reference's triangles(3,ntri*3)), de-duplicate with the reference's own match
 * predicate and first-occurrence numbering.  nodes: work array of 9*ntri floats receiving nodesT(3,k); surfElem:
 * (ntri,3) column-major, 1-based; the caller then fills surfX(k,c) = nodes(c,k) as subs.f90:99-103 does. */
int lsf_stl_count(const char *path, int *ntri);
int lsf_stl_read_triangles(const char *path, int ntri, float *tri);
int lsf_stl_dedup(const float *tri, int ntri, float *nodes, int32_t *surfElem, int *nSurfNode);

/* ---- device-resident pipeline (no host round trip between stages) ------- */
int lsf_grid_create(lsf_grid **g, int nx, int ny, int nz);
/* fp32 mode: phi is stored as float on the device; the same lsf_grid_* calls apply (upload / download convert
 * from / to the caller's REAL(8) arrays).  reinit runs natively in fp32; sign search and min/max flow -- a
 * negligible share of a run -- execute the fp64 kernels on a transient fp64 copy and round the result (node projection reads
 * that copy).  On a sharded fp32 grid (lsf_sgrid_create_f32) the transient copy is a sharded fp64 grid that the ranks create
 * and connect among themselves inside the call (16 B per point while it lives).  Not available: lsf_grid_download_phiN. */
int lsf_grid_create_f32(lsf_grid **g, int nx, int ny, int nz);
int lsf_grid_is_f32(lsf_grid *g);
int lsf_grid_destroy(lsf_grid *g);
int lsf_grid_fill(lsf_grid *g, double value);                       /* phi = value, set3d.f90:161 */
int lsf_grid_upload(lsf_grid *g, const double *phi_host);           /* H2D, dense Fortran layout */
int lsf_grid_download(lsf_grid *g, double *phi_host);               /* D2H */
int lsf_grid_download_phiN(lsf_grid *g, double *phiN_host);
void *lsf_grid_device_ptr(lsf_grid *g);                             /* device address of phi(0,0,0) of the local array; valid until
                                                                       the next lsf_grid_minmax call on g (that call ping-pongs
                                                                       between two buffers and may leave phi in the other one) */
/* Partition-independent digest of the OWNED points of phi: digest[0] = sum of bits(phi(q)) * (2q+1) mod 2^64 over the
 * points, q = global linear index (i + (nx+1)(j + (ny+1)k)); digest[1] = xor of the bit patterns.  The per-rank digests
 * of a sharded grid add / xor up to the digest of the same field on one GPU -- how bench.py and the multi-GPU tests
 * show that an N-GPU result is bit-identical to the single-GPU one without moving the field. */
int lsf_grid_checksum(lsf_grid *g, uint64_t digest[2]);
int lsf_grid_sign_init(lsf_grid *g, const double xLo[3], double dx,
                       const double *surfX, int nSurfNode, const int32_t *surfElem, int nSurfElem,
                       int im, int ip, int jm, int jp, int km, int kp);
int lsf_grid_reinit(lsf_grid *g, int iter, double dx, double h, double tol,
                    int *n_exit, double *rms_hist);
int lsf_grid_narrowband(lsf_grid *g, double dx, int32_t *phiNB_host, int32_t *phiSB_host);
int lsf_grid_minmax(lsf_grid *g, int iter, double dx, double h1, double tol,
                    int *n_exit, double *rms_hist);
/* K2' throughput mode -- NOT the reference's algorithm (the reference is forward-Euler Gauss-Seidel, subs.f90:737-855; its
 * RK scaffolding, set3d.f90:28,285-287, is never used): `steps` TVD Runge-Kutta-3 steps of the Jacobi WENO5 reinitialisation
 * equation phi_t = sgn(phi0)(1 - |grad phi|), the scheme BASELINE.json's north_star names.  Differs from lsf_grid_reinit by
 * ~4e-4 and in iteration count; same per-cell arithmetic, high-order window, boundary block (after every
 * stage) and RMS / EXIT / NaN tests (per step).  fp64; one GPU or z-slabs (there a plain halo problem: the stage buffers'
 * ghost planes are exchanged after every stage; phi bit-identical to the single-GPU run of the mode).  n_exit: 0-based index of
 * the last executed step. */
int lsf_grid_reinit_rk3(lsf_grid *g, int steps, double dx, double dt, double tol, int *n_exit, double *rms_hist);
/* lsf_advect_nodes on the resident phi; phiSB is the band the reference holds at that point: that of the field the
 * last narrowBand call saw (phi, or the previous iterate after a tolerance EXIT of lsf_grid_minmax).  On a sharded grid every
 * rank passes the SAME (whole) node list and receives the same result: each rank projects all nodes, gathering phi from the
 * slabs of whichever ranks a node crosses (peer loads over NVLink); xLo is the origin of the GLOBAL grid. */
int lsf_grid_advect_nodes(lsf_grid *g, const double xLo[3], double dx, double *surfXX, int nSurfNode,
                          double *phiSurf, double *gradPhiSurf, int iter, long long *n_moves);

/* ---- z-slab sharding over the GPUs of one node -----------------------------------
 * The reference is serial ("Parallel version is in the works", README.md:17).  Here phi(0:nx,0:ny,0:nz) is
 * cut along k -- the slowest index, so a slab is a contiguous range of the reference array -- into one
 * slab per process/GPU.  A sharded grid is used through the same lsf_grid_* calls as a whole one
 * (sign_init, reinit, narrowband, minmax, advect_nodes, fill, upload, download): all ranks make the same calls in the same order
 * (SPMD), host arrays hold the rank's OWNED planes k0..k1-1 only, n_exit / rms_hist are identical on all
 * ranks, and the result is bit-identical to the single-GPU one (the in-place Gauss-Seidel sweeps run as a
 * software pipeline along k: peer stores over NVLink from inside the sweep kernels, no collective on the
 * data path).  Set-up: every rank creates its slab, obtains its handle, the host program all-gathers the
 * handles (MPI_Allgather / torch.distributed.all_gather) and every rank attaches. */
int lsf_slab_range(int nz, int nranks, int rank, int *k0, int *k1);   /* owned planes [k0, k1) of rank */
int lsf_sgrid_create(lsf_grid **g, int nx, int ny, int nz, int rank, int nranks);   /* nz: GLOBAL extent */
int lsf_sgrid_create_f32(lsf_grid **g, int nx, int ny, int nz, int rank, int nranks);   /* the same slab in the optional fp32 mode */
int lsf_sgrid_ipc_handle(lsf_grid *g, void *handle /* LSF_IPC_HANDLE_BYTES */);
int lsf_sgrid_attach(lsf_grid *g, const void *handles /* nranks * LSF_IPC_HANDLE_BYTES, rank order */);
int lsf_sgrid_sync_ghosts(lsf_grid *g);   /* refresh the ghost planes after writing phi through lsf_grid_device_ptr */

#ifdef __cplusplus
}
#endif
#endif /* LSF_B200_H */
