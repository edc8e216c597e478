// lsf_internal.cuh -- internal state shared by the translation units of liblsf_b200.so.
#pragma once
#include "../../include/lsf_b200.h"
#include "lsf_common.cuh"
#include "lsf_slab.cuh"

struct lsf_grid {
    lsf::Dims dm;             // LOCAL array extents (a z-slab: owned + ghost planes)
    long long np;             // (nx+1)(ny+1)(nz+1) of the local array
    double *phi;              // the level set, reference layout
    double *phiS;             // frozen sign source of reinit (subs.f90:731)
    double *phiN;             // previous iterate for the RMS test (subs.f90:732,921; set3d.f90:377,454)
    double *lap;              // min/max: Laplacian of the iteration's old phi on band cells
    uint8_t *mask;            // min/max: band mask of the current iteration
    double *partial;          // per-block partial sums of the RMS reduction
    double *hist;             // device copy of the per-iteration RMS history
    int hist_cap;
    lsf::Ctrl *ctrl;          // device control block
    // march schedule state (lsf_march.cu)
    unsigned int *march_ticket;   // tile ticket counter
    int *march_colnext;           // dynamic tile scheduler: next tile row of every tile column (lsf_march.cuh: march_pick)
    int march_colnext_cap;
    long long *march_progress;    // per column-tile progress, epoch-encoded
    int march_tiles_cap;
    long long march_epoch;
    // z-slab sharding (lsf_slab.cu); sg.nranks == 1: the whole grid lives on this GPU
    lsf::SlabGeom sg;
    void *shared_base;            // sharded: ONE peer-visible allocation [SlabSync | phi | phiN]
    size_t shared_bytes;
    lsf::SlabSync *sync;          // = shared_base
    void *peer_base[lsf::SLAB_MAX_RANKS];   // the ranks' shared allocations mapped into this process (own entry = shared_base)
    bool attached;
    long long phase;              // number of ghost-plane exchanges so far (lockstep on all ranks)
    long long sum_seq;            // number of cross-rank reductions so far
    long long box_round;          // number of host-side mailbox rendezvous so far (lockstep on all ranks)
    unsigned int *exch_counter;
    // active-list min/max flow (lsf_mm_list.cu)
    long long *mml_list;          // linear indices of the cells that can still change, ascending
    long long mml_n, mml_cap;
    int *mml_counts;              // scratch of the ordered compaction (per 2048-point chunk), kept between calls
    long long *mml_offsets;
    long long mml_scratch_cap;
    uint8_t *mml_unres;           // per point: undecided in the current iteration (all zero between iterations)
    long long *mml_work;          // queue of undecided cells
    int *mml_work_count;
    // fp32 mode (lsf_f32.cu): f32 != 0 -> the fields live in these float arrays and the double ones are null
    int f32;
    float *phi_f, *phiS_f, *phiN_f;
    // the stencil band phiSB of the reference after the min/max loop is the band of the field the LAST narrowBand
    // call saw (set3d.f90:460): phiN after a tolerance EXIT, phi otherwise (lsf_nodes.cu)
    bool sb_from_phiN;
    // overlapped sweeps (lsf_march.cu: launch_reinit_sweeps_overlapped)
    long long *ov_progress;       // per sweep slot and tile
    double *ov_partial;           // per sweep slot: interior partials [ntiles] then boundary partials [ntiles]
    int ov_tiles_cap;
    double *ov_snap;              // phi before the current batch (tolerance EXIT inside a batch -> roll back and replay)
    bool prev_sweep_valid;        // the previous launch on this grid was a sweep of the same free-running sequence
    long long prev_sweep_epoch;
    int prev_sweep_fb;
};

namespace lsf {

typedef lsf_grid Grid;

constexpr int RMS_BLOCKS = 1184;   // 8 x 148 SMs
constexpr int BC_BLOCKS = 592;     // 4 x 148 SMs
constexpr int PARTIAL_CAP = 65536 + RMS_BLOCKS;

struct Global {
    bool inited = false;
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int arith = LSF_ARITH_AUTO;    // requested mode
    int arith_run = LSF_ARITH_FAST; // arithmetic the kernels currently use (FAST or EXACT)
    int arith_last = LSF_ARITH_FAST;
    int sched = LSF_SCHED_MARCH;
    int mm_algo = LSF_MINMAX_LIST;
    int prec = LSF_PREC_F64;       // precision of the host-buffer lsf_reinit entry point
    bool overlap = false;          // reinit: run the sweeps in overlapped batches (lsf_set_overlap)
    long long mm_active = 0;       // length of the active list of the most recent min/max call (this rank)
    int n_launch = 0;
    double last_ms = 0.;
    bool profile = false;
    double sweep_ms = 0.;
    int n_sweeps = 0;
    char err[512] = {0};
};
extern Global G;

int set_error(int code, const char *fmt, ...);
#define LSF_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return lsf::set_error(LSF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,        \
                                  cudaGetErrorString(e__));                                         \
    } while (0)

// lsf_kernels.cu
void launch_reinit_sweep_plane(Grid *g, int raster, const CellConst &cc, double *gradPhi, double *gradPhiMag, double *phi = nullptr);
void launch_reinit_bc(Grid *g, double dx);
void launch_reinit_bc_buf(Grid *g, double *buf, double dx);
void launch_reinit_bc_rms(Grid *g, double dx, int partial_off);
void launch_reinit_bc_rms_buf(Grid *g, double *buf, double dx, double *partial);
void launch_rms(Grid *g, bool copy);
void launch_finalize(Grid *g, int npart, int hist_off, double tol, const double *partial = nullptr);
void launch_copy_boundary(Grid *g, double *dst, const double *src);
void launch_copy_if_running(Grid *g, double *dst, const double *src);
void launch_narrowband(Grid *g, const double *phi, double dx, int32_t *nb, int32_t *sb);
void launch_minmax_iteration_plane(Grid *g, double dx, double h1, bool mask_given);
void launch_mask_from_i32(Grid *g, const int32_t *nb);
void launch_fill(Grid *g, double *p, double v);
void launch_checksum(const void *p, size_t elem_bytes, long long n, long long first_global_index, unsigned long long *d_out);
void launch_sign_init(Grid *g, const double xLo[3], double dx, const double *d_surfX, int nNode,
                      const int32_t *d_surfElem, int nElem, double *d_cen,
                      int im, int ip, int jm, int jp, int km, int kp);

// lsf_sign_bvh.cu
bool launch_sign_search_bvh(Grid *g, const double xLo[3], double dx, const double *d_surfX, int nNode, const int32_t *d_surfElem,
                            int nElem, const double *d_cen, int im, int jm, int km, int ni, int nj, int nk);

inline long long global_cells(const Grid *g) { return (long long)g->dm.nx * g->dm.ny * g->sg.NZ; }
inline bool sharded(const Grid *g) { return g->sg.nranks > 1; }

// lsf_slab.cu
int slab_check_attached(Grid *g);
// ghost-plane refresh of buf (default phi); no-op on one GPU.  handshake = false: the caller guarantees that
// the neighbours no longer read the ghost planes being overwritten (lsf_api.cu, active-list min/max)
void slab_exchange(Grid *g, bool in_loop, double *buf = nullptr, bool handshake = true);   // fp32 grid: exchanges phi_f
void slab_exchange_raw(Grid *g, bool in_loop, char *buf, size_t esize, bool handshake);
void launch_finalize_slab(Grid *g, int npart, int hist_off, double tol, int n);
long long slab_publish_sum(Grid *g, int npart);  // this rank's sum of partials -> every rank (fire and forget); returns its sequence number
void slab_decide(Grid *g, long long seq_first, int count, int n_first, int hist_off, double tol);   // EXIT / NaN tests of `count` iterations
// host-side rendezvous of all ranks through the SlabSync mailboxes: every rank contributes `bytes` (<= SLAB_BOX_BYTES, may be
// 0) and, if `all` is given, receives the ranks' contributions at all + SLAB_BOX_BYTES * rank.  Doubles as a host-level barrier.
int slab_host_exchange(Grid *g, const void *mine, size_t bytes, void *all);
// all ranks' streams meet here (device-side; the host does not wait)
void slab_device_barrier(Grid *g);
// view of a field of the global grid; `mine` = this rank's local array of that field (phi or phiN)
SlabView slab_view(const Grid *g, const double *mine);
int sgrid_shadow_f64(Grid *g, lsf_grid **shadow);      // sharded fp64 twin of a sharded (fp32) grid, attached; collective
int sgrid_shadow_release(Grid *g, lsf_grid *shadow);   // collective
template <class T> inline T *peer_ptr(const Grid *g, int rank, T *mine)
{
    return (T *)((char *)g->peer_base[rank] + ((char *)mine - (char *)g->shared_base));
}

// lsf_march.cu
int march_prepare(Grid *g);
int march_ntiles(const Grid *g);
void launch_reinit_sweep_march(Grid *g, int raster, const CellConst &cc);
void launch_reinit_sweep_march_f32(Grid *g, int raster, const CellConst &cc);
constexpr int OV_BATCH = 8;      // sweeps per overlapped launch (even: the boundary values are back in phi after a full batch)
int launch_reinit_sweeps_overlapped(Grid *g, int n_first, int nsweeps, const CellConst &cc, double tol);
const int *march_order();

// lsf_f32.cu -- fp32 grids (g->f32)
int f32_upload(Grid *g, const double *host, float *dev, long long n);      // n elements, host -> dev
int f32_download(Grid *g, const float *dev, double *host, long long n);
void launch_reinit_bc_rms_f32(Grid *g, double dx, int partial_off);
int f32_sign_init(Grid *g, const double xLo[3], double dx, const double *d_surfX, int nNode, const int32_t *d_surfElem, int nElem,
                  double *d_cen, int im, int ip, int jm, int jp, int km, int kp);
int f32_fill(Grid *g, double value);
int f32_narrowband(Grid *g, double dx, int32_t *d_nb, int32_t *d_sb);
int f32_reinit(Grid *g, int iter, double dx, double h, double tol, int *n_exit, double *rms_hist);
int f32_shadow_open(Grid *g, lsf_grid **shadow);
int f32_shadow_close(Grid *g, lsf_grid *shadow, bool write_back);

// lsf_rk.cu -- K2' throughput mode (Jacobi WENO5 + TVD-RK3; not the reference's algorithm)
long long rk_nblocks(const Grid *g);
void launch_rk_stage(Grid *g, const double *in, const double *phin, double *out, const CellConst &cc, double a, double b,
                     double *scratch_partial, double *rms_out, Grid *halo_grid = nullptr);

// lsf_mm_march.cu
int mm_march_prepare(Grid *g);
void launch_mm_check_boundary(Grid *g, const uint8_t *mask, double dx, bool check_abs);
void launch_minmax_iteration_march(Grid *g, const double *A, double *B, const uint8_t *mask, double dx, double h1);

// lsf_mm_list.cu
int mml_prepare(Grid *g, const uint8_t *mask, double dx);
int mml_npart();
void launch_minmax_iteration_list(Grid *g, const double *A, double *B, const uint8_t *mask, double dx, double h1);

}  // namespace lsf
