// lsf_slab.cu -- z-slab sharding over the GPUs of one node (design: lsf_slab.cuh): creation of the
// peer-visible slab, CUDA-IPC attachment, the ghost-plane exchange kernel and the cross-rank reduction of
// the RMS exit test.  All inter-GPU traffic is peer stores over NVLink issued by these kernels and by the
// sweep kernels themselves (streaming halo); the host only exchanges the IPC handles once.
#include <stdlib.h>
#include <string.h>

#include "lsf_internal.cuh"
#include "lsf_march.cuh"

namespace lsf {

static_assert(sizeof(cudaIpcMemHandle_t) <= LSF_IPC_HANDLE_BYTES, "IPC handle does not fit the ABI's buffer");

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int slab_check_attached(Grid *g)
{
    if (sharded(g) && !g->attached) return set_error(LSF_ERR_ARG, "sharded grid used before lsf_sgrid_attach");
    return LSF_OK;
}

// =====================================================================================
// Ghost-plane exchange.  Side s = 0: low neighbour, 1: high neighbour.
//  1. tell both neighbours "I have finished compute phase e" (they may now overwrite my ghost planes'
//     SOURCE -- i.e. nothing -- and, more to the point, I may overwrite THEIR ghost planes only once they
//     have said the same to me: their sweep may still be reading the old snapshot);
//  2. wait for the neighbours' announcements;
//  3. copy my 3 outermost owned planes on each side into the neighbour's ghost planes (peer stores);
//  4. the last CTA to finish publishes halo_seq = e on both neighbours (fence.sys + release).
// =====================================================================================
struct ExchArgs {
    SlabSync *self;
    SlabSync *nbr[2];
    const uint32_t *src[2];    // the planes are moved as 32-bit words (fp64 and fp32 fields alike)
    uint32_t *dst[2];
    long long n;               // words per side (3 planes)
    long long phase;
    unsigned *counter;
    Ctrl *ctrl;
    int in_loop;
    int handshake;
};

__global__ void __launch_bounds__(256)
k_slab_exchange(const ExchArgs a)
{
    if (a.in_loop && a.ctrl->done) return;
    if (a.handshake) {
        if (blockIdx.x == 0 && threadIdx.x < 2 && a.nbr[threadIdx.x]) {
            p_fence_sys();
            p_st_release_sys(&a.nbr[threadIdx.x]->phase_done[1 - threadIdx.x], a.phase);
        }
        if (threadIdx.x < 2 && a.nbr[threadIdx.x]) wait_ge<true>(&a.self->phase_done[threadIdx.x], a.phase, a.ctrl);
        __syncthreads();
    }
    const long long n4 = a.n / 4;     // 16-byte chunks; the tail (< 4 words) is handled below
    for (int s = 0; s < 2; ++s) {
        if (!a.nbr[s]) continue;
        const uint4 *src = reinterpret_cast<const uint4 *>(a.src[s]);
        uint4 *dst = reinterpret_cast<uint4 *>(a.dst[s]);
        const bool vec = ((((uintptr_t)a.src[s]) | ((uintptr_t)a.dst[s])) & 15) == 0;
        if (vec) {
            for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x)
                __stcg(dst + q, __ldcg(src + q));
            if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) __stcg(a.dst[s] + 4 * n4 + threadIdx.x, __ldcg(a.src[s] + 4 * n4 + threadIdx.x));
        } else {
            for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < a.n; q += (long long)gridDim.x * blockDim.x)
                __stcg(a.dst[s] + q, __ldcg(a.src[s] + q));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        p_fence_sys();
        const unsigned prev = atomicAdd(a.counter, 1u);
        if (prev == gridDim.x - 1) {
            *a.counter = 0;
            p_fence_sys();
            for (int s = 0; s < 2; ++s)
                if (a.nbr[s]) p_st_release_sys(&a.nbr[s]->halo_seq[1 - s], a.phase);
        }
    }
}

void slab_exchange(Grid *g, bool in_loop, double *buf, bool handshake)
{
    if (!sharded(g)) return;
    if (g->f32) { slab_exchange_raw(g, in_loop, (char *)g->phi_f, sizeof(float), handshake); return; }
    if (!buf) buf = g->phi;
    slab_exchange_raw(g, in_loop, (char *)buf, sizeof(double), handshake);
}

// buf: a field inside the rank's shared allocation (phi / phiN, fp64 or fp32), esize: bytes per element
void slab_exchange_raw(Grid *g, bool in_loop, char *buf, size_t esize, bool handshake)
{
    if (!sharded(g)) return;
    const SlabGeom &sg = g->sg;
    const long long pb = (long long)g->dm.sxy * (long long)esize;       // bytes per plane
    ExchArgs a;
    memset(&a, 0, sizeof(a));
    a.self = g->sync;
    a.n = (long long)SLAB_GHOST * pb / 4;
    a.phase = ++g->phase;
    a.counter = g->exch_counter;
    a.ctrl = g->ctrl;
    a.in_loop = in_loop ? 1 : 0;
    a.handshake = handshake ? 1 : 0;
    if (sg.rank > 0) {
        SlabGeom ng;
        slab_geom(sg.NZ, sg.nranks, sg.rank - 1, ng);
        a.nbr[0] = peer_ptr(g, sg.rank - 1, g->sync);
        a.src[0] = (const uint32_t *)(buf + (long long)sg.own_lo * pb);                                   // my global planes k0..k0+2
        a.dst[0] = (uint32_t *)(peer_ptr(g, sg.rank - 1, buf) + (long long)(sg.k0 - ng.kbase) * pb);      // = its upper ghost planes
    }
    if (sg.rank < sg.nranks - 1) {
        SlabGeom ng;
        slab_geom(sg.NZ, sg.nranks, sg.rank + 1, ng);
        a.nbr[1] = peer_ptr(g, sg.rank + 1, g->sync);
        a.src[1] = (const uint32_t *)(buf + (long long)(sg.own_hi - SLAB_GHOST + 1) * pb);                // my global planes k1-3..k1-1
        a.dst[1] = (uint32_t *)(peer_ptr(g, sg.rank + 1, buf) + (long long)(sg.k1 - SLAB_GHOST - ng.kbase) * pb);      // = its lower ghost planes
    }
    k_slab_exchange<<<2 * G.num_sms, 256, 0, G.stream>>>(a);
    G.n_launch++;
    g->prev_sweep_valid = false;      // the per-tile completion flags of earlier sweeps are history now
}

// =====================================================================================
// Loop control on a sharded grid: k_finalize (lsf_kernels.cu) with the sum taken over all ranks, split in two
// so that the reduction never stalls the pipeline of sweeps:
//   k_rank_publish : adds up this rank's partials in a fixed order and writes the result (and its status
//                    flags) into slot seq % SLOTS of EVERY rank's SlabSync block -- fire and forget;
//   k_decide_slab  : waits until all P contributions of `count` consecutive sequence numbers have arrived,
//                    sums each in rank order and replays the reference's per-sweep test (subs.f90:914-926) on
//                    them in order.  All ranks compute bit-identical phiErr values and take the same EXIT /
//                    NaN / error decision on the device: no host round trip, no collective library call.
// The reinit loop decides only where the sweep direction along k flips (after rasters 1 and 5,
// subs.f90:743-852): there the ranks have to wait for each other anyway (the pipeline reverses).
// =====================================================================================
struct PeerSyncs { SlabSync *s[SLAB_MAX_RANKS]; };

__global__ void __launch_bounds__(256)
k_rank_publish(const double *__restrict__ partial, int npart, Ctrl *ctrl, PeerSyncs peers, int rank, int nranks, long long seq)
{
    if (ctrl->done) return;
    __shared__ double sh[256];
    double acc = 0.;
    for (int q = threadIdx.x; q < npart; q += 256) acc = __dadd_rn(acc, partial[q]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    const int slot = (int)(seq % SLAB_SUM_SLOTS);
    if (threadIdx.x < nranks) {
        const int st = *(volatile int *)&ctrl->status;
        const int flags = (ctrl->guard ? 1 : 0) | (st == LSF_ERR_BAND_ON_BOUNDARY ? 2 : 0) | (st == LSF_ERR_TIMEOUT ? 4 : 0) |
                          ((st < 0 && st != LSF_ERR_BAND_ON_BOUNDARY && st != LSF_ERR_TIMEOUT) ? 8 : 0);
        SlabSync *q = peers.s[threadIdx.x];
        *(volatile double *)&q->rank_sum[slot][rank] = sh[0];
        *(volatile int *)&q->rank_flag[slot][rank] = flags;
        p_fence_sys();
        p_st_release_sys(&q->sum_seq[slot][rank], seq);
    }
}

__global__ void __launch_bounds__(256)
k_decide_slab(Ctrl *ctrl, double *__restrict__ hist, int hist_off, double denom, double tol, SlabSync *self, int nranks,
              long long seq_first, int count, int n_first)
{
    if (ctrl->done) return;
    for (int t = threadIdx.x; t < count * nranks; t += blockDim.x) {
        const long long seq = seq_first + t / nranks;
        wait_ge<true>(&self->sum_seq[seq % SLAB_SUM_SLOTS][t % nranks], seq, ctrl);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const bool timed_out = *(volatile int *)&ctrl->status == LSF_ERR_TIMEOUT;
    for (int m = 0; m < count; ++m) {
        const int slot = (int)((seq_first + m) % SLAB_SUM_SLOTS);
        const int n = n_first + m;
        double tot = 0.;
        int flags = timed_out ? 4 : 0;
        for (int r = 0; r < nranks; ++r) {
            tot = __dadd_rn(tot, *(volatile double *)&self->rank_sum[slot][r]);
            flags |= *(volatile int *)&self->rank_flag[slot][r];
        }
        if (flags & 1) ctrl->guard = 1;
        if (flags & 14) {
            ctrl->status = (flags & 4) ? LSF_ERR_TIMEOUT : (flags & 2) ? LSF_ERR_BAND_ON_BOUNDARY : LSF_ERR_ARG;
            ctrl->done = 1; ctrl->n_exit = n; ctrl->n = n;
            return;
        }
        const double err = sqrt(tot / denom);
        hist[n - hist_off] = err;
        if (err < tol) { ctrl->done = 1; ctrl->status = 0; ctrl->n_exit = n; ctrl->n = n; return; }
        if (err != err) { ctrl->done = 1; ctrl->status = 1; ctrl->n_exit = n; ctrl->n = n; return; }
    }
    ctrl->n = n_first + count;
}

static PeerSyncs peer_syncs(const Grid *g)
{
    PeerSyncs ps;
    memset(&ps, 0, sizeof(ps));
    for (int r = 0; r < g->sg.nranks; ++r) ps.s[r] = (SlabSync *)g->peer_base[r];
    return ps;
}

long long slab_publish_sum(Grid *g, int npart)
{
    const long long seq = ++g->sum_seq;
    k_rank_publish<<<1, 256, 0, G.stream>>>(g->partial, npart, g->ctrl, peer_syncs(g), g->sg.rank, g->sg.nranks, seq);
    G.n_launch++;
    return seq;
}

void slab_decide(Grid *g, long long seq_first, int count, int n_first, int hist_off, double tol)
{
    k_decide_slab<<<1, 256, 0, G.stream>>>(g->ctrl, g->hist, hist_off, (double)global_cells(g), tol, g->sync, g->sg.nranks,
                                           seq_first, count, n_first);
    G.n_launch++;
}

// per-iteration form (min/max flow): publish + decide at once; the caller passes the iteration index
void launch_finalize_slab(Grid *g, int npart, int hist_off, double tol, int n)
{
    const long long seq = slab_publish_sum(g, npart);
    slab_decide(g, seq, 1, n, hist_off, tol);
}

// =====================================================================================
// Device-side barrier over all ranks (through the reduction slots: a contribution without a sum)
// =====================================================================================
__global__ void k_slab_barrier(PeerSyncs peers, SlabSync *self, Ctrl *ctrl, int rank, int nranks, long long seq)
{
    const int slot = (int)(seq % SLAB_SUM_SLOTS);
    if (threadIdx.x < nranks) {
        p_fence_sys();
        p_st_release_sys(&peers.s[threadIdx.x]->sum_seq[slot][rank], seq);
        wait_ge<true>(&self->sum_seq[slot][threadIdx.x], seq, ctrl);
    }
}

void slab_device_barrier(Grid *g)
{
    if (!sharded(g)) return;
    const long long seq = ++g->sum_seq;
    k_slab_barrier<<<1, 32, 0, G.stream>>>(peer_syncs(g), g->sync, g->ctrl, g->sg.rank, g->sg.nranks, seq);
    G.n_launch++;
}

SlabView slab_view(const Grid *g, const double *mine)
{
    SlabView v;
    memset(&v, 0, sizeof(v));
    v.nranks = g->sg.nranks;
    v.sxy = g->dm.sxy;
    for (int r = 0; r < g->sg.nranks; ++r) {
        SlabGeom o;
        slab_geom(g->sg.NZ, g->sg.nranks, r, o);
        v.base[r] = sharded(g) ? peer_ptr(g, r, mine) : mine;
        v.shift[r] = (long long)o.kbase * g->dm.sxy;
        v.kend[r] = o.k1;
    }
    return v;
}

// =====================================================================================
// Host-side rendezvous through the mailboxes of the SlabSync blocks (plain cudaMemcpy into the peer-mapped
// allocations): the library has no communicator of its own, and a collective call that has to hand something to
// the other ranks after creation time -- the IPC handle of a transient shadow grid -- uses the memory the host
// program connected once, at lsf_sgrid_attach.
// =====================================================================================
int slab_host_exchange(Grid *g, const void *mine, size_t bytes, void *all)
{
    if (!sharded(g)) { if (all && mine) memcpy(all, mine, bytes); return LSF_OK; }
    if (bytes > (size_t)SLAB_BOX_BYTES) return set_error(LSF_ERR_ARG, "slab_host_exchange: %zu bytes do not fit a mailbox", bytes);
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    const long long seq = ++g->box_round;
    const int me = g->sg.rank, P = g->sg.nranks;
    for (int r = 0; r < P; ++r) {
        SlabSync *q = (SlabSync *)g->peer_base[r];
        if (bytes) LSF_CUDA(cudaMemcpy(q->box[me], mine, bytes, cudaMemcpyDefault));
    }
    for (int r = 0; r < P; ++r) {                      // the payload is in place everywhere this call put it: now the flags
        SlabSync *q = (SlabSync *)g->peer_base[r];
        LSF_CUDA(cudaMemcpy(&q->box_seq[me], &seq, sizeof(seq), cudaMemcpyDefault));
    }
    long long seen[SLAB_MAX_RANKS];
    for (long long spins = 0;; ++spins) {
        LSF_CUDA(cudaMemcpy(seen, g->sync->box_seq, sizeof(long long) * P, cudaMemcpyDeviceToHost));
        bool ok = true;
        for (int r = 0; r < P; ++r) ok = ok && seen[r] >= seq;
        if (ok) break;
        if (spins > 3000000) return set_error(LSF_ERR_TIMEOUT, "slab_host_exchange: a rank did not arrive at rendezvous %lld", seq);
    }
    if (all) LSF_CUDA(cudaMemcpy(all, g->sync->box, (size_t)SLAB_BOX_BYTES * (size_t)P, cudaMemcpyDeviceToHost));
    return LSF_OK;
}

}  // namespace lsf

using namespace lsf;

// =============================================================================================
extern "C" {

int lsf_slab_range(int nz, int nranks, int rank, int *k0, int *k1)
{
    SlabGeom sg;
    if (!slab_geom(nz, nranks, rank, sg))
        return set_error(LSF_ERR_ARG, "slab_range: %d planes cannot be cut into %d slabs of >= %d planes", nz + 1, nranks, SLAB_MIN_PLANES);
    if (k0) *k0 = sg.k0;
    if (k1) *k1 = sg.k1;
    return LSF_OK;
}

static int sgrid_create(lsf_grid **out, int nx, int ny, int nz, int rank, int nranks, bool f32);

int lsf_sgrid_create(lsf_grid **out, int nx, int ny, int nz, int rank, int nranks)
{
    return sgrid_create(out, nx, ny, nz, rank, nranks, false);
}

int lsf_sgrid_create_f32(lsf_grid **out, int nx, int ny, int nz, int rank, int nranks)
{
    return sgrid_create(out, nx, ny, nz, rank, nranks, true);
}

static int sgrid_create(lsf_grid **out, int nx, int ny, int nz, int rank, int nranks, bool f32)
{
    if (!out) return set_error(LSF_ERR_ARG, "null handle");
    *out = nullptr;
    if (nranks == 1) return f32 ? lsf_grid_create_f32(out, nx, ny, nz) : lsf_grid_create(out, nx, ny, nz);
    if (!G.inited) { int rc = lsf_init(-1); if (rc) return rc; }
    if (nx < 2 || ny < 2) return set_error(LSF_ERR_ARG, "grid extents must be >= 2");
    SlabGeom sg;
    if (!slab_geom(nz, nranks, rank, sg))
        return set_error(LSF_ERR_ARG, "sgrid_create: %d planes cannot be cut into %d slabs of >= %d planes (rank %d)", nz + 1, nranks,
                         SLAB_MIN_PLANES, rank);
    Grid *g = (Grid *)calloc(1, sizeof(Grid));
    if (!g) return set_error(LSF_ERR_ARG, "out of host memory");
    g->sg = sg;
    g->dm.nx = nx; g->dm.ny = ny; g->dm.nz = sg.nzl;
    g->dm.sx = (long long)nx + 1;
    g->dm.sxy = g->dm.sx * ((long long)ny + 1);
    g->np = g->dm.sxy * ((long long)sg.nzl + 1);
    // the layout of the shared allocation must be the same on every rank (peer_ptr maps a local address
    // to a peer's by its offset): size the two fields for the thickest slab
    int nzl_max = sg.nzl;
    for (int r = 0; r < nranks; ++r) { SlabGeom o; slab_geom(nz, nranks, r, o); if (o.nzl > nzl_max) nzl_max = o.nzl; }
    const size_t esize = f32 ? sizeof(float) : sizeof(double);
    g->f32 = f32 ? 1 : 0;
    const size_t fbytes = align_up(esize * (size_t)g->dm.sxy * ((size_t)nzl_max + 1) + 64, 256);
    const size_t sbytes = align_up(sizeof(SlabSync), 256);
    g->shared_bytes = sbytes + 2 * fbytes;
    cudaError_t e;
    if ((e = cudaMalloc(&g->shared_base, g->shared_bytes)) != cudaSuccess ||
        (e = cudaMemset(g->shared_base, 0, sbytes)) != cudaSuccess ||
        (e = cudaMalloc(f32 ? (void **)&g->phiS_f : (void **)&g->phiS, esize * (size_t)g->np + 64)) != cudaSuccess ||
        (e = cudaMalloc(&g->partial, sizeof(double) * PARTIAL_CAP)) != cudaSuccess ||
        (e = cudaMalloc(&g->ctrl, sizeof(Ctrl))) != cudaSuccess ||
        (e = cudaMalloc(&g->exch_counter, sizeof(unsigned))) != cudaSuccess ||
        (e = cudaMemset(g->exch_counter, 0, sizeof(unsigned))) != cudaSuccess ||
        (e = cudaMemset(g->ctrl, 0, sizeof(Ctrl))) != cudaSuccess) {
        lsf_grid_destroy(g);
        return set_error(LSF_ERR_CUDA, "sgrid_create: %s", cudaGetErrorString(e));
    }
    g->sync = (SlabSync *)g->shared_base;
    if (f32) {
        g->phi_f = (float *)((char *)g->shared_base + sbytes);
        g->phiN_f = (float *)((char *)g->shared_base + sbytes + fbytes);
    } else {
        g->phi = (double *)((char *)g->shared_base + sbytes);
        g->phiN = (double *)((char *)g->shared_base + sbytes + fbytes);
    }
    g->peer_base[rank] = g->shared_base;
    *out = g;
    return LSF_OK;
}

int lsf_sgrid_ipc_handle(lsf_grid *g, void *handle)
{
    if (!g || !handle) return set_error(LSF_ERR_ARG, "null argument");
    if (!sharded(g)) return set_error(LSF_ERR_ARG, "not a sharded grid");
    cudaIpcMemHandle_t h;
    LSF_CUDA(cudaIpcGetMemHandle(&h, g->shared_base));
    memset(handle, 0, LSF_IPC_HANDLE_BYTES);
    memcpy(handle, &h, sizeof(h));
    return LSF_OK;
}

int lsf_sgrid_attach(lsf_grid *g, const void *handles)
{
    if (!g || !handles) return set_error(LSF_ERR_ARG, "null argument");
    if (!sharded(g)) return LSF_OK;
    if (g->attached) return set_error(LSF_ERR_ARG, "sgrid_attach: already attached");
    for (int r = 0; r < g->sg.nranks; ++r) {
        if (r == g->sg.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)r * LSF_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return set_error(LSF_ERR_CUDA, "sgrid_attach: cudaIpcOpenMemHandle(rank %d) -> %s (GPUs of one node with peer access are required)",
                             r, cudaGetErrorString(e));
        g->peer_base[r] = p;
    }
    g->attached = true;
    return LSF_OK;
}

int lsf_sgrid_sync_ghosts(lsf_grid *g)
{
    if (!g) return set_error(LSF_ERR_ARG, "null grid");
    int rc = slab_check_attached(g);
    if (rc) return rc;
    slab_exchange(g, false);
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    return LSF_OK;
}

}  // extern "C"

namespace lsf {

// A sharded fp64 grid with the geometry of g, created and attached by all ranks together: the IPC handles travel through g's
// mailboxes.  A rank may overwrite a mailbox only after every rank has read the previous content: the rendezvous that
// follows every use (here: the second exchange) guarantees it.
int sgrid_shadow_f64(Grid *g, lsf_grid **shadow)
{
    *shadow = nullptr;
    lsf_grid *sh = nullptr;
    int rc = lsf_sgrid_create(&sh, g->dm.nx, g->dm.ny, g->sg.NZ, g->sg.rank, g->sg.nranks);
    static_assert(LSF_IPC_HANDLE_BYTES < SLAB_BOX_BYTES, "handle + status byte must fit a mailbox");
    unsigned char mine[SLAB_BOX_BYTES], all[SLAB_BOX_BYTES * SLAB_MAX_RANKS], handles[LSF_IPC_HANDLE_BYTES * SLAB_MAX_RANKS];
    memset(mine, 0, sizeof(mine));
    if (!rc) rc = lsf_sgrid_ipc_handle(sh, mine);
    mine[LSF_IPC_HANDLE_BYTES] = rc ? 1 : 0;                 // a rank that failed tells the others instead of leaving them waiting
    int rc2 = slab_host_exchange(g, mine, LSF_IPC_HANDLE_BYTES + 1, all);
    if (!rc2) rc2 = slab_host_exchange(g, nullptr, 0, nullptr);     // everyone has read the boxes
    bool peer_failed = false;
    for (int r = 0; r < g->sg.nranks && !rc2; ++r) {
        peer_failed = peer_failed || all[SLAB_BOX_BYTES * r + LSF_IPC_HANDLE_BYTES];
        memcpy(handles + LSF_IPC_HANDLE_BYTES * r, all + SLAB_BOX_BYTES * r, LSF_IPC_HANDLE_BYTES);
    }
    if (!rc && !rc2 && peer_failed) rc = set_error(LSF_ERR_CUDA, "shadow grid: another rank could not allocate its slab");
    if (!rc && !rc2) rc = lsf_sgrid_attach(sh, handles);
    if (rc || rc2) { lsf_grid_destroy(sh); return rc ? rc : rc2; }
    *shadow = sh;
    return LSF_OK;
}

// Every rank has stopped using the peers' slabs (stream drained + rendezvous), unmaps them, and only after a second rendezvous
// frees its own: memory exported through CUDA IPC must outlive every mapping of it.
int sgrid_shadow_release(Grid *g, lsf_grid *sh)
{
    if (!sh) return LSF_OK;
    int rc = slab_host_exchange(g, nullptr, 0, nullptr);
    for (int r = 0; r < sh->sg.nranks; ++r)
        if (r != sh->sg.rank && sh->peer_base[r]) { cudaIpcCloseMemHandle(sh->peer_base[r]); sh->peer_base[r] = nullptr; }
    const int rc2 = slab_host_exchange(g, nullptr, 0, nullptr);
    lsf_grid_destroy(sh);
    return rc ? rc : rc2;
}

}  // namespace lsf
