// lsf_api.cu -- the C ABI (include/lsf_b200.h): library state, device-resident grid objects,
// the iteration loops of reinit (subs.f90:735-928) and min/max flow (set3d.f90:394-462), and the
// host-buffer drop-in entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "lsf_internal.cuh"

namespace lsf {

Global G;

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(G.err, sizeof(G.err), fmt, ap);
    va_end(ap);
    return code;
}

static size_t owned_off(const Grid *g) { return (size_t)g->sg.own_lo * (size_t)g->dm.sxy; }
static size_t owned_elems(const Grid *g) { return (size_t)(g->sg.k1 - g->sg.k0) * (size_t)g->dm.sxy; }

static int ensure_init()
{
    if (G.inited) return LSF_OK;
    return lsf_init(-1);
}

struct Timer {
    void start() { G.n_launch = 0; cudaEventRecord(G.ev0, G.stream); }
    int stop()
    {
        LSF_CUDA(cudaEventRecord(G.ev1, G.stream));
        LSF_CUDA(cudaEventSynchronize(G.ev1));
        float ms = 0.f;
        LSF_CUDA(cudaEventElapsedTime(&ms, G.ev0, G.ev1));
        G.last_ms = ms;
        return LSF_OK;
    }
};

static int ensure_hist(Grid *g, int n)
{
    if (g->hist_cap >= n) return LSF_OK;
    if (g->hist) cudaFree(g->hist);
    g->hist = nullptr;
    g->hist_cap = 0;
    int cap = n < 16384 ? 16384 : n;
    LSF_CUDA(cudaMalloc(&g->hist, sizeof(double) * (size_t)cap));
    g->hist_cap = cap;
    return LSF_OK;
}

static int read_ctrl(Grid *g, Ctrl *h)
{
    LSF_CUDA(cudaMemcpyAsync(h, g->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, G.stream));
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    return LSF_OK;
}

// per-sweep event pairs of the profiling mode (lsf_set_profile): recorded around every sweep launch, read
// back whenever the host synchronises anyway
struct SweepEvents {
    cudaEvent_t ev[16][2];
    bool init = false;
    int npend = 0;
    void begin() {
        if (!G.profile) return;
        if (!init) { for (int q = 0; q < 16; ++q) { cudaEventCreate(&ev[q][0]); cudaEventCreate(&ev[q][1]); } init = true; }
        cudaEventRecord(ev[npend][0], G.stream);
    }
    void end() { if (G.profile) { cudaEventRecord(ev[npend][1], G.stream); ++npend; } }
    void collect() {
        static const bool log = getenv("LSF_SWEEP_LOG") != nullptr;      // experiments: per-sweep kernel times on stderr
        for (int q = 0; q < npend; ++q) {
            float ms = 0.f, gap = 0.f;
            if (cudaEventElapsedTime(&ms, ev[q][0], ev[q][1]) == cudaSuccess) { G.sweep_ms += ms; G.n_sweeps++; }
            if (log) {
                if (q + 1 < npend) cudaEventElapsedTime(&gap, ev[q][1], ev[q + 1][0]);
                fprintf(stderr, "[lsf sweep] dev %d #%d kernel %.3f ms, then %.3f ms to the next sweep\n", G.device, G.n_sweeps, ms, gap);
            }
        }
        npend = 0;
    }
};
static SweepEvents SE;

// reinit, subs.f90:717-931.  d_gradPhi / d_gradPhiMag: optional device arrays.
// One attempt in the arithmetic G.arith_run.  *guard_hit is set when a FAST attempt met an ill-conditioned
// cell update (then phi is NOT valid and the caller restarts from phiS in EXACT).
static int reinit_attempt(Grid *g, int iter, double dx, double h, double tol, double *d_gradPhi, double *d_gradPhiMag,
                          int *n_exit, double *rms_hist, bool watch_guard, bool *guard_hit)
{
    const size_t bytes = sizeof(double) * (size_t)g->np;
    LSF_CUDA(cudaMemsetAsync(g->ctrl, 0, sizeof(Ctrl), G.stream));
    CellConst cc;
    cc.dx = dx; cc.inv_dx = 1. / dx; cc.k12 = 1. / (12. * dx); cc.dx2 = dx * dx; cc.h = h;
    const bool want_grad = d_gradPhi || d_gradPhiMag;
    const bool march = G.sched == LSF_SCHED_MARCH;
    // gradPhi / gradPhiMag hold the weno outputs of the LAST executed sweep (subs.f90:696-703).  On the march
    // schedule the sweeps themselves do not write them; instead phi is snapshotted before every sweep (the copy
    // turns into a no-op once the loop has left, so the snapshot that survives is the input of the last executed
    // sweep) and that one sweep is replayed afterwards on the snapshot by the plane-schedule kernel, which does.
    const bool grad_replay = march && want_grad;
    int rc;
    if (march) { rc = march_prepare(g); if (rc) return rc; }
    else LSF_CUDA(cudaMemcpyAsync(g->phiN, g->phi, bytes, cudaMemcpyDeviceToDevice, G.stream));   // subs.f90:732
    const int check = march ? 8 : (g->np > 2000000 ? 1 : 8);
    Ctrl hc = {0, 0, 0, 0, 0};
    SE.npend = 0;
    *guard_hit = false;
    // Overlapped sweeps (lsf_set_overlap): batches of OV_BATCH sweeps in one launch each.  With a
    // tolerance the loop can leave inside a batch: phi is snapshotted before every batch and, if that happens, restored
    // and the sweeps up to the exit are replayed one by one, so phi, n_exit and rms_hist are those of the plain loop.
    const bool overlap = march && G.overlap && !want_grad && !sharded(g);
    if (overlap) {
        const bool snap = tol > 0.;
        if (snap && !g->ov_snap) LSF_CUDA(cudaMalloc(&g->ov_snap, bytes + 64));
        for (int n = 0; n <= iter; n += OV_BATCH) {
            const int nb = iter + 1 - n < OV_BATCH ? iter + 1 - n : OV_BATCH;
            if (snap) LSF_CUDA(cudaMemcpyAsync(g->ov_snap, g->phi, bytes, cudaMemcpyDeviceToDevice, G.stream));
            SE.begin();
            rc = launch_reinit_sweeps_overlapped(g, n, nb, cc, tol);
            SE.end();
            if (rc) return rc;
            rc = read_ctrl(g, &hc);
            if (rc) return rc;
            SE.collect();
            if (G.profile) G.n_sweeps += nb - 1;                         // one event pair covered nb sweeps
            if (watch_guard && hc.guard) { *guard_hit = true; return LSF_OK; }
            if (hc.done) {
                if (hc.status >= 0 && hc.n_exit < n + nb - 1 && snap) {
                    LSF_CUDA(cudaMemcpyAsync(g->phi, g->ov_snap, bytes, cudaMemcpyDeviceToDevice, G.stream));
                    LSF_CUDA(cudaMemsetAsync(g->ctrl, 0, sizeof(Ctrl), G.stream));
                    for (int m = n; m <= hc.n_exit; ++m) {
                        launch_reinit_sweep_march(g, m % 8 + 1, cc);
                        launch_reinit_bc_rms(g, dx, march_ntiles(g));
                    }
                    LSF_CUDA(cudaMemcpyAsync(g->ctrl, &hc, sizeof(Ctrl), cudaMemcpyHostToDevice, G.stream));
                    LSF_CUDA(cudaStreamSynchronize(G.stream));
                }
                break;
            }
        }
    } else
    for (int n = 0; n <= iter; ++n) {                                   // subs.f90:735
        const int raster = n % 8 + 1;                                   // subs.f90:740,855
        if (grad_replay) launch_copy_if_running(g, g->phiN, g->phi);
        SE.begin();
        if (march) launch_reinit_sweep_march(g, raster, cc);
        else launch_reinit_sweep_plane(g, raster, cc, d_gradPhi, d_gradPhiMag);
        SE.end();
        if (march) {
            // RMS fused: interior sums come from the sweep (one per column tile), boundary sums from the
            // BC kernel; phiN (subs.f90:732,921) is never materialised on this path
            launch_reinit_bc_rms(g, dx, march_ntiles(g));               // subs.f90:858-897 + boundary part of :902-914
            launch_finalize(g, march_ntiles(g) + BC_BLOCKS, 0, tol);    // subs.f90:914-926
        } else {
            launch_reinit_bc(g, dx);                                    // subs.f90:858-897
            launch_rms(g, true);                                        // subs.f90:902-914,921
            launch_finalize(g, RMS_BLOCKS, 0, tol);                     // subs.f90:914-926
        }
        if ((n + 1) % check == 0 || n == iter) {
            rc = read_ctrl(g, &hc);
            if (rc) return rc;
            SE.collect();
            if (watch_guard && hc.guard) { *guard_hit = true; return LSF_OK; }
            if (hc.done) break;
        }
    }
    LSF_CUDA(cudaGetLastError());
    const int ne = hc.done ? hc.n_exit : iter;
    if (grad_replay && !*guard_hit) {
        // replay sweep `ne` on its input (phiN) with the kernel that writes the weno outputs; both schedules are
        // exact re-orderings of the same loop, so phiN ends up equal to phi
        LSF_CUDA(cudaMemsetAsync(g->ctrl, 0, sizeof(Ctrl), G.stream));
        launch_reinit_sweep_plane(g, ne % 8 + 1, cc, d_gradPhi, d_gradPhiMag, g->phiN);
        LSF_CUDA(cudaMemcpyAsync(g->ctrl, &hc, sizeof(Ctrl), cudaMemcpyHostToDevice, G.stream));
        LSF_CUDA(cudaStreamSynchronize(G.stream));
    }
    if (G.profile && G.n_sweeps > ne + 1) G.n_sweeps = ne + 1;   // sweeps enqueued after the exit were no-ops
    if (n_exit) *n_exit = ne;
    if (rms_hist) LSF_CUDA(cudaMemcpy(rms_hist, g->hist, sizeof(double) * (size_t)(ne + 1), cudaMemcpyDeviceToHost));
    return hc.done ? hc.status : LSF_OK;
}

// reinit on a z-slab (lsf_slab.cuh).  The sweeps of all ranks form one free-running Gauss-Seidel pipeline
// along k: no bulk halo exchange and no reduction barrier between sweeps.  Every rank publishes its RMS
// partial sum after each sweep, but the EXIT / NaN tests (subs.f90:914-926) are evaluated only after rasters
// 1 and 5 -- where the k direction of the sweep flips and the ranks have to wait for each other anyway --
// for all sweeps since the previous evaluation, in order.  If the loop turns out to have left in the middle
// of such a batch, phi is rolled back to the snapshot taken at the previous evaluation (phiN is free on
// this path) and the sweeps up to the exit are replayed, so phi, n_exit and rms_hist are exactly those of
// the sweep-by-sweep loop.  Snapshots are only taken when an EXIT is possible (tol > 0).
// (fp64 and fp32 slabs alike: only the field pointers, the element size and the two launches differ)
static void *field_phi(Grid *g) { return g->f32 ? (void *)g->phi_f : (void *)g->phi; }
static void *field_phiN(Grid *g) { return g->f32 ? (void *)g->phiN_f : (void *)g->phiN; }
static void sweep_any(Grid *g, int raster, const CellConst &cc)
{
    if (g->f32) launch_reinit_sweep_march_f32(g, raster, cc); else launch_reinit_sweep_march(g, raster, cc);
}
static void bc_rms_any(Grid *g, double dx, int off)
{
    if (g->f32) launch_reinit_bc_rms_f32(g, dx, off); else launch_reinit_bc_rms(g, dx, off);
}

static int reinit_attempt_slab(Grid *g, int iter, double dx, double h, double tol, int *n_exit, double *rms_hist,
                               bool watch_guard, bool *guard_hit)
{
    if (G.sched != LSF_SCHED_MARCH) return set_error(LSF_ERR_ARG, "reinit: a sharded grid supports the march schedule only");
    const size_t bytes = (g->f32 ? sizeof(float) : sizeof(double)) * (size_t)g->np;
    LSF_CUDA(cudaMemsetAsync(g->ctrl, 0, sizeof(Ctrl), G.stream));
    CellConst cc;
    cc.dx = dx; cc.inv_dx = 1. / dx; cc.k12 = 1. / (12. * dx); cc.dx2 = dx * dx; cc.h = h;
    int rc = march_prepare(g);
    if (rc) return rc;
    *guard_hit = false;
    const bool snap = tol > 0.;
    const int npart = march_ntiles(g) + BC_BLOCKS;
    Ctrl hc = {0, 0, 0, 0, 0};
    slab_exchange(g, false);                                            // ghost planes = neighbours' current phi
    if (snap) LSF_CUDA(cudaMemcpyAsync(field_phiN(g), field_phi(g), bytes, cudaMemcpyDeviceToDevice, G.stream));
    int batch_first = 0;
    long long seq_first = 0;
    SE.npend = 0;
    for (int n = 0; n <= iter; ++n) {                                   // subs.f90:735
        const int raster = n % 8 + 1;                                   // subs.f90:740,855
        SE.begin();
        sweep_any(g, raster, cc);
        SE.end();
        bc_rms_any(g, dx, march_ntiles(g));                   // subs.f90:858-897 + boundary part of :902-914
        const long long seq = slab_publish_sum(g, npart);
        if (n == batch_first) seq_first = seq;
        if (n % 4 == 0 || n == iter) {                                  // after raster 1 / 5: the k direction flips next
            slab_decide(g, seq_first, n - batch_first + 1, batch_first, 0, tol);
            rc = read_ctrl(g, &hc);
            if (rc) return rc;
            SE.collect();
            if (watch_guard && hc.guard) { *guard_hit = true; break; }
            if (hc.done) {
                if (hc.status >= 0 && hc.n_exit < n && snap) {
                    // left in the middle of the batch: roll back and replay sweeps batch_first..n_exit
                    LSF_CUDA(cudaMemcpyAsync(field_phi(g), field_phiN(g), bytes, cudaMemcpyDeviceToDevice, G.stream));
                    LSF_CUDA(cudaMemsetAsync(g->ctrl, 0, sizeof(Ctrl), G.stream));
                    slab_exchange(g, false);
                    for (int m = batch_first; m <= hc.n_exit; ++m) {
                        sweep_any(g, m % 8 + 1, cc);
                        bc_rms_any(g, dx, march_ntiles(g));
                    }
                    LSF_CUDA(cudaMemcpyAsync(g->ctrl, &hc, sizeof(Ctrl), cudaMemcpyHostToDevice, G.stream));
                    LSF_CUDA(cudaStreamSynchronize(G.stream));
                }
                break;
            }
            if (snap && n < iter) LSF_CUDA(cudaMemcpyAsync(field_phiN(g), field_phi(g), bytes, cudaMemcpyDeviceToDevice, G.stream));
            batch_first = n + 1;
        }
    }
    slab_exchange(g, false);                                            // leave the ghost planes (faces included) current
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    LSF_CUDA(cudaGetLastError());
    if (hc.done && hc.status < 0)
        return set_error(hc.status, hc.status == LSF_ERR_TIMEOUT ? "reinit: a neighbouring rank stopped answering" : "reinit: device-side error");
    const int ne = hc.done ? hc.n_exit : iter;
    if (n_exit) *n_exit = ne;
    if (rms_hist && !*guard_hit) LSF_CUDA(cudaMemcpy(rms_hist, g->hist, sizeof(double) * (size_t)(ne + 1), cudaMemcpyDeviceToHost));
    return hc.done ? hc.status : LSF_OK;
}

static int reinit_core(Grid *g, int iter, double dx, double h, double tol, double *d_gradPhi, double *d_gradPhiMag,
                       int *n_exit, double *rms_hist)
{
    if (iter < 0 || !(dx > 0.)) return set_error(LSF_ERR_ARG, "reinit: bad iter/dx");
    if (g->dm.nx < 2 || g->dm.ny < 2 || g->dm.nz < 2) return set_error(LSF_ERR_ARG, "reinit: grid too small");
    int rc = slab_check_attached(g);
    if (rc) return rc;
    rc = ensure_hist(g, iter + 1);
    if (rc) return rc;
    const size_t bytes = sizeof(double) * (size_t)g->np;
    LSF_CUDA(cudaMemcpyAsync(g->phiS, g->phi, bytes, cudaMemcpyDeviceToDevice, G.stream));   // subs.f90:731
    Timer tm;
    tm.start();
    G.sweep_ms = 0.; G.n_sweeps = 0;
    G.arith_run = (G.arith == LSF_ARITH_EXACT) ? LSF_ARITH_EXACT : LSF_ARITH_FAST;
    bool guard_hit = false;
    const bool auto_arith = G.arith == LSF_ARITH_AUTO;
    if (sharded(g) && (d_gradPhi || d_gradPhiMag)) return set_error(LSF_ERR_ARG, "reinit: gradPhi outputs are not available on a sharded grid");
    int st = sharded(g) ? reinit_attempt_slab(g, iter, dx, h, tol, n_exit, rms_hist, auto_arith, &guard_hit)
                        : reinit_attempt(g, iter, dx, h, tol, d_gradPhi, d_gradPhiMag, n_exit, rms_hist, auto_arith, &guard_hit);
    if (st >= 0 && guard_hit) {
        // LSF_ARITH_AUTO: an ill-conditioned update was met -> the FAST result cannot be trusted to 1e-10;
        // start over from the frozen input (phiS still holds it) in the reference's exact arithmetic
        LSF_CUDA(cudaMemcpyAsync(g->phi, g->phiS, bytes, cudaMemcpyDeviceToDevice, G.stream));
        G.arith_run = LSF_ARITH_EXACT;
        G.sweep_ms = 0.; G.n_sweeps = 0;
        st = sharded(g) ? reinit_attempt_slab(g, iter, dx, h, tol, n_exit, rms_hist, false, &guard_hit)
                        : reinit_attempt(g, iter, dx, h, tol, d_gradPhi, d_gradPhiMag, n_exit, rms_hist, false, &guard_hit);
    }
    G.arith_last = G.arith_run;
    rc = tm.stop();
    if (rc) return rc;
    return st;
}

// reinit on a sharded fp32 grid: the loop of reinit_attempt_slab on the float fields (no conditioning guard in fp32)
static int f32_reinit_slab(Grid *g, int iter, double dx, double h, double tol, int *n_exit, double *rms_hist)
{
    if (iter < 0 || !(dx > 0.)) return set_error(LSF_ERR_ARG, "reinit: bad iter/dx");
    if (g->dm.nx < 2 || g->dm.ny < 2 || g->dm.nz < 2) return set_error(LSF_ERR_ARG, "reinit: grid too small");
    int rc = slab_check_attached(g);
    if (rc) return rc;
    rc = ensure_hist(g, iter + 1);
    if (rc) return rc;
    LSF_CUDA(cudaMemcpyAsync(g->phiS_f, g->phi_f, sizeof(float) * (size_t)g->np, cudaMemcpyDeviceToDevice, G.stream));   // subs.f90:731
    Timer tm;
    tm.start();
    G.sweep_ms = 0.; G.n_sweeps = 0;
    bool guard_hit = false;
    const int st = reinit_attempt_slab(g, iter, dx, h, tol, n_exit, rms_hist, false, &guard_hit);
    G.arith_last = LSF_ARITH_FAST;
    rc = tm.stop();
    if (rc) return rc;
    return st;
}

static int ensure_minmax_buffers(Grid *g, bool want_lap)
{
    if (want_lap && !g->lap) LSF_CUDA(cudaMalloc(&g->lap, sizeof(double) * (size_t)g->np));
    if (!g->mask) LSF_CUDA(cudaMalloc(&g->mask, (size_t)g->np));
    return LSF_OK;
}

// The min/max loop, set3d.f90:394-462.  phiN must already equal phi (set3d.f90:377).
// Plane schedule (cross-check): Jacobi Laplacian kernel + one launch per hyperplane, in place.
static int minmax_core_plane(Grid *g, int iter, double dx, double h1, double tol, bool mask_given, Ctrl *hc_out)
{
    int rc = ensure_minmax_buffers(g, true);
    if (rc) return rc;
    Ctrl hc = {0, 0, 1, 0, 0};
    const int check = g->np > 2000000 ? 1 : 8;
    for (int n = 1; n <= iter; ++n) {                                   // set3d.f90:394
        launch_minmax_iteration_plane(g, dx, h1, mask_given && n == 1); // :399-431 (+ narrowBand :460 of n-1)
        launch_rms(g, false);                                           // :435-447
        launch_finalize(g, RMS_BLOCKS, 1, tol);                         // :447-458
        launch_copy_if_running(g, g->phiN, g->phi);                     // :454
        if (n % check == 0 || n == iter) {
            rc = read_ctrl(g, &hc);
            if (rc) return rc;
            if (hc.done || hc.status < 0) break;
        }
    }
    *hc_out = hc;
    // on a tolerance EXIT the reference leaves phiN at the previous iterate; the guarded copy above did not
    // run for the exiting iteration (done was already set), so phiN is right in every case
    return LSF_OK;
}

// March schedule (production): one fused kernel per iteration, ping-pong between the phi and phiN buffers
// (iteration n reads the buffer holding phi_{n-1} and writes phi_n into the other), so phiN = phi
// (set3d.f90:454) is free and the RMS is the sum of the kernel's per-tile partials.
static int minmax_core_march(Grid *g, int iter, double dx, double h1, double tol, bool mask_given, Ctrl *hc_out)
{
    int rc = mm_march_prepare(g);
    if (rc) return rc;
    Ctrl hc = {0, 0, 1, 0, 0};
    if (iter >= 1) {
        launch_mm_check_boundary(g, mask_given ? g->mask : nullptr, dx, !mask_given || iter >= 2);
        if (sharded(g)) launch_finalize_slab(g, 0, 1, -1., 1);          // all ranks must agree on the verdict (no sum, no test)
        rc = read_ctrl(g, &hc);
        if (rc) return rc;
        if (hc.status < 0) { *hc_out = hc; return LSF_OK; }
        if (sharded(g)) {                                               // the agreement round advanced n: restore it
            Ctrl init = {0, 0, 1, 0, 0};
            LSF_CUDA(cudaMemcpyAsync(g->ctrl, &init, sizeof(Ctrl), cudaMemcpyHostToDevice, G.stream));
        }
    }
    const int ntiles = march_ntiles(g);
    double *buf[2] = {g->phi, g->phiN};                                 // buf[0] = phi_0, buf[1] = copy of it
    for (int n = 1; n <= iter; ++n) {                                   // set3d.f90:394
        const double *A = buf[(n - 1) & 1];
        double *B = buf[n & 1];
        launch_minmax_iteration_march(g, A, B, (mask_given && n == 1) ? g->mask : nullptr, dx, h1);   // :399-431
        if (sharded(g)) {
            slab_exchange(g, true, B);                                  // B's boundary planes -> the neighbours' ghost planes
            launch_finalize_slab(g, ntiles, 1, tol, n);                 // :435-458, sum over all ranks
        } else
        launch_finalize(g, ntiles, 1, tol);                             // :435-458
        if (n % 8 == 0 || n == iter) {
            rc = read_ctrl(g, &hc);
            if (rc) return rc;
            if (hc.done) break;
        }
    }
    const int ne = hc.done ? hc.n_exit : iter;
    // phi_ne lives in buf[ne & 1], phi_{ne-1} in the other buffer
    g->phi = buf[ne & 1];
    g->phiN = buf[(ne & 1) ^ 1];
    *hc_out = hc;
    return LSF_OK;
}

// Active-list schedule (production, lsf_mm_list.cuh): the set of cells that can ever change is compacted once,
// every iteration is an order-free pass over that list plus a (normally empty) settle step.
static int minmax_core_list(Grid *g, int iter, double dx, double h1, double tol, bool mask_given, Ctrl *hc_out)
{
    int rc = mml_prepare(g, mask_given ? g->mask : nullptr, dx);
    if (rc) return rc;
    Ctrl hc = {0, 0, 1, 0, 0};
    if (iter >= 1) {
        launch_mm_check_boundary(g, mask_given ? g->mask : nullptr, dx, !mask_given || iter >= 2);
        if (sharded(g)) launch_finalize_slab(g, 0, 1, -1., 1);          // all ranks agree on the verdict -- and have all
                                                                        // finished setting up phiN before anyone pushes into it
        rc = read_ctrl(g, &hc);
        if (rc) return rc;
        if (hc.status < 0) { *hc_out = hc; return LSF_OK; }
        if (sharded(g)) {
            Ctrl init = {0, 0, 1, 0, 0};
            LSF_CUDA(cudaMemcpyAsync(g->ctrl, &init, sizeof(Ctrl), cudaMemcpyHostToDevice, G.stream));
        }
    }
    const int npart = mml_npart();
    double *buf[2] = {g->phi, g->phiN};                                 // buf[0] = phi_0, buf[1] = copy of it
    for (int n = 1; n <= iter; ++n) {                                   // set3d.f90:394
        const double *A = buf[(n - 1) & 1];
        double *B = buf[n & 1];
        launch_minmax_iteration_list(g, A, B, (mask_given && n == 1) ? g->mask : nullptr, dx, h1);   // :399-431
        if (sharded(g)) {
            // B's boundary planes -> the neighbours' ghost planes.  No handshake: a neighbour starts iteration n
            // only after this rank's planes of iteration n-1 have arrived, i.e. it is past everything that read
            // the buffer being overwritten here.
            slab_exchange(g, true, B, false);
            launch_finalize_slab(g, npart, 1, tol, n);                  // :435-458, sum over all ranks
        } else
            launch_finalize(g, npart, 1, tol);                          // :435-458
        if (n % 8 == 0 || n == iter) {
            rc = read_ctrl(g, &hc);
            if (rc) return rc;
            if (hc.done) break;
        }
    }
    const int ne = hc.done ? hc.n_exit : iter;
    g->phi = buf[ne & 1];
    g->phiN = buf[(ne & 1) ^ 1];
    if (hc.done && hc.status == LSF_ERR_ARG) {                          // queue overflow: leave the scratch state clean
        cudaMemsetAsync(g->mml_unres, 0, (size_t)g->np, G.stream);
        cudaMemsetAsync(g->mml_work_count, 0, sizeof(int), G.stream);
    }
    *hc_out = hc;
    return LSF_OK;
}

static int minmax_core(Grid *g, int iter, double dx, double h1, double tol, bool mask_given,
                       int *n_exit, double *rms_hist, int *converged)
{
    if (iter < 0 || !(dx > 0.)) return set_error(LSF_ERR_ARG, "minmax: bad iter/dx");
    if (g->dm.nx < 2 || g->dm.ny < 2 || g->dm.nz < 2) return set_error(LSF_ERR_ARG, "minmax: grid too small");
    if (sharded(g) && (G.sched != LSF_SCHED_MARCH || mask_given))
        return set_error(LSF_ERR_ARG, "minmax: a sharded grid supports the march schedule through lsf_grid_minmax only");
    int rc = slab_check_attached(g);
    if (rc) return rc;
    rc = ensure_hist(g, iter + 1);
    if (rc) return rc;
    Ctrl init = {0, 0, 1, 0, 0};
    LSF_CUDA(cudaMemcpyAsync(g->ctrl, &init, sizeof(Ctrl), cudaMemcpyHostToDevice, G.stream));
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    Timer tm;
    tm.start();
    Ctrl hc = init;
    if (G.sched == LSF_SCHED_MARCH && G.mm_algo == LSF_MINMAX_LIST) rc = minmax_core_list(g, iter, dx, h1, tol, mask_given, &hc);
    else if (G.sched == LSF_SCHED_MARCH) rc = minmax_core_march(g, iter, dx, h1, tol, mask_given, &hc);
    else rc = minmax_core_plane(g, iter, dx, h1, tol, mask_given, &hc);
    if (rc) return rc;
    rc = tm.stop();
    if (rc) return rc;
    LSF_CUDA(cudaGetLastError());
    if (hc.status == LSF_ERR_BAND_ON_BOUNDARY)
        return set_error(LSF_ERR_BAND_ON_BOUNDARY, "minmax: narrow band touches the grid boundary");
    if (hc.status == LSF_ERR_ARG)
        return set_error(LSF_ERR_ARG, "minmax: more than 2^22 undecided cells in one iteration; use LSF_MINMAX_MARCH for this input");
    if (hc.status == LSF_ERR_TIMEOUT) return set_error(LSF_ERR_TIMEOUT, "minmax: a neighbouring rank stopped answering");
    const int ne = hc.done ? hc.n_exit : iter;
    const bool conv = hc.done && hc.status == 0;
    if (!conv && ne >= 1)      // the reference executed phiN = phi (set3d.f90:454) in the last iteration it ran
        LSF_CUDA(cudaMemcpyAsync(g->phiN, g->phi, sizeof(double) * (size_t)g->np, cudaMemcpyDeviceToDevice, G.stream));
    if (n_exit) *n_exit = ne;
    if (converged) *converged = conv ? 1 : 0;
    if (rms_hist && ne >= 1) LSF_CUDA(cudaMemcpy(rms_hist, g->hist, sizeof(double) * (size_t)ne, cudaMemcpyDeviceToHost));
    return hc.done ? hc.status : LSF_OK;
}

static int sign_core(Grid *g, const double xLo[3], double dx, const double *surfX, int nNode,
                     const int32_t *surfElem, int nElem, int im, int ip, int jm, int jp, int km, int kp)
{
    const Dims &dm = g->dm;
    if (nNode < 1 || nElem < 1) return set_error(LSF_ERR_ARG, "sign_init: empty surface");
    { int rc0 = slab_check_attached(g); if (rc0) return rc0; }
    if (im < 0 || jm < 0 || km < 0 || ip > dm.nx || jp > dm.ny || kp > g->sg.NZ || ip < im || jp < jm || kp < km)
        return set_error(LSF_ERR_ARG, "sign_init: sub-box outside the grid");
    for (long long q = 0; q < 3LL * nElem; ++q)
        if (surfElem[q] < 1 || surfElem[q] > nNode) return set_error(LSF_ERR_ARG, "sign_init: surfElem index out of range");
    double *d_X = nullptr, *d_cen = nullptr;
    int32_t *d_E = nullptr;
    LSF_CUDA(cudaMalloc(&d_X, sizeof(double) * 3 * (size_t)nNode));
    LSF_CUDA(cudaMalloc(&d_E, sizeof(int32_t) * 3 * (size_t)nElem));
    LSF_CUDA(cudaMalloc(&d_cen, sizeof(double) * 3 * (size_t)nElem));
    LSF_CUDA(cudaMemcpyAsync(d_X, surfX, sizeof(double) * 3 * (size_t)nNode, cudaMemcpyHostToDevice, G.stream));
    LSF_CUDA(cudaMemcpyAsync(d_E, surfElem, sizeof(int32_t) * 3 * (size_t)nElem, cudaMemcpyHostToDevice, G.stream));
    Timer tm;
    tm.start();
    int rcs = LSF_OK;
    if (g->f32) rcs = f32_sign_init(g, xLo, dx, d_X, nNode, d_E, nElem, d_cen, im, ip, jm, jp, km, kp);
    else launch_sign_init(g, xLo, dx, d_X, nNode, d_E, nElem, d_cen, im, ip, jm, jp, km, kp);
    slab_exchange(g, false);
    int rc = tm.stop();
    if (!rc) rc = rcs;
    cudaError_t e = cudaGetLastError();
    cudaFree(d_X); cudaFree(d_E); cudaFree(d_cen);
    if (rc) return rc;
    if (e != cudaSuccess) return set_error(LSF_ERR_CUDA, "sign_init: %s", cudaGetErrorString(e));
    return LSF_OK;
}

}  // namespace lsf

using namespace lsf;

// The host-buffer entry points (the Fortran drop-in calls) run on a library-owned device grid.  Allocating
// and freeing three fields per call costs ~0.1 s at 1024^3, so the grid of the most recent call is kept and
// reused when the next call has the same extents (set3d.f90 calls sign search, reinit, narrowBand and the
// min/max loop on the same grid).  It is released by lsf_finalize, by a call with other extents, or never kept
// at all with LSF_NO_GRID_CACHE=1.
static lsf_grid *g_cached = nullptr;

static int host_grid_acquire(lsf_grid **out, int nx, int ny, int nz, bool f32 = false)
{
    if (g_cached && g_cached->dm.nx == nx && g_cached->dm.ny == ny && g_cached->dm.nz == nz && !sharded(g_cached) &&
        (g_cached->f32 != 0) == f32) {
        *out = g_cached;
        g_cached = nullptr;
        return LSF_OK;
    }
    if (g_cached) { lsf_grid_destroy(g_cached); g_cached = nullptr; }
    return f32 ? lsf_grid_create_f32(out, nx, ny, nz) : lsf_grid_create(out, nx, ny, nz);
}

static void host_grid_release(lsf_grid *g)
{
    if (!g) return;
    static const bool no_cache = getenv("LSF_NO_GRID_CACHE") != nullptr;
    if (no_cache || g_cached) { lsf_grid_destroy(g); return; }
    g_cached = g;
}

// =============================================================================================
extern "C" {

int lsf_init(int device)
{
    if (G.inited) return LSF_OK;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_error(LSF_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0) {
        const char *lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % ndev : 0;
    }
    if (device >= ndev) return set_error(LSF_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    LSF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LSF_CUDA(cudaGetDeviceProperties(&prop, device));
    G.device = device;
    G.num_sms = prop.multiProcessorCount;
    LSF_CUDA(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
    LSF_CUDA(cudaEventCreate(&G.ev0));
    LSF_CUDA(cudaEventCreate(&G.ev1));
    // environment overrides for hosts that cannot easily call the setters (e.g. a Fortran driver)
    if (const char *a = getenv("LSF_ARITH"))
        G.arith = (strcmp(a, "exact") == 0) ? LSF_ARITH_EXACT : (strcmp(a, "fast") == 0) ? LSF_ARITH_FAST : LSF_ARITH_AUTO;
    if (const char *s = getenv("LSF_SCHED")) G.sched = (strcmp(s, "plane") == 0) ? LSF_SCHED_PLANE : LSF_SCHED_MARCH;
    if (const char *m = getenv("LSF_MINMAX")) G.mm_algo = (strcmp(m, "march") == 0) ? LSF_MINMAX_MARCH : LSF_MINMAX_LIST;
    if (const char *o = getenv("LSF_SWEEP_OVERLAP")) G.overlap = atoi(o) != 0;
    if (const char *q = getenv("LSF_PRECISION")) G.prec = (strcmp(q, "f32") == 0) ? LSF_PREC_F32 : LSF_PREC_F64;
    G.inited = true;
    return LSF_OK;
}

int lsf_finalize(void)
{
    if (!G.inited) return LSF_OK;
    if (g_cached) { lsf_grid_destroy(g_cached); g_cached = nullptr; }
    cudaStreamSynchronize(G.stream);
    cudaEventDestroy(G.ev0);
    cudaEventDestroy(G.ev1);
    cudaStreamDestroy(G.stream);
    G.inited = false;
    return LSF_OK;
}

const char *lsf_last_error(void) { return G.err; }

int lsf_set_arith(int arith)
{
    if (arith != LSF_ARITH_FAST && arith != LSF_ARITH_EXACT && arith != LSF_ARITH_AUTO)
        return set_error(LSF_ERR_ARG, "bad arith %d", arith);
    G.arith = arith;
    return LSF_OK;
}

int lsf_last_arith(void) { return G.arith_last; }

int lsf_set_sched(int sched)
{
    if (sched != LSF_SCHED_MARCH && sched != LSF_SCHED_PLANE) return set_error(LSF_ERR_ARG, "bad sched %d", sched);
    G.sched = sched;
    return LSF_OK;
}

int lsf_set_minmax_algo(int algo)
{
    if (algo != LSF_MINMAX_LIST && algo != LSF_MINMAX_MARCH) return set_error(LSF_ERR_ARG, "bad minmax algo %d", algo);
    G.mm_algo = algo;
    return LSF_OK;
}

long long lsf_last_minmax_active(void) { return G.mm_active; }

int lsf_set_overlap(int on)
{
    G.overlap = on != 0;
    return LSF_OK;
}

int lsf_set_precision(int prec)
{
    if (prec != LSF_PREC_F64 && prec != LSF_PREC_F32) return set_error(LSF_ERR_ARG, "bad precision %d", prec);
    G.prec = prec;
    return LSF_OK;
}

int lsf_set_profile(int on)
{
    G.profile = on != 0;
    return LSF_OK;
}

int lsf_last_sweep_timing(double *sweep_ms, int *n_sweeps)
{
    if (sweep_ms) *sweep_ms = G.sweep_ms;
    if (n_sweeps) *n_sweeps = G.n_sweeps;
    return LSF_OK;
}

int lsf_last_timing(double *kernel_ms, int *n_launches)
{
    if (kernel_ms) *kernel_ms = G.last_ms;
    if (n_launches) *n_launches = G.n_launch;
    return LSF_OK;
}

// ---------------------------------------------------------------------------------------------
int lsf_grid_create(lsf_grid **out, int nx, int ny, int nz)
{
    if (!out) return set_error(LSF_ERR_ARG, "null handle");
    *out = nullptr;
    int rc = ensure_init();
    if (rc) return rc;
    if (nx < 1 || ny < 1 || nz < 1) return set_error(LSF_ERR_ARG, "grid extents must be >= 1");
    Grid *g = (Grid *)calloc(1, sizeof(Grid));
    if (!g) return set_error(LSF_ERR_ARG, "out of host memory");
    g->dm.nx = nx; g->dm.ny = ny; g->dm.nz = nz;
    g->dm.sx = (long long)nx + 1;
    g->dm.sxy = g->dm.sx * ((long long)ny + 1);
    g->np = g->dm.sxy * ((long long)nz + 1);
    slab_geom(nz, 1, 0, g->sg);                                         // one slab: the whole grid
    const size_t bytes = sizeof(double) * (size_t)g->np + 64;          // + one chunk: vector row readers may fetch a whole aligned chunk at the end
    cudaError_t e;
    if ((e = cudaMalloc(&g->phi, bytes)) != cudaSuccess || (e = cudaMalloc(&g->phiS, bytes)) != cudaSuccess ||
        (e = cudaMalloc(&g->phiN, bytes)) != cudaSuccess ||
        (e = cudaMalloc(&g->partial, sizeof(double) * PARTIAL_CAP)) != cudaSuccess ||
        (e = cudaMalloc(&g->ctrl, sizeof(Ctrl))) != cudaSuccess) {
        lsf_grid_destroy(g);
        return set_error(LSF_ERR_CUDA, "grid_create: %s", cudaGetErrorString(e));
    }
    *out = g;
    return LSF_OK;
}

int lsf_grid_destroy(lsf_grid *g)
{
    if (!g) return LSF_OK;
    if (G.inited) cudaStreamSynchronize(G.stream);
    if (g->shared_base) {                                               // sharded: phi/phiN live in the shared allocation
        for (int r = 0; r < g->sg.nranks; ++r)
            if (r != g->sg.rank && g->peer_base[r]) cudaIpcCloseMemHandle(g->peer_base[r]);
        cudaFree(g->shared_base);
        cudaFree(g->exch_counter);
    } else { cudaFree(g->phi); cudaFree(g->phiN); }
    if (!g->shared_base) { cudaFree(g->phi_f); cudaFree(g->phiN_f); }
    cudaFree(g->phiS_f);
    cudaFree(g->phiS); cudaFree(g->lap); cudaFree(g->mask);
    cudaFree(g->partial); cudaFree(g->hist); cudaFree(g->ctrl);
    cudaFree(g->march_ticket); cudaFree(g->march_progress); cudaFree(g->march_colnext);
    cudaFree(g->ov_progress); cudaFree(g->ov_partial); cudaFree(g->ov_snap);
    cudaFree(g->mml_list); cudaFree(g->mml_unres); cudaFree(g->mml_work); cudaFree(g->mml_work_count);
    cudaFree(g->mml_counts); cudaFree(g->mml_offsets);
    free(g);
    return LSF_OK;
}

int lsf_grid_fill(lsf_grid *g, double value)
{
    if (!g) return set_error(LSF_ERR_ARG, "null grid");
    g->sb_from_phiN = false;
    if (g->f32) return f32_fill(g, value);
    launch_fill(g, g->phi, value);                                      // ghost planes included: consistent on all ranks
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    return LSF_OK;
}

int lsf_grid_upload(lsf_grid *g, const double *phi_host)
{
    if (!g || !phi_host) return set_error(LSF_ERR_ARG, "null argument");
    g->sb_from_phiN = false;
    int rc = slab_check_attached(g);
    if (rc) return rc;
    if (g->f32) {
        rc = f32_upload(g, phi_host, g->phi_f + owned_off(g), (long long)owned_elems(g));
        if (rc) return rc;
        slab_exchange(g, false);
        LSF_CUDA(cudaStreamSynchronize(G.stream));
        return LSF_OK;
    }
    // z-slab: the host array holds this rank's owned planes k0..k1-1 (a contiguous range of the global array)
    LSF_CUDA(cudaMemcpyAsync(g->phi + owned_off(g), phi_host, sizeof(double) * owned_elems(g), cudaMemcpyHostToDevice, G.stream));
    slab_exchange(g, false);
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    return LSF_OK;
}

int lsf_grid_download(lsf_grid *g, double *phi_host)
{
    if (!g || !phi_host) return set_error(LSF_ERR_ARG, "null argument");
    if (g->f32) return f32_download(g, g->phi_f + owned_off(g), phi_host, (long long)owned_elems(g));
    LSF_CUDA(cudaMemcpyAsync(phi_host, g->phi + owned_off(g), sizeof(double) * owned_elems(g), cudaMemcpyDeviceToHost, G.stream));
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    return LSF_OK;
}

int lsf_grid_download_phiN(lsf_grid *g, double *phiN_host)
{
    if (!g || !phiN_host) return set_error(LSF_ERR_ARG, "null argument");
    if (g->f32) return set_error(LSF_ERR_ARG, "download_phiN: an fp32 grid keeps no phiN");
    LSF_CUDA(cudaMemcpyAsync(phiN_host, g->phiN + owned_off(g), sizeof(double) * owned_elems(g), cudaMemcpyDeviceToHost, G.stream));
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    return LSF_OK;
}

int lsf_grid_checksum(lsf_grid *g, uint64_t digest[2])
{
    if (!g || !digest) return set_error(LSF_ERR_ARG, "null argument");
    unsigned long long *d = nullptr;
    LSF_CUDA(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), G.stream);
    const long long base = (long long)g->sg.k0 * g->dm.sxy;            // global linear index of this rank's first owned point
    if (g->f32) launch_checksum(g->phi_f + owned_off(g), 4, (long long)owned_elems(g), base, d);
    else launch_checksum(g->phi + owned_off(g), 8, (long long)owned_elems(g), base, d);
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, G.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
    cudaFree(d);
    if (e != cudaSuccess) return set_error(LSF_ERR_CUDA, "checksum: %s", cudaGetErrorString(e));
    digest[0] = h[0]; digest[1] = h[1];
    return LSF_OK;
}

// Page-locking of caller-owned host arrays (the Fortran driver's ALLOCATEd phi): explicit, because only the caller knows
// the lifetime of the allocation -- memory must be unregistered before it is freed.
int lsf_host_register(void *ptr, size_t nbytes)
{
    if (!ptr || !nbytes) return set_error(LSF_ERR_ARG, "null argument");
    int rc = ensure_init();
    if (rc) return rc;
    LSF_CUDA(cudaHostRegister(ptr, nbytes, cudaHostRegisterDefault));
    return LSF_OK;
}

int lsf_host_unregister(void *ptr)
{
    if (!ptr) return set_error(LSF_ERR_ARG, "null argument");
    LSF_CUDA(cudaHostUnregister(ptr));
    return LSF_OK;
}

void *lsf_grid_device_ptr(lsf_grid *g) { return g ? (g->f32 ? (void *)g->phi_f : (void *)g->phi) : nullptr; }

int lsf_grid_is_f32(lsf_grid *g) { return g && g->f32 ? 1 : 0; }

int lsf_grid_sign_init(lsf_grid *g, const double xLo[3], double dx, const double *surfX, int nSurfNode,
                       const int32_t *surfElem, int nSurfElem, int im, int ip, int jm, int jp, int km, int kp)
{
    if (!g || !xLo || !surfX || !surfElem) return set_error(LSF_ERR_ARG, "null argument");
    g->sb_from_phiN = false;
    return sign_core(g, xLo, dx, surfX, nSurfNode, surfElem, nSurfElem, im, ip, jm, jp, km, kp);
}

int lsf_grid_reinit(lsf_grid *g, int iter, double dx, double h, double tol, int *n_exit, double *rms_hist)
{
    if (!g) return set_error(LSF_ERR_ARG, "null grid");
    if (g->f32 && !sharded(g)) return f32_reinit(g, iter, dx, h, tol, n_exit, rms_hist);
    if (g->f32) return f32_reinit_slab(g, iter, dx, h, tol, n_exit, rms_hist);
    g->sb_from_phiN = false;
    return reinit_core(g, iter, dx, h, tol, nullptr, nullptr, n_exit, rms_hist);
}

// K2' throughput mode (lsf_rk.cu): `steps` TVD-RK3 steps of the Jacobi WENO5 reinitialisation equation.  NOT the reference's
// scheme (see lsf_rk.cu); offered because the north_star names it.  Per step: 3 stage kernels, the boundary block after each,
// RMS of (phi_new - phi_old) over all points / EXIT / NaN tests as in subs.f90:902-926.
int lsf_grid_reinit_rk3(lsf_grid *g, int steps, double dx, double dt, double tol, int *n_exit, double *rms_hist)
{
    if (!g) return set_error(LSF_ERR_ARG, "null grid");
    if (g->f32) return set_error(LSF_ERR_ARG, "reinit_rk3: fp64 grids only");
    if (steps < 1 || !(dx > 0.)) return set_error(LSF_ERR_ARG, "reinit_rk3: bad steps/dx");
    if (g->dm.nx < 2 || g->dm.ny < 2 || g->dm.nz < 2) return set_error(LSF_ERR_ARG, "reinit_rk3: grid too small");
    int rc = slab_check_attached(g);
    if (rc) return rc;
    rc = ensure_hist(g, steps);
    if (rc) return rc;
    g->sb_from_phiN = false;
    const bool mg = sharded(g);
    const size_t bytes = sizeof(double) * (size_t)g->np;
    // Stage buffers: phi1 = phiN; phi2 is extra.  On z-slabs every stage buffer has to be peer-visible (its ghost planes are
    // filled by the neighbours after each stage): phi2 is the phi of a transient sharded grid of the same geometry that the
    // ranks create and connect together (lsf_slab.cu: sgrid_shadow_f64).  A plain halo problem -- no dependence between the cells
    // of a stage -- so unlike the Gauss-Seidel path there is no pipeline: stage, boundary block, 3-plane exchange, next stage.
    double *phi2 = nullptr, *scratch = nullptr;
    lsf_grid *sh = nullptr;
    const long long nblk = rk_nblocks(g);
    if (mg) {
        rc = sgrid_shadow_f64(g, &sh);
        if (rc) return rc;
        phi2 = sh->phi;
    } else {
        LSF_CUDA(cudaMalloc(&phi2, bytes));
    }
    cudaError_t e = cudaMalloc(&scratch, sizeof(double) * (size_t)(nblk + BC_BLOCKS));
    if (e != cudaSuccess) {
        if (mg) sgrid_shadow_release(g, sh); else cudaFree(phi2);
        return set_error(LSF_ERR_CUDA, "reinit_rk3: %s", cudaGetErrorString(e));
    }
    double *phi1 = g->phiN;
    slab_exchange(g, false);                                                                     // z-slabs: ghost planes of phi current
    LSF_CUDA(cudaMemcpyAsync(g->phiS, g->phi, bytes, cudaMemcpyDeviceToDevice, G.stream));       // frozen sign source
    LSF_CUDA(cudaMemcpyAsync(phi1, g->phi, bytes, cudaMemcpyDeviceToDevice, G.stream));          // boundary points of the stage buffers
    LSF_CUDA(cudaMemcpyAsync(phi2, g->phi, bytes, cudaMemcpyDeviceToDevice, G.stream));
    LSF_CUDA(cudaMemsetAsync(g->ctrl, 0, sizeof(Ctrl), G.stream));
    CellConst cc;
    cc.dx = dx; cc.inv_dx = 1. / dx; cc.k12 = 1. / (12. * dx); cc.dx2 = dx * dx; cc.h = dt;
    G.arith_run = (G.arith == LSF_ARITH_EXACT) ? LSF_ARITH_EXACT : LSF_ARITH_FAST;
    Timer tm;
    tm.start();
    Ctrl hc = {0, 0, 0, 0, 0};
    for (int n = 0; n < steps; ++n) {
        launch_rk_stage(g, g->phi, g->phi, phi1, cc, 0., 1., scratch, nullptr, mg ? g : nullptr);
        if (mg) { launch_reinit_bc_rms_buf(g, phi1, dx, scratch + nblk); slab_exchange(g, true, phi1); }
        else launch_reinit_bc_buf(g, phi1, dx);
        launch_rk_stage(g, phi1, g->phi, phi2, cc, 0.75, 0.25, scratch, nullptr, mg ? g : nullptr);
        if (mg) { launch_reinit_bc_rms_buf(g, phi2, dx, scratch + nblk); slab_exchange(sh, true, phi2); }
        else launch_reinit_bc_buf(g, phi2, dx);
        launch_rk_stage(g, phi2, g->phi, g->phi, cc, 1. / 3., 2. / 3., scratch, g->partial, mg ? sh : nullptr);    // interior part of the RMS -> partial[0]
        launch_reinit_bc_rms(g, dx, 1);                                                          // boundary block + its part -> partial[1..]
        if (mg) { launch_finalize_slab(g, 1 + BC_BLOCKS, 0, tol, n); slab_exchange(g, true); }
        else launch_finalize(g, 1 + BC_BLOCKS, 0, tol);
        if ((n + 1) % 16 == 0 || n == steps - 1) {
            rc = read_ctrl(g, &hc);
            if (rc) break;
            if (hc.done) break;
        }
    }
    const int rc_t = tm.stop();                                        // before the (slow, synchronising) cudaFree calls
    if (mg) { slab_exchange(g, false); cudaStreamSynchronize(G.stream); const int rc_s = sgrid_shadow_release(g, sh); if (!rc) rc = rc_s; }
    else cudaFree(phi2);
    cudaFree(scratch);
    if (rc) return rc;
    if (rc_t) return rc_t;
    LSF_CUDA(cudaGetLastError());
    G.arith_last = G.arith_run;
    const int ne = hc.done ? hc.n_exit : steps - 1;
    if (n_exit) *n_exit = ne;
    if (rms_hist) LSF_CUDA(cudaMemcpy(rms_hist, g->hist, sizeof(double) * (size_t)(ne + 1), cudaMemcpyDeviceToHost));
    return hc.done ? hc.status : LSF_OK;
}

static int narrowband_to_host(Grid *g, const double *d_phi, double dx, int32_t *nb_host, int32_t *sb_host)
{
    int32_t *d_nb = nullptr, *d_sb = nullptr;
    const size_t bytes = sizeof(int32_t) * (size_t)g->np;
    LSF_CUDA(cudaMalloc(&d_nb, bytes));
    cudaError_t e = cudaMalloc(&d_sb, bytes);
    if (e != cudaSuccess) { cudaFree(d_nb); return set_error(LSF_ERR_CUDA, "narrowband: %s", cudaGetErrorString(e)); }
    if (g->f32) f32_narrowband(g, dx, d_nb, d_sb);
    else launch_narrowband(g, d_phi, dx, d_nb, d_sb);
    const size_t obytes = sizeof(int32_t) * owned_elems(g);
    if (nb_host) cudaMemcpyAsync(nb_host, d_nb + owned_off(g), obytes, cudaMemcpyDeviceToHost, G.stream);
    if (sb_host) cudaMemcpyAsync(sb_host, d_sb + owned_off(g), obytes, cudaMemcpyDeviceToHost, G.stream);
    e = cudaStreamSynchronize(G.stream);
    cudaFree(d_nb); cudaFree(d_sb);
    if (e != cudaSuccess) return set_error(LSF_ERR_CUDA, "narrowband: %s", cudaGetErrorString(e));
    return LSF_OK;
}

int lsf_grid_narrowband(lsf_grid *g, double dx, int32_t *phiNB_host, int32_t *phiSB_host)
{
    if (!g) return set_error(LSF_ERR_ARG, "null grid");
    return narrowband_to_host(g, g->phi, dx, phiNB_host, phiSB_host);
}

int lsf_grid_minmax(lsf_grid *g, int iter, double dx, double h1, double tol, int *n_exit, double *rms_hist)
{
    if (!g) return set_error(LSF_ERR_ARG, "null grid");
    { const int rc0 = slab_check_attached(g); if (rc0) return rc0; }
    if (g->f32) {                                                       // fp64 flow on a transient shadow, rounded to fp32
        lsf_grid *sh = nullptr;
        int rc = f32_shadow_open(g, &sh);
        if (rc) return rc;
        const int st = lsf_grid_minmax(sh, iter, dx, h1, tol, n_exit, rms_hist);
        rc = f32_shadow_close(g, sh, st >= 0);
        return st ? st : rc;
    }
    slab_exchange(g, false);                                            // z-slabs: phi's ghost planes current before they are copied
    LSF_CUDA(cudaMemcpyAsync(g->phiN, g->phi, sizeof(double) * (size_t)g->np, cudaMemcpyDeviceToDevice, G.stream)); // set3d.f90:377
    int conv = 0;
    const int st = minmax_core(g, iter, dx, h1, tol, false, n_exit, rms_hist, &conv);
    g->sb_from_phiN = st >= 0 && conv;
    if (st >= 0) slab_exchange(g, false);
    return st;
}

// ---------------------------------------------------------------------------------------------
// host-buffer drop-in entry points
// ---------------------------------------------------------------------------------------------
int lsf_sign_init(double *phi, int nx, int ny, int nz, const double xLo[3], double dx, const double *surfX,
                  int nSurfNode, const int32_t *surfElem, int nSurfElem, int im, int ip, int jm, int jp, int km, int kp)
{
    if (!phi || !xLo || !surfX || !surfElem) return set_error(LSF_ERR_ARG, "null argument");
    lsf_grid *g = nullptr;
    int rc = host_grid_acquire(&g, nx, ny, nz);
    if (rc) return rc;
    rc = lsf_grid_upload(g, phi);
    if (!rc) rc = sign_core(g, xLo, dx, surfX, nSurfNode, surfElem, nSurfElem, im, ip, jm, jp, km, kp);
    if (!rc) rc = lsf_grid_download(g, phi);
    host_grid_release(g);
    return rc;
}

int lsf_reinit(double *phi, double *gradPhi, double *gradPhiMag, int nx, int ny, int nz, int iter, double dx, double h,
               int *n_exit, double *rms_hist)
{
    if (!phi) return set_error(LSF_ERR_ARG, "null phi");
    lsf_grid *g = nullptr;
    if (G.prec == LSF_PREC_F32) {
        // fp32 mode (lsf_set_precision): host arrays stay REAL(8); gradPhi / gradPhiMag (dead downstream,
        // set3d.f90:372-375) are not written
        int rc = host_grid_acquire(&g, nx, ny, nz, true);
        if (rc) return rc;
        rc = lsf_grid_upload(g, phi);
        int st = LSF_OK;
        if (!rc) { st = f32_reinit(g, iter, dx, h, 1.E-5, n_exit, rms_hist); if (st < 0) rc = st; }
        if (!rc) rc = lsf_grid_download(g, phi);
        host_grid_release(g);
        return rc ? rc : st;
    }
    int rc = host_grid_acquire(&g, nx, ny, nz);
    if (rc) return rc;
    double *d_g = nullptr, *d_gm = nullptr;
    const size_t bytes = sizeof(double) * (size_t)g->np;
    cudaError_t e = cudaSuccess;
    if (gradPhi && (e = cudaMalloc(&d_g, 3 * bytes)) == cudaSuccess) e = cudaMemcpy(d_g, gradPhi, 3 * bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && gradPhiMag && (e = cudaMalloc(&d_gm, bytes)) == cudaSuccess) e = cudaMemcpy(d_gm, gradPhiMag, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = set_error(LSF_ERR_CUDA, "reinit: %s", cudaGetErrorString(e));
    if (!rc) rc = lsf_grid_upload(g, phi);
    int st = LSF_OK;
    if (!rc) {
        st = reinit_core(g, iter, dx, h, 1.E-5, d_g, d_gm, n_exit, rms_hist);   // tol: subs.f90:915
        if (st < 0) rc = st;
    }
    if (!rc) rc = lsf_grid_download(g, phi);
    if (!rc && d_g && cudaMemcpy(gradPhi, d_g, 3 * bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_error(LSF_ERR_CUDA, "reinit: D2H gradPhi");
    if (!rc && d_gm && cudaMemcpy(gradPhiMag, d_gm, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_error(LSF_ERR_CUDA, "reinit: D2H gradPhiMag");
    cudaFree(d_g); cudaFree(d_gm);
    host_grid_release(g);
    return rc ? rc : st;
}

int lsf_narrowband(int nx, int ny, int nz, double dx, const double *phi, int32_t *phiNB, int32_t *phiSB)
{
    if (!phi || !phiNB || !phiSB) return set_error(LSF_ERR_ARG, "null argument");
    lsf_grid *g = nullptr;
    int rc = host_grid_acquire(&g, nx, ny, nz);
    if (rc) return rc;
    rc = lsf_grid_upload(g, phi);
    if (!rc) rc = narrowband_to_host(g, g->phi, dx, phiNB, phiSB);
    host_grid_release(g);
    return rc;
}

int lsf_minmax(double *phi, double *phiN, int32_t *phiNB, int32_t *phiSB, int nx, int ny, int nz, int iter, double dx,
               double h1, double tol, int *n_exit, double *rms_hist)
{
    if (!phi || !phiN || !phiNB || !phiSB) return set_error(LSF_ERR_ARG, "null argument");
    lsf_grid *g = nullptr;
    int rc = host_grid_acquire(&g, nx, ny, nz);
    if (rc) return rc;
    int st = LSF_OK, conv = 0, ne = 0;
    const size_t bytes = sizeof(double) * (size_t)g->np;
    int32_t *d_nb = nullptr;
    do {
        if ((rc = lsf_grid_upload(g, phi))) break;
        if (cudaMemcpy(g->phiN, phiN, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { rc = set_error(LSF_ERR_CUDA, "minmax: H2D phiN"); break; }
        if ((rc = ensure_minmax_buffers(g, false))) break;
        if (cudaMalloc(&d_nb, sizeof(int32_t) * (size_t)g->np) != cudaSuccess ||
            cudaMemcpy(d_nb, phiNB, sizeof(int32_t) * (size_t)g->np, cudaMemcpyHostToDevice) != cudaSuccess) {
            rc = set_error(LSF_ERR_CUDA, "minmax: H2D phiNB"); break;
        }
        launch_mask_from_i32(g, d_nb);
        cudaStreamSynchronize(G.stream);
        cudaFree(d_nb); d_nb = nullptr;
        st = minmax_core(g, iter, dx, h1, tol, true, &ne, rms_hist, &conv);
        if (st < 0) { rc = st; break; }
        if (n_exit) *n_exit = ne;
        if ((rc = lsf_grid_download(g, phi))) break;
        if ((rc = lsf_grid_download_phiN(g, phiN))) break;
        // masks: the last narrowBand call (set3d.f90:460) saw phi of the last non-exiting iteration,
        // i.e. phiN when the loop EXITed on tolerance, phi otherwise; none if no iteration ran it.
        if (iter >= 1 && !(conv && ne == 1) && st != LSF_NAN)
            rc = narrowband_to_host(g, conv ? g->phiN : g->phi, dx, phiNB, phiSB);
    } while (0);
    cudaFree(d_nb);
    host_grid_release(g);
    return rc ? rc : st;
}

}  // extern "C"
