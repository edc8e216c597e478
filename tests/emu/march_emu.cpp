// march_emu.cpp -- runs levelsetfortran_b200/csrc/lsf_march.cuh on the CPU: every CUDA thread is
// an OS thread, every CTA a pthread barrier domain, CTAs run concurrently and synchronise through
// the same ticket/progress flags as on the GPU.  Test infrastructure only (tests/test_march_emu.py).
#define LSF_EMU 1
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <vector>

#include "../../levelsetfortran_b200/csrc/lsf_march.cuh"
#include "../../levelsetfortran_b200/csrc/lsf_mm_march.cuh"
#include "../../levelsetfortran_b200/csrc/lsf_slab.cuh"
#include "../../levelsetfortran_b200/csrc/lsf_mm_list.cuh"

namespace lsf { thread_local EmuCta *emu_cta = nullptr; int emu_stall_us = 0; }
// dynamic tile scheduler (march_pick) instead of static tickets: switched on by the test harness
static int emu_dynamic = 0;
extern "C" void emu_set_dynamic(int on) { emu_dynamic = on; }
#include <deque>
#include <mutex>
static std::deque<std::vector<int>> emu_colnext_store;
static std::mutex emu_colnext_mu;
static int *emu_colnext(int ntb)
{
    if (!emu_dynamic) return nullptr;
    std::lock_guard<std::mutex> lk(emu_colnext_mu);
    emu_colnext_store.emplace_back((size_t)ntb, 0);
    return emu_colnext_store.back().data();
}
extern "C" void emu_set_stall(int us) { lsf::emu_stall_us = us; }
using namespace lsf;

typedef MarchCfgDefault CFG;
typedef MarchSmem<CFG> Smem;
constexpr int M_THREADS = CFG::THREADS;
struct ThreadArg { const MarchParams *p; Smem *sm; EmuCta *cta; int tid; int arith; };

static void *thread_main(void *v)
{
    ThreadArg *a = (ThreadArg *)v;
    emu_cta = a->cta;
    if (a->arith == 1) march_cta_any<ExactArith, CFG>(*a->p, *a->sm, a->tid);
    else march_cta_any<FastArith, CFG>(*a->p, *a->sm, a->tid);
    return nullptr;
}

// One in-place sweep of raster 1..8 on a dense Fortran-layout grid.  Returns sum of the per-tile
// partials (sum over interior cells of (new-old)^2), or -1 on failure.
extern "C" double emu_march_sweep(double *phi, const double *phiS, int nx, int ny, int nz, int raster,
                                  double dx, double h, int arith, int ncta)
{
    MarchParams p;
    memset(&p, 0, sizeof(p));
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    march_orient<CFG>(p, nx, ny, nz, sx, sxy, raster);
    p.phi = phi; p.phiS = phiS;
    p.cc.dx = dx; p.cc.inv_dx = 1. / dx; p.cc.k12 = 1. / (12. * dx); p.cc.dx2 = dx * dx; p.cc.h = h;
    std::vector<double> partial(p.ntiles, 0.);
    std::vector<int> order(p.ntiles);
    std::vector<long long> progress(p.ntiles, 0);
    march_fill_order(p.ntb, p.ntc, order.data());
    unsigned ticket = 0;
    Ctrl ctrl = {0, 0, 0, 0, 0};
    p.partial = partial.data(); p.order = order.data(); p.progress = progress.data();
    p.ticket = &ticket; p.ctrl = &ctrl; p.epoch = 1;
    p.col_next = emu_colnext(p.ntb);
    if (ncta > p.ntiles) ncta = p.ntiles;
    std::vector<Smem> sm(ncta);
    std::vector<EmuCta> ctas(ncta);
    std::vector<ThreadArg> args((size_t)ncta * M_THREADS);
    std::vector<pthread_t> th((size_t)ncta * M_THREADS);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int c = 0; c < ncta; ++c) pthread_barrier_init(&ctas[c].bar, nullptr, M_THREADS);
    for (int c = 0; c < ncta; ++c)
        for (int t = 0; t < M_THREADS; ++t) {
            ThreadArg &a = args[(size_t)c * M_THREADS + t];
            a.p = &p; a.sm = &sm[c]; a.cta = &ctas[c]; a.tid = t; a.arith = arith;
            if (pthread_create(&th[(size_t)c * M_THREADS + t], &attr, thread_main, &a) != 0) return -1.;
        }
    for (size_t q = 0; q < th.size(); ++q) pthread_join(th[q], nullptr);
    for (int c = 0; c < ncta; ++c) pthread_barrier_destroy(&ctas[c].bar);
    double s = 0.;
    for (int q = 0; q < p.ntiles; ++q) s += partial[q];
    return s;
}

// ---------------------------------------------------------------------------------------------------
// The same sweep on a grid cut into `nranks` z-slabs (lsf_slab.cuh), all ranks running concurrently in this
// process: every rank has its own local array with ghost planes, its own tickets / progress flags / CTAs,
// and talks to its neighbours only through the streaming-halo protocol of lsf_march.cuh (peer stores into
// the downstream rank's ghost planes + in_progress flags).  phi/phiS are the GLOBAL arrays; the ghost
// planes are filled from the pre-sweep state (what k_slab_exchange leaves there) and the owned planes are
// gathered back.  Returns the sum of all ranks' partials, -1 on failure, -2 on a bad partition.
extern "C" double emu_march_sweep_slabs(double *phi, const double *phiS, int nx, int ny, int NZ, int nranks, int raster,
                                        double dx, double h, int arith, int ncta, int order_m)
{
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    std::vector<SlabGeom> geo(nranks);
    for (int r = 0; r < nranks; ++r) if (!slab_geom(NZ, nranks, r, geo[r])) return -2.;
    std::vector<std::vector<double>> lphi(nranks), lphiS(nranks), partial(nranks);
    std::vector<std::vector<int>> order(nranks);
    std::vector<std::vector<long long>> progress(nranks);
    std::vector<SlabSync> sync(nranks);
    std::vector<MarchParams> P(nranks);
    std::vector<unsigned> ticket(nranks, 0u);
    std::vector<Ctrl> ctrl(nranks, Ctrl{0, 0, 0, 0, 0});
    memset(sync.data(), 0, sizeof(SlabSync) * nranks);
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        const size_t n = (size_t)sxy * (g.nzl + 1);
        lphi[r].assign(phi + (size_t)g.kbase * sxy, phi + (size_t)g.kbase * sxy + n);
        lphiS[r].assign(phiS + (size_t)g.kbase * sxy, phiS + (size_t)g.kbase * sxy + n);
    }
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        MarchParams &p = P[r];
        memset(&p, 0, sizeof(p));
        march_orient<CFG>(p, nx, ny, g.nzl, sx, sxy, raster, g.kupd_lo, g.kupd_hi, g.kbase, NZ);
        p.phi = lphi[r].data(); p.phiS = lphiS[r].data();
        p.cc.dx = dx; p.cc.inv_dx = 1. / dx; p.cc.k12 = 1. / (12. * dx); p.cc.dx2 = dx * dx; p.cc.h = h;
        partial[r].assign(p.ntiles, 0.); order[r].resize(p.ntiles); progress[r].assign(p.ntiles, 0);
        march_fill_order(p.ntb, p.ntc, order[r].data(), order_m);
        p.partial = partial[r].data(); p.order = order[r].data(); p.progress = progress[r].data();
        p.ticket = &ticket[r]; p.ctrl = &ctrl[r]; p.epoch = 1;
        p.col_next = emu_colnext(p.ntb);
        const int up = p.fc ? r + 1 : r - 1, down = p.fc ? r - 1 : r + 1;
        if (up >= 0 && up < nranks) p.in_progress = sync[r].in_progress;
        if (down >= 0 && down < nranks) {
            p.push_delta = (lphi[down].data() + (long long)(g.kbase - geo[down].kbase) * sxy) - lphi[r].data();
            p.push_progress = sync[down].in_progress;
        }
    }
    std::vector<int> nc(nranks);
    size_t nthreads = 0;
    for (int r = 0; r < nranks; ++r) { nc[r] = ncta < P[r].ntiles ? ncta : P[r].ntiles; nthreads += (size_t)nc[r] * M_THREADS; }
    std::vector<std::vector<Smem>> sm(nranks);
    std::vector<std::vector<EmuCta>> ctas(nranks);
    std::vector<ThreadArg> args(nthreads);
    std::vector<pthread_t> th(nthreads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int r = 0; r < nranks; ++r) {
        sm[r] = std::vector<Smem>(nc[r]);
        ctas[r] = std::vector<EmuCta>(nc[r]);
        for (int c = 0; c < nc[r]; ++c) pthread_barrier_init(&ctas[r][c].bar, nullptr, M_THREADS);
    }
    size_t q = 0;
    for (int r = 0; r < nranks; ++r)
        for (int c = 0; c < nc[r]; ++c)
            for (int t = 0; t < M_THREADS; ++t, ++q) {
                ThreadArg &a = args[q];
                a.p = &P[r]; a.sm = &sm[r][c]; a.cta = &ctas[r][c]; a.tid = t; a.arith = arith;
                if (pthread_create(&th[q], &attr, thread_main, &a) != 0) return -1.;
            }
    for (size_t i = 0; i < th.size(); ++i) pthread_join(th[i], nullptr);
    double s = 0.;
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        for (int c = 0; c < nc[r]; ++c) pthread_barrier_destroy(&ctas[r][c].bar);
        if (ctrl[r].status != 0) return -3.;
        memcpy(phi + (size_t)g.k0 * sxy, lphi[r].data() + (size_t)g.own_lo * sxy, sizeof(double) * (size_t)sxy * (g.k1 - g.k0));
        for (int i = 0; i < P[r].ntiles; ++i) s += partial[r][i];
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------
// Several consecutive sweeps on z-slabs with NO synchronisation between the ranks other than the flags of
// lsf_march.cuh (streaming halo, write-through to the upstream rank, per-tile completion flags): every rank
// runs its sweeps back to back from its own controller thread, so ranks drift apart by whole sweeps exactly
// as the GPUs do.  `skew_us` delays the ranks differently before every sweep to provoke that drift.
struct RankCtl {
    int r, nranks, nx, ny, NZ, nsweeps, first_raster, arith, ncta, order_m, skew_us;
    double dx, h;
    std::vector<SlabGeom> *geo;
    std::vector<std::vector<double>> *lphi, *lphiS;
    std::vector<SlabSync> *sync;
    double sum;
    int status;
};

static void *rank_main(void *v)
{
    RankCtl &R = *(RankCtl *)v;
    const int r = R.r, nranks = R.nranks;
    const long long sx = R.nx + 1, sxy = sx * (R.ny + 1);
    const SlabGeom &g = (*R.geo)[r];
    std::vector<SlabSync> &sync = *R.sync;
    R.sum = 0.; R.status = 0;
    int prev_fb = 0;
    std::vector<long long> progress;
    for (int n = 0; n < R.nsweeps; ++n) {
        const int raster = (R.first_raster - 1 + n) % 8 + 1;
        if (R.skew_us) usleep((useconds_t)R.skew_us * (unsigned)((n & 1) ? r : nranks - 1 - r));
        MarchParams p;
        memset(&p, 0, sizeof(p));
        march_orient<CFG>(p, R.nx, R.ny, g.nzl, sx, sxy, raster, g.kupd_lo, g.kupd_hi, g.kbase, R.NZ);
        p.phi = (*R.lphi)[r].data(); p.phiS = (*R.lphiS)[r].data();
        p.cc.dx = R.dx; p.cc.inv_dx = 1. / R.dx; p.cc.k12 = 1. / (12. * R.dx); p.cc.dx2 = R.dx * R.dx; p.cc.h = R.h;
        std::vector<double> partial(p.ntiles, 0.);
        std::vector<int> order(p.ntiles);
        if (progress.empty()) progress.assign(p.ntiles, 0);
        march_fill_order(p.ntb, p.ntc, order.data(), R.order_m);
        unsigned ticket = 0;
        Ctrl ctrl = {0, 0, 0, 0, 0};
        p.partial = partial.data(); p.order = order.data(); p.progress = progress.data();
        p.ticket = &ticket; p.ctrl = &ctrl; p.epoch = n + 1;
        p.col_next = emu_colnext(p.ntb);
        const int up = p.fc ? r + 1 : r - 1, down = p.fc ? r - 1 : r + 1;
        const int side_up = p.fc ? 1 : 0, side_down = p.fc ? 0 : 1;      // which of MY sides that neighbour is on
        if (up >= 0 && up < nranks) {
            p.in_progress = sync[r].in_progress;
            p.push_up_delta = ((*R.lphi)[up].data() + (long long)(g.kbase - (*R.geo)[up].kbase) * sxy) - p.phi;
            p.edge_pub[0] = sync[up].edge_done[1 - side_up];
        }
        if (down >= 0 && down < nranks) {
            p.push_delta = ((*R.lphi)[down].data() + (long long)(g.kbase - (*R.geo)[down].kbase) * sxy) - p.phi;
            p.push_progress = sync[down].in_progress;
            p.edge_pub[1] = sync[down].edge_done[1 - side_down];
            if (n > 0 && !getenv("EMU_NO_EDGE")) { p.edge_wait = sync[r].edge_done[side_down]; p.edge_need = n; p.edge_prev_fb = prev_fb; }
        }
        prev_fb = p.fb;
        const int nc = R.ncta < p.ntiles ? R.ncta : p.ntiles;
        std::vector<Smem> sm(nc);
        std::vector<EmuCta> ctas(nc);
        std::vector<ThreadArg> args((size_t)nc * M_THREADS);
        std::vector<pthread_t> th((size_t)nc * M_THREADS);
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        pthread_attr_setstacksize(&attr, 256 * 1024);
        for (int c = 0; c < nc; ++c) pthread_barrier_init(&ctas[c].bar, nullptr, M_THREADS);
        for (int c = 0; c < nc; ++c)
            for (int t = 0; t < M_THREADS; ++t) {
                ThreadArg &a = args[(size_t)c * M_THREADS + t];
                a.p = &p; a.sm = &sm[c]; a.cta = &ctas[c]; a.tid = t; a.arith = R.arith;
                if (pthread_create(&th[(size_t)c * M_THREADS + t], &attr, thread_main, &a) != 0) { R.status = -1; return nullptr; }
            }
        for (size_t i = 0; i < th.size(); ++i) pthread_join(th[i], nullptr);
        for (int c = 0; c < nc; ++c) pthread_barrier_destroy(&ctas[c].bar);
        if (ctrl.status != 0) { R.status = -3; return nullptr; }
        for (int i = 0; i < p.ntiles; ++i) R.sum += partial[i];
    }
    return nullptr;
}

extern "C" double emu_march_multi_slabs(double *phi, const double *phiS, int nx, int ny, int NZ, int nranks, int nsweeps,
                                        int first_raster, double dx, double h, int arith, int ncta, int order_m, int skew_us)
{
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    std::vector<SlabGeom> geo(nranks);
    for (int r = 0; r < nranks; ++r) if (!slab_geom(NZ, nranks, r, geo[r])) return -2.;
    std::vector<std::vector<double>> lphi(nranks), lphiS(nranks);
    std::vector<SlabSync> sync(nranks);
    memset(sync.data(), 0, sizeof(SlabSync) * nranks);
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        const size_t n = (size_t)sxy * (g.nzl + 1);
        lphi[r].assign(phi + (size_t)g.kbase * sxy, phi + (size_t)g.kbase * sxy + n);
        lphiS[r].assign(phiS + (size_t)g.kbase * sxy, phiS + (size_t)g.kbase * sxy + n);
    }
    std::vector<RankCtl> ctl(nranks);
    std::vector<pthread_t> th(nranks);
    for (int r = 0; r < nranks; ++r) {
        RankCtl &R = ctl[r];
        R.r = r; R.nranks = nranks; R.nx = nx; R.ny = ny; R.NZ = NZ; R.nsweeps = nsweeps; R.first_raster = first_raster;
        R.arith = arith; R.ncta = ncta; R.order_m = order_m; R.skew_us = skew_us; R.dx = dx; R.h = h;
        R.geo = &geo; R.lphi = &lphi; R.lphiS = &lphiS; R.sync = &sync;
        if (pthread_create(&th[r], nullptr, rank_main, &R) != 0) return -1.;
    }
    for (int r = 0; r < nranks; ++r) pthread_join(th[r], nullptr);
    double s = 0.;
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        if (ctl[r].status != 0) return (double)ctl[r].status;
        memcpy(phi + (size_t)g.k0 * sxy, lphi[r].data() + (size_t)g.own_lo * sxy, sizeof(double) * (size_t)sxy * (g.k1 - g.k0));
        s += ctl[r].sum;
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------
// One fused min/max-flow iteration (lsf_mm_march.cuh): reads A (phi_old), writes B (phi_new; must hold a
// copy of A's boundary values), mask may be NULL (band = abs(old) < 4.1*dx).  Returns the sum of the
// per-tile partials = sum over band cells of (new-old)^2, or -1 on failure.
typedef MmCfgDefault MCFG;
typedef MmSmem<MCFG> MSmem;
struct MmThreadArg { const MmParams *p; MSmem *sm; EmuCta *cta; int tid; };

static void *mm_thread_main(void *v)
{
    MmThreadArg *a = (MmThreadArg *)v;
    emu_cta = a->cta;
    mm_cta<MCFG>(*a->p, *a->sm, a->tid);
    return nullptr;
}

extern "C" double emu_mm_iteration(const double *A, double *B, const uint8_t *mask, int nx, int ny, int nz,
                                   double dx, double h1, int ncta)
{
    MmParams p;
    memset(&p, 0, sizeof(p));
    mm_orient<MCFG>(p, nx, ny, nz);
    p.A = A; p.B = B; p.mask = mask;
    p.bNB = 4.1 * dx; p.dxx = 1. / (dx * dx); p.h1 = h1;
    std::vector<double> partial(p.ntiles, 0.);
    std::vector<int> order(p.ntiles);
    std::vector<long long> progress(p.ntiles, 0);
    march_fill_order(p.ntb, p.ntc, order.data());
    unsigned ticket = 0;
    Ctrl ctrl = {0, 0, 0, 0, 0};
    p.partial = partial.data(); p.order = order.data(); p.progress = progress.data();
    p.ticket = &ticket; p.ctrl = &ctrl; p.epoch = 1;
    if (ncta > p.ntiles) ncta = p.ntiles;
    const int NT = MCFG::THREADS;
    std::vector<MSmem> sm(ncta);
    std::vector<EmuCta> ctas(ncta);
    std::vector<MmThreadArg> args((size_t)ncta * NT);
    std::vector<pthread_t> th((size_t)ncta * NT);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int c = 0; c < ncta; ++c) pthread_barrier_init(&ctas[c].bar, nullptr, NT);
    for (int c = 0; c < ncta; ++c)
        for (int t = 0; t < NT; ++t) {
            MmThreadArg &a = args[(size_t)c * NT + t];
            a.p = &p; a.sm = &sm[c]; a.cta = &ctas[c]; a.tid = t;
            if (pthread_create(&th[(size_t)c * NT + t], &attr, mm_thread_main, &a) != 0) return -1.;
        }
    for (size_t q = 0; q < th.size(); ++q) pthread_join(th[q], nullptr);
    for (int c = 0; c < ncta; ++c) pthread_barrier_destroy(&ctas[c].bar);
    double s = 0.;
    for (int q = 0; q < p.ntiles; ++q) s += partial[q];
    return s;
}

// ---------------------------------------------------------------------------------------------------
// The fused min/max iteration on z-slabs, all ranks concurrently (streaming halo of NEW values into the
// downstream rank's ghost plane of B; ghost planes of A = OLD values of the neighbours, as the bulk exchange
// between iterations leaves them).  A, B: GLOBAL arrays (B a copy of A on entry).  Returns the summed partials.
extern "C" double emu_mm_iteration_slabs(const double *A, double *B, int nx, int ny, int NZ, int nranks, double dx, double h1,
                                         int ncta, int order_m)
{
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    std::vector<SlabGeom> geo(nranks);
    for (int r = 0; r < nranks; ++r) if (!slab_geom(NZ, nranks, r, geo[r])) return -2.;
    std::vector<std::vector<double>> lA(nranks), lB(nranks), partial(nranks);
    std::vector<std::vector<int>> order(nranks);
    std::vector<std::vector<long long>> progress(nranks);
    std::vector<SlabSync> sync(nranks);
    std::vector<MmParams> P(nranks);
    std::vector<unsigned> ticket(nranks, 0u);
    std::vector<Ctrl> ctrl(nranks, Ctrl{0, 0, 0, 0, 0});
    memset(sync.data(), 0, sizeof(SlabSync) * nranks);
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        const size_t n = (size_t)sxy * (g.nzl + 1);
        lA[r].assign(A + (size_t)g.kbase * sxy, A + (size_t)g.kbase * sxy + n);
        lB[r].assign(B + (size_t)g.kbase * sxy, B + (size_t)g.kbase * sxy + n);
    }
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        MmParams &p = P[r];
        memset(&p, 0, sizeof(p));
        mm_orient<MCFG>(p, nx, ny, g.nzl, g.kupd_lo, g.kupd_hi);
        p.A = lA[r].data(); p.B = lB[r].data(); p.mask = nullptr;
        p.bNB = 4.1 * dx; p.dxx = 1. / (dx * dx); p.h1 = h1;
        partial[r].assign(p.ntiles, 0.); order[r].resize(p.ntiles); progress[r].assign(p.ntiles, 0);
        march_fill_order(p.ntb, p.ntc, order[r].data(), order_m);
        p.partial = partial[r].data(); p.order = order[r].data(); p.progress = progress[r].data();
        p.ticket = &ticket[r]; p.ctrl = &ctrl[r]; p.epoch = 1;
        if (r > 0) p.in_progress = sync[r].in_progress;
        if (r < nranks - 1) {
            p.push_delta = (lB[r + 1].data() + (long long)(g.kbase - geo[r + 1].kbase) * sxy) - lB[r].data();
            p.push_progress = sync[r + 1].in_progress;
        }
    }
    const int NT = MCFG::THREADS;
    std::vector<int> nc(nranks);
    size_t nthreads = 0;
    for (int r = 0; r < nranks; ++r) { nc[r] = ncta < P[r].ntiles ? ncta : P[r].ntiles; nthreads += (size_t)nc[r] * NT; }
    std::vector<std::vector<MSmem>> sm(nranks);
    std::vector<std::vector<EmuCta>> ctas(nranks);
    std::vector<MmThreadArg> args(nthreads);
    std::vector<pthread_t> th(nthreads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int r = 0; r < nranks; ++r) {
        sm[r] = std::vector<MSmem>(nc[r]);
        ctas[r] = std::vector<EmuCta>(nc[r]);
        for (int c = 0; c < nc[r]; ++c) pthread_barrier_init(&ctas[r][c].bar, nullptr, NT);
    }
    size_t q = 0;
    for (int r = 0; r < nranks; ++r)
        for (int c = 0; c < nc[r]; ++c)
            for (int t = 0; t < NT; ++t, ++q) {
                MmThreadArg &a = args[q];
                a.p = &P[r]; a.sm = &sm[r][c]; a.cta = &ctas[r][c]; a.tid = t;
                if (pthread_create(&th[q], &attr, mm_thread_main, &a) != 0) return -1.;
            }
    for (size_t i = 0; i < th.size(); ++i) pthread_join(th[i], nullptr);
    double s = 0.;
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        for (int c = 0; c < nc[r]; ++c) pthread_barrier_destroy(&ctas[r][c].bar);
        if (ctrl[r].status != 0) return -3.;
        memcpy(B + (size_t)g.k0 * sxy, lB[r].data() + (size_t)g.own_lo * sxy, sizeof(double) * (size_t)sxy * (g.k1 - g.k0));
        for (int i = 0; i < P[r].ntiles; ++i) s += partial[r][i];
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------
// The active-list min/max iteration (lsf_mm_list.cuh) run serially: speculate every band cell from OLD
// values in an arbitrary (here: descending!) order, then settle the undecided ones in dependence order.
// B must be a copy of A on entry.  stats[0] = band cells, [1] = cells that needed the exact 8-combination
// check, [2] = undecided cells.  Returns the sum over band cells of (new-old)^2.
extern "C" double emu_mm_list_iteration(const double *A, double *B, const unsigned char *mask, int nx, int ny, int nz,
                                        double dx, double h1, long long *stats)
{
    MmListConst c;
    c.sx = nx + 1; c.sxy = c.sx * (ny + 1); c.nx = nx; c.ny = ny; c.k_lo = 1; c.k_hi = nz - 1;
    c.bNB = 4.1 * dx; c.dxx = 1. / (dx * dx); c.h1 = h1;
    std::vector<long long> work;
    double s = 0.;
    long long nband = 0, ncomb = 0;
    for (int k = nz - 1; k >= 1; --k)
        for (int j = ny - 1; j >= 1; --j)
            for (int i = nx - 1; i >= 1; --i) {
                const long long q = i + c.sx * j + c.sxy * k;
                if (!mm_inband(c, A, mask, q)) { B[q] = A[q]; continue; }
                ++nband;
                double pn;
                bool comb = false;
                if (mm_cell_speculate(c, A, mask, q, i, j, k, pn, &comb)) { B[q] = pn; const double d = pn - A[q]; s += d * d; }
                else work.push_back(q);
                ncomb += comb ? 1 : 0;
            }
    std::vector<long long> w(work.rbegin(), work.rend());        // ascending q = a topological order
    for (long long q : w) { B[q] = mm_cell_settle(c, A, B, q); const double d = B[q] - A[q]; s += d * d; }
    if (stats) { stats[0] = nband; stats[1] = ncomb; stats[2] = (long long)w.size(); }
    return s;
}

// ---------------------------------------------------------------------------------------------------
// fp32 mode (F32Arith, float slot ring, MarchCfgF32): one sweep through the production schedule, and the
// literal lexicographic in-place loop (subs.f90:742-852) over the same cell arithmetic for comparison --
// the schedule must be an exact re-ordering in fp32 too (bitwise equal on the CPU).
typedef MarchCfg<16, 16, LSF_ROWS, float> CFGF;
typedef MarchSmem<CFGF> SmemF;
struct ThreadArgF { const MarchParamsT<float> *p; SmemF *sm; EmuCta *cta; int tid; int mg; };

template <bool MG>
static void run_cta_f32(const MarchParamsT<float> &p, SmemF &sm, int tid)
{
    const int o = (p.fa ? 1 : 0) | (p.fb ? 2 : 0) | (p.fc ? 4 : 0);
    switch (o) {
    case 0: march_cta<F32Arith, false, false, false, CFGF, MG>(p, sm, tid); break;
    case 1: march_cta<F32Arith, true, false, false, CFGF, MG>(p, sm, tid); break;
    case 2: march_cta<F32Arith, false, true, false, CFGF, MG>(p, sm, tid); break;
    case 3: march_cta<F32Arith, true, true, false, CFGF, MG>(p, sm, tid); break;
    case 4: march_cta<F32Arith, false, false, true, CFGF, MG>(p, sm, tid); break;
    case 5: march_cta<F32Arith, true, false, true, CFGF, MG>(p, sm, tid); break;
    case 6: march_cta<F32Arith, false, true, true, CFGF, MG>(p, sm, tid); break;
    default: march_cta<F32Arith, true, true, true, CFGF, MG>(p, sm, tid); break;
    }
}

static void *thread_main_f32(void *v)
{
    ThreadArgF *a = (ThreadArgF *)v;
    emu_cta = a->cta;
    if (a->mg) run_cta_f32<true>(*a->p, *a->sm, a->tid); else run_cta_f32<false>(*a->p, *a->sm, a->tid);
    return nullptr;
}

static void cc_f32(CellConstT<float> &cc, double dx, double h)
{
    cc.dx = (float)dx; cc.inv_dx = (float)(1. / dx); cc.k12 = (float)(1. / (12. * dx)); cc.dx2 = (float)(dx * dx); cc.h = (float)h;
}

extern "C" double emu_march_sweep_f32(float *phi, const float *phiS, int nx, int ny, int nz, int raster,
                                      double dx, double h, int ncta)
{
    MarchParamsT<float> p;
    memset(&p, 0, sizeof(p));
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    march_orient<CFGF>(p, nx, ny, nz, sx, sxy, raster);
    // vector loads fetch whole aligned chunks: work on 16-byte aligned copies with a chunk of padding at the end
    const size_t npts = (size_t)sxy * (nz + 1);
    float *wphi = nullptr, *wphiS = nullptr;
    if (posix_memalign((void **)&wphi, 64, sizeof(float) * (npts + 16)) || posix_memalign((void **)&wphiS, 64, sizeof(float) * (npts + 16))) return -1.;
    memcpy(wphi, phi, sizeof(float) * npts); memcpy(wphiS, phiS, sizeof(float) * npts);
    for (int q = 0; q < 16; ++q) { wphi[npts + q] = 1.e30f; wphiS[npts + q] = 1.e30f; }
    p.phi = wphi; p.phiS = wphiS;
    cc_f32(p.cc, dx, h);
    std::vector<double> partial(p.ntiles, 0.);
    std::vector<int> order(p.ntiles);
    std::vector<long long> progress(p.ntiles, 0);
    march_fill_order(p.ntb, p.ntc, order.data());
    unsigned ticket = 0;
    Ctrl ctrl = {0, 0, 0, 0, 0};
    p.partial = partial.data(); p.order = order.data(); p.progress = progress.data();
    p.ticket = &ticket; p.ctrl = &ctrl; p.epoch = 1;
    p.col_next = emu_colnext(p.ntb);
    if (ncta > p.ntiles) ncta = p.ntiles;
    constexpr int NT = CFGF::THREADS;
    std::vector<SmemF> sm(ncta);
    std::vector<EmuCta> ctas(ncta);
    std::vector<ThreadArgF> args((size_t)ncta * NT);
    std::vector<pthread_t> th((size_t)ncta * NT);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int c = 0; c < ncta; ++c) pthread_barrier_init(&ctas[c].bar, nullptr, NT);
    for (int c = 0; c < ncta; ++c)
        for (int t = 0; t < NT; ++t) {
            ThreadArgF &a = args[(size_t)c * NT + t];
            a.p = &p; a.sm = &sm[c]; a.cta = &ctas[c]; a.tid = t; a.mg = 0;
            if (pthread_create(&th[(size_t)c * NT + t], &attr, thread_main_f32, &a) != 0) return -1.;
        }
    for (size_t q = 0; q < th.size(); ++q) pthread_join(th[q], nullptr);
    for (int c = 0; c < ncta; ++c) pthread_barrier_destroy(&ctas[c].bar);
    memcpy(phi, wphi, sizeof(float) * npts);
    free(wphi); free(wphiS);
    double s = 0.;
    for (int q = 0; q < p.ntiles; ++q) s += partial[q];
    return s;
}

// literal in-place raster sweep with the F32Arith cell (plus the closed-form boundary block when bc != 0)
extern "C" void emu_lex_sweep_f32(float *phi, const float *phiS, int nx, int ny, int nz, int raster, double dx, double h, int bc)
{
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    CellConstT<float> cc;
    cc_f32(cc, dx, h);
    int d[3];
    raster_dirs(raster, d);
    for (int a = 1; a <= nx - 1; ++a)
        for (int b = 1; b <= ny - 1; ++b)
            for (int c = 1; c <= nz - 1; ++c) {
                const int i = d[0] > 0 ? a : nx - a, j = d[1] > 0 ? b : ny - b, k = d[2] > 0 ? c : nz - c;
                const long long q = i + sx * j + sxy * k;
                const bool hi = (i > 3) && (i < nx - 4) && (j > 3) && (j < ny - 4) && (k > 3) && (k < nz - 4);
                float vx[7], vy[7], vz[7];
                for (int m = -3; m <= 3; ++m) {
                    const bool ok = hi || (m >= -1 && m <= 1);
                    vx[m + 3] = ok ? phi[q + m] : 0.f;
                    vy[m + 3] = ok ? phi[q + m * sx] : 0.f;
                    vz[m + 3] = ok ? phi[q + m * sxy] : 0.f;
                }
                float g[3], gM;
                bool sens;
                phi[q] = reinit_cell<F32Arith>(vx, vy, vz, phiS[q], hi, cc, g, gM, sens);
            }
    if (!bc) return;
    const float dxf = (float)dx;
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i) {
                const int B = (i == 0 || i == nx) + (j == 0 || j == ny) + (k == 0 || k == nz);
                if (!B) continue;
                const int H = (i == nx) + (j == ny) + (k == nz);
                const int m = 1 + H < B ? 1 + H : B;
                const int ci = i < 1 ? 1 : (i > nx - 1 ? nx - 1 : i), cj = j < 1 ? 1 : (j > ny - 1 ? ny - 1 : j),
                          ck = k < 1 ? 1 : (k > nz - 1 ? nz - 1 : k);
                float v = phi[ci + sx * cj + sxy * ck];
                for (int r = 0; r < m; ++r) v = v + dxf;
                phi[i + sx * j + sxy * k] = v;
            }
}

// ---------------------------------------------------------------------------------------------------
// Surface-node projection (lsf_nodes.cuh): the per-node code of the product kernel, run serially on the CPU.
#include "../../levelsetfortran_b200/csrc/lsf_nodes.cuh"
extern "C" int emu_advect_nodes(const double *phi, const double *sbsrc, int nx, int ny, int nz, const double *xLo, double dx,
                                double *X, int nNode, double *phiSurf, double *gradPhiSurf, int iter, long long *moves)
{
    NodeConst c;
    c.sx = nx + 1; c.sxy = c.sx * (ny + 1); c.nx = nx; c.ny = ny; c.nz = nz;
    c.xLo[0] = xLo[0]; c.xLo[1] = xLo[1]; c.xLo[2] = xLo[2]; c.dx = dx; c.bSB = 8.1 * dx;
    int worst = 0;
    long long nm = 0;
    for (int n = 0; n < nNode; ++n) {
        double x[3] = {X[n], X[n + (size_t)nNode], X[n + 2 * (size_t)nNode]}, ps, gs[3];
        int mv;
        const int st = node_project(c, phi, sbsrc, x, ps, gs, iter, mv);
        if (st > worst) worst = st;
        if (st) continue;
        X[n] = x[0]; X[n + (size_t)nNode] = x[1]; X[n + 2 * (size_t)nNode] = x[2];
        phiSurf[n] = ps;
        for (int q = 0; q < 3; ++q) gradPhiSurf[n + (size_t)q * nNode] = gs[q];
        nm += mv;
    }
    if (moves) *moves = nm;
    return worst;
}

// The same on z-slabs: the global field is cut into `nranks` local arrays (owned planes + 3 ghost planes per neighbour, geometry
// of lsf_slab.cuh) and every value is fetched through a SlabView -- the accessor the z-slab kernel uses on the peers' slabs.
// Ghost planes are filled with NaN: the view must only ever read a plane from the rank that OWNS it.
extern "C" int emu_advect_nodes_slabs(const double *phi, const double *sbsrc, int nx, int ny, int NZ, int nranks, const double *xLo,
                                      double dx, double *X, int nNode, double *phiSurf, double *gradPhiSurf, int iter, long long *moves)
{
    NodeConst c;
    c.sx = nx + 1; c.sxy = c.sx * (ny + 1); c.nx = nx; c.ny = ny; c.nz = NZ;
    c.xLo[0] = xLo[0]; c.xLo[1] = xLo[1]; c.xLo[2] = xLo[2]; c.dx = dx; c.bSB = 8.1 * dx;
    std::vector<std::vector<double>> lp(nranks), ls(nranks);
    SlabView vp, vs;
    memset(&vp, 0, sizeof(vp)); memset(&vs, 0, sizeof(vs));
    vp.nranks = vs.nranks = nranks; vp.sxy = vs.sxy = c.sxy;
    for (int r = 0; r < nranks; ++r) {
        SlabGeom g;
        if (!slab_geom(NZ, nranks, r, g)) return -1;
        const size_t n = (size_t)c.sxy * (g.nzl + 1);
        lp[r].assign(n, NAN); ls[r].assign(n, NAN);
        for (int k = g.k0; k < g.k1; ++k) {
            memcpy(&lp[r][(size_t)(k - g.kbase) * c.sxy], phi + (size_t)k * c.sxy, sizeof(double) * c.sxy);
            memcpy(&ls[r][(size_t)(k - g.kbase) * c.sxy], sbsrc + (size_t)k * c.sxy, sizeof(double) * c.sxy);
        }
        vp.base[r] = lp[r].data(); vs.base[r] = ls[r].data();
        vp.shift[r] = vs.shift[r] = (long long)g.kbase * c.sxy;
        vp.kend[r] = vs.kend[r] = g.k1;
    }
    int worst = 0;
    long long nm = 0;
    for (int n = 0; n < nNode; ++n) {
        double x[3] = {X[n], X[n + (size_t)nNode], X[n + 2 * (size_t)nNode]}, ps, gs[3];
        int mv;
        const int st = node_project(c, vp, vs, x, ps, gs, iter, mv);
        if (st > worst) worst = st;
        if (st) continue;
        X[n] = x[0]; X[n + (size_t)nNode] = x[1]; X[n + 2 * (size_t)nNode] = x[2];
        phiSurf[n] = ps;
        for (int q = 0; q < 3; ++q) gradPhiSurf[n + (size_t)q * nNode] = gs[q];
        nm += mv;
    }
    if (moves) *moves = nm;
    return worst;
}

// fp32 twin of emu_march_sweep_slabs: one sweep on `nranks` z-slabs running concurrently, coupled only through the
// streaming-halo protocol (peer stores of float values + in_progress flags).
extern "C" double emu_march_sweep_slabs_f32(float *phi, const float *phiS, int nx, int ny, int NZ, int nranks, int raster,
                                            double dx, double h, int ncta, int order_m)
{
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    std::vector<SlabGeom> geo(nranks);
    for (int r = 0; r < nranks; ++r) if (!slab_geom(NZ, nranks, r, geo[r])) return -2.;
    std::vector<std::vector<float>> lphi(nranks), lphiS(nranks);
    std::vector<std::vector<double>> partial(nranks);
    std::vector<std::vector<int>> order(nranks);
    std::vector<std::vector<long long>> progress(nranks);
    std::vector<SlabSync> sync(nranks);
    std::vector<MarchParamsT<float>> P(nranks);
    std::vector<unsigned> ticket(nranks, 0u);
    std::vector<Ctrl> ctrl(nranks, Ctrl{0, 0, 0, 0, 0});
    memset(sync.data(), 0, sizeof(SlabSync) * nranks);
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        const size_t n = (size_t)sxy * (g.nzl + 1);
        lphi[r].assign(phi + (size_t)g.kbase * sxy, phi + (size_t)g.kbase * sxy + n);
        lphiS[r].assign(phiS + (size_t)g.kbase * sxy, phiS + (size_t)g.kbase * sxy + n);
        lphi[r].resize(n + 16, 1.e30f); lphiS[r].resize(n + 16, 1.e30f);
    }
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        MarchParamsT<float> &p = P[r];
        memset(&p, 0, sizeof(p));
        march_orient<CFGF>(p, nx, ny, g.nzl, sx, sxy, raster, g.kupd_lo, g.kupd_hi, g.kbase, NZ);
        p.phi = lphi[r].data(); p.phiS = lphiS[r].data();
        cc_f32(p.cc, dx, h);
        partial[r].assign(p.ntiles, 0.); order[r].resize(p.ntiles); progress[r].assign(p.ntiles, 0);
        march_fill_order(p.ntb, p.ntc, order[r].data(), order_m);
        p.partial = partial[r].data(); p.order = order[r].data(); p.progress = progress[r].data();
        p.ticket = &ticket[r]; p.ctrl = &ctrl[r]; p.epoch = 1;
        p.col_next = emu_colnext(p.ntb);
        const int up = p.fc ? r + 1 : r - 1, down = p.fc ? r - 1 : r + 1;
        if (up >= 0 && up < nranks) p.in_progress = sync[r].in_progress;
        if (down >= 0 && down < nranks) {
            p.push_delta = (lphi[down].data() + (long long)(g.kbase - geo[down].kbase) * sxy) - lphi[r].data();
            p.push_progress = sync[down].in_progress;
        }
    }
    constexpr int NT = CFGF::THREADS;
    std::vector<int> nc(nranks);
    size_t nthreads = 0;
    for (int r = 0; r < nranks; ++r) { nc[r] = ncta < P[r].ntiles ? ncta : P[r].ntiles; nthreads += (size_t)nc[r] * NT; }
    std::vector<std::vector<SmemF>> sm(nranks);
    std::vector<std::vector<EmuCta>> ctas(nranks);
    std::vector<ThreadArgF> args(nthreads);
    std::vector<pthread_t> th(nthreads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int r = 0; r < nranks; ++r) {
        sm[r] = std::vector<SmemF>(nc[r]);
        ctas[r] = std::vector<EmuCta>(nc[r]);
        for (int c = 0; c < nc[r]; ++c) pthread_barrier_init(&ctas[r][c].bar, nullptr, NT);
    }
    size_t q = 0;
    for (int r = 0; r < nranks; ++r)
        for (int c = 0; c < nc[r]; ++c)
            for (int t = 0; t < NT; ++t, ++q) {
                ThreadArgF &a = args[q];
                a.p = &P[r]; a.sm = &sm[r][c]; a.cta = &ctas[r][c]; a.tid = t; a.mg = 1;
                if (pthread_create(&th[q], &attr, thread_main_f32, &a) != 0) return -1.;
            }
    for (size_t i = 0; i < th.size(); ++i) pthread_join(th[i], nullptr);
    double s = 0.;
    for (int r = 0; r < nranks; ++r) {
        const SlabGeom &g = geo[r];
        for (int c = 0; c < nc[r]; ++c) pthread_barrier_destroy(&ctas[r][c].bar);
        if (ctrl[r].status != 0) return -3.;
        memcpy(phi + (size_t)g.k0 * sxy, lphi[r].data() + (size_t)g.own_lo * sxy, sizeof(float) * (size_t)sxy * (g.k1 - g.k0));
        for (int i = 0; i < P[r].ntiles; ++i) s += partial[r][i];
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------
// Overlapped sweeps (march_multi_cta): `nsweeps` consecutive sweeps starting with raster `first_raster` in ONE pass
// over a shared ticket counter, CTAs moving on to the next sweep's tiles while the previous one drains; boundary block
// folded into the tiles, boundary values alternating between phi and a shell array.  rms_sum[s] receives the sum of
// (new-old)^2 over ALL points of sweep s.  Returns 0, or a negative code.
template <class AR>
struct MultiArg { const MarchParams *sweeps; int nsweeps; unsigned *ticket; Smem *sm; EmuCta *cta; int tid; };

template <class AR>
static void *thread_main_multi(void *v)
{
    MultiArg<AR> *a = (MultiArg<AR> *)v;
    emu_cta = a->cta;
    march_multi_cta<AR, CFG>(a->sweeps, a->nsweeps, a->ticket, *a->sm, a->tid);
    return nullptr;
}

extern "C" int emu_march_overlapped(double *phi, const double *phiS, int nx, int ny, int nz, int nsweeps, int first_raster,
                                    double dx, double h, int arith, int ncta, double *rms_sum)
{
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    const size_t np = (size_t)sxy * (nz + 1);
    std::vector<double> shell(np, NAN);                       // only boundary points are ever read, after being written
    std::vector<MarchParams> P(nsweeps);
    MarchParams p0;
    memset(&p0, 0, sizeof(p0));
    march_orient<CFG>(p0, nx, ny, nz, sx, sxy, 1);
    const int ntiles = p0.ntiles;
    std::vector<double> partial((size_t)nsweeps * ntiles, 0.), partial_bc((size_t)nsweeps * ntiles, 0.);
    std::vector<long long> progress((size_t)nsweeps * ntiles, 0);
    std::vector<int> order(ntiles);
    march_fill_order(p0.ntb, p0.ntc, order.data());
    unsigned ticket = 0;
    Ctrl ctrl = {0, 0, 0, 0, 0};
    for (int s = 0; s < nsweeps; ++s) {
        MarchParams &p = P[s];
        memset(&p, 0, sizeof(p));
        march_orient<CFG>(p, nx, ny, nz, sx, sxy, (first_raster - 1 + s) % 8 + 1);
        p.phi = phi; p.phiS = phiS;
        p.cc.dx = dx; p.cc.inv_dx = 1. / dx; p.cc.k12 = 1. / (12. * dx); p.cc.dx2 = dx * dx; p.cc.h = h;
        p.partial = partial.data() + (size_t)s * ntiles; p.partial_bc = partial_bc.data() + (size_t)s * ntiles;
        p.order = order.data(); p.progress = progress.data() + (size_t)s * ntiles;
        p.ticket = &ticket; p.ctrl = &ctrl; p.epoch = 1;
        p.shell_rd_delta = (s & 1) ? shell.data() - phi : 0;
        p.shell_wr_delta = ((s + 1) & 1) ? shell.data() - phi : 0;
        p.fold_bc = 1;
        if (s > 0) {
            p.prev_progress = progress.data() + (size_t)(s - 1) * ntiles;
            p.prev_fin = (p.epoch << 32) + M_BIAS + M_FIN;
            p.prev_fb = P[s - 1].fb; p.prev_fc = P[s - 1].fc;
        }
    }
    if (ncta > ntiles) ncta = ntiles;
    std::vector<Smem> sm(ncta);
    std::vector<EmuCta> ctas(ncta);
    std::vector<pthread_t> th((size_t)ncta * M_THREADS);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 512 * 1024);
    for (int c = 0; c < ncta; ++c) pthread_barrier_init(&ctas[c].bar, nullptr, M_THREADS);
    std::vector<MultiArg<ExactArith>> ax((size_t)ncta * M_THREADS);
    std::vector<MultiArg<FastArith>> af((size_t)ncta * M_THREADS);
    for (int c = 0; c < ncta; ++c)
        for (int t = 0; t < M_THREADS; ++t) {
            const size_t q = (size_t)c * M_THREADS + t;
            int rc;
            if (arith == 1) {
                ax[q] = MultiArg<ExactArith>{P.data(), nsweeps, &ticket, &sm[c], &ctas[c], t};
                rc = pthread_create(&th[q], &attr, thread_main_multi<ExactArith>, &ax[q]);
            } else {
                af[q] = MultiArg<FastArith>{P.data(), nsweeps, &ticket, &sm[c], &ctas[c], t};
                rc = pthread_create(&th[q], &attr, thread_main_multi<FastArith>, &af[q]);
            }
            if (rc != 0) return -1;
        }
    for (size_t q = 0; q < th.size(); ++q) pthread_join(th[q], nullptr);
    for (int c = 0; c < ncta; ++c) pthread_barrier_destroy(&ctas[c].bar);
    if (ctrl.status != 0) return -3;
    if (nsweeps & 1)                                          // the last boundary block went to the shell array
        for (int k = 0; k <= nz; ++k)
            for (int j = 0; j <= ny; ++j)
                for (int i = 0; i <= nx; ++i)
                    if (i == 0 || i == nx || j == 0 || j == ny || k == 0 || k == nz) phi[i + sx * j + sxy * k] = shell[i + sx * j + sxy * k];
    if (rms_sum)
        for (int s = 0; s < nsweeps; ++s) {
            double t = 0.;
            for (int q = 0; q < ntiles; ++q) t += partial[(size_t)s * ntiles + q] + partial_bc[(size_t)s * ntiles + q];
            rms_sum[s] = t;
        }
    return 0;
}

// The second packaging of the overlapped sweeps (one kernel per sweep, chained by programmatic dependent launch on the
// GPU): here every sweep gets its own CTAs and its own ticket counter and ALL sweeps' CTAs run at once -- the most
// aggressive overlap the flags have to cope with (on the GPU at most two sweeps are resident together).
struct SweepArg { const MarchParams *p; Smem *sm; EmuCta *cta; int tid; int arith; };

static void *thread_main_ov_sweep(void *v)
{
    SweepArg *a = (SweepArg *)v;
    emu_cta = a->cta;
    const MarchParams &p = *a->p;
    const int o = (p.fa ? 1 : 0) | (p.fb ? 2 : 0) | (p.fc ? 4 : 0);
#define OVRUN(AR)                                                                                   \
    switch (o) {                                                                                    \
    case 0: march_cta<AR, false, false, false, CFG, false, true>(p, *a->sm, a->tid); break;         \
    case 1: march_cta<AR, true, false, false, CFG, false, true>(p, *a->sm, a->tid); break;          \
    case 2: march_cta<AR, false, true, false, CFG, false, true>(p, *a->sm, a->tid); break;          \
    case 3: march_cta<AR, true, true, false, CFG, false, true>(p, *a->sm, a->tid); break;           \
    case 4: march_cta<AR, false, false, true, CFG, false, true>(p, *a->sm, a->tid); break;          \
    case 5: march_cta<AR, true, false, true, CFG, false, true>(p, *a->sm, a->tid); break;           \
    case 6: march_cta<AR, false, true, true, CFG, false, true>(p, *a->sm, a->tid); break;           \
    default: march_cta<AR, true, true, true, CFG, false, true>(p, *a->sm, a->tid); break;           \
    }
    if (a->arith == 1) { OVRUN(ExactArith) } else { OVRUN(FastArith) }
#undef OVRUN
    return nullptr;
}

extern "C" int emu_march_overlapped_per_sweep(double *phi, const double *phiS, int nx, int ny, int nz, int nsweeps, int first_raster,
                                              double dx, double h, int arith, int ncta, double *rms_sum)
{
    const long long sx = nx + 1, sxy = sx * (ny + 1);
    const size_t np = (size_t)sxy * (nz + 1);
    std::vector<double> shell(np, NAN);
    std::vector<MarchParams> P(nsweeps);
    MarchParams p0;
    memset(&p0, 0, sizeof(p0));
    march_orient<CFG>(p0, nx, ny, nz, sx, sxy, 1);
    const int ntiles = p0.ntiles;
    std::vector<double> partial((size_t)nsweeps * ntiles, 0.), partial_bc((size_t)nsweeps * ntiles, 0.);
    std::vector<long long> progress((size_t)nsweeps * ntiles, 0);
    std::vector<int> order(ntiles);
    march_fill_order(p0.ntb, p0.ntc, order.data());
    std::vector<unsigned> tickets(nsweeps, 0u);
    Ctrl ctrl = {0, 0, 0, 0, 0};
    for (int s = 0; s < nsweeps; ++s) {
        MarchParams &p = P[s];
        memset(&p, 0, sizeof(p));
        march_orient<CFG>(p, nx, ny, nz, sx, sxy, (first_raster - 1 + s) % 8 + 1);
        p.phi = phi; p.phiS = phiS;
        p.cc.dx = dx; p.cc.inv_dx = 1. / dx; p.cc.k12 = 1. / (12. * dx); p.cc.dx2 = dx * dx; p.cc.h = h;
        p.partial = partial.data() + (size_t)s * ntiles; p.partial_bc = partial_bc.data() + (size_t)s * ntiles;
        p.order = order.data(); p.progress = progress.data() + (size_t)s * ntiles;
        p.ticket = &tickets[s]; p.ctrl = &ctrl; p.epoch = 1;
        p.shell_rd_delta = (s & 1) ? shell.data() - phi : 0;
        p.shell_wr_delta = ((s + 1) & 1) ? shell.data() - phi : 0;
        p.fold_bc = 1;
        if (s > 0) {
            p.prev_progress = progress.data() + (size_t)(s - 1) * ntiles;
            p.prev_fin = (p.epoch << 32) + M_BIAS + M_FIN;
            p.prev_fb = P[s - 1].fb; p.prev_fc = P[s - 1].fc;
        }
    }
    if (ncta > ntiles) ncta = ntiles;
    const size_t nthreads = (size_t)nsweeps * ncta * M_THREADS;
    std::vector<Smem> sm((size_t)nsweeps * ncta);
    std::vector<EmuCta> ctas((size_t)nsweeps * ncta);
    std::vector<SweepArg> args(nthreads);
    std::vector<pthread_t> th(nthreads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 512 * 1024);
    for (size_t c = 0; c < ctas.size(); ++c) pthread_barrier_init(&ctas[c].bar, nullptr, M_THREADS);
    size_t q = 0;
    for (int s = 0; s < nsweeps; ++s)
        for (int c = 0; c < ncta; ++c)
            for (int t = 0; t < M_THREADS; ++t, ++q) {
                args[q] = SweepArg{&P[s], &sm[(size_t)s * ncta + c], &ctas[(size_t)s * ncta + c], t, arith};
                if (pthread_create(&th[q], &attr, thread_main_ov_sweep, &args[q]) != 0) return -1;
            }
    for (size_t i = 0; i < th.size(); ++i) pthread_join(th[i], nullptr);
    for (size_t c = 0; c < ctas.size(); ++c) pthread_barrier_destroy(&ctas[c].bar);
    if (ctrl.status != 0) return -3;
    if (nsweeps & 1)
        for (int k = 0; k <= nz; ++k)
            for (int j = 0; j <= ny; ++j)
                for (int i = 0; i <= nx; ++i)
                    if (i == 0 || i == nx || j == 0 || j == ny || k == 0 || k == nz) phi[i + sx * j + sxy * k] = shell[i + sx * j + sxy * k];
    if (rms_sum)
        for (int s = 0; s < nsweeps; ++s) {
            double t = 0.;
            for (int i = 0; i < ntiles; ++i) t += partial[(size_t)s * ntiles + i] + partial_bc[(size_t)s * ntiles + i];
            rms_sum[s] = t;
        }
    return 0;
}
