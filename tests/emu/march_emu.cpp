// march_emu.cpp -- runs levelsetfortran_b200/csrc/lsf_march.cuh on the CPU: every CUDA thread is
// an OS thread, every CTA a pthread barrier domain, CTAs run concurrently and synchronise through
// the same ticket/progress flags as on the GPU.  Test infrastructure only (tests/test_march_emu.py).
#define LSF_EMU 1
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../levelsetfortran_b200/csrc/lsf_march.cuh"
#include "../../levelsetfortran_b200/csrc/lsf_mm_march.cuh"

namespace lsf { thread_local EmuCta *emu_cta = nullptr; }
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
