// lsf_mm_march.cu -- GPU launch of the fused min/max-flow iteration kernel (lsf_mm_march.cuh).
#include "lsf_internal.cuh"
#include "lsf_mm_march.cuh"

namespace lsf {

typedef MmCfgDefault MCFG;

__global__ void __launch_bounds__(MCFG::THREADS, 4)
k_minmax_march(const MmParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MmSmem<MCFG> &sm = *reinterpret_cast<MmSmem<MCFG> *>(smem_raw);
    mm_cta<MCFG>(p, sm, threadIdx.x);
}

// A narrow-band cell on the grid boundary makes the reference read phi(-1,..) (set3d.f90:402-403 ->
// subs.f90:387): undefined there, an error here.  Boundary values never change during the flow, so one
// check before the loop covers every iteration.
__global__ void k_mm_check_boundary(const double *__restrict__ phi, const uint8_t *__restrict__ mask, Dims dm,
                                    double bNB, int check_abs, Ctrl *ctrl, int kA, int kB, int kbase, int NZ, int hasLo, int hasHi)
{
    // boundary points of this rank's OWNED planes (z-slab: the global k faces only where it holds them)
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzm = kB - kA + 1, nym = dm.ny - 1;
    const long long fk = nxp * nyp, fj = nxp * nzm, fi = nym * nzm;
    const long long nkf = (long long)(hasLo + hasHi) * fk;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int i, j, k;
    if (t < nkf) { k = (t >= fk || !hasLo) ? NZ - kbase : -kbase; t %= fk; i = (int)(t % nxp); j = (int)(t / nxp); }
    else if ((t -= nkf) < 2 * fj) { j = (t >= fj) ? dm.ny : 0; t %= fj; i = (int)(t % nxp); k = kA + (int)(t / nxp); }
    else if ((t -= 2 * fj) < 2 * fi) { i = (t >= fi) ? dm.nx : 0; t %= fi; j = 1 + (int)(t % nym); k = kA + (int)(t / nym); }
    else return;
    const long long q = i + dm.sx * j + dm.sxy * k;
    const bool hit = (mask && mask[q]) || (check_abs && fabs(phi[q]) < bNB);
    if (hit) ctrl->status = LSF_ERR_BAND_ON_BOUNDARY;
}

void launch_mm_check_boundary(Grid *g, const uint8_t *mask, double dx, bool check_abs)
{
    const Dims &dm = g->dm;
    const SlabGeom &sg = g->sg;
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzm = sg.kupd_hi - sg.kupd_lo + 1, nym = dm.ny - 1;
    const int hasLo = sg.k0 == 0, hasHi = sg.k1 == sg.NZ + 1;
    const long long tot = (hasLo + hasHi) * nxp * nyp + 2 * (nxp * nzm + nym * nzm);
    k_mm_check_boundary<<<(unsigned)((tot + 255) / 256), 256, 0, G.stream>>>(g->phi, mask, dm, 4.1 * dx, check_abs ? 1 : 0, g->ctrl,
                                                                              sg.kupd_lo, sg.kupd_hi, sg.kbase, sg.NZ, hasLo, hasHi);
    G.n_launch++;
}

int mm_march_prepare(Grid *g)
{
    int rc = march_prepare(g);   // ticket, per-tile progress flags, tile order: shared with the reinit sweep
    if (rc) return rc;
    MmParams p;
    mm_orient<MCFG>(p, g->dm.nx, g->dm.ny, g->dm.nz, g->sg.kupd_lo, g->sg.kupd_hi);
    if (p.ntiles != march_ntiles(g)) return set_error(LSF_ERR_ARG, "minmax: tile grid mismatch");
    static bool attr_done = false;
    if (!attr_done) {
        LSF_CUDA(cudaFuncSetAttribute(k_minmax_march, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MmSmem<MCFG>)));
        attr_done = true;
    }
    return LSF_OK;
}

// One iteration: reads A (phi_old), writes B (phi_new), per-tile RMS partials into g->partial.
void launch_minmax_iteration_march(Grid *g, const double *A, double *B, const uint8_t *mask, double dx, double h1)
{
    MmParams p;
    mm_orient<MCFG>(p, g->dm.nx, g->dm.ny, g->dm.nz, g->sg.kupd_lo, g->sg.kupd_hi);
    p.A = A; p.B = B; p.mask = mask;
    p.bNB = 4.1 * dx; p.dxx = 1. / (dx * dx); p.h1 = h1;
    p.partial = g->partial; p.ticket = g->march_ticket; p.order = march_order();
    p.progress = g->march_progress; p.epoch = ++g->march_epoch; p.ctrl = g->ctrl;
    p.in_progress = nullptr; p.push_delta = 0; p.push_progress = nullptr; p.halo_seq = nullptr;
    p.halo_need[0] = p.halo_need[1] = 0;
    if (sharded(g)) {
        // pass B is an ascending-k Gauss-Seidel sweep (subs.f90:473): the rank below is upstream
        const SlabGeom &sg = g->sg;
        if (sg.rank > 0) p.in_progress = g->sync->in_progress;
        if (sg.rank < sg.nranks - 1) {
            SlabGeom dg;
            slab_geom(sg.NZ, sg.nranks, sg.rank + 1, dg);
            p.push_delta = (peer_ptr(g, sg.rank + 1, B) + (long long)(sg.kbase - dg.kbase) * g->dm.sxy) - B;
            p.push_progress = peer_ptr(g, sg.rank + 1, g->sync)->in_progress;
        }
        p.halo_seq = g->sync->halo_seq;
        if (sg.rank > 0) p.halo_need[0] = g->phase;
        if (sg.rank < sg.nranks - 1) p.halo_need[1] = g->phase;
        g->prev_sweep_valid = false;
    }
    cudaMemsetAsync(g->march_ticket, 0, sizeof(unsigned), G.stream);
    const int ncta = p.ntiles < 4 * G.num_sms ? p.ntiles : 4 * G.num_sms;
    k_minmax_march<<<ncta, MCFG::THREADS, sizeof(MmSmem<MCFG>), G.stream>>>(p);
    G.n_launch++;
}

}  // namespace lsf
