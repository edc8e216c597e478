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
                                    double bNB, int check_abs, Ctrl *ctrl)
{
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzp = dm.nz + 1;
    const long long fxy = nxp * nyp, fxz = nxp * nzp, fyz = nyp * nzp;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int i, j, k;
    if (t < 2 * fxy) { k = (t >= fxy) ? dm.nz : 0; t %= fxy; i = (int)(t % nxp); j = (int)(t / nxp); }
    else if ((t -= 2 * fxy) < 2 * fxz) { j = (t >= fxz) ? dm.ny : 0; t %= fxz; i = (int)(t % nxp); k = (int)(t / nxp); }
    else if ((t -= 2 * fxz) < 2 * fyz) { i = (t >= fyz) ? dm.nx : 0; t %= fyz; j = (int)(t % nyp); k = (int)(t / nyp); }
    else return;
    const long long q = i + dm.sx * j + dm.sxy * k;
    const bool hit = (mask && mask[q]) || (check_abs && fabs(phi[q]) < bNB);
    if (hit) ctrl->status = LSF_ERR_BAND_ON_BOUNDARY;
}

void launch_mm_check_boundary(Grid *g, const uint8_t *mask, double dx, bool check_abs)
{
    const Dims &dm = g->dm;
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzp = dm.nz + 1;
    const long long tot = 2 * (nxp * nyp + nxp * nzp + nyp * nzp);
    k_mm_check_boundary<<<(unsigned)((tot + 255) / 256), 256, 0, G.stream>>>(g->phi, mask, dm, 4.1 * dx, check_abs ? 1 : 0, g->ctrl);
    G.n_launch++;
}

int mm_march_prepare(Grid *g)
{
    int rc = march_prepare(g);   // ticket, per-tile progress flags, tile order: shared with the reinit sweep
    if (rc) return rc;
    MmParams p;
    mm_orient<MCFG>(p, g->dm.nx, g->dm.ny, g->dm.nz);
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
    mm_orient<MCFG>(p, g->dm.nx, g->dm.ny, g->dm.nz);
    p.A = A; p.B = B; p.mask = mask;
    p.bNB = 4.1 * dx; p.dxx = 1. / (dx * dx); p.h1 = h1;
    p.partial = g->partial; p.ticket = g->march_ticket; p.order = march_order();
    p.progress = g->march_progress; p.epoch = ++g->march_epoch; p.ctrl = g->ctrl;
    cudaMemsetAsync(g->march_ticket, 0, sizeof(unsigned), G.stream);
    const int ncta = p.ntiles < 4 * G.num_sms ? p.ntiles : 4 * G.num_sms;
    k_minmax_march<<<ncta, MCFG::THREADS, sizeof(MmSmem<MCFG>), G.stream>>>(p);
    G.n_launch++;
}

}  // namespace lsf
