// lsf_kernels.cu -- simple whole-grid kernels: hyperplane-scheduled Gauss-Seidel sweeps (the
// cross-check schedule), boundary extrapolation, RMS reduction, loop control, narrow band,
// min/max flow and the inside/outside sign search.  The production reinit schedule (skewed
// x-marching column tiles) lives in lsf_march.cu.
#include <stdlib.h>

#include "lsf_internal.cuh"

namespace lsf {

// =====================================================================================
// K2 (plane schedule): one launch per hyperplane a+b+c = s of the sweep-oriented indices.
// Cells of one hyperplane do not depend on each other (7-point star stencil, +-3), cells at
// -m were finished by earlier launches and cells at +m are untouched: exactly the data the
// reference's in-place raster loop (subs.f90:742-852) sees.
// =====================================================================================
template <class AR, bool WG>
__global__ void __launch_bounds__(128)
k_reinit_plane(double *__restrict__ phi, const double *__restrict__ phiS, Dims dm, int d0, int d1, int d2,
               int s, CellConst cc, double *__restrict__ gradPhi, double *__restrict__ gradPhiMag,
               Ctrl *__restrict__ ctrl)
{
    if (ctrl->done) return;
    const int a = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int b = 1 + blockIdx.y;
    const int c = s - a - b;
    if (a > dm.nx - 1 || c < 1 || c > dm.nz - 1) return;
    const int i = d0 > 0 ? a : dm.nx - a;
    const int j = d1 > 0 ? b : dm.ny - b;
    const int k = d2 > 0 ? c : dm.nz - c;
    const long long q = i + dm.sx * j + dm.sxy * k;
    const bool hi = (i > 3) && (i < dm.nx - 4) && (j > 3) && (j < dm.ny - 4) && (k > 3) && (k < dm.nz - 4);
    double vx[7], vy[7], vz[7];
    const double pc = phi[q];
    if (hi) {
#pragma unroll
        for (int m = -3; m <= 3; ++m) {
            vx[m + 3] = phi[q + m];
            vy[m + 3] = phi[q + m * dm.sx];
            vz[m + 3] = phi[q + m * dm.sxy];
        }
    } else {
        vx[3] = vy[3] = vz[3] = pc;
        vx[2] = phi[q - 1]; vx[4] = phi[q + 1];
        vy[2] = phi[q - dm.sx]; vy[4] = phi[q + dm.sx];
        vz[2] = phi[q - dm.sxy]; vz[4] = phi[q + dm.sxy];
    }
    double g[3], gM;
    bool sens;
    const double pn = reinit_cell<AR>(vx, vy, vz, phiS[q], hi, cc, g, gM, sens);
    if (sens) ctrl->guard = 1;
    phi[q] = pn;
    if (WG) {
        const long long np = dm.sxy * (dm.nz + 1);
        if (gradPhi) { gradPhi[q] = g[0]; gradPhi[q + np] = g[1]; gradPhi[q + 2 * np] = g[2]; }
        if (gradPhiMag) gradPhiMag[q] = gM;
    }
}

void launch_reinit_sweep_plane(Grid *g, int raster, const CellConst &cc, double *gradPhi, double *gradPhiMag, double *phi)
{
    if (!phi) phi = g->phi;
    int d[3];
    raster_dirs(raster, d);
    const Dims &dm = g->dm;
    dim3 blk(128), grd((dm.nx - 1 + 127) / 128, dm.ny - 1);
    const bool wg = gradPhi || gradPhiMag;
    for (int s = 3; s <= (dm.nx - 1) + (dm.ny - 1) + (dm.nz - 1); ++s) {
        if (G.arith_run == LSF_ARITH_EXACT) {
            if (wg) k_reinit_plane<ExactArith, true><<<grd, blk, 0, G.stream>>>(phi, g->phiS, dm, d[0], d[1], d[2], s, cc, gradPhi, gradPhiMag, g->ctrl);
            else k_reinit_plane<ExactArith, false><<<grd, blk, 0, G.stream>>>(phi, g->phiS, dm, d[0], d[1], d[2], s, cc, nullptr, nullptr, g->ctrl);
        } else {
            if (wg) k_reinit_plane<FastArith, true><<<grd, blk, 0, G.stream>>>(phi, g->phiS, dm, d[0], d[1], d[2], s, cc, gradPhi, gradPhiMag, g->ctrl);
            else k_reinit_plane<FastArith, false><<<grd, blk, 0, G.stream>>>(phi, g->phiS, dm, d[0], d[1], d[2], s, cc, nullptr, nullptr, g->ctrl);
        }
        G.n_launch++;
    }
}

// =====================================================================================
// K3: extrapolation boundary block of reinit, subs.f90:858-897, in closed form: for a boundary
// point c with B boundary axes of which H are on the high side,
//     phi(c) = phi(clamp(c, 1..n-1)) + dx   applied min(1+H, B) times.
// (Derived by symbolic execution of the literal statement order; verified against the literal
// loop in tests/test_oracle_pins.py.)  Sources are interior cells only, so there is no hazard.
// One thread per point of the six faces; edge/corner points are written more than once with
// the same value.
// =====================================================================================
__global__ void k_reinit_bc(double *__restrict__ phi, Dims dm, double dx, const Ctrl *__restrict__ ctrl)
{
    if (ctrl->done) return;
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzp = dm.nz + 1;
    const long long fxy = nxp * nyp, fxz = nxp * nzp, fyz = nyp * nzp;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int i, j, k;
    if (t < 2 * fxy) { k = (t >= fxy) ? dm.nz : 0; t %= fxy; i = (int)(t % nxp); j = (int)(t / nxp); }
    else if ((t -= 2 * fxy) < 2 * fxz) { j = (t >= fxz) ? dm.ny : 0; t %= fxz; i = (int)(t % nxp); k = (int)(t / nxp); }
    else if ((t -= 2 * fxz) < 2 * fyz) { i = (t >= fyz) ? dm.nx : 0; t %= fyz; j = (int)(t % nyp); k = (int)(t / nyp); }
    else return;
    const int B = (i == 0 || i == dm.nx) + (j == 0 || j == dm.ny) + (k == 0 || k == dm.nz);
    const int H = (i == dm.nx) + (j == dm.ny) + (k == dm.nz);
    const int m = min(1 + H, B);
    const int ci = min(max(i, 1), dm.nx - 1), cj = min(max(j, 1), dm.ny - 1), ck = min(max(k, 1), dm.nz - 1);
    double v = phi[ci + dm.sx * cj + dm.sxy * ck];
    for (int r = 0; r < m; ++r) v = __dadd_rn(v, dx);
    phi[i + dm.sx * j + dm.sxy * k] = v;
}

void launch_reinit_bc_buf(Grid *g, double *buf, double dx)      // the same block on any field of the grid's shape (lsf_rk.cu stages)
{
    const Dims &dm = g->dm;
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzp = dm.nz + 1;
    const long long tot = 2 * (nxp * nyp + nxp * nzp + nyp * nzp);
    k_reinit_bc<<<(unsigned)((tot + 255) / 256), 256, 0, G.stream>>>(buf, dm, dx, g->ctrl);
    G.n_launch++;
}

void launch_reinit_bc(Grid *g, double dx)
{
    const Dims &dm = g->dm;
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzp = dm.nz + 1;
    const long long tot = 2 * (nxp * nyp + nxp * nzp + nyp * nzp);
    k_reinit_bc<<<(unsigned)((tot + 255) / 256), 256, 0, G.stream>>>(g->phi, dm, dx, g->ctrl);
    G.n_launch++;
}

// K3+K4 fused: the same closed form, every boundary point visited exactly once (k faces whole,
// j faces without the k faces' points, i faces without both), and the boundary part of the RMS sum
// (subs.f90:902-914) accumulated on the fly: the sweep never writes boundary points, so the value
// found in phi IS phiN there.  Together with the per-tile sums of the sweep kernel this makes the
// separate RMS pass and the phiN array unnecessary in reinit.  Per-block partials, fixed order.
// z-slabs (lsf_slab.cuh): a rank visits the boundary points of its OWNED planes only -- the global k faces if
// it holds them (hasLo / hasHi), and the i/j faces of its updated planes kA..kB; B, H and the clamp are
// properties of the GLOBAL plane index k + kbase in 0..NZ.  One GPU: kA = 1, kB = nz-1, kbase = 0, NZ = nz.
__global__ void __launch_bounds__(256)
k_reinit_bc_rms(double *__restrict__ phi, Dims dm, double dx, double *__restrict__ partial,
                const Ctrl *__restrict__ ctrl, int kA, int kB, int kbase, int NZ, int hasLo, int hasHi)
{
    if (ctrl->done) return;
    __shared__ double sh[256];
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzm = kB - kA + 1, nym = dm.ny - 1;
    const long long fk = nxp * nyp, fj = nxp * nzm, fi = nym * nzm;
    const long long nkf = (long long)(hasLo + hasHi) * fk;
    const long long tot = nkf + 2 * (fj + fi);
    double acc = 0.;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < tot; t0 += (long long)gridDim.x * blockDim.x) {
        long long t = t0;
        int i, j, k;
        if (t < nkf) { k = (t >= fk || !hasLo) ? NZ - kbase : -kbase; t %= fk; i = (int)(t % nxp); j = (int)(t / nxp); }
        else if ((t -= nkf) < 2 * fj) { j = (t >= fj) ? dm.ny : 0; t %= fj; i = (int)(t % nxp); k = kA + (int)(t / nxp); }
        else { t -= 2 * fj; i = (t >= fi) ? dm.nx : 0; t %= fi; j = 1 + (int)(t % nym); k = kA + (int)(t / nym); }
        const int kg = k + kbase;
        const int B = (i == 0 || i == dm.nx) + (j == 0 || j == dm.ny) + (kg == 0 || kg == NZ);
        const int H = (i == dm.nx) + (j == dm.ny) + (kg == NZ);
        const int m = min(1 + H, B);
        const int ci = min(max(i, 1), dm.nx - 1), cj = min(max(j, 1), dm.ny - 1), ck = min(max(kg, 1), NZ - 1) - kbase;
        double v = phi[ci + dm.sx * cj + dm.sxy * ck];
        for (int r = 0; r < m; ++r) v = __dadd_rn(v, dx);
        const long long q = i + dm.sx * j + dm.sxy * k;
        const double d = __dsub_rn(v, phi[q]);
        acc = __dadd_rn(acc, __dmul_rn(d, d));
        phi[q] = v;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// the same on any field of the grid's shape (stage buffers of lsf_grid_reinit_rk3 on z-slabs); the RMS part goes to `partial`
void launch_reinit_bc_rms_buf(Grid *g, double *buf, double dx, double *partial)
{
    const SlabGeom &sg = g->sg;
    k_reinit_bc_rms<<<BC_BLOCKS, 256, 0, G.stream>>>(buf, g->dm, dx, partial, g->ctrl, sg.kupd_lo, sg.kupd_hi, sg.kbase, sg.NZ, sg.k0 == 0,
                                                     sg.k1 == sg.NZ + 1);
    G.n_launch++;
}

void launch_reinit_bc_rms(Grid *g, double dx, int partial_off)
{
    const SlabGeom &sg = g->sg;
    k_reinit_bc_rms<<<BC_BLOCKS, 256, 0, G.stream>>>(g->phi, g->dm, dx, g->partial + partial_off, g->ctrl, sg.kupd_lo, sg.kupd_hi,
                                                     sg.kbase, sg.NZ, sg.k0 == 0, sg.k1 == sg.NZ + 1);
    G.n_launch++;
}

// =====================================================================================
// K4: RMS of (phi - phiN) over ALL points (subs.f90:902-914, set3d.f90:435-447) as per-block
// partial sums in a fixed order (deterministic run to run), optionally fused with phiN = phi
// (subs.f90:921; legal for reinit where phiN is a local).
// =====================================================================================
template <bool COPY>
__global__ void __launch_bounds__(256)
k_rms_partial(const double *__restrict__ phi, double *__restrict__ phiN, long long np, double *__restrict__ partial,
              const Ctrl *__restrict__ ctrl)
{
    if (ctrl->done) return;
    __shared__ double sh[256];
    const long long per = (np + gridDim.x - 1) / gridDim.x;
    const long long lo = per * blockIdx.x, hi = min(np, lo + per);
    double acc = 0.;
    for (long long q = lo + threadIdx.x; q < hi; q += blockDim.x) {
        const double p = phi[q];
        const double d = __dsub_rn(p, phiN[q]);
        acc = __dadd_rn(acc, __dmul_rn(d, d));
        if (COPY) phiN[q] = p;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

void launch_rms(Grid *g, bool copy)
{
    if (copy) k_rms_partial<true><<<RMS_BLOCKS, 256, 0, G.stream>>>(g->phi, g->phiN, g->np, g->partial, g->ctrl);
    else k_rms_partial<false><<<RMS_BLOCKS, 256, 0, G.stream>>>(g->phi, g->phiN, g->np, g->partial, g->ctrl);
    G.n_launch++;
}

// Loop control: sum the partials in a fixed order, phiErr = sqrt(sum / (nx*ny*nz))
// (subs.f90:914; 64-bit product instead of the reference's overflowing int32), record it, and
// apply the EXIT (:915) / NaN STOP (:926) tests.
__global__ void __launch_bounds__(256)
k_finalize(const double *__restrict__ partial, int npart, Ctrl *ctrl, double *__restrict__ hist, int hist_off,
           double denom, double tol)
{
    if (ctrl->done) return;
    if (ctrl->status < 0) {          // an error flagged by a kernel of this iteration ends the loop as it is
        if (threadIdx.x == 0) { ctrl->done = 1; ctrl->n_exit = ctrl->n; }
        return;
    }
    __shared__ double sh[256];
    double acc = 0.;
    for (int q = threadIdx.x; q < npart; q += 256) acc = __dadd_rn(acc, partial[q]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double err = sqrt(sh[0] / denom);
        const int n = ctrl->n;
        hist[n - hist_off] = err;
        if (err < tol) { ctrl->done = 1; ctrl->status = 0; ctrl->n_exit = n; }
#if !defined(LSF_EXP_NOSYNC)           // (timing experiment without the step barrier: results are garbage, keep the loop running)
        else if (err != err) { ctrl->done = 1; ctrl->status = 1; ctrl->n_exit = n; }
#endif
        ctrl->n = n + 1;
    }
}

void launch_finalize(Grid *g, int npart, int hist_off, double tol, const double *partial)
{
    const double denom = (double)((long long)g->dm.nx * g->dm.ny * g->dm.nz);
    k_finalize<<<1, 256, 0, G.stream>>>(partial ? partial : g->partial, npart, g->ctrl, g->hist, hist_off, denom, tol);
    G.n_launch++;
}

// overlapped sweeps: after an odd number of sweeps the current boundary values live in the shell array
__global__ void k_copy_boundary(double *__restrict__ dst, const double *__restrict__ src, Dims dm)
{
    const long long nxp = dm.nx + 1, nyp = dm.ny + 1, nzp = dm.nz + 1;
    const long long fxy = nxp * nyp, fxz = nxp * nzp, fyz = nyp * nzp;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < 2 * (fxy + fxz + fyz); t0 += (long long)gridDim.x * blockDim.x) {
        long long t = t0;
        long long i, j, k;
        if (t < 2 * fxy) { k = (t >= fxy) ? dm.nz : 0; t %= fxy; i = t % nxp; j = t / nxp; }
        else if ((t -= 2 * fxy) < 2 * fxz) { j = (t >= fxz) ? dm.ny : 0; t %= fxz; i = t % nxp; k = t / nxp; }
        else { t -= 2 * fxz; i = (t >= fyz) ? dm.nx : 0; t %= fyz; j = t % nyp; k = t / nyp; }
        const long long q = i + dm.sx * j + dm.sxy * k;
        dst[q] = src[q];
    }
}

void launch_copy_boundary(Grid *g, double *dst, const double *src)
{
    k_copy_boundary<<<BC_BLOCKS, 256, 0, G.stream>>>(dst, src, g->dm);
    G.n_launch++;
}

__global__ void k_copy_if_running(double *__restrict__ dst, const double *__restrict__ src, long long np,
                                  const Ctrl *__restrict__ ctrl)
{
    if (ctrl->done) return;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < np; q += (long long)gridDim.x * blockDim.x)
        dst[q] = src[q];
}

void launch_copy_if_running(Grid *g, double *dst, const double *src)
{
    k_copy_if_running<<<RMS_BLOCKS, 256, 0, G.stream>>>(dst, src, g->np, g->ctrl);
    G.n_launch++;
}

// =====================================================================================
// K5: narrowBand, subs.f90:178-207
// =====================================================================================
__global__ void k_narrowband(const double *__restrict__ phi, long long np, double bNB, double bSB,
                             int32_t *__restrict__ nb, int32_t *__restrict__ sb)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < np; q += (long long)gridDim.x * blockDim.x) {
        const double a = fabs(phi[q]);
        nb[q] = a < bNB ? 1 : 0;
        sb[q] = a < bSB ? 1 : 0;
    }
}

void launch_narrowband(Grid *g, const double *phi, double dx, int32_t *nb, int32_t *sb)
{
    k_narrowband<<<RMS_BLOCKS, 256, 0, G.stream>>>(phi, g->np, 4.1 * dx, 8.1 * dx, nb, sb);
    G.n_launch++;
}

// =====================================================================================
// K6: pass A of a min/max iteration (set3d.f90:399-414 + secondDeriv subs.f90:384-389):
// band mask from the iteration's own "old" phi (narrowBand of the previous iteration's result,
// set3d.f90:460) and the Laplacian XX+YY+ZZ (subs.f90:461) of that old phi on band cells.
// The mixed derivatives and the weno call of pass B feed nothing that is read.
// All arithmetic is the reference's order with no FMA, so this path is bit-exact.
// =====================================================================================
template <bool MASK_GIVEN>
__global__ void __launch_bounds__(256)
k_minmax_lap(const double *__restrict__ phi, double *__restrict__ lap, uint8_t *__restrict__ mask, Dims dm,
             double bNB, double dxx, Ctrl *ctrl)
{
    if (ctrl->done) return;
    const long long np = dm.sxy * (dm.nz + 1);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < np; q += (long long)gridDim.x * blockDim.x) {
        const double p = phi[q];
        uint8_t mk;
        if (MASK_GIVEN) mk = mask[q];
        else { mk = fabs(p) < bNB ? 1 : 0; mask[q] = mk; }
        if (!mk) continue;
        const int i = (int)(q % dm.sx), j = (int)((q / dm.sx) % (dm.ny + 1)), k = (int)(q / dm.sxy);
        if (i == 0 || j == 0 || k == 0 || i == dm.nx || j == dm.ny || k == dm.nz) {
            ctrl->status = LSF_ERR_BAND_ON_BOUNDARY;   // reference would read out of bounds here
            continue;
        }
        const double m2 = __dmul_rn(-2., p);
        const double xx = __dmul_rn(__dadd_rn(__dadd_rn(m2, phi[q + 1]), phi[q - 1]), dxx);
        const double yy = __dmul_rn(__dadd_rn(__dadd_rn(m2, phi[q + dm.sx]), phi[q - dm.sx]), dxx);
        const double zz = __dmul_rn(__dadd_rn(__dadd_rn(m2, phi[q + dm.sxy]), phi[q - dm.sxy]), dxx);
        lap[q] = __dadd_rn(__dadd_rn(xx, yy), zz);
    }
}

// K7 (plane schedule): pass B, set3d.f90:417-431 + minMax subs.f90:461-481, in place.  The
// reference visits band cells in ascending (i,j,k); pAve reads the live array, so the -1
// neighbours are already updated.  Hyperplanes i+j+k = s are an exact re-ordering (+-1 star).
__global__ void __launch_bounds__(128)
k_minmax_plane(double *__restrict__ phi, const double *__restrict__ lap, const uint8_t *__restrict__ mask, Dims dm,
               int s, double h1, const Ctrl *__restrict__ ctrl)
{
    if (ctrl->done) return;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y;
    const int k = s - i - j;
    if (i > dm.nx - 1 || k < 1 || k > dm.nz - 1) return;
    const long long q = i + dm.sx * j + dm.sxy * k;
    if (!mask[q]) return;
    const double p = phi[q];
    double pAve = __dadd_rn(p, phi[q - 1]);                 // subs.f90:473, left to right
    pAve = __dadd_rn(pAve, phi[q + 1]);
    pAve = __dadd_rn(pAve, phi[q + dm.sx]);
    pAve = __dadd_rn(pAve, phi[q - dm.sx]);
    pAve = __dadd_rn(pAve, phi[q + dm.sxy]);
    pAve = __dadd_rn(pAve, phi[q - dm.sxy]);
    pAve = __ddiv_rn(pAve, 7.);
    const double curv = lap[q];
    const double F = (pAve < 0.) ? fmin_f(curv, 0.0) : fmax_f(curv, 0.0);
    phi[q] = __dadd_rn(p, __dmul_rn(h1, F));
}

void launch_minmax_iteration_plane(Grid *g, double dx, double h1, bool mask_given)
{
    const Dims &dm = g->dm;
    const double dxx = 1. / (dx * dx);
    if (mask_given) k_minmax_lap<true><<<RMS_BLOCKS, 256, 0, G.stream>>>(g->phi, g->lap, g->mask, dm, 4.1 * dx, dxx, g->ctrl);
    else k_minmax_lap<false><<<RMS_BLOCKS, 256, 0, G.stream>>>(g->phi, g->lap, g->mask, dm, 4.1 * dx, dxx, g->ctrl);
    G.n_launch++;
    dim3 blk(128), grd((dm.nx - 1 + 127) / 128, dm.ny - 1);
    for (int s = 3; s <= (dm.nx - 1) + (dm.ny - 1) + (dm.nz - 1); ++s) {
        k_minmax_plane<<<grd, blk, 0, G.stream>>>(g->phi, g->lap, g->mask, dm, s, h1, g->ctrl);
        G.n_launch++;
    }
}

__global__ void k_mask_from_i32(const int32_t *__restrict__ nb, uint8_t *__restrict__ mask, long long np)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < np; q += (long long)gridDim.x * blockDim.x)
        mask[q] = nb[q] == 1 ? 1 : 0;
}

void launch_mask_from_i32(Grid *g, const int32_t *nb)
{
    k_mask_from_i32<<<RMS_BLOCKS, 256, 0, G.stream>>>(nb, g->mask, g->np);
    G.n_launch++;
}

__global__ void k_fill(double *__restrict__ p, long long np, double v)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < np; q += (long long)gridDim.x * blockDim.x) p[q] = v;
}

// Partition-independent digest of a field: out[0] = sum over the points of bits(v) * (2*q + 1) modulo 2^64 with q the
// GLOBAL linear index of the point (odd weights: a permutation of values changes the sum), out[1] = xor of the bit
// patterns.  Integer addition is associative, so neither depends on the reduction order or on how the grid is cut
// into slabs: the digests of the ranks' slabs add / xor up to the digest of the whole grid.
template <class U>
__global__ void k_checksum(const U *__restrict__ p, long long n, long long base, unsigned long long *out)
{
    unsigned long long s = 0, x = 0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const unsigned long long v = (unsigned long long)p[q];
        s += v * (2ull * (unsigned long long)(q + base) + 1ull);
        x ^= v;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); x ^= __shfl_down_sync(0xffffffffu, x, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, s); atomicXor(out + 1, x); }
}

void launch_checksum(const void *p, size_t elem_bytes, long long n, long long first_global_index, unsigned long long *d_out)
{
    if (elem_bytes == 8) k_checksum<unsigned long long><<<RMS_BLOCKS, 256, 0, G.stream>>>((const unsigned long long *)p, n, first_global_index, d_out);
    else k_checksum<unsigned int><<<RMS_BLOCKS, 256, 0, G.stream>>>((const unsigned int *)p, n, first_global_index, d_out);
    G.n_launch++;
}

void launch_fill(Grid *g, double *p, double v)
{
    k_fill<<<RMS_BLOCKS, 256, 0, G.stream>>>(p, g->np, v);
    G.n_launch++;
}

// =====================================================================================
// K1: inside/outside sign search, set3d.f90:196-268.
//  - centroids (p1+p2+p3)/3. per triangle (:212-214)
//  - per grid point of the sub-box: first-index argmin over triangles of
//    dis = sqrt(dx^2+dy^2+dz^2) with strict '<' and minD = 100000. (:223-236).  sqrt is
//    monotone, so dis < minD implies q < q(minD): the squared distance is a safe pre-filter
//    and the sqrt + exact comparison run only on the rare improving candidates; ties after
//    rounding keep the earlier index exactly as the sequential loop does.
//  - scalar triple product in the reference's operation order (:242-258), smeared sign with
//    gM = 1 (:260-264).  No FMA anywhere: the sign field (incl. its exact zeros) is bit-exact.
// One thread per point (i fastest), triangles streamed through shared memory in tiles.
// =====================================================================================
__global__ void k_sign_centroids(const double *__restrict__ surfX, int nNode, const int32_t *__restrict__ surfElem,
                                 int nElem, double *__restrict__ cen)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nElem) return;
    const int n1 = surfElem[n] - 1, n2 = surfElem[n + nElem] - 1, n3 = surfElem[n + 2 * (long long)nElem] - 1;
    for (int c = 0; c < 3; ++c) {
        const double *X = surfX + (long long)c * nNode;
        cen[n + (long long)c * nElem] = __ddiv_rn(__dadd_rn(__dadd_rn(X[n1], X[n2]), X[n3]), 3.);
    }
}

constexpr int SIGN_TILE = 256;

__global__ void __launch_bounds__(SIGN_TILE)
k_sign_search(double *__restrict__ phi, Dims dm, double x0, double y0, double z0, double dx,
              const double *__restrict__ surfX, int nNode, const int32_t *__restrict__ surfElem, int nElem,
              const double *__restrict__ cen, int im, int jm, int km, int ni, int nj, int nk, int kbase)
{
    __shared__ double sc[3][SIGN_TILE];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npts = (long long)ni * nj * nk;
    const bool live = t < npts;
    const long long tt = live ? t : 0;
    const int i = im + (int)(tt % ni), j = jm + (int)((tt / ni) % nj), k = km + (int)(tt / ((long long)ni * nj));
    const double gX = __dadd_rn(x0, __dmul_rn((double)i, dx));       // set3d.f90:168-170
    const double gY = __dadd_rn(y0, __dmul_rn((double)j, dx));
    const double gZ = __dadd_rn(z0, __dmul_rn((double)k, dx));
    double minD = 100000., qbest = __longlong_as_double(0x7ff0000000000000LL);
    int fN = 0;
    for (int base = 0; base < nElem; base += SIGN_TILE) {
        const int n = base + threadIdx.x;
        __syncthreads();
        if (n < nElem) {
            sc[0][threadIdx.x] = cen[n];
            sc[1][threadIdx.x] = cen[n + (long long)nElem];
            sc[2][threadIdx.x] = cen[n + 2 * (long long)nElem];
        }
        __syncthreads();
        const int cnt = min(SIGN_TILE, nElem - base);
#pragma unroll 4
        for (int r = 0; r < cnt; ++r) {
            const double ex = __dsub_rn(sc[0][r], gX), ey = __dsub_rn(sc[1][r], gY), ez = __dsub_rn(sc[2][r], gZ);
            const double q = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
            if (q < qbest) {
                const double dis = __dsqrt_rn(q);
                if (dis < minD) { minD = dis; qbest = q; fN = base + r; }
            }
        }
    }
    if (!live) return;
    const int n1 = surfElem[fN] - 1, n2 = surfElem[fN + nElem] - 1, n3 = surfElem[fN + 2 * (long long)nElem] - 1;
    const double *X = surfX, *Y = surfX + nNode, *Z = surfX + 2 * (long long)nNode;
    const double A1 = __dsub_rn(X[n1], gX), A2 = __dsub_rn(Y[n1], gY), A3 = __dsub_rn(Z[n1], gZ);
    const double B1 = __dsub_rn(X[n2], gX), B2 = __dsub_rn(Y[n2], gY), B3 = __dsub_rn(Z[n2], gZ);
    const double C1 = __dsub_rn(X[n3], gX), C2 = __dsub_rn(Y[n3], gY), C3 = __dsub_rn(Z[n3], gZ);
    const double pSx = __dsub_rn(__dmul_rn(A2, B3), __dmul_rn(A3, B2));
    const double pSy = -__dsub_rn(__dmul_rn(A1, B3), __dmul_rn(B1, A3));
    const double pSz = __dsub_rn(__dmul_rn(A1, B2), __dmul_rn(B1, A2));
    const double pS = -__dadd_rn(__dadd_rn(__dmul_rn(pSx, C1), __dmul_rn(pSy, C2)), __dmul_rn(pSz, C3));
    const double den = __dsqrt_rn(__dadd_rn(__dmul_rn(pS, pS), __dmul_rn(__dmul_rn(dx, dx), 1.)));
    phi[i + dm.sx * j + dm.sxy * (k - kbase)] = __ddiv_rn(pS, den);   // k is the GLOBAL plane (z-slab: local plane k - kbase)
}

// K1 (production): the same search with exact tile-level culling of the triangle list.
// A CTA owns a block of SB_I x SB_J x SB_K grid points with centre P and half-diagonal rho.
//   pass 1: d_min = min over all centroids of |c - P| (cooperative, plain fp64);
//   pass 2: every point p of the block has its nearest centroid within |c - p| <= d_min + rho, hence
//           within |c - P| <= d_min + 2 rho: only those triangles can win or tie.  They are streamed
//           through shared memory in ASCENDING index order (ballot compaction keeps the order), and each
//           thread runs the reference's loop -- same explicitly rounded arithmetic, strict '<', first index
//           wins -- over that sub-list.  A relative margin of 1e-9 on the bound covers the rounding of the
//           filter itself and every candidate that could tie after the rounded sqrt, so the result is
//           bit-identical to the brute-force loop (tests: against k_sign_search and the oracle).
// Cost per point drops from nElem to nElem/128 + |candidates| distance evaluations.
constexpr int SB_I = 8, SB_J = 8, SB_K = 4, SB_THREADS = SB_I * SB_J * SB_K;

__global__ void __launch_bounds__(SB_THREADS)
k_sign_search_tiled(double *__restrict__ phi, Dims dm, double x0, double y0, double z0, double dx,
                    const double *__restrict__ surfX, int nNode, const int32_t *__restrict__ surfElem, int nElem,
                    const double *__restrict__ cen, int im, int jm, int km, int ni, int nj, int nk, int kbase,
                    int nbi, int nbj)
{
    __shared__ double sc[3][SB_THREADS];
    __shared__ int sidx[SB_THREADS];
    __shared__ double sred[SB_THREADS / 32];
    __shared__ int swcnt[SB_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int bi = blockIdx.x % nbi, bj = (blockIdx.x / nbi) % nbj, bk = blockIdx.x / (nbi * nbj);
    const int li = tid % SB_I, lj = (tid / SB_I) % SB_J, lk = tid / (SB_I * SB_J);
    const int i = im + bi * SB_I + li, j = jm + bj * SB_J + lj, k = km + bk * SB_K + lk;
    const bool live = (i < im + ni) && (j < jm + nj) && (k < km + nk);
    const double gX = __dadd_rn(x0, __dmul_rn((double)i, dx));       // set3d.f90:168-170
    const double gY = __dadd_rn(y0, __dmul_rn((double)j, dx));
    const double gZ = __dadd_rn(z0, __dmul_rn((double)k, dx));
    // block centre and half-diagonal (of the full block: an upper bound for clipped blocks)
    const double PX = x0 + (im + bi * SB_I + 0.5 * (SB_I - 1)) * dx;
    const double PY = y0 + (jm + bj * SB_J + 0.5 * (SB_J - 1)) * dx;
    const double PZ = z0 + (km + bk * SB_K + 0.5 * (SB_K - 1)) * dx;
    const double rho = 0.5 * dx * sqrt((double)((SB_I - 1) * (SB_I - 1) + (SB_J - 1) * (SB_J - 1) + (SB_K - 1) * (SB_K - 1)));
    // ---- pass 1: distance of the nearest centroid to the block centre -----------------------------
    double m2 = __longlong_as_double(0x7ff0000000000000LL);
    for (int n = tid; n < nElem; n += SB_THREADS) {
        const double ex = cen[n] - PX, ey = cen[n + (long long)nElem] - PY, ez = cen[n + 2 * (long long)nElem] - PZ;
        m2 = fmin(m2, ex * ex + ey * ey + ez * ez);
    }
    for (int o = 16; o > 0; o >>= 1) m2 = fmin(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    if (lane == 0) sred[wid] = m2;
    __syncthreads();
    m2 = sred[0];
    for (int w = 1; w < SB_THREADS / 32; ++w) m2 = fmin(m2, sred[w]);
    const double bound = (sqrt(m2) + 2.0 * rho) * (1.0 + 1.0e-9);
    const double bound2 = bound * bound;
    // ---- pass 2: stream the candidates (ascending index) through shared memory --------------------
    double minD = 100000., qbest = __longlong_as_double(0x7ff0000000000000LL);
    int fN = 0;
    int fill = 0;                                     // entries waiting in the shared buffer
    for (int base = 0; base < nElem; base += SB_THREADS) {
        const int n = base + tid;
        double cx = 0., cy = 0., cz = 0.;
        bool cand = false;
        if (n < nElem) {
            cx = cen[n]; cy = cen[n + (long long)nElem]; cz = cen[n + 2 * (long long)nElem];
            const double ex = cx - PX, ey = cy - PY, ez = cz - PZ;
            cand = (ex * ex + ey * ey + ez * ez) <= bound2;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) swcnt[wid] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
        for (int w = 0; w < SB_THREADS / 32; ++w) { const int cw = swcnt[w]; if (w < wid) off += cw; tot += cw; }
        // flush first if the newcomers do not fit behind what is already buffered
        if (fill + tot > SB_THREADS) {
            for (int r = 0; r < fill; ++r) {
                const double ex = __dsub_rn(sc[0][r], gX), ey = __dsub_rn(sc[1][r], gY), ez = __dsub_rn(sc[2][r], gZ);
                const double q = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
                if (q < qbest) {
                    const double dis = __dsqrt_rn(q);
                    if (dis < minD) { minD = dis; qbest = q; fN = sidx[r]; }
                }
            }
            fill = 0;
            __syncthreads();
        }
        if (cand) {
            const int pos = fill + off + __popc(bal & ((1u << lane) - 1u));
            sc[0][pos] = cx; sc[1][pos] = cy; sc[2][pos] = cz; sidx[pos] = n;
        }
        fill += tot;
        __syncthreads();
    }
    for (int r = 0; r < fill; ++r) {
        const double ex = __dsub_rn(sc[0][r], gX), ey = __dsub_rn(sc[1][r], gY), ez = __dsub_rn(sc[2][r], gZ);
        const double q = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
        if (q < qbest) {
            const double dis = __dsqrt_rn(q);
            if (dis < minD) { minD = dis; qbest = q; fN = sidx[r]; }
        }
    }
    if (!live) return;
    const int n1 = surfElem[fN] - 1, n2 = surfElem[fN + nElem] - 1, n3 = surfElem[fN + 2 * (long long)nElem] - 1;
    const double *X = surfX, *Y = surfX + nNode, *Z = surfX + 2 * (long long)nNode;
    const double A1 = __dsub_rn(X[n1], gX), A2 = __dsub_rn(Y[n1], gY), A3 = __dsub_rn(Z[n1], gZ);
    const double B1 = __dsub_rn(X[n2], gX), B2 = __dsub_rn(Y[n2], gY), B3 = __dsub_rn(Z[n2], gZ);
    const double C1 = __dsub_rn(X[n3], gX), C2 = __dsub_rn(Y[n3], gY), C3 = __dsub_rn(Z[n3], gZ);
    const double pSx = __dsub_rn(__dmul_rn(A2, B3), __dmul_rn(A3, B2));
    const double pSy = -__dsub_rn(__dmul_rn(A1, B3), __dmul_rn(B1, A3));
    const double pSz = __dsub_rn(__dmul_rn(A1, B2), __dmul_rn(B1, A2));
    const double pS = -__dadd_rn(__dadd_rn(__dmul_rn(pSx, C1), __dmul_rn(pSy, C2)), __dmul_rn(pSz, C3));
    const double den = __dsqrt_rn(__dadd_rn(__dmul_rn(pS, pS), __dmul_rn(__dmul_rn(dx, dx), 1.)));
    phi[i + dm.sx * j + dm.sxy * (k - kbase)] = __ddiv_rn(pS, den);
}

void launch_sign_init(Grid *g, const double xLo[3], double dx, const double *d_surfX, int nNode,
                      const int32_t *d_surfElem, int nElem, double *d_cen,
                      int im, int ip, int jm, int jp, int km, int kp)
{
    k_sign_centroids<<<(nElem + 255) / 256, 256, 0, G.stream>>>(d_surfX, nNode, d_surfElem, nElem, d_cen);
    G.n_launch++;
    // z-slab: this rank searches the planes of the sub-box it owns (points are independent: no collective)
    km = km > g->sg.k0 ? km : g->sg.k0;
    kp = kp < g->sg.k1 - 1 ? kp : g->sg.k1 - 1;
    if (kp < km) return;
    const int ni = ip - im + 1, nj = jp - jm + 1, nk = kp - km + 1;
    const long long npts = (long long)ni * nj * nk;
    const bool brute = getenv("LSF_SIGN_BRUTE") != nullptr;            // cross-check: the un-culled kernel
    const bool tiled = getenv("LSF_SIGN_TILED") != nullptr;            // cross-check: round-1 kernel (exact culling, linear scans)
    if (!brute && !tiled && nElem > 64 &&
        launch_sign_search_bvh(g, xLo, dx, d_surfX, nNode, d_surfElem, nElem, d_cen, im, jm, km, ni, nj, nk))
        return;                                                        // production: tree walks instead of the two linear scans
    const long long nbi = (ni + SB_I - 1) / SB_I, nbj = (nj + SB_J - 1) / SB_J, nbk = (nk + SB_K - 1) / SB_K;
    if (brute || nbi * nbj * nbk > 0x7fffffffLL)
        k_sign_search<<<(unsigned)((npts + SIGN_TILE - 1) / SIGN_TILE), SIGN_TILE, 0, G.stream>>>(
            g->phi, g->dm, xLo[0], xLo[1], xLo[2], dx, d_surfX, nNode, d_surfElem, nElem, d_cen, im, jm, km, ni, nj, nk, g->sg.kbase);
    else
        k_sign_search_tiled<<<(unsigned)(nbi * nbj * nbk), SB_THREADS, 0, G.stream>>>(
            g->phi, g->dm, xLo[0], xLo[1], xLo[2], dx, d_surfX, nNode, d_surfElem, nElem, d_cen, im, jm, km, ni, nj, nk, g->sg.kbase,
            (int)nbi, (int)nbj);
    G.n_launch++;
}

}  // namespace lsf
