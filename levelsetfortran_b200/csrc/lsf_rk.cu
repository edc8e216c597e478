// lsf_rk.cu -- K2' (SURVEY.md section 2): the north_star's LITERAL reinitialisation scheme -- Jacobi WENO5 right-hand side with
// TVD Runge-Kutta-3 pseudo-time stages -- as a separately reported throughput mode.
//
// THIS IS NOT THE REFERENCE'S ALGORITHM.  The reference (subs.f90:717-931) is forward Euler with in-place Gauss-Seidel
// sweeps in 8 rotating raster orders; its RK scaffolding (phi1/phi2/phi3, k1..k4, set3d.f90:28,285-287) is allocated and
// never used.  A Jacobi update differs from the reference by ~4e-4 and exits at a different iteration,
// so nothing computed here can claim reference parity; it is checked against a Jacobi/RK3 restatement built from the
// oracle's own per-cell `weno` / `phiSign` (tests/test_gpu_rk.py).  What it shares with the parity path: the per-cell
// arithmetic (lsf_cell.cuh: ExactArith / FastArith), the high-order window of subs.f90:506, the extrapolation boundary
// block (subs.f90:858-897, applied after every stage) and the RMS / EXIT / NaN tests (subs.f90:902-926, per RK step).
//
//   phi1 = phi + dt L(phi);   phi2 = 3/4 phi + 1/4 (phi1 + dt L(phi1));   phi <- 1/3 phi + 2/3 (phi2 + dt L(phi2))
//   L(u) = sgn(phiS) (1 - |grad u|_Godunov-WENO5),  phiS = phi at entry (the frozen sign source, subs.f90:731)
//
// Kernel: no data dependence between cells of a stage, so the schedule is the plain one -- a 32 x 8 thread tile in
// (x, y) marches along z with the 7 z-stencil values in a register queue; the x / y neighbours are read through L1
// (the x line of a warp is one 256-B row; the y lines are the rows of the tile's other warps).  One read of u and phiS
// and one write per cell and stage reach HBM (24 B, 32 B in stages 2 and 3 which also read phi).  fp64 WENO5 stays
// FP64-pipe bound here as in the sweep kernel (about 260 FP64 instructions per cell, lsf_cell.cuh).
#include "lsf_internal.cuh"
#include "lsf_cell.cuh"
#include "lsf_march.cuh"      // wait_ge (system-scope flag wait)

namespace lsf {

#ifndef LSF_RK_ZC
#define LSF_RK_ZC 64
#endif
constexpr int RK_TX = 32, RK_TY = 8, RK_ZC = LSF_RK_ZC;   // planes a block marches through (its z-queue warm-up re-reads 6 planes)

#ifndef LSF_RK_OCC
#define LSF_RK_OCC 4            // resident CTAs per SM the stage kernel is compiled for (64 registers, 64 B of spills).  Session 31, Gcell-stage-updates/s at
                                // 1024^3: 2 CTAs (126 registers) 38.9, 3 (79) 46.3, 4 47.8 -- the kernel waits on its in-plane loads (ncu r2l: long scoreboard
                                // 3.5 per issue at 16 warps per SM), so more resident warps pay
#endif
template <class AR>
__global__ void __launch_bounds__(RK_TX *RK_TY, LSF_RK_OCC)
k_rk_stage(const double *in, const double *phin, const double *__restrict__ phiS, double *out, Dims dm, CellConst cc,
           double a, double b, double *__restrict__ partial, const Ctrl *__restrict__ ctrl, int want_rms,
           int kA, int kB, int kbase, int NZ,       // z-slab: local planes kA..kB are updated, local plane k is global plane k + kbase of 0..NZ
           const long long *halo_seq, long long need0, long long need1, Ctrl *ctrl_w)   // z-slab: `in`'s ghost planes must have arrived
{
    if (ctrl->done) return;
    if (halo_seq) {
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            if (need0) wait_ge<true>(halo_seq + 0, need0, ctrl_w);
            if (need1) wait_ge<true>(halo_seq + 1, need1, ctrl_w);
        }
        __syncthreads();
    }
    const int i = 1 + blockIdx.x * RK_TX + threadIdx.x;
    const int j = 1 + blockIdx.y * RK_TY + threadIdx.y;
    const int k0 = kA + blockIdx.z * RK_ZC;
    const int k1 = min(k0 + RK_ZC - 1, kB);
    double acc = 0.;
    if (i <= dm.nx - 1 && j <= dm.ny - 1) {
        const bool hij = (i > 3) && (i < dm.nx - 4) && (j > 3) && (j < dm.ny - 4);      // subs.f90:506
        const long long base = i + dm.sx * j;
        double q[7];
#pragma unroll
        for (int m = -3; m <= 3; ++m) {
            const int kk = min(max(k0 + m, 0), dm.nz);
            q[m + 3] = __ldg(in + base + dm.sxy * kk);
        }
        for (int k = k0; k <= k1; ++k) {
            const long long c = base + dm.sxy * k;
            double vx[7], vy[7];
            if (hij) {
#pragma unroll
                for (int m = -3; m <= 3; ++m) {
                    vx[m + 3] = (m == 0) ? q[3] : __ldg(in + c + m);
                    vy[m + 3] = (m == 0) ? q[3] : __ldg(in + c + m * dm.sx);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 7; ++m) { vx[m] = 0.; vy[m] = 0.; }
                vx[2] = __ldg(in + c - 1); vx[3] = q[3]; vx[4] = __ldg(in + c + 1);
                vy[2] = __ldg(in + c - dm.sx); vy[3] = q[3]; vy[4] = __ldg(in + c + dm.sx);
            }
            const bool hi = hij && (k + kbase > 3) && (k + kbase < NZ - 4);
            double g[3], gM;
            bool sens;
            const double e = reinit_cell<AR>(vx, vy, q, __ldg(phiS + c), hi, cc, g, gM, sens);      // u + dt sgn (1 - |grad u|)
            const double pold = (a != 0. || want_rms) ? phin[c] : 0.;
            const double o = (a != 0.) ? fma(a, pold, b * e) : e;
            out[c] = o;
            if (want_rms) { const double d = o - pold; acc = fma(d, d, acc); }
#pragma unroll
            for (int m = 0; m < 6; ++m) q[m] = q[m + 1];
            q[6] = __ldg(in + base + dm.sxy * min(k + 4, dm.nz));
        }
    }
    if (want_rms) {
        __shared__ double sh[RK_TX * RK_TY];
        const int t = threadIdx.x + RK_TX * threadIdx.y;
        sh[t] = acc;
        __syncthreads();
        for (int w = RK_TX * RK_TY / 2; w > 0; w >>= 1) {
            if (t < w) sh[t] = sh[t] + sh[t + w];
            __syncthreads();
        }
        if (t == 0) partial[blockIdx.x + gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z)] = sh[0];
    }
}

// The same stage with the in-plane neighbours through shared memory (-DLSF_RK_TILE=1; measured alternative, off): per plane every thread deposits the
// value of its own column (it sits in the z-queue already) into a (TY+6) x (TX+6) tile, the 3-wide ring around the tile is fetched
// from global memory ONE PLANE AHEAD into two registers per thread, and the 12 x / y stencil values are shared-memory loads --
// instead of 12 L1 loads per cell whose latency the few resident warps cannot hide (ncu r2l: long scoreboard 3.5 per issue).
// Two tiles alternate, so one __syncthreads per plane.  Same values, same arithmetic: bit-identical to k_rk_stage.
// Session 33, 1024^3: 42.3 Gcell-stage-updates/s (4 CTAs/SM; 42.1 at 3) against 47.8 for the L1 version -- the deposits and the
// barrier per plane cost more than the L1 hits they replace; the L1 version with 4 resident CTAs stays the default.
template <class AR>
__global__ void __launch_bounds__(RK_TX *RK_TY, LSF_RK_OCC)
k_rk_stage_tile(const double *in, const double *phin, const double *__restrict__ phiS, double *out, Dims dm, CellConst cc,
                double a, double b, double *__restrict__ partial, const Ctrl *__restrict__ ctrl, int want_rms,
                int kA, int kB, int kbase, int NZ, const long long *halo_seq, long long need0, long long need1, Ctrl *ctrl_w)
{
    if (ctrl->done) return;
    constexpr int SW = RK_TX + 6, SH = RK_TY + 6, NT = RK_TX * RK_TY, NH = SW * SH - NT;   // tile pitch / rows, threads, ring cells
    constexpr int HPT = (NH + NT - 1) / NT;                                                  // ring cells per thread (2)
    __shared__ double tile[2][SH][SW];
    __shared__ double sh[NT];
    const int t = threadIdx.x + RK_TX * threadIdx.y;
    if (halo_seq) {
        if (t == 0) {
            if (need0) wait_ge<true>(halo_seq + 0, need0, ctrl_w);
            if (need1) wait_ge<true>(halo_seq + 1, need1, ctrl_w);
        }
        __syncthreads();
    }
    const int i0 = 1 + blockIdx.x * RK_TX, j0 = 1 + blockIdx.y * RK_TY;
    const int i = i0 + threadIdx.x, j = j0 + threadIdx.y;
    const int k0 = kA + blockIdx.z * RK_ZC;
    const int k1 = min(k0 + RK_ZC - 1, kB);
    const bool inArr = (i <= dm.nx) && (j <= dm.ny);                 // the column exists (its values feed the neighbours)
    const bool comp = (i <= dm.nx - 1) && (j <= dm.ny - 1);          // ... and is updated
    const bool hij = (i > 3) && (i < dm.nx - 4) && (j > 3) && (j < dm.ny - 4);      // subs.f90:506
    const long long base = i + dm.sx * j;
    // ring cells of this thread: index h in 0..NH-1 enumerates the tile's cells that are not one of the NT centres, row-major
    long long hoff[HPT];
    int hrow[HPT], hcol[HPT];
    bool hok[HPT];
#pragma unroll
    for (int r = 0; r < HPT; ++r) {
        const int h = t + r * NT;
        int row = 0, col = 0;
        bool ok = h < NH;
        if (ok) {
            // rows 0..2 and SH-3..SH-1 are whole ring rows (SW cells each); the middle rows contribute 3 cells left and 3 right
            if (h < 3 * SW) { row = h / SW; col = h % SW; }
            else if (h < 3 * SW + RK_TY * 6) { const int q2 = h - 3 * SW; row = 3 + q2 / 6; const int c6 = q2 % 6; col = c6 < 3 ? c6 : RK_TX + c6; }
            else { const int q2 = h - 3 * SW - RK_TY * 6; row = 3 + RK_TY + q2 / SW; col = q2 % SW; }
            const int gi = i0 - 3 + col, gj = j0 - 3 + row;
            ok = gi >= 0 && gi <= dm.nx && gj >= 0 && gj <= dm.ny;
            hoff[r] = gi + dm.sx * (long long)gj;
        } else hoff[r] = 0;
        hrow[r] = row; hcol[r] = col; hok[r] = ok;
    }
    double acc = 0.;
    double q[7];
#pragma unroll
    for (int m = -3; m <= 3; ++m) {
        const int kk = min(max(k0 + m, 0), dm.nz);
        q[m + 3] = inArr ? __ldg(in + base + dm.sxy * kk) : 0.;
    }
    double hv[HPT];
#pragma unroll
    for (int r = 0; r < HPT; ++r) hv[r] = hok[r] ? __ldg(in + hoff[r] + dm.sxy * k0) : 0.;
    for (int k = k0; k <= k1; ++k) {
        double(*T)[SW] = tile[(k - k0) & 1];
        T[threadIdx.y + 3][threadIdx.x + 3] = q[3];
#pragma unroll
        for (int r = 0; r < HPT; ++r)
            if (t + r * NT < NH) T[hrow[r]][hcol[r]] = hv[r];
        __syncthreads();
        if (k < k1) {                                                 // the ring of the next plane: in flight during this plane's arithmetic
#pragma unroll
            for (int r = 0; r < HPT; ++r) hv[r] = hok[r] ? __ldg(in + hoff[r] + dm.sxy * (k + 1)) : 0.;
        }
        if (comp) {
            const long long c = base + dm.sxy * k;
            double vx[7], vy[7];
            if (hij) {
#pragma unroll
                for (int m = -3; m <= 3; ++m) {
                    vx[m + 3] = (m == 0) ? q[3] : T[threadIdx.y + 3][threadIdx.x + 3 + m];
                    vy[m + 3] = (m == 0) ? q[3] : T[threadIdx.y + 3 + m][threadIdx.x + 3];
                }
            } else {
#pragma unroll
                for (int m = 0; m < 7; ++m) { vx[m] = 0.; vy[m] = 0.; }
                vx[2] = T[threadIdx.y + 3][threadIdx.x + 2]; vx[3] = q[3]; vx[4] = T[threadIdx.y + 3][threadIdx.x + 4];
                vy[2] = T[threadIdx.y + 2][threadIdx.x + 3]; vy[3] = q[3]; vy[4] = T[threadIdx.y + 4][threadIdx.x + 3];
            }
            const bool hi = hij && (k + kbase > 3) && (k + kbase < NZ - 4);
            double g[3], gM;
            bool sens;
            const double e = reinit_cell<AR>(vx, vy, q, __ldg(phiS + c), hi, cc, g, gM, sens);
            const double pold = (a != 0. || want_rms) ? phin[c] : 0.;
            const double o = (a != 0.) ? fma(a, pold, b * e) : e;
            out[c] = o;
            if (want_rms) { const double d = o - pold; acc = fma(d, d, acc); }
        }
#pragma unroll
        for (int m = 0; m < 6; ++m) q[m] = q[m + 1];
        q[6] = inArr ? __ldg(in + base + dm.sxy * min(k + 4, dm.nz)) : 0.;
    }
    if (want_rms) {
        __syncthreads();
        sh[t] = acc;
        __syncthreads();
        for (int w = NT / 2; w > 0; w >>= 1) {
            if (t < w) sh[t] = sh[t] + sh[t + w];
            __syncthreads();
        }
        if (t == 0) partial[blockIdx.x + gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z)] = sh[0];
    }
}

// fixed-order sum of the stage's per-block partials into ONE partial slot (k_finalize then adds the boundary part)
__global__ void k_rk_sum(const double *__restrict__ partial, long long n, double *__restrict__ out, const Ctrl *__restrict__ ctrl)
{
    if (ctrl->done) return;
    __shared__ double sh[1024];
    double acc = 0.;
    for (long long q = threadIdx.x; q < n; q += 1024) acc += partial[q];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

long long rk_nblocks(const Grid *g)
{
    const long long nk = g->sg.kupd_hi - g->sg.kupd_lo + 1;
    const long long bx = (g->dm.nx - 1 + RK_TX - 1) / RK_TX, by = (g->dm.ny - 1 + RK_TY - 1) / RK_TY, bz = (nk + RK_ZC - 1) / RK_ZC;
    return bx * by * bz;
}

// one stage: out = a*phin + b*(in + dt L(in)); with want_rms the sum over the interior of (out - phin)^2 lands in rms_out[0]
// hg: the sharded grid whose last ghost-plane exchange refreshed `in` (g itself, or the transient grid that owns phi2); null on one GPU
void launch_rk_stage(Grid *g, const double *in, const double *phin, double *out, const CellConst &cc, double a, double b,
                     double *scratch_partial, double *rms_out, Grid *hg)
{
    const long long *hs = (hg && sharded(hg)) ? hg->sync->halo_seq : nullptr;
    const long long n0 = (hs && hg->sg.rank > 0) ? hg->phase : 0, n1 = (hs && hg->sg.rank < hg->sg.nranks - 1) ? hg->phase : 0;
    const SlabGeom &sg = g->sg;      // one GPU: kupd_lo = 1, kupd_hi = nz-1, kbase = 0, NZ = nz
    dim3 grid((g->dm.nx - 1 + RK_TX - 1) / RK_TX, (g->dm.ny - 1 + RK_TY - 1) / RK_TY, (sg.kupd_hi - sg.kupd_lo + 1 + RK_ZC - 1) / RK_ZC);
    dim3 block(RK_TX, RK_TY);
    const int want = rms_out != nullptr;
#ifndef LSF_RK_TILE
#define LSF_RK_TILE 0
#endif
    if (LSF_RK_TILE) {
        if (G.arith_run == LSF_ARITH_EXACT)
            k_rk_stage_tile<ExactArith><<<grid, block, 0, G.stream>>>(in, phin, g->phiS, out, g->dm, cc, a, b, scratch_partial, g->ctrl, want,
                                                                      sg.kupd_lo, sg.kupd_hi, sg.kbase, sg.NZ, hs, n0, n1, g->ctrl);
        else
            k_rk_stage_tile<FastArith><<<grid, block, 0, G.stream>>>(in, phin, g->phiS, out, g->dm, cc, a, b, scratch_partial, g->ctrl, want,
                                                                     sg.kupd_lo, sg.kupd_hi, sg.kbase, sg.NZ, hs, n0, n1, g->ctrl);
    } else if (G.arith_run == LSF_ARITH_EXACT)
        k_rk_stage<ExactArith><<<grid, block, 0, G.stream>>>(in, phin, g->phiS, out, g->dm, cc, a, b, scratch_partial, g->ctrl, want,
                                                             sg.kupd_lo, sg.kupd_hi, sg.kbase, sg.NZ, hs, n0, n1, g->ctrl);
    else
        k_rk_stage<FastArith><<<grid, block, 0, G.stream>>>(in, phin, g->phiS, out, g->dm, cc, a, b, scratch_partial, g->ctrl, want,
                                                            sg.kupd_lo, sg.kupd_hi, sg.kbase, sg.NZ, hs, n0, n1, g->ctrl);
    G.n_launch++;
    if (want) {
        k_rk_sum<<<1, 1024, 0, G.stream>>>(scratch_partial, rk_nblocks(g), rms_out, g->ctrl);
        G.n_launch++;
    }
}

}  // namespace lsf
