// lsf_nodes.cu -- GPU side of the surface-node projection (algorithm and exactness argument: lsf_nodes.cuh).
// One thread per surface node; the nodes of a mesh are few (1e4..1e6) next to the grid, so the kernel is
// latency-bound on its gathers (8 corners x 25 stencil points per step, L2 hits) and takes microseconds
// where the reference's all-nodes re-interpolation takes minutes.
#include <stdlib.h>
#include <string.h>

#include "lsf_internal.cuh"
#include "lsf_nodes.cuh"

namespace lsf {

// PV: `const double *` (the grid lives on this GPU) or SlabView (z-slabs: gathers from the peers' slabs over NVLink)
template <class PV>
__global__ void __launch_bounds__(128)
k_advect_nodes(NodeConst c, const PV phi, const PV sbsrc, double *__restrict__ X, int nNode,
               double *__restrict__ phiSurf, double *__restrict__ gradPhiSurf, int iter, int *__restrict__ status,
               unsigned long long *__restrict__ moves)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNode) return;
    double x[3] = {X[n], X[n + (long long)nNode], X[n + 2 * (long long)nNode]};
    double ps, gs[3];
    int mv;
    const int st = node_project(c, phi, sbsrc, x, ps, gs, iter, mv);
    if (st != NODE_OK) { atomicMax(status, st); return; }
    X[n] = x[0]; X[n + (long long)nNode] = x[1]; X[n + 2 * (long long)nNode] = x[2];
    phiSurf[n] = ps;
    gradPhiSurf[n] = gs[0]; gradPhiSurf[n + (long long)nNode] = gs[1]; gradPhiSurf[n + 2 * (long long)nNode] = gs[2];
    if (mv) atomicAdd(moves, (unsigned long long)mv);
}

// the caller's phiSB as a stand-in field: abs(.) < 8.1*dx exactly where phiSB == 1
__global__ void k_sb_to_field(const int32_t *__restrict__ sb, double *__restrict__ out, long long n)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        out[q] = sb[q] == 1 ? 0. : 1.0e300;
}

// d_phi: the level set; d_sbsrc: the field whose band abs(.) < 8.1*dx is phiSB (the last narrowBand call's input)
int advect_nodes_core(Grid *g, const double *d_phi, const double *d_sbsrc, const double xLo[3], double dx, double *surfXX,
                      int nNode, double *phiSurf, double *gradPhiSurf, int iter, long long *n_moves)
{
    if (nNode < 1 || iter < 0 || !(dx > 0.)) return set_error(LSF_ERR_ARG, "advect_nodes: bad nSurfNode/iter/dx");
    NodeConst c;
    c.sx = g->dm.sx; c.sxy = g->dm.sxy; c.nx = g->dm.nx; c.ny = g->dm.ny; c.nz = sharded(g) ? g->sg.NZ : g->dm.nz;
    c.xLo[0] = xLo[0]; c.xLo[1] = xLo[1]; c.xLo[2] = xLo[2];
    c.dx = dx; c.bSB = 8.1 * dx;
    double *d_X = nullptr, *d_ps = nullptr, *d_gs = nullptr;
    int *d_st = nullptr;
    unsigned long long *d_mv = nullptr;
    const size_t nb = sizeof(double) * (size_t)nNode;
    cudaError_t e;
    if ((e = cudaMalloc(&d_X, 3 * nb)) != cudaSuccess || (e = cudaMalloc(&d_ps, nb)) != cudaSuccess ||
        (e = cudaMalloc(&d_gs, 3 * nb)) != cudaSuccess || (e = cudaMalloc(&d_st, sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&d_mv, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_X, surfXX, 3 * nb, cudaMemcpyHostToDevice, G.stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(d_st, 0, sizeof(int), G.stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(d_mv, 0, sizeof(unsigned long long), G.stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(d_ps, 0, nb, G.stream)) != cudaSuccess || (e = cudaMemsetAsync(d_gs, 0, 3 * nb, G.stream)) != cudaSuccess) {
        cudaFree(d_X); cudaFree(d_ps); cudaFree(d_gs); cudaFree(d_st); cudaFree(d_mv);
        return set_error(LSF_ERR_CUDA, "advect_nodes: %s", cudaGetErrorString(e));
    }
    G.n_launch = 0;
    cudaEventRecord(G.ev0, G.stream);
    if (sharded(g)) {
        // Every rank projects ALL nodes (they are few next to the grid and the mesh is replicated like the triangles of the sign
        // search), reading phi wherever a node goes through the peer-mapped slabs: identical results on all ranks.  The field
        // must be final on every rank before anyone reads it and must stay untouched until everyone has finished reading.
        slab_device_barrier(g);
        k_advect_nodes<SlabView><<<(nNode + 127) / 128, 128, 0, G.stream>>>(c, slab_view(g, d_phi), slab_view(g, d_sbsrc), d_X, nNode, d_ps,
                                                                            d_gs, iter, d_st, d_mv);
        slab_device_barrier(g);
    } else {
        k_advect_nodes<const double *><<<(nNode + 127) / 128, 128, 0, G.stream>>>(c, d_phi, d_sbsrc, d_X, nNode, d_ps, d_gs, iter, d_st, d_mv);
    }
    G.n_launch++;
    cudaEventRecord(G.ev1, G.stream);
    int st = 0;
    unsigned long long mv = 0;
    e = cudaMemcpyAsync(&st, d_st, sizeof(int), cudaMemcpyDeviceToHost, G.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&mv, d_mv, sizeof(mv), cudaMemcpyDeviceToHost, G.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
    if (e == cudaSuccess && st == NODE_OK) {
        e = cudaMemcpy(surfXX, d_X, 3 * nb, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(phiSurf, d_ps, nb, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(gradPhiSurf, d_gs, 3 * nb, cudaMemcpyDeviceToHost);
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, G.ev0, G.ev1) == cudaSuccess) G.last_ms = ms;
    cudaFree(d_X); cudaFree(d_ps); cudaFree(d_gs); cudaFree(d_st); cudaFree(d_mv);
    if (e != cudaSuccess) return set_error(LSF_ERR_CUDA, "advect_nodes: %s", cudaGetErrorString(e));
    if (st == NODE_OFF_GRID)
        return set_error(LSF_ERR_NODE_OFF_GRID, "advect_nodes: a surface node lies outside the grid (the reference reads out of bounds, subs.f90:1104)");
    if (st == NODE_BAND_ON_BOUNDARY)
        return set_error(LSF_ERR_BAND_ON_BOUNDARY, "advect_nodes: a stencil-band point next to a node is closer than 4 points to the grid boundary (subs.f90:341)");
    if (n_moves) *n_moves = (long long)mv;
    return LSF_OK;
}

}  // namespace lsf

using namespace lsf;

extern "C" {

int lsf_grid_advect_nodes(lsf_grid *g, const double xLo[3], double dx, double *surfXX, int nSurfNode,
                          double *phiSurf, double *gradPhiSurf, int iter, long long *n_moves)
{
    if (!g || !xLo || !surfXX || !phiSurf || !gradPhiSurf) return set_error(LSF_ERR_ARG, "null argument");
    { const int rc0 = slab_check_attached(g); if (rc0) return rc0; }
    if (g->f32) {                                                       // fp64 evaluation on a transient widened copy of phi
        lsf_grid *sh = nullptr;
        int rc = f32_shadow_open(g, &sh);
        if (rc) return rc;
        rc = advect_nodes_core(sh, sh->phi, sh->phi, xLo, dx, surfXX, nSurfNode, phiSurf, gradPhiSurf, iter, n_moves);
        const int rc2 = f32_shadow_close(g, sh, false);
        return rc ? rc : rc2;
    }
    return advect_nodes_core(g, g->phi, g->sb_from_phiN ? g->phiN : g->phi, xLo, dx, surfXX, nSurfNode, phiSurf, gradPhiSurf, iter, n_moves);
}

int lsf_advect_nodes(const double *phi, const int32_t *phiSB, int nx, int ny, int nz, const double xLo[3], double dx,
                     double *surfXX, int nSurfNode, double *phiSurf, double *gradPhiSurf, int iter, long long *n_moves)
{
    if (!phi || !phiSB || !xLo || !surfXX || !phiSurf || !gradPhiSurf) return set_error(LSF_ERR_ARG, "null argument");
    if (!G.inited) { int rc0 = lsf_init(-1); if (rc0) return rc0; }
    lsf_grid *g = nullptr;
    int rc = lsf_grid_create(&g, nx, ny, nz);
    if (rc) return rc;
    // the caller's phiSB decides band membership: phiS becomes a stand-in field with abs(.) < 8.1*dx exactly on the band
    // (masks go up in chunks through a small staging buffer)
    const long long np = g->np, cap = np < (1LL << 25) ? np : (1LL << 25);
    int32_t *stage = nullptr;
    cudaError_t e = cudaMalloc(&stage, sizeof(int32_t) * (size_t)cap);
    for (long long o = 0; o < np && e == cudaSuccess; o += cap) {
        const long long n = np - o < cap ? np - o : cap;
        e = cudaMemcpyAsync(stage, phiSB + o, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, G.stream);
        k_sb_to_field<<<RMS_BLOCKS, 256, 0, G.stream>>>(stage, g->phiS + o, n);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
    cudaFree(stage);
    if (e != cudaSuccess) rc = set_error(LSF_ERR_CUDA, "advect_nodes: %s", cudaGetErrorString(e));
    if (!rc) rc = lsf_grid_upload(g, phi);
    if (!rc) rc = advect_nodes_core(g, g->phi, g->phiS, xLo, dx, surfXX, nSurfNode, phiSurf, gradPhiSurf, iter, n_moves);
    lsf_grid_destroy(g);
    return rc;
}

}  // extern "C"
