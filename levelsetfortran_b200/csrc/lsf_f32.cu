// lsf_f32.cu -- the optional single-precision mode (contract: 1e-4 relative to the fp64
// reference, tests/test_gpu_f32.py).
//
// An fp32 grid (lsf_grid_create_f32) keeps phi / phiS as float arrays in the reference's dense layout:
// 12 B per cell update instead of 24, and the WENO5 sweep runs on the FP32 pipe (F32Arith, lsf_cell.cuh)
// with the very same Gauss-Seidel column-tile schedule as the fp64 path (lsf_march.cuh is templated on the
// element type).  Host arrays stay REAL(8): upload / download convert on the device.
//   reinit      : native fp32 (sweep kernel, boundary block, fused RMS); RMS sums are accumulated in fp64
//   narrowBand  : the reference's fp64 comparison on the widened value
//   sign search and min/max flow are <1 % of the work of a run: they execute on a transient fp64 shadow grid
//   (the fp64 kernels, results rounded to fp32) so that the sign field stays the fp64 one and the active-list
//   exactness argument of lsf_mm_list.cuh is untouched.
// Sharded fp32 grids (lsf_sgrid_create_f32) run fill / upload / download / sign search / reinit / narrowBand: the
// sweep kernel, the ghost-plane exchange and the loop control are the fp64 path's, instantiated for float.
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lsf_internal.cuh"

namespace lsf {

__global__ void k_d2f(const double *__restrict__ src, float *__restrict__ dst, long long n)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        dst[q] = (float)src[q];
}

__global__ void k_f2d(const float *__restrict__ src, double *__restrict__ dst, long long n)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        dst[q] = (double)src[q];
}

__global__ void k_fill_f(float *__restrict__ p, long long np, float v)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < np; q += (long long)gridDim.x * blockDim.x) p[q] = v;
}

// Boundary block of reinit (closed form) + boundary part of the RMS sum: the fp32 twin of k_reinit_bc_rms
// (lsf_kernels.cu), including its z-slab form (a rank visits the boundary points of its owned planes only).
__global__ void __launch_bounds__(256)
k_reinit_bc_rms_f32(float *__restrict__ phi, Dims dm, float dx, double *__restrict__ partial, const Ctrl *__restrict__ ctrl,
                    int kA, int kB, int kbase, int NZ, int hasLo, int hasHi)
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
        float v = phi[ci + dm.sx * cj + dm.sxy * ck];
        for (int r = 0; r < m; ++r) v = __fadd_rn(v, dx);
        const long long q = i + dm.sx * j + dm.sxy * k;
        const double d = (double)v - (double)phi[q];
        acc += d * d;
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

void launch_reinit_bc_rms_f32(Grid *g, double dx, int partial_off)
{
    const SlabGeom &sg = g->sg;
    k_reinit_bc_rms_f32<<<BC_BLOCKS, 256, 0, G.stream>>>(g->phi_f, g->dm, (float)dx, g->partial + partial_off, g->ctrl, sg.kupd_lo,
                                                         sg.kupd_hi, sg.kbase, sg.NZ, sg.k0 == 0, sg.k1 == sg.NZ + 1);
    G.n_launch++;
}

__global__ void k_narrowband_f32(const float *__restrict__ phi, long long np, double bNB, double bSB,
                                 int32_t *__restrict__ nb, int32_t *__restrict__ sb)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < np; q += (long long)gridDim.x * blockDim.x) {
        const double a = fabs((double)phi[q]);
        nb[q] = a < bNB ? 1 : 0;
        sb[q] = a < bSB ? 1 : 0;
    }
}

static void convert_d2f(const double *src, float *dst, long long n)
{
    k_d2f<<<RMS_BLOCKS, 256, 0, G.stream>>>(src, dst, n);
    G.n_launch++;
}

static void convert_f2d(const float *src, double *dst, long long n)
{
    k_f2d<<<RMS_BLOCKS, 256, 0, G.stream>>>(src, dst, n);
    G.n_launch++;
}

// host REAL(8) <-> device float through a device staging buffer, in chunks (the copy of chunk c+1 overlaps
// nothing: simple and bounded in memory; PCIe is the limit either way)
constexpr long long F32_STAGE = 1LL << 25;   // elements per chunk (256 MB of doubles)

// LSF_F32_IO_LOG=1: wall time of the conversions on stderr (diagnostics)
static double io_now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static bool io_log() { static const bool on = getenv("LSF_F32_IO_LOG") != nullptr; return on; }

int f32_upload(Grid *g, const double *host, float *dev, long long np)
{
    (void)g;
    const double t_io = io_now();
    double *stage = nullptr;
    const long long cap = np < F32_STAGE ? np : F32_STAGE;
    LSF_CUDA(cudaMalloc(&stage, sizeof(double) * (size_t)cap));
    cudaError_t e = cudaSuccess;
    for (long long o = 0; o < np && e == cudaSuccess; o += cap) {
        const long long n = np - o < cap ? np - o : cap;
        e = cudaMemcpyAsync(stage, host + o, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, G.stream);
        convert_d2f(stage, dev + o, n);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
    const double t_sync = io_now();
    cudaFree(stage);
    if (io_log()) fprintf(stderr, "[lsf f32] upload %.1f MB: %.3f s to sync (%.1f GB/s), %.3f s cudaFree\n", np * 8e-6, t_sync - t_io,
                          np * 8e-9 / (t_sync - t_io), io_now() - t_sync);
    if (e != cudaSuccess) return set_error(LSF_ERR_CUDA, "f32 upload: %s", cudaGetErrorString(e));
    return LSF_OK;
}

int f32_download(Grid *g, const float *dev, double *host, long long np)
{
    (void)g;
    const double t_io = io_now();
    double *stage = nullptr;
    const long long cap = np < F32_STAGE ? np : F32_STAGE;
    LSF_CUDA(cudaMalloc(&stage, sizeof(double) * (size_t)cap));
    cudaError_t e = cudaSuccess;
    for (long long o = 0; o < np && e == cudaSuccess; o += cap) {
        const long long n = np - o < cap ? np - o : cap;
        convert_f2d(dev + o, stage, n);
        e = cudaMemcpyAsync(host + o, stage, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, G.stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
    if (io_log()) fprintf(stderr, "[lsf f32] download %.1f MB: %.3f s (%.1f GB/s)\n", np * 8e-6, io_now() - t_io, np * 8e-9 / (io_now() - t_io));
    cudaFree(stage);
    if (e != cudaSuccess) return set_error(LSF_ERR_CUDA, "f32 download: %s", cudaGetErrorString(e));
    return LSF_OK;
}

int f32_fill(Grid *g, double value)
{
    k_fill_f<<<RMS_BLOCKS, 256, 0, G.stream>>>(g->phi_f, g->np, (float)value);
    G.n_launch++;
    LSF_CUDA(cudaStreamSynchronize(G.stream));
    return LSF_OK;
}

int f32_narrowband(Grid *g, double dx, int32_t *d_nb, int32_t *d_sb)
{
    k_narrowband_f32<<<RMS_BLOCKS, 256, 0, G.stream>>>(g->phi_f, g->np, 4.1 * dx, 8.1 * dx, d_nb, d_sb);
    G.n_launch++;
    return LSF_OK;
}

// reinit, subs.f90:717-931, in fp32: the march schedule with the fused boundary block / RMS, device-side
// loop control exactly as on the fp64 path (lsf_api.cu: reinit_attempt).
int f32_reinit(Grid *g, int iter, double dx, double h, double tol, int *n_exit, double *rms_hist)
{
    if (iter < 0 || !(dx > 0.)) return set_error(LSF_ERR_ARG, "reinit: bad iter/dx");
    if (g->dm.nx < 2 || g->dm.ny < 2 || g->dm.nz < 2) return set_error(LSF_ERR_ARG, "reinit: grid too small");
    if (g->hist_cap < iter + 1) {
        cudaFree(g->hist);
        g->hist = nullptr; g->hist_cap = 0;
        const int cap = iter + 1 < 16384 ? 16384 : iter + 1;
        LSF_CUDA(cudaMalloc(&g->hist, sizeof(double) * (size_t)cap));
        g->hist_cap = cap;
    }
    LSF_CUDA(cudaMemcpyAsync(g->phiS_f, g->phi_f, sizeof(float) * (size_t)g->np, cudaMemcpyDeviceToDevice, G.stream));   // subs.f90:731
    LSF_CUDA(cudaMemsetAsync(g->ctrl, 0, sizeof(Ctrl), G.stream));
    CellConst cc;
    cc.dx = dx; cc.inv_dx = 1. / dx; cc.k12 = 1. / (12. * dx); cc.dx2 = dx * dx; cc.h = h;
    int rc = march_prepare(g);
    if (rc) return rc;
    G.n_launch = 0;
    G.sweep_ms = 0.; G.n_sweeps = 0;
    LSF_CUDA(cudaEventRecord(G.ev0, G.stream));
    static cudaEvent_t pe[16][2];
    static bool pe_init = false;
    if (G.profile && !pe_init) { for (int q = 0; q < 16; ++q) { cudaEventCreate(&pe[q][0]); cudaEventCreate(&pe[q][1]); } pe_init = true; }
    int npend = 0;
    Ctrl hc = {0, 0, 0, 0, 0};
    const int ntiles = march_ntiles(g);
    for (int n = 0; n <= iter; ++n) {                                   // subs.f90:735
        const int raster = n % 8 + 1;                                   // subs.f90:740,855
        if (G.profile) cudaEventRecord(pe[npend][0], G.stream);
        launch_reinit_sweep_march_f32(g, raster, cc);
        if (G.profile) { cudaEventRecord(pe[npend][1], G.stream); ++npend; }
        launch_reinit_bc_rms_f32(g, dx, ntiles);                        // :858-897
        launch_finalize(g, ntiles + BC_BLOCKS, 0, tol);                 // :914-926
        if ((n + 1) % 8 == 0 || n == iter) {
            LSF_CUDA(cudaMemcpyAsync(&hc, g->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, G.stream));
            LSF_CUDA(cudaStreamSynchronize(G.stream));
            for (int q = 0; q < npend; ++q) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, pe[q][0], pe[q][1]) == cudaSuccess) { G.sweep_ms += ms; G.n_sweeps++; }
            }
            npend = 0;
            if (hc.done) break;
        }
    }
    LSF_CUDA(cudaGetLastError());
    const int ne = hc.done ? hc.n_exit : iter;
    if (G.profile && G.n_sweeps > ne + 1) G.n_sweeps = ne + 1;
    if (n_exit) *n_exit = ne;
    if (rms_hist) LSF_CUDA(cudaMemcpy(rms_hist, g->hist, sizeof(double) * (size_t)(ne + 1), cudaMemcpyDeviceToHost));
    LSF_CUDA(cudaEventRecord(G.ev1, G.stream));
    LSF_CUDA(cudaEventSynchronize(G.ev1));
    float ms = 0.f;
    LSF_CUDA(cudaEventElapsedTime(&ms, G.ev0, G.ev1));
    G.last_ms = ms;
    G.arith_last = LSF_ARITH_FAST;
    return hc.done ? hc.status : LSF_OK;
}

// Sign search on an fp32 grid (whole or z-slab): the fp64 kernels write into one temporary fp64 field holding the
// widened phi, the result is rounded to fp32 -- so the sign field is the fp64 one, exact zeros and -0.0 included.
int f32_sign_init(Grid *g, const double xLo[3], double dx, const double *d_surfX, int nNode, const int32_t *d_surfElem, int nElem,
                  double *d_cen, int im, int ip, int jm, int jp, int km, int kp)
{
    double *tmp = nullptr;
    LSF_CUDA(cudaMalloc(&tmp, sizeof(double) * (size_t)g->np));
    convert_f2d(g->phi_f, tmp, g->np);
    Grid view = *g;                       // same extents and slab geometry, phi -> the temporary
    view.f32 = 0;
    view.phi = tmp;
    launch_sign_init(&view, xLo, dx, d_surfX, nNode, d_surfElem, nElem, d_cen, im, ip, jm, jp, km, kp);
    G.n_launch += 0;
    convert_d2f(tmp, g->phi_f, g->np);
    cudaError_t e = cudaStreamSynchronize(G.stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return set_error(LSF_ERR_CUDA, "sign_init (fp32 grid): %s", cudaGetErrorString(e));
    return LSF_OK;
}

// transient fp64 shadow of an fp32 grid: phi widened on creation, rounded back by f32_shadow_close.  On z-slabs the shadow
// is a sharded fp64 grid of the same geometry that all ranks create and connect together (lsf_slab.cu: sgrid_shadow_f64), so
// every multi-GPU stage of the fp64 path -- min/max flow with its ghost-plane exchange, node projection -- runs on it unchanged.
int f32_shadow_open(Grid *g, lsf_grid **shadow)
{
    int rc = sharded(g) ? sgrid_shadow_f64(g, shadow) : lsf_grid_create(shadow, g->dm.nx, g->dm.ny, g->dm.nz);
    if (rc) return rc;
    convert_f2d(g->phi_f, (*shadow)->phi, g->np);          // ghost planes included: they hold the neighbours' current values
    return LSF_OK;
}

int f32_shadow_close(Grid *g, lsf_grid *shadow, bool write_back)
{
    if (write_back) convert_d2f(shadow->phi, g->phi_f, g->np);
    cudaError_t e = cudaStreamSynchronize(G.stream);
    int rc = LSF_OK;
    if (sharded(g)) rc = sgrid_shadow_release(g, shadow); else lsf_grid_destroy(shadow);
    if (e != cudaSuccess) return set_error(LSF_ERR_CUDA, "f32 shadow: %s", cudaGetErrorString(e));
    return rc;
}

}  // namespace lsf

using namespace lsf;

extern "C" int lsf_grid_create_f32(lsf_grid **out, int nx, int ny, int nz)
{
    if (!out) return set_error(LSF_ERR_ARG, "null handle");
    *out = nullptr;
    if (!G.inited) { int rc = lsf_init(-1); if (rc) return rc; }
    if (nx < 1 || ny < 1 || nz < 1) return set_error(LSF_ERR_ARG, "grid extents must be >= 1");
    Grid *g = (Grid *)calloc(1, sizeof(Grid));
    if (!g) return set_error(LSF_ERR_ARG, "out of host memory");
    g->f32 = 1;
    g->dm.nx = nx; g->dm.ny = ny; g->dm.nz = nz;
    g->dm.sx = (long long)nx + 1;
    g->dm.sxy = g->dm.sx * ((long long)ny + 1);
    g->np = g->dm.sxy * ((long long)nz + 1);
    slab_geom(nz, 1, 0, g->sg);
    const size_t bytes = sizeof(float) * (size_t)g->np + 64;           // + one chunk: the sweep kernel reads whole aligned vectors (RowReader)
    cudaError_t e;
    if ((e = cudaMalloc(&g->phi_f, bytes)) != cudaSuccess || (e = cudaMalloc(&g->phiS_f, bytes)) != cudaSuccess ||
        (e = cudaMalloc(&g->partial, sizeof(double) * PARTIAL_CAP)) != cudaSuccess ||
        (e = cudaMalloc(&g->ctrl, sizeof(Ctrl))) != cudaSuccess) {
        lsf_grid_destroy(g);
        return set_error(LSF_ERR_CUDA, "grid_create_f32: %s", cudaGetErrorString(e));
    }
    *out = g;
    return LSF_OK;
}
