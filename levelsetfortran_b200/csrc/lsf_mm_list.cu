// lsf_mm_list.cu -- GPU side of the active-list min/max flow (algorithm and exactness argument:
// lsf_mm_list.cuh).  Kernels:
//   k_mml_count / k_mml_scan / k_mml_fill : deterministic, ordered compaction of the active set S into a list
//   k_mml_iter    : one iteration over the list -- speculative, order-free update of every band cell; the few
//                   undecided cells are marked and queued
//   k_mml_settle  : settles the queue in dependence order (one CTA; normally finds it empty)
// HBM traffic per iteration is proportional to |S| (the narrow band, ~1.5 % of a 1024^3 grid), not to the
// grid: the stencil reads of neighbouring list entries hit L1/L2.
#include <stdlib.h>

#include "lsf_internal.cuh"
#include "lsf_march.cuh"
#include "lsf_mm_list.cuh"

namespace lsf {

constexpr int ML_THREADS = 256;
constexpr int ML_PER_THREAD = 8;
constexpr int ML_CHUNK = ML_THREADS * ML_PER_THREAD;     // consecutive points per CTA of the compaction
constexpr int ML_WORK_CAP = 1 << 22;                     // undecided cells per iteration the queue can hold

// membership in S = {phiNB_1 == 1} U {abs(phi_0) < 4.1*dx}, restricted to the cells this rank updates
__device__ __forceinline__ bool mml_member(const double *__restrict__ phi, const uint8_t *__restrict__ mask, const Dims &dm,
                                           long long q, int kA, int kB, double bNB)
{
    const int i = (int)(q % dm.sx), j = (int)((q / dm.sx) % (dm.ny + 1)), k = (int)(q / dm.sxy);
    if (i < 1 || i > dm.nx - 1 || j < 1 || j > dm.ny - 1 || k < kA || k > kB) return false;
    return (mask && mask[q]) || fabs(phi[q]) < bNB;
}

__global__ void __launch_bounds__(ML_THREADS)
k_mml_count(const double *__restrict__ phi, const uint8_t *__restrict__ mask, Dims dm, long long q0, long long q1, int kA, int kB,
            double bNB, int *__restrict__ counts)
{
    __shared__ int sh[ML_THREADS / 32];
    const long long base = q0 + (long long)blockIdx.x * ML_CHUNK;
    int n = 0;
    for (int r = 0; r < ML_PER_THREAD; ++r) {
        const long long q = base + r * ML_THREADS + threadIdx.x;
        if (q < q1 && mml_member(phi, mask, dm, q, kA, kB, bNB)) ++n;
    }
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < ML_THREADS / 32; ++w) t += sh[w];
        counts[blockIdx.x] = t;
    }
}

// exclusive scan of the per-CTA counts (one CTA, chunked); offsets[nblocks] = total
__global__ void __launch_bounds__(1024)
k_mml_scan(const int *__restrict__ counts, long long *__restrict__ offsets, int nblocks)
{
    __shared__ long long sh[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int n = base + threadIdx.x;
        const long long v = n < nblocks ? counts[n] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (n < nblocks) offsets[n] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[nblocks] = carry;
}

__global__ void __launch_bounds__(ML_THREADS)
k_mml_fill(const double *__restrict__ phi, const uint8_t *__restrict__ mask, Dims dm, long long q0, long long q1, int kA, int kB,
           double bNB, const long long *__restrict__ offsets, long long *__restrict__ list)
{
    __shared__ int sh[ML_THREADS / 32];
    const long long base = q0 + (long long)blockIdx.x * ML_CHUNK;
    long long out = offsets[blockIdx.x];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int r = 0; r < ML_PER_THREAD; ++r) {                  // ascending q: r-major, then thread
        const long long q = base + r * ML_THREADS + threadIdx.x;
        const bool m = q < q1 && mml_member(phi, mask, dm, q, kA, kB, bNB);
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        if (lane == 0) sh[wid] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
        for (int w = 0; w < ML_THREADS / 32; ++w) { if (w < wid) off += sh[w]; tot += sh[w]; }
        if (m) list[out + off + __popc(bal & ((1u << lane) - 1u))] = q;
        out += tot;
        __syncthreads();
    }
}

struct MmlArgs {
    const long long *list;
    long long n;
    const double *A;
    double *B;
    const uint8_t *mask;          // band of iteration 1 when the caller gave one, else null
    MmListConst c;
    uint8_t *unres;               // 1 = undecided (queued)
    long long *work;
    int *work_count;
    double *partial;
    Ctrl *ctrl;
    const long long *halo_seq;    // z-slab: ghost planes of A must be in place before the iteration starts
    long long halo_need[2];
    int k_first;                  // z-slab: first updated local plane of a rank that has a lower neighbour, else -1
    long long settle_need;        // ... whose queue entries on that plane wait for the lower rank's planes of THIS iteration
};

__global__ void __launch_bounds__(ML_THREADS)
k_mml_iter(const MmlArgs a)
{
    if (a.ctrl->done) return;
    __shared__ double sh[ML_THREADS];
    if (a.halo_seq) {
        if (threadIdx.x == 0) {
            if (a.halo_need[0]) wait_ge<true>(a.halo_seq + 0, a.halo_need[0], a.ctrl);
            if (a.halo_need[1]) wait_ge<true>(a.halo_seq + 1, a.halo_need[1], a.ctrl);
        }
        __syncthreads();
    }
    double acc = 0.;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (long long)gridDim.x * blockDim.x) {
        const long long q = a.list[e];
        if (!mm_inband(a.c, a.A, a.mask, q)) { a.B[q] = a.A[q]; continue; }      // left the band: frozen from now on
        const int i = (int)(q % a.c.sx), j = (int)((q / a.c.sx) % (a.c.ny + 1)), k = (int)(q / a.c.sxy);
        double pn;
        if (mm_cell_speculate(a.c, a.A, a.mask, q, i, j, k, pn)) {
            a.B[q] = pn;
            const double d = __dsub_rn(pn, a.A[q]);
            acc = __dadd_rn(acc, __dmul_rn(d, d));
        } else {
            a.unres[q] = 1;
            const int pos = atomicAdd(a.work_count, 1);
            if (pos < ML_WORK_CAP) a.work[pos] = q;
            else a.ctrl->status = LSF_ERR_ARG;          // queue overflow: reported by the host (never silently wrong)
        }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = ML_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) a.partial[blockIdx.x] = sh[0];
}

// Settle the queue: a cell is ready once none of its three upstream neighbours is still undecided.
// Passes are repeated until the queue is empty (dependence chains among undecided cells are short).
__global__ void __launch_bounds__(ML_THREADS)
k_mml_settle(const MmlArgs a, double *partial_slot)
{
    if (a.ctrl->done) return;
    __shared__ double sh[ML_THREADS];
    __shared__ int remaining, needs_lower;
    const int n = min(*a.work_count, ML_WORK_CAP);
    double acc = 0.;
    if (n > 0) {
        if (threadIdx.x == 0) needs_lower = 0;
        __syncthreads();
        if (a.k_first >= 0) {
            for (int e = threadIdx.x; e < n; e += ML_THREADS)
                if ((int)(a.work[e] / a.c.sxy) == a.k_first) needs_lower = 1;
            __syncthreads();
            // the k-1 neighbour of such a cell belongs to the lower rank: its final value arrives with that
            // rank's boundary planes of this iteration
            if (needs_lower && threadIdx.x == 0) wait_ge<true>(a.halo_seq + 0, a.settle_need, a.ctrl);
            __syncthreads();
        }
        for (int pass = 0; pass < n + 1; ++pass) {
            if (threadIdx.x == 0) remaining = 0;
            __syncthreads();
            for (int e = threadIdx.x; e < n; e += ML_THREADS) {
                const long long q = a.work[e];
                if (a.unres[q] == 1 && a.unres[q - 1] == 0 && a.unres[q - a.c.sx] == 0 && a.unres[q - a.c.sxy] == 0) a.unres[q] = 2;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < n; e += ML_THREADS) {
                const long long q = a.work[e];
                if (a.unres[q] == 2) {
                    const double pn = mm_cell_settle(a.c, a.A, a.B, q);
                    a.B[q] = pn;
                    const double d = __dsub_rn(pn, a.A[q]);
                    acc = __dadd_rn(acc, __dmul_rn(d, d));
                    a.unres[q] = 0;
                } else if (a.unres[q] == 1) remaining = 1;
            }
            __threadfence_block();
            __syncthreads();
            if (!remaining) break;
            __syncthreads();
        }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = ML_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) { *partial_slot = sh[0]; *a.work_count = 0; }
}

// ---------------------------------------------------------------------------------------------------
int mml_prepare(Grid *g, const uint8_t *mask, double dx)
{
    const Dims &dm = g->dm;
    const SlabGeom &sg = g->sg;
    const long long q0 = (long long)sg.kupd_lo * dm.sxy, q1 = (long long)(sg.kupd_hi + 1) * dm.sxy;
    const long long nblk = (q1 - q0 + ML_CHUNK - 1) / ML_CHUNK;
    if (nblk > 0x7fffffffLL) return set_error(LSF_ERR_ARG, "minmax: grid too large for the list builder");
    // scratch of the list builder, kept with the grid: a cudaMalloc / cudaFree pair per call costs 0.05 - 0.8 s of host time
    // next to 14 ms of kernels (measured, round 2 session 9: the same 64 iterations took 60 / 337 / 132 / 115 ms)
    if (g->mml_scratch_cap < nblk + 1) {
        cudaFree(g->mml_counts); cudaFree(g->mml_offsets);
        g->mml_counts = nullptr; g->mml_offsets = nullptr; g->mml_scratch_cap = 0;
        LSF_CUDA(cudaMalloc(&g->mml_counts, sizeof(int) * (size_t)(nblk + 1)));
        LSF_CUDA(cudaMalloc(&g->mml_offsets, sizeof(long long) * (size_t)(nblk + 1)));
        g->mml_scratch_cap = nblk + 1;
    }
    int *counts = g->mml_counts;
    long long *offsets = g->mml_offsets;
    cudaError_t e = cudaSuccess;
    k_mml_count<<<(unsigned)nblk, ML_THREADS, 0, G.stream>>>(g->phi, mask, dm, q0, q1, sg.kupd_lo, sg.kupd_hi, 4.1 * dx, counts);
    k_mml_scan<<<1, 1024, 0, G.stream>>>(counts, offsets, (int)nblk);
    long long total = 0;
    e = cudaMemcpyAsync(&total, offsets + nblk, sizeof(long long), cudaMemcpyDeviceToHost, G.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
    G.n_launch += 2;
    int rc = LSF_OK;
    if (e != cudaSuccess) rc = set_error(LSF_ERR_CUDA, "minmax: %s", cudaGetErrorString(e));
    if (!rc && g->mml_cap < total) {
        cudaFree(g->mml_list);
        g->mml_list = nullptr; g->mml_cap = 0;
        const long long cap = total + total / 8 + 1024;
        if (cudaMalloc(&g->mml_list, sizeof(long long) * (size_t)cap) != cudaSuccess) rc = set_error(LSF_ERR_CUDA, "minmax: out of device memory (list)");
        else g->mml_cap = cap;
    }
    if (!rc && !g->mml_unres) {
        if (cudaMalloc(&g->mml_unres, (size_t)g->np) != cudaSuccess || cudaMemsetAsync(g->mml_unres, 0, (size_t)g->np, G.stream) != cudaSuccess ||
            cudaMalloc(&g->mml_work, sizeof(long long) * ML_WORK_CAP) != cudaSuccess || cudaMalloc(&g->mml_work_count, sizeof(int)) != cudaSuccess ||
            cudaMemsetAsync(g->mml_work_count, 0, sizeof(int), G.stream) != cudaSuccess)
            rc = set_error(LSF_ERR_CUDA, "minmax: out of device memory (queue)");
    }
    if (!rc && total > 0) {
        k_mml_fill<<<(unsigned)nblk, ML_THREADS, 0, G.stream>>>(g->phi, mask, dm, q0, q1, sg.kupd_lo, sg.kupd_hi, 4.1 * dx, offsets, g->mml_list);
        G.n_launch++;
    }
    g->mml_n = total;
    G.mm_active = total;
    return rc;
}

int mml_npart() { return RMS_BLOCKS + 1; }

// One iteration: reads A (phi_old), writes B (phi_new) on the active list; partials -> g->partial[0..mml_npart())
void launch_minmax_iteration_list(Grid *g, const double *A, double *B, const uint8_t *mask, double dx, double h1)
{
    const SlabGeom &sg = g->sg;
    MmlArgs a;
    memset(&a, 0, sizeof(a));
    a.list = g->mml_list; a.n = g->mml_n; a.A = A; a.B = B; a.mask = mask;
    a.c.sx = g->dm.sx; a.c.sxy = g->dm.sxy; a.c.nx = g->dm.nx; a.c.ny = g->dm.ny;
    a.c.k_lo = 1 - sg.kbase; a.c.k_hi = sg.NZ - 1 - sg.kbase;
    a.c.bNB = 4.1 * dx; a.c.dxx = 1. / (dx * dx); a.c.h1 = h1;
    a.unres = g->mml_unres; a.work = g->mml_work; a.work_count = g->mml_work_count;
    a.partial = g->partial; a.ctrl = g->ctrl;
    a.k_first = -1;
    if (sharded(g)) {
        a.halo_seq = g->sync->halo_seq;
        if (sg.rank > 0) { a.halo_need[0] = g->phase; a.k_first = sg.kupd_lo; a.settle_need = g->phase + 1; }
        if (sg.rank < sg.nranks - 1) a.halo_need[1] = g->phase;
        g->prev_sweep_valid = false;
    }
    k_mml_iter<<<RMS_BLOCKS, ML_THREADS, 0, G.stream>>>(a);
    k_mml_settle<<<1, ML_THREADS, 0, G.stream>>>(a, g->partial + RMS_BLOCKS);
    G.n_launch += 2;
}

}  // namespace lsf
