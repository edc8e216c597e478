// lsf_sign_bvh.cu -- K1 with a bounding-volume hierarchy over the triangle centroids (round 2).
//
// Reference: the inline loop set3d.f90:218-268 -- per grid point the FIRST index of the nearest triangle centroid
// (`dis = sqrt(...)`, strict `dis < minD`, minD = 100000. to start with), then the scalar triple product and the
// smeared sign.  The tiled kernel of round 1 (lsf_kernels.cu: k_sign_search_tiled) culls exactly but scans ALL centroids
// twice per 8x8x4 block of points: O(blocks x nTri), 180 ms at 1024^3 with 10 k triangles and 10x that with 100 k.
// Here both scans become tree walks:
//   host   : a binary tree over the centroids (median split along the longest axis, leaves of <= 8), built once per
//            call from the device-computed centroids (so the boxes bound exactly the values the kernel compares);
//   pass 1 : d_min = distance of the nearest centroid to the block centre P -- breadth-first over the tree, a node
//            survives while its box can still hold something nearer than the best upper bound (min over visited boxes
//            of the farthest-corner distance, tightened by exact leaf evaluations);
//   pass 2 : every point of the block has its winner within |c - P| <= d_min + 2 rho (x (1 + 1e-9), as in the tiled
//            kernel): a range query collects those centroids leaf by leaf into shared memory and every thread runs the
//            reference's comparison over them with the same explicitly rounded arithmetic.
// The candidates arrive in TREE order, not index order, so "first index wins" is applied explicitly:
//            replace when dis < minD, or dis == minD (< 100000.) and n < fN
// -- the result of the ascending strict-< loop for any visiting order.  sqrt is only taken where it can matter: for
// q <= qbest (1 + 1e-15); two squared distances can only round to the same dis when they differ by <= 2^-51 relative.
// A frontier that outgrows its shared-memory array (never seen; a pathological all-equidistant cloud could) makes the
// block fall back to the linear scan, so the result is bit-identical to k_sign_search in every case.
#include <algorithm>
#include <vector>

#include "lsf_internal.cuh"

namespace lsf {

constexpr int BV_LEAF = 8;                 // centroids per leaf
constexpr int BV_FCAP = 1024;              // frontier capacity (nodes) per level
constexpr int BV_LCAP = 2048;              // leaves of one range query
constexpr int BI = 8, BJ = 8, BK = 4, BT = BI * BJ * BK;

struct BvhNode {
    double lo[3], hi[3];
    int a, b;                              // internal: children a, b (b > 0); leaf: first centroid a, count -b
};

// ---------------------------------------------------------------------------------------------- host build
static int bvh_build_rec(std::vector<BvhNode> &nodes, std::vector<int> &perm, const double *cx, const double *cy, const double *cz,
                         int first, int count)
{
    const int me = (int)nodes.size();
    nodes.push_back(BvhNode());
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int q = first; q < first + count; ++q) {
        const int n = perm[q];
        const double c[3] = {cx[n], cy[n], cz[n]};
        for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], c[d]); hi[d] = std::max(hi[d], c[d]); }
    }
    for (int d = 0; d < 3; ++d) { nodes[me].lo[d] = lo[d]; nodes[me].hi[d] = hi[d]; }
    if (count <= BV_LEAF) { nodes[me].a = first; nodes[me].b = -count; return me; }
    int ax = 0;
    if (hi[1] - lo[1] > hi[ax] - lo[ax]) ax = 1;
    if (hi[2] - lo[2] > hi[ax] - lo[ax]) ax = 2;
    const double *c = ax == 0 ? cx : (ax == 1 ? cy : cz);
    const int half = count / 2;
    std::nth_element(perm.begin() + first, perm.begin() + first + half, perm.begin() + first + count,
                     [c](int u, int v) { return c[u] < c[v] || (c[u] == c[v] && u < v); });
    const int l = bvh_build_rec(nodes, perm, cx, cy, cz, first, half);
    const int r = bvh_build_rec(nodes, perm, cx, cy, cz, first + half, count - half);
    nodes[me].a = l; nodes[me].b = r;
    return me;
}

// ---------------------------------------------------------------------------------------------- device
__device__ __forceinline__ void box_dist2(const BvhNode &nd, double PX, double PY, double PZ, double &mind2, double &maxd2)
{
    const double P[3] = {PX, PY, PZ};
    mind2 = 0.; maxd2 = 0.;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double a = nd.lo[d] - P[d], b = P[d] - nd.hi[d];
        const double out = fmax(fmax(a, b), 0.);
        const double far = fmax(fabs(a), fabs(b));
        mind2 += out * out;
        maxd2 += far * far;
    }
}

__device__ __forceinline__ void atomic_min_nonneg(double *addr, double v)
{
    atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));     // bit order == value order for v >= 0
}

struct SignPoint {          // per-thread state of the reference's argmin loop
    double gX, gY, gZ, minD, qbest, qlim;
    int fN;
    __device__ __forceinline__ void offer(double cx, double cy, double cz, int n)
    {
        const double ex = __dsub_rn(cx, gX), ey = __dsub_rn(cy, gY), ez = __dsub_rn(cz, gZ);
        const double q = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
        if (q <= qlim) {
            const double dis = __dsqrt_rn(q);
            if (dis < minD || (dis == minD && dis < 100000. && n < fN)) {
                if (dis < minD || q < qbest) { qbest = q; qlim = q * (1.0 + 1.0e-15); }
                minD = dis; fN = n;
            }
        }
    }
};

__global__ void __launch_bounds__(BT)
k_sign_search_bvh(double *__restrict__ phi, Dims dm, double x0, double y0, double z0, double dx,
                  const double *__restrict__ surfX, int nNode, const int32_t *__restrict__ surfElem, int nElem,
                  const BvhNode *__restrict__ nodes, const double *__restrict__ pc /* permuted centroids, SoA [3][nElem] */,
                  const int *__restrict__ pidx /* original index of permuted centroid */,
                  const double *__restrict__ cen /* original order (linear fallback) */,
                  int im, int jm, int km, int ni, int nj, int nk, int kbase, int nbi, int nbj)
{
    __shared__ int fr[2][BV_FCAP];
    __shared__ int leaves[BV_LCAP];
    __shared__ int fcnt[2], lcnt, overflow;
    __shared__ double s_ub;
    __shared__ double sc[3][BT];
    __shared__ int sidx[BT];
    __shared__ int s_fill;
    const int tid = threadIdx.x;
    const int bi = blockIdx.x % nbi, bj = (blockIdx.x / nbi) % nbj, bk = blockIdx.x / (nbi * nbj);
    const int li = tid % BI, lj = (tid / BI) % BJ, lk = tid / (BI * BJ);
    const int i = im + bi * BI + li, j = jm + bj * BJ + lj, k = km + bk * BK + lk;
    const bool live = (i < im + ni) && (j < jm + nj) && (k < km + nk);
    SignPoint pt;
    pt.gX = __dadd_rn(x0, __dmul_rn((double)i, dx));          // set3d.f90:168-170
    pt.gY = __dadd_rn(y0, __dmul_rn((double)j, dx));
    pt.gZ = __dadd_rn(z0, __dmul_rn((double)k, dx));
    pt.minD = 100000.; pt.qbest = 1.0e10; pt.qlim = 1.0e10 * (1.0 + 1.0e-15); pt.fN = 0;
    const double PX = x0 + (im + bi * BI + 0.5 * (BI - 1)) * dx;
    const double PY = y0 + (jm + bj * BJ + 0.5 * (BJ - 1)) * dx;
    const double PZ = z0 + (km + bk * BK + 0.5 * (BK - 1)) * dx;
    const double rho = 0.5 * dx * sqrt((double)((BI - 1) * (BI - 1) + (BJ - 1) * (BJ - 1) + (BK - 1) * (BK - 1)));

    // ---- pass 1: nearest centroid distance to P --------------------------------------------------------------
    if (tid == 0) { fr[0][0] = 0; fcnt[0] = 1; fcnt[1] = 0; overflow = 0; s_ub = __longlong_as_double(0x7fefffffffffffffLL); }
    __syncthreads();
    int cur = 0;
    while (true) {
        const int nf = fcnt[cur];
        if (nf == 0 || overflow) break;
        // (a) tighten the upper bound from this level's nodes
        for (int q = tid; q < nf; q += BT) {
            const BvhNode nd = nodes[fr[cur][q]];
            if (nd.b < 0) {
                double m = __longlong_as_double(0x7fefffffffffffffLL);
                for (int r = 0; r < -nd.b; ++r) {
                    const int c = nd.a + r;
                    const double ex = pc[c] - PX, ey = pc[c + (long long)nElem] - PY, ez = pc[c + 2 * (long long)nElem] - PZ;
                    m = fmin(m, ex * ex + ey * ey + ez * ez);
                }
                atomic_min_nonneg(&s_ub, m);
            } else {
                double mn, mx;
                box_dist2(nodes[nd.a], PX, PY, PZ, mn, mx); atomic_min_nonneg(&s_ub, mx);
                box_dist2(nodes[nd.b], PX, PY, PZ, mn, mx); atomic_min_nonneg(&s_ub, mx);
            }
        }
        __syncthreads();
        // (b) children that can still hold something nearer go to the next level
        const double ub = s_ub * (1.0 + 1.0e-12);
        for (int q = tid; q < nf; q += BT) {
            const BvhNode nd = nodes[fr[cur][q]];
            if (nd.b >= 0) {
                const int ch[2] = {nd.a, nd.b};
                for (int s = 0; s < 2; ++s) {
                    double mn, mx;
                    box_dist2(nodes[ch[s]], PX, PY, PZ, mn, mx);
                    if (mn <= ub) {
                        const int pos = atomicAdd(&fcnt[cur ^ 1], 1);
                        if (pos < BV_FCAP) fr[cur ^ 1][pos] = ch[s]; else overflow = 1;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) fcnt[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
    bool linear = overflow != 0;
    double bound2 = 0.;
    if (!linear) {
        const double bound = (sqrt(s_ub) + 2.0 * rho) * (1.0 + 1.0e-9);
        bound2 = bound * bound;
        // ---- pass 2a: leaves that intersect the ball |c - P| <= bound -----------------------------------------
        __syncthreads();
        if (tid == 0) { fr[0][0] = 0; fcnt[0] = 1; fcnt[1] = 0; lcnt = 0; }
        __syncthreads();
        cur = 0;
        while (true) {
            const int nf = fcnt[cur];
            if (nf == 0 || overflow) break;
            for (int q = tid; q < nf; q += BT) {
                const int id = fr[cur][q];
                const BvhNode nd = nodes[id];
                if (nd.b < 0) {
                    const int pos = atomicAdd(&lcnt, 1);
                    if (pos < BV_LCAP) leaves[pos] = id; else overflow = 1;
                } else {
                    const int ch[2] = {nd.a, nd.b};
                    for (int s = 0; s < 2; ++s) {
                        double mn, mx;
                        box_dist2(nodes[ch[s]], PX, PY, PZ, mn, mx);
                        if (mn <= bound2) {
                            const int pos = atomicAdd(&fcnt[cur ^ 1], 1);
                            if (pos < BV_FCAP) fr[cur ^ 1][pos] = ch[s]; else overflow = 1;
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) fcnt[cur] = 0;
            cur ^= 1;
            __syncthreads();
        }
        linear = overflow != 0;
    }
    if (!linear) {
        // ---- pass 2b: the leaves' centroids inside the ball, BT / BV_LEAF leaves per round --------------------
        const int nl = lcnt;
        for (int base = 0; base < nl; base += BT / BV_LEAF) {
            __syncthreads();
            if (tid == 0) s_fill = 0;
            __syncthreads();
            const int lq = base + tid / BV_LEAF, r = tid % BV_LEAF;
            if (lq < nl) {
                const BvhNode nd = nodes[leaves[lq]];
                if (r < -nd.b) {
                    const int c = nd.a + r;
                    const double cx = pc[c], cy = pc[c + (long long)nElem], cz = pc[c + 2 * (long long)nElem];
                    const double ex = cx - PX, ey = cy - PY, ez = cz - PZ;
                    if (ex * ex + ey * ey + ez * ez <= bound2) {
                        const int pos = atomicAdd(&s_fill, 1);
                        sc[0][pos] = cx; sc[1][pos] = cy; sc[2][pos] = cz; sidx[pos] = pidx[c];
                    }
                }
            }
            __syncthreads();
            const int fill = s_fill;
            for (int q = 0; q < fill; ++q) pt.offer(sc[0][q], sc[1][q], sc[2][q], sidx[q]);
        }
    } else {
        // ---- fallback: the reference's loop over ALL centroids in index order (bit-identical by construction) --
        for (int base = 0; base < nElem; base += BT) {
            const int n = base + tid;
            __syncthreads();
            if (n < nElem) { sc[0][tid] = cen[n]; sc[1][tid] = cen[n + (long long)nElem]; sc[2][tid] = cen[n + 2 * (long long)nElem]; }
            __syncthreads();
            const int cnt = min(BT, nElem - base);
            for (int q = 0; q < cnt; ++q) pt.offer(sc[0][q], sc[1][q], sc[2][q], base + q);
        }
    }
    if (!live) return;
    const int fN = pt.fN;
    const double gX = pt.gX, gY = pt.gY, gZ = pt.gZ;
    const int n1 = surfElem[fN] - 1, n2 = surfElem[fN + nElem] - 1, n3 = surfElem[fN + 2 * (long long)nElem] - 1;
    const double *X = surfX, *Y = surfX + nNode, *Z = surfX + 2 * (long long)nNode;
    const double A1 = __dsub_rn(X[n1], gX), A2 = __dsub_rn(Y[n1], gY), A3 = __dsub_rn(Z[n1], gZ);      // set3d.f90:242-250
    const double B1 = __dsub_rn(X[n2], gX), B2 = __dsub_rn(Y[n2], gY), B3 = __dsub_rn(Z[n2], gZ);
    const double C1 = __dsub_rn(X[n3], gX), C2 = __dsub_rn(Y[n3], gY), C3 = __dsub_rn(Z[n3], gZ);
    const double pSx = __dsub_rn(__dmul_rn(A2, B3), __dmul_rn(A3, B2));                               // :253-255
    const double pSy = -__dsub_rn(__dmul_rn(A1, B3), __dmul_rn(B1, A3));
    const double pSz = __dsub_rn(__dmul_rn(A1, B2), __dmul_rn(B1, A2));
    const double pS = -__dadd_rn(__dadd_rn(__dmul_rn(pSx, C1), __dmul_rn(pSy, C2)), __dmul_rn(pSz, C3));   // :258
    const double den = __dsqrt_rn(__dadd_rn(__dmul_rn(pS, pS), __dmul_rn(__dmul_rn(dx, dx), 1.)));   // phiSign, gM = 1 (:260-264)
    phi[i + dm.sx * j + dm.sxy * (k - kbase)] = __ddiv_rn(pS, den);
}

// Builds the tree from the device-computed centroids and runs the search.  Returns false if the BVH path cannot be used
// (allocation failure): the caller then runs the tiled kernel.
bool launch_sign_search_bvh(Grid *g, const double xLo[3], double dx, const double *d_surfX, int nNode, const int32_t *d_surfElem,
                            int nElem, const double *d_cen, int im, int jm, int km, int ni, int nj, int nk)
{
    std::vector<double> cen(3 * (size_t)nElem);
    if (cudaMemcpyAsync(cen.data(), d_cen, sizeof(double) * cen.size(), cudaMemcpyDeviceToHost, G.stream) != cudaSuccess) return false;
    if (cudaStreamSynchronize(G.stream) != cudaSuccess) return false;
    const double *cx = cen.data(), *cy = cx + nElem, *cz = cy + nElem;
    std::vector<int> perm(nElem);
    for (int n = 0; n < nElem; ++n) perm[n] = n;
    std::vector<BvhNode> nodes;
    nodes.reserve(2 * (size_t)nElem / BV_LEAF * 2 + 16);
    bvh_build_rec(nodes, perm, cx, cy, cz, 0, nElem);
    std::vector<double> pc(3 * (size_t)nElem);
    for (int q = 0; q < nElem; ++q) { pc[q] = cx[perm[q]]; pc[q + (size_t)nElem] = cy[perm[q]]; pc[q + 2 * (size_t)nElem] = cz[perm[q]]; }
    BvhNode *d_nodes = nullptr;
    double *d_pc = nullptr;
    int *d_pidx = nullptr;
    bool ok = cudaMalloc(&d_nodes, sizeof(BvhNode) * nodes.size()) == cudaSuccess && cudaMalloc(&d_pc, sizeof(double) * pc.size()) == cudaSuccess &&
              cudaMalloc(&d_pidx, sizeof(int) * (size_t)nElem) == cudaSuccess;
    if (ok) {
        cudaMemcpyAsync(d_nodes, nodes.data(), sizeof(BvhNode) * nodes.size(), cudaMemcpyHostToDevice, G.stream);
        cudaMemcpyAsync(d_pc, pc.data(), sizeof(double) * pc.size(), cudaMemcpyHostToDevice, G.stream);
        cudaMemcpyAsync(d_pidx, perm.data(), sizeof(int) * (size_t)nElem, cudaMemcpyHostToDevice, G.stream);
        const long long nbi = (ni + BI - 1) / BI, nbj = (nj + BJ - 1) / BJ, nbk = (nk + BK - 1) / BK;
        k_sign_search_bvh<<<(unsigned)(nbi * nbj * nbk), BT, 0, G.stream>>>(g->phi, g->dm, xLo[0], xLo[1], xLo[2], dx, d_surfX, nNode, d_surfElem,
                                                                       nElem, d_nodes, d_pc, d_pidx, d_cen, im, jm, km, ni, nj, nk, g->sg.kbase,
                                                                       (int)nbi, (int)nbj);
        G.n_launch++;
        ok = cudaStreamSynchronize(G.stream) == cudaSuccess;      // the host vectors and the tree live until here
    }
    cudaFree(d_nodes); cudaFree(d_pc); cudaFree(d_pidx);
    return ok;
}

}  // namespace lsf
