// lsf_nodes.cuh -- surface-node projection ("Advect Nodes", set3d.f90:465-501).
//
// Reference: gradPhi (zeroed at set3d.f90:372) receives the 8th-order first derivative (firstDeriv order 8,
// subs.f90:311-347, with the jp1 typo of :346) on the stencil band phiSB == 1; setPhiSurf (subs.f90:1057-1170)
// interpolates phi and gradPhi trilinearly at every node; then, for up to `iter` passes, every node with
// phiSurf > 1E-13 is moved by phiSurf * gradPhiSurf and setPhiSurf is run again over ALL nodes -- O(iter *
// nNode^2) interpolations.  setPhiSurf is a pure function of each node's own position, so re-interpolating
// only the moved node gives identical results, and the nodes are independent of each other: one thread per
// node runs the reference's loop for its node.  gradPhi is not materialised: the <= 24 derivative values a
// node needs per step are evaluated on the fly, with the reference's operation order and explicit rounding,
// so positions, phiSurf and gradPhiSurf are bit-identical to the reference loop.
//
// Plain host/device code (also compiled by g++ for tests/emu).
#pragma once
#include "lsf_cell.cuh"

namespace lsf {

struct NodeConst {
    long long sx, sxy;
    int nx, ny, nz;
    double xLo[3];
    double dx;
    double bSB;          // 8.1*dx (subs.f90:198)
};

constexpr int NODE_OK = 0, NODE_OFF_GRID = 1, NODE_BAND_ON_BOUNDARY = 2;

// firstDeriv order 8 at point q along one axis (stride st); YTYPO: subs.f90:346 reads j+1 where j+2 is meant
// PV: anything indexable by the global linear index -- `const double *` on one GPU, lsf::SlabView on a sharded grid
template <bool YTYPO, class PV>
LSF_HD double node_d8(const PV &phi, long long q, long long st, double dx)
{
    typedef ExactArith X;
    const double aa1 = 1. / 280., aa2 = -4. / 105., aa3 = 1. / 5., aa4 = -4. / 5.;
    const double aa6 = 4. / 5, aa7 = -1. / 5., aa8 = 4. / 105., aa9 = -1. / 280.;
    double s = X::mul(phi[q - 4 * st], aa1);
    s = X::add(s, X::mul(phi[q - 3 * st], aa2));
    s = X::add(s, X::mul(phi[q - 2 * st], aa3));
    s = X::add(s, X::mul(phi[q - st], aa4));
    s = X::add(s, X::mul(phi[q + st], aa6));
    s = X::add(s, X::mul(phi[q + (YTYPO ? 1 : 2) * st], aa7));
    s = X::add(s, X::mul(phi[q + 3 * st], aa8));
    s = X::add(s, X::mul(phi[q + 4 * st], aa9));
    return X::div(s, dx);
}

// gradPhi(i,j,k,1:3) as the reference holds it when setPhiSurf runs: the order-8 derivative on the stencil
// band (band of `sbsrc`: the field the last narrowBand call saw), 0 elsewhere.
template <class PV>
LSF_HD int node_grad(const NodeConst &c, const PV &phi, const PV &sbsrc, int i, int j, int k, double g[3])
{
    const long long q = i + c.sx * j + c.sxy * k;
    g[0] = g[1] = g[2] = 0.;
    if (!(fabs(sbsrc[q]) < c.bSB)) return NODE_OK;
    if (i < 4 || j < 4 || k < 4 || i > c.nx - 4 || j > c.ny - 4 || k > c.nz - 4) return NODE_BAND_ON_BOUNDARY;
    g[0] = node_d8<false>(phi, q, 1, c.dx);
    g[1] = node_d8<true>(phi, q, c.sx, c.dx);
    g[2] = node_d8<false>(phi, q, c.sxy, c.dx);
    return NODE_OK;
}

// setPhiSurf for one node, subs.f90:1078-1166
template <class PV>
LSF_HD int node_interp(const NodeConst &c, const PV &phi, const PV &sbsrc, const double x[3], double &phiSurf, double gs[3])
{
    typedef ExactArith X;
    const double fi = floor(X::div(X::sub(x[0], c.xLo[0]), c.dx));
    const double fj = floor(X::div(X::sub(x[1], c.xLo[1]), c.dx));
    const double fk = floor(X::div(X::sub(x[2], c.xLo[2]), c.dx));
    if (!(fi >= 0 && fi <= c.nx - 1 && fj >= 0 && fj <= c.ny - 1 && fk >= 0 && fk <= c.nz - 1)) return NODE_OFF_GRID;
    const int i0 = (int)fi, j0 = (int)fj, k0 = (int)fk;
    const double x0 = X::add(X::mul((double)i0, c.dx), c.xLo[0]), x1 = X::add(X::mul((double)(i0 + 1), c.dx), c.xLo[0]);
    const double y0 = X::add(X::mul((double)j0, c.dx), c.xLo[1]), y1 = X::add(X::mul((double)(j0 + 1), c.dx), c.xLo[1]);
    const double z0 = X::add(X::mul((double)k0, c.dx), c.xLo[2]), z1 = X::add(X::mul((double)(k0 + 1), c.dx), c.xLo[2]);
    const double xd = X::div(X::sub(x[0], x0), X::sub(x1, x0));
    const double yd = X::div(X::sub(x[1], y0), X::sub(y1, y0));
    const double zd = X::div(X::sub(x[2], z0), X::sub(z1, z0));
    const double ux = X::sub(1., xd), uy = X::sub(1., yd), uz = X::sub(1., zd);
    // corner order of subs.f90:1104-1107: (i0,j0,k0) (i1,j0,k0) | (i0,j1,k0) (i1,j1,k0) | (i0,j0,k1) (i1,j0,k1) | (i0,j1,k1) (i1,j1,k1)
    double f[4][8];
    int st = NODE_OK;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int di = m & 1, dj = (m >> 1) & 1, dk = (m >> 2) & 1;
        const long long q = (i0 + di) + c.sx * (j0 + dj) + c.sxy * (k0 + dk);
        f[0][m] = phi[q];
        double g[3];
        const int r = node_grad(c, phi, sbsrc, i0 + di, j0 + dj, k0 + dk, g);
        if (r) st = r;
        f[1][m] = g[0]; f[2][m] = g[1]; f[3][m] = g[2];
    }
    if (st) return st;
    double v[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double c00 = X::add(X::mul(f[a][0], ux), X::mul(f[a][1], xd));
        const double c10 = X::add(X::mul(f[a][2], ux), X::mul(f[a][3], xd));
        const double c01 = X::add(X::mul(f[a][4], ux), X::mul(f[a][5], xd));
        const double c11 = X::add(X::mul(f[a][6], ux), X::mul(f[a][7], xd));
        const double c0 = X::add(X::mul(c00, uy), X::mul(c10, yd));
        const double c1 = X::add(X::mul(c01, uy), X::mul(c11, yd));
        v[a] = X::add(X::mul(c0, uz), X::mul(c1, zd));
    }
    phiSurf = v[0];
    gs[0] = -v[1]; gs[1] = -v[2]; gs[2] = -v[3];
    const double gm2 = X::add(X::add(X::mul(gs[0], gs[0]), X::mul(gs[1], gs[1])), X::mul(gs[2], gs[2]));
    if (gm2 < 1.E-7) {
        gs[0] = gs[1] = gs[2] = 0.;
    } else {
        const double gm = X::sqr(gm2);
        gs[0] = X::div(gs[0], gm); gs[1] = X::div(gs[1], gm); gs[2] = X::div(gs[2], gm);
    }
    return NODE_OK;
}

// The reference's loop for one node (set3d.f90:483-501): x in/out, returns the status and the number of moves.
template <class PV>
LSF_HD int node_project(const NodeConst &c, const PV &phi, const PV &sbsrc, double x[3], double &phiSurf, double gs[3],
                        int iter, int &moves)
{
    typedef ExactArith X;
    moves = 0;
    phiSurf = 0.;
    gs[0] = gs[1] = gs[2] = 0.;
    int st = node_interp(c, phi, sbsrc, x, phiSurf, gs);
    for (int k = 1; k <= iter && st == NODE_OK; ++k) {
        if (!(phiSurf > 1E-13)) {
            // a node that fails the test once fails it in every later pass: its state no longer changes
            break;
        }
        x[0] = X::add(x[0], X::mul(phiSurf, gs[0]));
        x[1] = X::add(x[1], X::mul(phiSurf, gs[1]));
        x[2] = X::add(x[2], X::mul(phiSurf, gs[2]));
        ++moves;
        st = node_interp(c, phi, sbsrc, x, phiSurf, gs);
    }
    return st;
}

}  // namespace lsf
