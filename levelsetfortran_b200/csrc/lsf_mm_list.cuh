// lsf_mm_list.cuh -- the min/max-flow iteration (set3d.f90:399-431) on an ACTIVE LIST, fully parallel and
// still bit-identical to the reference's in-place Gauss-Seidel pass.
//
// Two facts about the reference loop make this possible:
//  1. A cell is touched in iteration n only if abs(phi_{n-1}) < 4.1*dx (narrowBand, subs.f90:194; the first
//     iteration uses the caller's phiNB).  A cell outside the band is never written, so its value -- and with
//     it its band status -- can never change again: from the second iteration on the band only SHRINKS.
//     The set S = {phiNB_1 == 1} U {abs(phi_0) < 4.1*dx} therefore contains every cell that will ever change,
//     and an iteration only has to visit S (about 1.5 % of a 1024^3 grid) instead of streaming the grid.
//  2. The Gauss-Seidel coupling of pass B is weak: F is either L or 0 (min/max of the Jacobi Laplacian with
//     0, subs.f90:477-481), selected by the sign of pAve, the 7-point average of the LIVE array
//     (subs.f90:473-474).  The three already-updated neighbours (i-1, j-1, k-1) enter pAve with their new
//     values, each of which is one of two known numbers: phi_old (not moved) or phi_old + h1*L (moved).
//     So s(c) = [pAve(c) < 0] can be evaluated for all <= 8 combinations from OLD values alone; if they
//     agree -- always, except where pAve is within ~h1*|L| of zero -- the cell is decided without knowing
//     its neighbours' outcome.  The handful of undecided cells go to a worklist and are settled afterwards
//     in dependence order from the then final neighbour values (mm_cell_settle).
// Every floating-point operation that produces a stored value is the reference's, in its order, explicitly
// rounded; the speculation only decides WHICH of the reference's two values a cell takes.
//
// This header is plain host/device code (also compiled by g++ for tests/emu).
#pragma once
#include "lsf_cell.cuh"

namespace lsf {

struct MmListConst {
    long long sx, sxy;     // strides of j and k
    int nx, ny;            // i, j extents (points 0..nx, 0..ny)
    int k_lo, k_hi;        // LOCAL planes that are interior planes 1..NZ-1 of the GLOBAL grid (z-slab: ghost planes included)
    double bNB;            // 4.1*dx (subs.f90:194)
    double dxx;            // 1./(dx*dx) (subs.f90:384)
    double h1;
};

// is q an updatable cell (interior of the global grid)?  The reference never writes boundary points.
LSF_HD bool mm_interior(const MmListConst &c, int i, int j, int k)
{
    return i >= 1 && i <= c.nx - 1 && j >= 1 && j <= c.ny - 1 && k >= c.k_lo && k <= c.k_hi;
}

// band status in this iteration: the caller's mask (iteration 1, set3d.f90:360) or abs(old) < 4.1*dx (:460)
LSF_HD bool mm_inband(const MmListConst &c, const double *A, const unsigned char *mask, long long q)
{
    return mask ? mask[q] != 0 : fabs(A[q]) < c.bNB;
}

// phiXX + phiYY + phiZZ of the old values (secondDeriv subs.f90:387-389, minMax :461)
LSF_HD double mm_lap(const MmListConst &c, const double *A, long long q)
{
    typedef ExactArith X;
    const double m2 = X::mul(-2., A[q]);
    const double xx = X::mul(X::add(X::add(m2, A[q + 1]), A[q - 1]), c.dxx);
    const double yy = X::mul(X::add(X::add(m2, A[q + c.sx]), A[q - c.sx]), c.dxx);
    const double zz = X::mul(X::add(X::add(m2, A[q + c.sxy]), A[q - c.sxy]), c.dxx);
    return X::add(X::add(xx, yy), zz);
}

// subs.f90:473-474, left to right
LSF_HD double mm_pave(double pc, double nxm, double oxp, double oyp, double nym, double ozp, double nzm)
{
    typedef ExactArith X;
    double s = X::add(pc, nxm);
    s = X::add(s, oxp);
    s = X::add(s, oyp);
    s = X::add(s, nym);
    s = X::add(s, ozp);
    s = X::add(s, nzm);
    return X::div(s, 7.);
}

// subs.f90:477-481 + set3d.f90:426
LSF_HD double mm_apply(double pc, double L, bool s, double h1)
{
    const double F = s ? fmin_f(L, 0.0) : fmax_f(L, 0.0);
    return ExactArith::add(pc, ExactArith::mul(h1, F));
}

// The two values an upstream neighbour u can have after its own update: v[0] = not moved, v[1] = moved.
// Returns the number of distinct candidates (1 or 2).
LSF_HD int mm_candidates(const MmListConst &c, const double *A, const unsigned char *mask, long long qu, int i, int j, int k,
                         double v[2])
{
    v[0] = v[1] = A[qu];
    if (!mm_interior(c, i, j, k) || !mm_inband(c, A, mask, qu)) return 1;
    const double L = mm_lap(c, A, qu);
    if (!(L < 0.) && !(L > 0.)) return 1;            // L == 0 or NaN: F = 0 either way
    v[1] = ExactArith::add(A[qu], ExactArith::mul(c.h1, L));
    return 2;
}

// Speculative update of band cell q = (i,j,k) from OLD values only.
// Returns true and the new value if the sign of pAve is the same for every possible outcome of the three
// upstream neighbours; false if the cell has to be settled later.
LSF_HD bool mm_cell_speculate(const MmListConst &c, const double *A, const unsigned char *mask, long long q, int i, int j, int k,
                              double &pnew, bool *combos = nullptr)
{
    const double pc = A[q];
    const double oxp = A[q + 1], oyp = A[q + c.sx], ozp = A[q + c.sxy];
    const double L = mm_lap(c, A, q);
    double vx[2], vy[2], vz[2];
    const int cx = mm_candidates(c, A, mask, q - 1, i - 1, j, k, vx);
    const int cy = mm_candidates(c, A, mask, q - c.sx, i, j - 1, k, vy);
    const int cz = mm_candidates(c, A, mask, q - c.sxy, i, j, k - 1, vz);
    const bool s0 = mm_pave(pc, vx[0], oxp, oyp, vy[0], ozp, vz[0]) < 0.;
    // cheap certificate: the sum moves by at most delta over the combinations; far from zero -> same sign
    const double delta = fabs(vx[1] - vx[0]) + fabs(vy[1] - vy[0]) + fabs(vz[1] - vz[0]);
    const double sJ = pc + vx[0] + oxp + oyp + vy[0] + ozp + vz[0];
    const double scale = fabs(pc) + fabs(vx[0]) + fabs(oxp) + fabs(oyp) + fabs(vy[0]) + fabs(ozp) + fabs(vz[0]);
    bool same = fabs(sJ) >= 4. * delta + 1.0e-12 * scale + 1.0e-290;
    if (combos) *combos = !same;
    if (!same) {                                     // exact check of every combination
        same = true;
        for (int m = 1; m < 8 && same; ++m) {
            const int ax = m & 1, ay = (m >> 1) & 1, az = (m >> 2) & 1;
            if ((ax && cx == 1) || (ay && cy == 1) || (az && cz == 1)) continue;
            same = (mm_pave(pc, vx[ax], oxp, oyp, vy[ay], ozp, vz[az]) < 0.) == s0;
        }
    }
    if (!same) return false;
    pnew = mm_apply(pc, L, s0, c.h1);
    return true;
}

// Settle an undecided cell once its three upstream neighbours hold their final values in B.
LSF_HD double mm_cell_settle(const MmListConst &c, const double *A, const double *B, long long q)
{
    const double pc = A[q];
    const double L = mm_lap(c, A, q);
    const bool s = mm_pave(pc, B[q - 1], A[q + 1], A[q + c.sx], B[q - c.sx], A[q + c.sxy], B[q - c.sxy]) < 0.;
    return mm_apply(pc, L, s, c.h1);
}

}  // namespace lsf
