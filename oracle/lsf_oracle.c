/*
 * lsf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code)
 *
 * A serial, plain-C restatement of the grid hot path of musheen/LevelSetFortran
 * (sign search, phiSign, narrowBand, secondDeriv, minMax, weno, reinit, the
 * min/max time loop and the STL reader that feeds them).  Every function cites
 * the reference file:line it follows.  It exists only so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * can check / time the reference algorithm; the product (levelsetfortran_b200/)
 * never links, imports or calls it.
 *
 * PARITY PINNED TO THE REFERENCE'S SOURCE TEXT (round 2): no Fortran compiler exists in the build
 * container, so the reference's two files are machine-translated to C statement by statement
 * (oracle/f90_to_c.py -> oracle/_ref/libref.so) and that translation -- the reference program
 * itself, run on its own cube40.stl / twoCube10.stl -- is compared with this restatement bit for
 * bit at every stage (stlRead, grid set-up, sign search, 2155 reinit sweeps, 406 min/max
 * iterations, node projection, reinit #2, NaN STOP at n = 272): tests/golden/make_golden.py
 * (verdicts in tests/golden/REF_PIN_REPORT.txt) and tests/test_ref_pins.py.  The golden fixtures
 * are generated from libref.so.  What remains unverified is the translator's reading of gfortran's
 * arithmetic (documented in f90_to_c.py), not this file.  Earlier pins stay: the known-answer
 * values of SURVEY.md section 6 and the internal consistency checks (tests/test_oracle_pins.py).
 *
 * Arithmetic contract (mirrors `gfortran -O3 -fdefault-real-8` on x86-64,
 * reference Makefile:4): every REAL and real literal is IEEE binary64, no FMA
 * contraction, no re-association, left-to-right evaluation, IEEE div/sqrt.
 * Build with:  gcc -O2 -ffp-contract=off -fno-fast-math  (oracle/Makefile).
 *
 * Array layout: Fortran column-major phi(0:nx,0:ny,0:nz), i fastest:
 *     idx(i,j,k) = i + (nx+1)*(j + (ny+1)*k)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j, k) ((size_t)(i) + sx * ((size_t)(j) + sy * (size_t)(k)))

/* gfortran MAX/MIN for reals: "mvar = a; if (b > mvar || isnan(mvar)) mvar = b" */
static inline double fmax_f(double a, double b) { return (b > a || isnan(a)) ? b : a; }
static inline double fmin_f(double a, double b) { return (b < a || isnan(a)) ? b : a; }

/* ------------------------------------------------------------------------- */
/* phiSign, subs.f90:152-172 (smeared sign; note gM is NOT squared, :169)     */
/* ------------------------------------------------------------------------- */
double orc_phisign(double pS, double dxx, double gM)
{
    return pS / sqrt(pS * pS + dxx * dxx * gM);
}

/* ------------------------------------------------------------------------- */
/* narrowBand, subs.f90:178-207                                               */
/* ------------------------------------------------------------------------- */
void orc_narrowband(int nx, int ny, int nz, double dx, const double *phi,
                    int32_t *phiNB, int32_t *phiSB)
{
    size_t n = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
    double bNB = 4.1 * dx, bSB = 8.1 * dx;   /* subs.f90:194,199 */
    for (size_t q = 0; q < n; ++q) {
        double a = fabs(phi[q]);
        phiNB[q] = (a < bNB) ? 1 : 0;
        phiSB[q] = (a < bSB) ? 1 : 0;
    }
}

/* ------------------------------------------------------------------------- */
/* weno, subs.f90:489-711.  One direction of the high-order branch            */
/* (x :509-552, y :555-598, z :601-644).  v[0..6] = phi at offsets -3..+3.    */
/* yquirk reproduces subs.f90:576 where p5 = (phi(j+3)-phi(j+3))/dx == 0.     */
/* ------------------------------------------------------------------------- */
static inline void weno_dir(const double v[7], double dx, int yquirk,
                            double *dm_out, double *dp_out)
{
    const double m3 = v[0], m2 = v[1], m1 = v[2], c0 = v[3], p1v = v[4], p2v = v[5], p3v = v[6];
    double ap = (p3v - 2. * p2v + p1v) / dx;
    double am = (m3 - 2. * m2 + m1) / dx;
    double bp = (p2v - 2. * p1v + c0) / dx;
    double bm = (m2 - 2. * m1 + c0) / dx;
    double cp = (p1v - 2. * c0 + m1) / dx;
    double cm = cp, dp = bm, dm = bp;

    double IS0p = 13. * (ap - bp) * (ap - bp) + 3. * (ap - 3. * bp) * (ap - 3. * bp);
    double IS0m = 13. * (am - bm) * (am - bm) + 3. * (am - 3. * bm) * (am - 3. * bm);
    double IS1p = 13. * (bp - cp) * (bp - cp) + 3. * (bp + cp) * (bp + cp);
    double IS1m = 13. * (bm - cm) * (bm - cm) + 3. * (bm + cm) * (bm + cm);
    double IS2p = 13. * (cp - dp) * (cp - dp) + 3. * (3. * cp - dp) * (3. * cp - dp);
    double IS2m = 13. * (cm - dm) * (cm - dm) + 3. * (3. * cm - dm) * (3. * cm - dm);

    double p0 = (m2 - m3) / dx;
    double p1 = (m1 - m2) / dx;
    double p2 = (c0 - m1) / dx;
    double p3 = (p1v - c0) / dx;
    double p4 = (p2v - p1v) / dx;
    double p5 = yquirk ? (p3v - p3v) / dx : (p3v - p2v) / dx;

    double epsp = (1.E-6) * fmax_f(p1 * p1, fmax_f(p2 * p2, fmax_f(p3 * p3, fmax_f(p4 * p4, p5 * p5)))) + 1.E-99;
    double epsm = (1.E-6) * fmax_f(p0 * p0, fmax_f(p1 * p1, fmax_f(p2 * p2, fmax_f(p3 * p3, p4 * p4)))) + 1.E-99;

    double a0p = 1. / ((epsp + IS0p) * (epsp + IS0p));
    double a0m = 1. / ((epsm + IS0m) * (epsm + IS0m));
    double a1p = 6. / ((epsp + IS1p) * (epsp + IS1p));
    double a1m = 6. / ((epsm + IS1m) * (epsm + IS1m));
    double a2p = 3. / ((epsp + IS2p) * (epsp + IS2p));
    double a2m = 3. / ((epsm + IS2m) * (epsm + IS2m));

    double w0p = a0p / (a0p + a1p + a2p);
    double w0m = a0m / (a0m + a1m + a2m);
    double w2p = a2p / (a0p + a1p + a2p);
    double w2m = a2m / (a0m + a1m + a2m);

    const double third = 1. / 3., sixth = 1. / 6., twelfth = 1. / 12.;
    double PWp = third * w0p * (ap - 2. * bp + cp) + sixth * (w2p - 0.5) * (bp - 2. * cp + dp);
    double PWm = third * w0m * (am - 2. * bm + cm) + sixth * (w2m - 0.5) * (bm - 2. * cm + dm);

    *dm_out = twelfth * (-p1 + 7. * p2 + 7. * p3 - p4) - PWm;
    *dp_out = twelfth * (-p1 + 7. * p2 + 7. * p3 - p4) + PWp;
}

/* weno, subs.f90:489-711: returns gM; optionally writes gradPhi(:,:,:,1:3)
 * (squared upwind terms, :696-698) and gradPhiMag (:703) when non-NULL.      */
double orc_weno(int i, int j, int k, int nx, int ny, int nz, double dx,
                const double *phi, double *gradPhi, double *gradPhiMag)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
    const size_t np = sx * sy * ((size_t)nz + 1);
    double a, b, c, d, e, f;
    if ((i > 3) && (i < nx - 4) && (j > 3) && (j < ny - 4) && (k > 3) && (k < nz - 4)) { /* :506 */
        double v[7];
        for (int m = -3; m <= 3; ++m) v[m + 3] = phi[IDX(i + m, j, k)];
        weno_dir(v, dx, 0, &a, &b);
        for (int m = -3; m <= 3; ++m) v[m + 3] = phi[IDX(i, j + m, k)];
        weno_dir(v, dx, 1, &c, &d);
        for (int m = -3; m <= 3; ++m) v[m + 3] = phi[IDX(i, j, k + m)];
        weno_dir(v, dx, 0, &e, &f);
    } else {                                                       /* :646-664 */
        double p = phi[IDX(i, j, k)];
        a = (p - phi[IDX(i - 1, j, k)]) / dx;
        b = (phi[IDX(i + 1, j, k)] - p) / dx;
        c = (p - phi[IDX(i, j - 1, k)]) / dx;
        d = (phi[IDX(i, j + 1, k)] - p) / dx;
        e = (p - phi[IDX(i, j, k - 1)]) / dx;
        f = (phi[IDX(i, j, k + 1)] - p) / dx;
    }
    double pa = fmax_f(a, 0.), pb = fmax_f(b, 0.), pc = fmax_f(c, 0.);
    double pd = fmax_f(d, 0.), pe = fmax_f(e, 0.), pf = fmax_f(f, 0.);
    double na = fmin_f(a, 0.), nb = fmin_f(b, 0.), nc = fmin_f(c, 0.);
    double nd = fmin_f(d, 0.), ne = fmin_f(e, 0.), nf = fmin_f(f, 0.);
    double gradX, gradY, gradZ;
    if (phi[IDX(i, j, k)] > 0.) {                                  /* :684-692 */
        gradX = fmax_f(pa * pa, nb * nb);
        gradY = fmax_f(pc * pc, nd * nd);
        gradZ = fmax_f(pe * pe, nf * nf);
    } else {
        gradX = fmax_f(pb * pb, na * na);
        gradY = fmax_f(pd * pd, nc * nc);
        gradZ = fmax_f(pf * pf, ne * ne);
    }
    double gM = sqrt(gradX + gradY + gradZ);                       /* :702 */
    if (gradPhi) {
        size_t q = IDX(i, j, k);
        gradPhi[q] = gradX; gradPhi[q + np] = gradY; gradPhi[q + 2 * np] = gradZ;
    }
    if (gradPhiMag) gradPhiMag[IDX(i, j, k)] = gM;
    return gM;
}

/* ------------------------------------------------------------------------- */
/* reinit boundary block, subs.f90:858-897.                                   */
/* Literal: the 26 statements re-executed for every (i,j,k), in order.        */
/* ------------------------------------------------------------------------- */
void orc_bc_literal(double *phi, int nx, int ny, int nz, double dx)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
#define P(i, j, k) phi[IDX(i, j, k)]
    for (int i = 0; i <= nx; ++i)
        for (int j = 0; j <= ny; ++j)
            for (int k = 0; k <= nz; ++k) {
                /* corners :864-871 */
                P(0, 0, 0) = P(1, 1, 1) + dx;
                P(nx, 0, 0) = P(nx - 1, 1, 1) + dx;
                P(0, ny, 0) = P(1, ny - 1, 1) + dx;
                P(0, 0, nz) = P(1, 1, nz - 1) + dx;
                P(nx, ny, 0) = P(nx - 1, ny - 1, 1) + dx;
                P(0, ny, nz) = P(1, ny - 1, nz - 1) + dx;
                P(nx, 0, nz) = P(nx - 1, 1, nz - 1) + dx;
                P(nx, ny, nz) = P(nx - 1, ny - 1, nz - 1) + dx;
                /* edges :874-885 */
                P(i, 0, 0) = P(i, 1, 1) + dx;
                P(0, j, 0) = P(1, j, 1) + dx;
                P(0, 0, k) = P(1, 1, k) + dx;
                P(i, ny, nz) = P(i, ny - 1, nz - 1) + dx;
                P(nx, j, nz) = P(nx - 1, j, nz - 1) + dx;
                P(nx, ny, k) = P(nx - 1, ny - 1, k) + dx;
                P(i, 0, nz) = P(i, 1, nz - 1) + dx;
                P(nx, j, 0) = P(nx - 1, j, 1) + dx;
                P(nx, 0, k) = P(nx - 1, 1, k) + dx;
                P(i, ny, 0) = P(i, ny - 1, 1) + dx;
                P(0, j, nz) = P(1, j, nz - 1) + dx;
                P(0, ny, k) = P(1, ny - 1, k) + dx;
                /* faces :888-893 */
                P(0, j, k) = P(1, j, k) + dx;
                P(i, 0, k) = P(i, 1, k) + dx;
                P(i, j, 0) = P(i, j, 1) + dx;
                P(nx, j, k) = P(nx - 1, j, k) + dx;
                P(i, ny, k) = P(i, ny - 1, k) + dx;
                P(i, j, nz) = P(i, j, nz - 1) + dx;
            }
#undef P
}

/* Closed form of the block above (re-verified
 * against orc_bc_literal in tests/test_oracle_pins.py): for a boundary point c,
 * with B = number of axes on which c is on the boundary and H = how many of
 * those are the high side,  phi(c) = phi(clamp(c,1..n-1)) + dx  applied
 * m = min(1+H, B) times (sequential adds).                                   */
void orc_bc_closed(double *phi, int nx, int ny, int nz, double dx)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i) {
                int bi = (i == 0 || i == nx), bj = (j == 0 || j == ny), bk = (k == 0 || k == nz);
                int B = bi + bj + bk;
                if (!B) continue;
                int H = (i == nx) + (j == ny) + (k == nz);
                int m = (1 + H < B) ? 1 + H : B;
                int ci = i < 1 ? 1 : (i > nx - 1 ? nx - 1 : i);
                int cj = j < 1 ? 1 : (j > ny - 1 ? ny - 1 : j);
                int ck = k < 1 ? 1 : (k > nz - 1 ? nz - 1 : k);
                double v = phi[IDX(ci, cj, ck)];
                for (int t = 0; t < m; ++t) v = v + dx;
                phi[IDX(i, j, k)] = v;
            }
}

/* raster -> sweep direction per axis (+1 ascending / -1 descending), subs.f90:742-852 */
static const int RASTER_DIR[8][3] = {
    {+1, +1, +1}, {+1, +1, -1}, {+1, -1, -1}, {-1, -1, -1},
    {-1, +1, -1}, {-1, -1, +1}, {-1, +1, +1}, {+1, -1, +1}};

static inline void reinit_cell(double *phi, const double *phiS, int i, int j, int k,
                               int nx, int ny, int nz, double dx, double h,
                               double *gradPhi, double *gradPhiMag)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
    double gM = orc_weno(i, j, k, nx, ny, nz, dx, phi, gradPhi, gradPhiMag);  /* :747 */
    double sgn = orc_phisign(phiS[IDX(i, j, k)], dx, gM);                     /* :748 */
    double k1 = sgn * (1. - gM);                                              /* :749 */
    phi[IDX(i, j, k)] = phi[IDX(i, j, k)] + h * k1;                           /* :750 */
}

/* One in-place Gauss-Seidel raster sweep, literal loop nest (i outer, j, k inner). */
void orc_reinit_sweep(double *phi, const double *phiS, int nx, int ny, int nz,
                      double dx, double h, int raster /*1..8*/,
                      double *gradPhi, double *gradPhiMag)
{
    const int *d = RASTER_DIR[raster - 1];
    for (int a = 1; a <= nx - 1; ++a) {
        int i = d[0] > 0 ? a : nx - a;
        for (int b = 1; b <= ny - 1; ++b) {
            int j = d[1] > 0 ? b : ny - b;
            for (int c = 1; c <= nz - 1; ++c) {
                int k = d[2] > 0 ? c : nz - c;
                reinit_cell(phi, phiS, i, j, k, nx, ny, nz, dx, h, gradPhi, gradPhiMag);
            }
        }
    }
}

/* Same sweep, visited in hyperplane order s = i'+j'+k' of the sweep-oriented
 * indices.  Not in the reference: exists to prove (bitwise) that wavefront
 * ordering -- what the CUDA kernels use -- is an exact reordering.            */
void orc_reinit_sweep_hyperplane(double *phi, const double *phiS, int nx, int ny, int nz,
                                 double dx, double h, int raster,
                                 double *gradPhi, double *gradPhiMag)
{
    const int *d = RASTER_DIR[raster - 1];
    for (int s = 3; s <= (nx - 1) + (ny - 1) + (nz - 1); ++s)
        for (int a = 1; a <= nx - 1; ++a)
            for (int b = 1; b <= ny - 1; ++b) {
                int c = s - a - b;
                if (c < 1 || c > nz - 1) continue;
                int i = d[0] > 0 ? a : nx - a;
                int j = d[1] > 0 ? b : ny - b;
                int k = d[2] > 0 ? c : nz - c;
                reinit_cell(phi, phiS, i, j, k, nx, ny, nz, dx, h, gradPhi, gradPhiMag);
            }
}

/* RMS of (phi - phiN) over ALL points in the reference's order (i outer, k inner),
 * divided by nx*ny*nz, subs.f90:902-914 and set3d.f90:435-447.  The reference's
 * divisor is an int32 product; this uses a 64-bit product (identical below 2^31). */
double orc_rms(const double *phi, const double *phiN, int nx, int ny, int nz)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
    double err = 0.;
    for (int i = 0; i <= nx; ++i)
        for (int j = 0; j <= ny; ++j)
            for (int k = 0; k <= nz; ++k) {
                size_t q = IDX(i, j, k);
                err = err + (phi[q] - phiN[q]) * (phi[q] - phiN[q]);
            }
    return sqrt(err / (double)((int64_t)nx * ny * nz));
}

/*
 * reinit, subs.f90:717-931.
 *   order: 0 = literal lexicographic loops, 1 = hyperplane order (test only)
 *   bc   : 0 = literal BC loop, 1 = closed form
 *   tol  : 1.E-5 in the reference (:915)
 * Returns 0 = converged (EXIT at :917), 1 = NaN STOP (:926), 2 = ran all iter+1 sweeps.
 * *n_exit = loop index n at which the routine left (iter when exhausted);
 * rms_hist[n] = phiErr of sweep n (the value the reference prints at :923, and
 * also the one that triggered EXIT).
 */
int orc_reinit(double *phi, double *gradPhi, double *gradPhiMag, int nx, int ny, int nz,
               int iter, double dx, double h, double tol, int order, int bc,
               int *n_exit, double *rms_hist)
{
    size_t np = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
    double *phiS = (double *)malloc(np * sizeof(double));
    double *phiN = (double *)malloc(np * sizeof(double));
    if (!phiS || !phiN) { free(phiS); free(phiN); return -1; }
    memcpy(phiS, phi, np * sizeof(double));                       /* :731 */
    memcpy(phiN, phi, np * sizeof(double));                       /* :732 */
    int raster = 0, status = 2, n;
    for (n = 0; n <= iter; ++n) {                                 /* :735 */
        raster = raster + 1;
        if (order == 0) orc_reinit_sweep(phi, phiS, nx, ny, nz, dx, h, raster, gradPhi, gradPhiMag);
        else orc_reinit_sweep_hyperplane(phi, phiS, nx, ny, nz, dx, h, raster, gradPhi, gradPhiMag);
        if (raster == 8) raster = 0;                              /* :855 */
        if (bc == 0) orc_bc_literal(phi, nx, ny, nz, dx);
        else orc_bc_closed(phi, nx, ny, nz, dx);
        double phiErr = orc_rms(phi, phiN, nx, ny, nz);
        if (rms_hist) rms_hist[n] = phiErr;
        if (phiErr < tol) { status = 0; break; }                  /* :915-918 */
        memcpy(phiN, phi, np * sizeof(double));                   /* :921 */
        if (isnan(phiErr)) { status = 1; break; }                 /* :926 */
    }
    if (n > iter) n = iter;
    if (n_exit) *n_exit = n;
    free(phiS); free(phiN);
    return status;
}

/* ------------------------------------------------------------------------- */
/* secondDeriv, subs.f90:370-407 (order 2 only).  out[6] = XX,YY,ZZ,XY,XZ,YZ  */
/* ------------------------------------------------------------------------- */
void orc_secondderiv(int i, int j, int k, int nx, int ny, int nz, double dx,
                     const double *phi, double out[6])
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
    (void)nz;
#define P(i, j, k) phi[IDX(i, j, k)]
    double dxx = 1. / (dx * dx);                                              /* :384 */
    out[0] = (-2. * P(i, j, k) + P(i + 1, j, k) + P(i - 1, j, k)) * dxx;      /* :387 */
    out[1] = (-2. * P(i, j, k) + P(i, j + 1, k) + P(i, j - 1, k)) * dxx;
    out[2] = (-2. * P(i, j, k) + P(i, j, k + 1) + P(i, j, k - 1)) * dxx;
    double xy = P(i + 1, j + 1, k) - P(i + 1, j - 1, k) - P(i - 1, j + 1, k) + P(i - 1, j - 1, k);
    double yz = P(i, j + 1, k + 1) - P(i, j + 1, k - 1) - P(i, j - 1, k + 1) + P(i, j - 1, k - 1);
    double xz = P(i + 1, j, k + 1) - P(i + 1, j, k - 1) - P(i - 1, j, k + 1) + P(i - 1, j, k - 1);
    out[3] = xy * dxx / 4.;
    out[4] = xz * dxx / 4.;
    out[5] = yz * dxx / 4.;
#undef P
}

/* minMax, subs.f90:413-483 (Laplacian branch :451-481).  lap3 = XX,YY,ZZ.     */
double orc_minmax_F(int i, int j, int k, int nx, int ny, int nz,
                    const double *phi, const double lap3[3])
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
    (void)nz;
#define P(i, j, k) phi[IDX(i, j, k)]
    double curv = lap3[0] + lap3[1] + lap3[2];                               /* :461 */
    double pAve = P(i, j, k) + P(i - 1, j, k) + P(i + 1, j, k) + P(i, j + 1, k)
                + P(i, j - 1, k) + P(i, j, k + 1) + P(i, j, k - 1);           /* :473 */
    pAve = pAve / 7.;                                                        /* :474 */
#undef P
    if (pAve < 0.) return fmin_f(curv, 0.0);                                 /* :477-481 */
    return fmax_f(curv, 0.0);
}

/*
 * The min/max time loop, set3d.f90:394-462 (DO n = 1,iter).
 * phiNB/phiSB must hold narrowBand(phi) on entry (set3d.f90:360) and hold the
 * last narrowBand result on exit.  phiN must equal phi on entry (:377).
 * The weno call at :422 and the mixed derivatives only feed values that are
 * overwritten before any use and are omitted.
 * Band cells on the grid boundary would make the reference read out of bounds
 * (undefined); this returns -2 in that case.
 * Returns 0 = steady state EXIT (:448-450), 1 = NaN STOP (:458), 2 = iter exhausted.
 * *n_exit = n at exit (iter when exhausted); rms_hist[n-1] = phiErr of iteration n.
 */
int orc_minmax(double *phi, double *phiN, int32_t *phiNB, int32_t *phiSB,
               int nx, int ny, int nz, int iter, double dx, double h1, double tol,
               int *n_exit, double *rms_hist)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
    size_t np = sx * sy * ((size_t)nz + 1);
    double *lap = (double *)calloc(3 * np, sizeof(double));      /* grad2Phi, set3d.f90:373 */
    if (!lap) return -1;
    int status = 2, n;
    for (n = 1; n <= iter; ++n) {
        /* pass A :399-414 */
        for (int i = 0; i <= nx; ++i)
            for (int j = 0; j <= ny; ++j)
                for (int k = 0; k <= nz; ++k) {
                    size_t q = IDX(i, j, k);
                    if (phiNB[q] == 1) {
                        if (i == 0 || j == 0 || k == 0 || i == nx || j == ny || k == nz) { free(lap); return -2; }
                        double o[6];
                        orc_secondderiv(i, j, k, nx, ny, nz, dx, phi, o);
                        lap[q] = o[0]; lap[q + np] = o[1]; lap[q + 2 * np] = o[2];
                    }
                }
        /* pass B :417-431 (in place) */
        for (int i = 0; i <= nx; ++i)
            for (int j = 0; j <= ny; ++j)
                for (int k = 0; k <= nz; ++k) {
                    size_t q = IDX(i, j, k);
                    if (phiNB[q] == 1) {
                        double l3[3] = {lap[q], lap[q + np], lap[q + 2 * np]};
                        double F = orc_minmax_F(i, j, k, nx, ny, nz, phi, l3);
                        phi[q] = phi[q] + h1 * F;                             /* :426 */
                    }
                }
        double phiErr = orc_rms(phi, phiN, nx, ny, nz);                       /* :435-447 */
        if (rms_hist) rms_hist[n - 1] = phiErr;
        if (phiErr < tol) { status = 0; break; }                              /* :448-451 */
        memcpy(phiN, phi, np * sizeof(double));                               /* :454 */
        if (isnan(phiErr)) { status = 1; break; }                             /* :458 */
        orc_narrowband(nx, ny, nz, dx, phi, phiNB, phiSB);                    /* :460 */
    }
    if (n > iter) n = iter;
    if (n_exit) *n_exit = n;
    free(lap);
    return status;
}

/* ------------------------------------------------------------------------- */
/* Grid definition, set3d.f90:90-186 and :301.                                */
/* surfX is Fortran (nSurfNode,3) column-major.  out: n[3]=nx,ny,nz;          */
/* xLo[3]; box[6]=im,ip,jm,jp,km,kp; *dxx_norm = dx/sqrt(ddx^2+ddy^2+ddz^2).  */
/* ------------------------------------------------------------------------- */
void orc_grid_from_surface(const double *surfX, int nSurfNode, double dx, int dd,
                           int n[3], double xLo[3], int box[6], double *dxx_norm)
{
    double mn[3], mx[3];
    for (int c = 0; c < 3; ++c) {
        mn[c] = mx[c] = surfX[(size_t)c * nSurfNode];
        for (int q = 1; q < nSurfNode; ++q) {
            double v = surfX[(size_t)c * nSurfNode + q];
            if (v > mx[c]) mx[c] = v;
            if (v < mn[c]) mn[c] = v;
        }
    }
    double dd3[3];
    for (int c = 0; c < 3; ++c) {
        dd3[c] = mx[c] - mn[c];                                   /* :135-137 */
        n[c] = (int)ceil((mx[c] - mn[c]) / dx) + 1;               /* :143-145 */
        n[c] = n[c] + 2 * dd;                                     /* :151-153 */
        xLo[c] = mn[c] - dd * dx;                                 /* :156 */
        box[2 * c] = (int)floor((mn[c] - xLo[c]) / dx) - 3;       /* :180-182 */
        box[2 * c + 1] = (int)floor((mx[c] - xLo[c]) / dx) + 3;   /* :184-186 */
    }
    *dxx_norm = dx / sqrt(dd3[0] * dd3[0] + dd3[1] * dd3[1] + dd3[2] * dd3[2]); /* :301 */
}

/* ------------------------------------------------------------------------- */
/* Inside/outside sign search, set3d.f90:196-268.                             */
/* surfX (nSurfNode,3) column-major fp64; surfElem (nSurfElem,3) column-major  */
/* int32, 1-based.  phi must be pre-filled by the caller (reference: 1.).     */
/* ------------------------------------------------------------------------- */
void orc_sign_init(double *phi, int nx, int ny, int nz, const double xLo[3], double dx,
                   const double *surfX, int nSurfNode, const int32_t *surfElem, int nSurfElem,
                   int im, int ip, int jm, int jp, int km, int kp)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
    (void)nz;
    double *cen = (double *)malloc((size_t)3 * nSurfElem * sizeof(double));
#define SX(n, c) surfX[(size_t)((n) - 1) + (size_t)(c) * nSurfNode]
#define SE(n, c) surfElem[(size_t)(n) + (size_t)(c) * nSurfElem]
    for (int n = 0; n < nSurfElem; ++n) {                          /* :199-215 */
        int n1 = SE(n, 0), n2 = SE(n, 1), n3 = SE(n, 2);
        for (int c = 0; c < 3; ++c)
            cen[(size_t)n + (size_t)c * nSurfElem] = (SX(n1, c) + SX(n2, c) + SX(n3, c)) / 3.;
    }
    for (int i = im; i <= ip; ++i)
        for (int j = jm; j <= jp; ++j)
            for (int k = km; k <= kp; ++k) {
                double gX = xLo[0] + i * dx, gY = xLo[1] + j * dx, gZ = xLo[2] + k * dx; /* :168-170 */
                double minD = 100000.;
                int fN = 0;
                for (int n = 0; n < nSurfElem; ++n) {              /* :224-236 */
                    double pX = cen[n], pY = cen[(size_t)n + nSurfElem], pZ = cen[(size_t)n + 2 * (size_t)nSurfElem];
                    double dis = sqrt((pX - gX) * (pX - gX) + (pY - gY) * (pY - gY) + (pZ - gZ) * (pZ - gZ));
                    if (dis < minD) { minD = dis; fN = n; }
                }
                int n1 = SE(fN, 0), n2 = SE(fN, 1), n3 = SE(fN, 2);
                double A1 = SX(n1, 0) - gX, A2 = SX(n1, 1) - gY, A3 = SX(n1, 2) - gZ;  /* :242-250 */
                double B1 = SX(n2, 0) - gX, B2 = SX(n2, 1) - gY, B3 = SX(n2, 2) - gZ;
                double C1 = SX(n3, 0) - gX, C2 = SX(n3, 1) - gY, C3 = SX(n3, 2) - gZ;
                double pSx = A2 * B3 - A3 * B2;                                        /* :253-255 */
                double pSy = -(A1 * B3 - B1 * A3);
                double pSz = A1 * B2 - B1 * A2;
                double pS = -(pSx * C1 + pSy * C2 + pSz * C3);                         /* :258 */
                phi[IDX(i, j, k)] = orc_phisign(pS, dx, 1.);                           /* :260-264 */
            }
#undef SX
#undef SE
    free(cen);
}

/* ------------------------------------------------------------------------- */
/* stlRead, subs.f90:17-121 (binary STL, vertex de-duplication with the        */
/* reference's search window: 1..nSurfNode where nSurfNode is refreshed only   */
/* after each triangle, :70-93; the sentinel 1000000. fills slots 1..ntri).    */
/* Two calls: orc_stl_ntri to size the buffers, orc_stl_read to fill them.     */
/* surfX_out: capacity 3*ntri*3 doubles, filled as (nSurfNode,3) column-major  */
/* surfElem_out: (ntri,3) column-major, 1-based.                               */
/* ------------------------------------------------------------------------- */
int orc_stl_ntri(const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    int32_t ntri = -1;
    if (fseek(fp, 80, SEEK_SET) != 0 || fread(&ntri, 4, 1, fp) != 1) ntri = -1;
    fclose(fp);
    return ntri;
}

int orc_stl_read(const char *path, double *surfX_out, int32_t *surfElem_out, int *nSurfNode_out)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    int32_t ntri;
    if (fseek(fp, 80, SEEK_SET) != 0 || fread(&ntri, 4, 1, fp) != 1) { fclose(fp); return -1; }
    float *tri = (float *)malloc((size_t)ntri * 9 * sizeof(float));
    for (int t = 0; t < ntri; ++t) {
        float rec[12]; uint16_t pad;
        if (fread(rec, 4, 12, fp) != 12 || fread(&pad, 2, 1, fp) != 1) { free(tri); fclose(fp); return -1; }
        memcpy(tri + (size_t)9 * t, rec + 3, 9 * sizeof(float));
    }
    fclose(fp);
    size_t cap = (size_t)ntri * 5 > 3 ? (size_t)ntri * 5 : 3;
    float *nodes = (float *)malloc(cap * 3 * sizeof(float));
    for (size_t q = 0; q < cap * 3; ++q) nodes[q] = 0.f;
    for (int q = 0; q < ntri; ++q) nodes[3 * q] = nodes[3 * q + 1] = nodes[3 * q + 2] = 1000000.f; /* :62-66 */
    int nSurfNode = 3, kcount = 0;                                                                 /* :70-71 */
    size_t iv = 0;
    for (int n = 0; n < ntri; ++n) {
        for (int p = 0; p < 3; ++p) {
            const float *v = tri + 3 * iv;
            int share = 0;
            for (int kk = 1; kk <= nSurfNode; ++kk) {                                              /* :75-82 */
                const float *w = nodes + 3 * (size_t)(kk - 1);
                if (((double)fabsf(w[0] - v[0]) < 1.e-13) && ((double)fabsf(w[1] - v[1]) < 1.e-13) &&
                    ((double)fabsf(w[2] - v[2]) < 1.e-13)) { share = kk; break; }
            }
            if (share > 0) surfElem_out[(size_t)n + (size_t)p * ntri] = share;
            else {
                kcount = kcount + 1;
                memcpy(nodes + 3 * (size_t)(kcount - 1), v, 3 * sizeof(float));
                surfElem_out[(size_t)n + (size_t)p * ntri] = kcount;
            }
            iv++;
        }
        nSurfNode = kcount;                                                                        /* :92 */
    }
    for (int q = 0; q < nSurfNode; ++q)
        for (int c = 0; c < 3; ++c)
            surfX_out[(size_t)q + (size_t)c * nSurfNode] = (double)nodes[3 * (size_t)q + c];       /* :99-103 */
    *nSurfNode_out = nSurfNode;
    free(tri); free(nodes);
    return ntri;
}

/* =========================================================================
 * Surface-node projection ("Advect Nodes"), set3d.f90:465-501  
 * ========================================================================= */

/* firstDeriv, order == 8 branch, subs.f90:311-347 + :357-364.  Literal, including the typo of
 * :346 (phi(i,jp1,k)*aa7 instead of jp2).  Writes gradPhi(i,j,k,1:3) and returns gMM. */
double orc_firstderiv8(int i, int j, int k, int nx, int ny, int nz, double dx, const double *phi, double *gradPhi)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1, np = sx * sy * ((size_t)nz + 1);
    const double aa1 = 1. / 280., aa2 = -4. / 105., aa3 = 1. / 5., aa4 = -4. / 5.;
    const double aa6 = 4. / 5, aa7 = -1. / 5., aa8 = 4. / 105., aa9 = -1. / 280.;
#define P(a, b, c) phi[IDX(a, b, c)]
    const double phiX = (P(i - 4, j, k) * aa1 + P(i - 3, j, k) * aa2 + P(i - 2, j, k) * aa3 + P(i - 1, j, k) * aa4 +
                         P(i + 1, j, k) * aa6 + P(i + 2, j, k) * aa7 + P(i + 3, j, k) * aa8 + P(i + 4, j, k) * aa9) / dx;
    const double phiY = (P(i, j - 4, k) * aa1 + P(i, j - 3, k) * aa2 + P(i, j - 2, k) * aa3 + P(i, j - 1, k) * aa4 +
                         P(i, j + 1, k) * aa6 + P(i, j + 1, k) * aa7 + P(i, j + 3, k) * aa8 + P(i, j + 4, k) * aa9) / dx;
    const double phiZ = (P(i, j, k - 4) * aa1 + P(i, j, k - 3) * aa2 + P(i, j, k - 2) * aa3 + P(i, j, k - 1) * aa4 +
                         P(i, j, k + 1) * aa6 + P(i, j, k + 2) * aa7 + P(i, j, k + 3) * aa8 + P(i, j, k + 4) * aa9) / dx;
#undef P
    gradPhi[IDX(i, j, k)] = phiX;
    gradPhi[IDX(i, j, k) + np] = phiY;
    gradPhi[IDX(i, j, k) + 2 * np] = phiZ;
    double gMM = phiX * phiX + phiY * phiY + phiZ * phiZ;
    return sqrt(gMM);
}

/* set3d.f90:469-478: gradPhi (zeroed at :372) gets the order-8 derivative on the stencil band.
 * A band cell closer than 4 points to the boundary makes the reference read outside phi's bounds (it does so
 * on its own cube40 input: the 8.1*dx band reaches to 2 points from the boundary); what it reads there is
 * undefined, so those entries are poisoned with NaN here and any node that would interpolate from one is
 * reported (-3) instead of being given a made-up value.  Returns the number of poisoned cells. */
int orc_gradphi_band8(int nx, int ny, int nz, double dx, const double *phi, const int32_t *phiSB, double *gradPhi)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1, np = sx * sy * ((size_t)nz + 1);
    int poisoned = 0;
    for (int i = 0; i <= nx; ++i)
        for (int j = 0; j <= ny; ++j)
            for (int k = 0; k <= nz; ++k)
                if (phiSB[IDX(i, j, k)] == 1) {
                    if (i < 4 || j < 4 || k < 4 || i > nx - 4 || j > ny - 4 || k > nz - 4) {
                        for (int c = 0; c < 3; ++c) gradPhi[IDX(i, j, k) + (size_t)c * np] = NAN;
                        ++poisoned;
                    } else
                        orc_firstderiv8(i, j, k, nx, ny, nz, dx, phi, gradPhi);
                }
    return poisoned;
}

/* setPhiSurf for ONE node, subs.f90:1078-1166.  surfX / gradPhiSurf are (nSurfNode,3) column-major.
 * Returns -5 if the node's cell lies outside the grid (out-of-bounds read in the reference). */
static int set_phi_surf_node(int n, const double xLo[3], int nx, int ny, int nz, double dx, double *phiSurf, const double *phi,
                             int nSurfNode, const double *surfX, double *gradPhiSurf, const double *gradPhi)
{
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1, np = sx * sy * ((size_t)nz + 1);
    const double minX = xLo[0], minY = xLo[1], minZ = xLo[2];
    const double x = surfX[n], y = surfX[n + (size_t)nSurfNode], z = surfX[n + 2 * (size_t)nSurfNode];
    const double fi = floor((x - minX) / dx), fj = floor((y - minY) / dx), fk = floor((z - minZ) / dx);
    if (!(fi >= 0 && fi <= nx - 1 && fj >= 0 && fj <= ny - 1 && fk >= 0 && fk <= nz - 1)) return -5;
    const int i0 = (int)fi, j0 = (int)fj, k0 = (int)fk;
    const double x0 = i0 * dx + minX, y0 = j0 * dx + minY, z0 = k0 * dx + minZ;
    const int i1 = i0 + 1, j1 = j0 + 1, k1 = k0 + 1;
    const double x1 = i1 * dx + minX, y1 = j1 * dx + minY, z1 = k1 * dx + minZ;
    const double xd = (x - x0) / (x1 - x0), yd = (y - y0) / (y1 - y0), zd = (z - z0) / (z1 - z0);
    for (int m = 0; m < 8; ++m)                                   /* poisoned corner: see orc_gradphi_band8 */
        if (isnan(gradPhi[IDX(i0 + (m & 1), j0 + ((m >> 1) & 1), k0 + ((m >> 2) & 1))])) return -3;
    double c00, c10, c01, c11, c0, c1;
#define TRI(F, off)                                                                         \
    c00 = F[IDX(i0, j0, k0) + off] * (1. - xd) + F[IDX(i1, j0, k0) + off] * xd;            \
    c10 = F[IDX(i0, j1, k0) + off] * (1. - xd) + F[IDX(i1, j1, k0) + off] * xd;            \
    c01 = F[IDX(i0, j0, k1) + off] * (1. - xd) + F[IDX(i1, j0, k1) + off] * xd;            \
    c11 = F[IDX(i0, j1, k1) + off] * (1. - xd) + F[IDX(i1, j1, k1) + off] * xd;            \
    c0 = c00 * (1. - yd) + c10 * yd;                                                        \
    c1 = c01 * (1. - yd) + c11 * yd;
    TRI(phi, 0)
    phiSurf[n] = c0 * (1. - zd) + c1 * zd;
    double gs[3];
    for (int c = 0; c < 3; ++c) {
        TRI(gradPhi, (size_t)c * np)
        gs[c] = -(c0 * (1. - zd) + c1 * zd);
    }
#undef TRI
    const double gradMag2 = gs[0] * gs[0] + gs[1] * gs[1] + gs[2] * gs[2];
    if (gradMag2 < 1.E-7) {
        gs[0] = gs[1] = gs[2] = 0.;
    } else {
        const double gradPhiMag = sqrt(gradMag2);
        gs[0] = gs[0] / gradPhiMag; gs[1] = gs[1] / gradPhiMag; gs[2] = gs[2] / gradPhiMag;
    }
    for (int c = 0; c < 3; ++c) gradPhiSurf[n + (size_t)c * nSurfNode] = gs[c];
    return 0;
}

/* SUBROUTINE setPhiSurf, subs.f90:1057-1170: all nodes. */
int orc_set_phi_surf(const double xLo[3], int nx, int ny, int nz, double dx, double *phiSurf, const double *phi,
                     int nSurfNode, const double *surfX, double *gradPhiSurf, const double *gradPhi)
{
    for (int n = 0; n < nSurfNode; ++n) {
        const int rc = set_phi_surf_node(n, xLo, nx, ny, nz, dx, phiSurf, phi, nSurfNode, surfX, gradPhiSurf, gradPhi);
        if (rc) return rc;
    }
    return 0;
}

/* The node loop of set3d.f90:480-501.  literal != 0: exactly as written -- setPhiSurf over ALL nodes after
 * every single move (O(iter * nNode^2): only for small inputs); literal == 0: only the moved node is
 * re-interpolated, which is equivalent because setPhiSurf is a pure function of each node's own position
 * (checked by tests/test_oracle_pins.py).  surfXX holds surfX on entry (:485).  Returns 0, or -5.
 * moves (may be NULL) receives the number of node moves executed. */
int orc_advect_nodes(const double xLo[3], int nx, int ny, int nz, double dx, const double *phi, const double *gradPhi,
                     int nSurfNode, double *surfXX, double *phiSurf, double *gradPhiSurf, int iter, int literal, long long *moves)
{
    long long nm = 0;
    for (int n = 0; n < nSurfNode; ++n) phiSurf[n] = 0.;                        /* set3d.f90:483 */
    int rc = orc_set_phi_surf(xLo, nx, ny, nz, dx, phiSurf, phi, nSurfNode, surfXX, gradPhiSurf, gradPhi);   /* :487 */
    if (rc) return rc;
    for (int k = 1; k <= iter; ++k)                                             /* :491 */
        for (int n = 0; n < nSurfNode; ++n)
            if (phiSurf[n] > 1E-13) {                                           /* :493 */
                for (int c = 0; c < 3; ++c)
                    surfXX[n + (size_t)c * nSurfNode] = surfXX[n + (size_t)c * nSurfNode] + phiSurf[n] * gradPhiSurf[n + (size_t)c * nSurfNode];
                ++nm;
                if (literal) rc = orc_set_phi_surf(xLo, nx, ny, nz, dx, phiSurf, phi, nSurfNode, surfXX, gradPhiSurf, gradPhi);
                else rc = set_phi_surf_node(n, xLo, nx, ny, nz, dx, phiSurf, phi, nSurfNode, surfXX, gradPhiSurf, gradPhi);
                if (rc) return rc;
            }
    if (moves) *moves = nm;
    return 0;
}
