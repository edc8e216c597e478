// lsf_cell.cuh -- per-cell arithmetic of the WENO5 Hamilton-Jacobi reinitialisation update.
//
// Two arithmetic policies compute the same update (reference: weno subs.f90:489-711, phiSign
// subs.f90:152-172, Euler update subs.f90:747-750):
//   ExactArith : the reference's operation order with round-to-nearest mul/add/div/sqrt and no
//                FMA contraction -> bit-identical to `gfortran -O3 -fdefault-real-8` semantics.
//   FastArith  : algebraically equivalent form built for the FP64 pipe: first/second differences
//                are kept unscaled (1/dx factored out), the six 1/(eps+IS)^2 divisions and four
//                weight divisions per direction collapse into ONE reciprocal per side
//                (w0 = q1 q2 / D, w2 = 3 q0 q1 / D, D = q1 q2 + 6 q0 q2 + 3 q0 q1, q = (eps+IS)^2),
//                reciprocals are MUFU seed + 2 Newton steps, and FMAs are allowed.  The 1.E-99
//                epsilon floor (subs.f90:533) becomes 1.E-60 (in dx^2-scaled units) so the products
//                of squares cannot underflow in exactly flat regions; where that floor matters the
//                WENO correction it weights is < 1e-30.  Deviation from the reference <= 1e-10
//                (measured ~1e-14 after 2155 sweeps, tests/test_gpu_parity.py).
//
// The header is also compiled by g++ (tests/emu) so the schedule logic can be checked on a CPU.
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define LSF_HD __host__ __device__ __forceinline__
#else
#define LSF_HD inline
#endif

namespace lsf {

// gfortran MAX/MIN on reals: a NaN first operand is replaced by the second.
LSF_HD double fmax_f(double a, double b) { return (b > a || a != a) ? b : a; }
LSF_HD double fmin_f(double a, double b) { return (b < a || a != a) ? b : a; }

template <class T>
struct CellConstT {
    T dx;       // grid spacing
    T inv_dx;   // 1/dx
    T k12;      // 1/(12 dx)
    T dx2;      // dx*dx            (phiSign: dxx*dxx, subs.f90:169)
    T h;        // pseudo-time step (subs.f90:750)
};
typedef CellConstT<double> CellConst;

// ------------------------------------------------------------------------------------------
struct ExactArith {
    typedef double real;
#if defined(__CUDA_ARCH__)
    static LSF_HD double mul(double a, double b) { return __dmul_rn(a, b); }
    static LSF_HD double add(double a, double b) { return __dadd_rn(a, b); }
    static LSF_HD double sub(double a, double b) { return __dsub_rn(a, b); }
    static LSF_HD double div(double a, double b) { return __ddiv_rn(a, b); }
    static LSF_HD double sqr(double a) { return __dsqrt_rn(a); }
#else
    static LSF_HD double mul(double a, double b) { return a * b; }   // host: build with -ffp-contract=off
    static LSF_HD double add(double a, double b) { return a + b; }
    static LSF_HD double sub(double a, double b) { return a - b; }
    static LSF_HD double div(double a, double b) { return a / b; }
    static LSF_HD double sqr(double a) { return sqrt(a); }
#endif

    // one direction of the high-order branch; v[0..6] = phi at physical offsets -3..+3.
    // YQ reproduces subs.f90:576 (y direction: p5 = (phi(j+3)-phi(j+3))/dx).
    template <bool YQ>
    static LSF_HD void weno_dir(const double v[7], const CellConst &cc, double &dminus, double &dplus)
    {
        const double dx = cc.dx;
        const double m3 = v[0], m2 = v[1], m1 = v[2], c0 = v[3], p1v = v[4], p2v = v[5], p3v = v[6];
        const double ap = div(add(sub(p3v, mul(2., p2v)), p1v), dx);
        const double am = div(add(sub(m3, mul(2., m2)), m1), dx);
        const double bp = div(add(sub(p2v, mul(2., p1v)), c0), dx);
        const double bm = div(add(sub(m2, mul(2., m1)), c0), dx);
        const double cp = div(add(sub(p1v, mul(2., c0)), m1), dx);
        const double cm = cp, dp = bm, dm = bp;
#define LSF_IS(x, y) add(mul(mul(13., (x)), (x)), mul(mul(3., (y)), (y)))
        const double IS0p = LSF_IS(sub(ap, bp), sub(ap, mul(3., bp)));
        const double IS0m = LSF_IS(sub(am, bm), sub(am, mul(3., bm)));
        const double IS1p = LSF_IS(sub(bp, cp), add(bp, cp));
        const double IS1m = LSF_IS(sub(bm, cm), add(bm, cm));
        const double IS2p = LSF_IS(sub(cp, dp), sub(mul(3., cp), dp));
        const double IS2m = LSF_IS(sub(cm, dm), sub(mul(3., cm), dm));
#undef LSF_IS
        const double p0 = div(sub(m2, m3), dx);
        const double p1 = div(sub(m1, m2), dx);
        const double p2 = div(sub(c0, m1), dx);
        const double p3 = div(sub(p1v, c0), dx);
        const double p4 = div(sub(p2v, p1v), dx);
        const double p5 = YQ ? div(sub(p3v, p3v), dx) : div(sub(p3v, p2v), dx);
        const double q0 = mul(p0, p0), q1 = mul(p1, p1), q2 = mul(p2, p2), q3 = mul(p3, p3),
                     q4 = mul(p4, p4), q5 = mul(p5, p5);
        const double epsp = add(mul(1.E-6, fmax_f(q1, fmax_f(q2, fmax_f(q3, fmax_f(q4, q5))))), 1.E-99);
        const double epsm = add(mul(1.E-6, fmax_f(q0, fmax_f(q1, fmax_f(q2, fmax_f(q3, q4))))), 1.E-99);
#define LSF_AL(c, eps, IS) div((c), mul(add((eps), (IS)), add((eps), (IS))))
        const double a0p = LSF_AL(1., epsp, IS0p), a0m = LSF_AL(1., epsm, IS0m);
        const double a1p = LSF_AL(6., epsp, IS1p), a1m = LSF_AL(6., epsm, IS1m);
        const double a2p = LSF_AL(3., epsp, IS2p), a2m = LSF_AL(3., epsm, IS2m);
#undef LSF_AL
        const double sp = add(add(a0p, a1p), a2p), sm = add(add(a0m, a1m), a2m);
        const double w0p = div(a0p, sp), w0m = div(a0m, sm);
        const double w2p = div(a2p, sp), w2m = div(a2m, sm);
        const double third = 1. / 3., sixth = 1. / 6., twelfth = 1. / 12.;
        const double PWp = add(mul(mul(third, w0p), add(sub(ap, mul(2., bp)), cp)),
                               mul(mul(sixth, sub(w2p, 0.5)), add(sub(bp, mul(2., cp)), dp)));
        const double PWm = add(mul(mul(third, w0m), add(sub(am, mul(2., bm)), cm)),
                               mul(mul(sixth, sub(w2m, 0.5)), add(sub(bm, mul(2., cm)), dm)));
        const double cen = mul(twelfth, sub(add(add(-p1, mul(7., p2)), mul(7., p3)), p4));
        dminus = sub(cen, PWm);
        dplus = add(cen, PWp);
    }

    // low-order one-sided differences, subs.f90:657-662
    static LSF_HD void lo_dir(double vm, double vc, double vp, const CellConst &cc, double &dminus, double &dplus)
    {
        dminus = div(sub(vc, vm), cc.dx);
        dplus = div(sub(vp, vc), cc.dx);
    }

    // Godunov selection + |grad phi|, subs.f90:667-702.  g[3] receives gradX,gradY,gradZ.
    static LSF_HD double godunov(double phic, double a, double b, double c, double d, double e, double f, double g[3], const CellConst &)
    {
        const double pa = fmax_f(a, 0.), pb = fmax_f(b, 0.), pc = fmax_f(c, 0.);
        const double pd = fmax_f(d, 0.), pe = fmax_f(e, 0.), pf = fmax_f(f, 0.);
        const double na = fmin_f(a, 0.), nb = fmin_f(b, 0.), nc = fmin_f(c, 0.);
        const double nd = fmin_f(d, 0.), ne = fmin_f(e, 0.), nf = fmin_f(f, 0.);
        if (phic > 0.) {
            g[0] = fmax_f(mul(pa, pa), mul(nb, nb));
            g[1] = fmax_f(mul(pc, pc), mul(nd, nd));
            g[2] = fmax_f(mul(pe, pe), mul(nf, nf));
        } else {
            g[0] = fmax_f(mul(pb, pb), mul(na, na));
            g[1] = fmax_f(mul(pd, pd), mul(nc, nc));
            g[2] = fmax_f(mul(pf, pf), mul(ne, ne));
        }
        return sqr(add(add(g[0], g[1]), g[2]));
    }

    // phiSign (subs.f90:169) + Euler update (subs.f90:749-750)
    static LSF_HD double update(double phic, double phiS, double gM, const CellConst &cc, bool &sens)
    {
        sens = false;
        const double sgn = div(phiS, sqr(add(mul(phiS, phiS), mul(mul(cc.dx, cc.dx), gM))));
        const double k1 = mul(sgn, sub(1., gM));
        return add(phic, mul(cc.h, k1));
    }
};

// ------------------------------------------------------------------------------------------
struct FastArith {
    typedef double real;
    // ---- select-style helpers that stay off the FP64 pipe ---------------------------------------
    // (a double fmax/fmin costs a DSETP on the FP64 pipe plus ~6 integer instructions of NaN
    // fix-up; here the operands are known to be sign-definite, so integer compares on the bit
    // patterns are exact.  NaN inputs are not canonicalised: a NaN still ends in a NaN RMS.)
#if defined(__CUDA_ARCH__)
    static LSF_HD double pos_part(double x) { return __double2hiint(x) < 0 ? 0.0 : x; }          // max(x, 0)
    static LSF_HD double neg_part(double x) { return __double2hiint(x) < 0 ? x : 0.0; }          // min(x, 0), -0 for x = 0
    static LSF_HD double dabs(double x) { return __longlong_as_double(__double_as_longlong(x) & 0x7fffffffffffffffLL); }
    static LSF_HD double max_nn(double a, double b)                                              // a, b >= 0
    {
#if defined(LSF_MAX_DSETP)      // one DSETP on the FP64 pipe + 2 selects instead of 2 ISETP + 2 selects (experiment)
        return a > b ? a : b;
#else
        const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
        return __longlong_as_double(ia > ib ? ia : ib);
#endif
    }
    // 1/sqrt(x) and sqrt(x) without the library's special-case path (a CALL that 15 % of the warps of the bench grid took:
    // far from the surface phi is constant and the sum of squares under the root is exactly 0): MUFU seed (relative error
    // < 2^-20) + one third-order step  y1 = y0 + y0 e (1/2 + 3/8 e),  e = 1 - x y0^2  ->  relative error ~ e^3, below 1 ulp.
    // x = 0 gives y0 = +inf and a NaN from 0 * inf: exactly what update() needs (0 * rsqrt(0) = NaN is the reference's 0/0,
    // subs.f90:169); fsqrt returns 0 for zero and denormal arguments.
    static LSF_HD double rsq(double x)
    {
#if defined(LSF_LIB_SQRT)
        return rsqrt(x);
#endif
        double y0;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
        const double e = fma(-x, y0 * y0, 1.0);
        return fma(fma(0.375, e, 0.5), y0 * e, y0);
    }
#if defined(LSF_LIB_SQRT)
    static LSF_HD double fsqrt(double x) { return sqrt(x); }
#else
    static LSF_HD double fsqrt(double x) { return __double2hiint(x) < 0x00100000 ? 0.0 : x * rsq(x); }
#endif
    static LSF_HD double flip_if(double x, bool f) { return __hiloint2double(__double2hiint(x) ^ (f ? (int)0x80000000 : 0), __double2loint(x)); }
#else
    static LSF_HD double flip_if(double x, bool f) { return f ? -x : x; }
    static LSF_HD double pos_part(double x) { return x > 0. ? x : 0.0; }
    static LSF_HD double neg_part(double x) { return x < 0. ? x : 0.0; }
    static LSF_HD double dabs(double x) { return fabs(x); }
    static LSF_HD double max_nn(double a, double b) { return a > b ? a : b; }
    static LSF_HD double rsq(double x) { return 1.0 / sqrt(x); }
    static LSF_HD double fsqrt(double x) { return sqrt(x); }
#endif

#if defined(__CUDA_ARCH__)
    static LSF_HD double dabs_i(double x)
    {
        int hi;
        asm("and.b32 %0, %1, 0x7fffffff;" : "=r"(hi) : "r"(__double2hiint(x)));
        return __hiloint2double(hi, __double2loint(x));
    }
#endif

    // ---- max |e| for the WENO epsilon (subs.f90:533: eps = 1e-6 max(e^2) + 1e-99) ---------------------------------
    // eps only regularises the weights: an error of relative size d in max|e| moves a weight by at most ~d and the
    // one-sided derivative by d times a third difference.  LSF_EPS_MODE 0 takes the exact maximum (64-bit integer
    // compares on the bit patterns: ISETP + ISETP.EX + 2 SEL per max, plus one LOP3 per |e|); 1 compares 32-bit keys
    // made of exponent + 21 mantissa bits (one funnel shift per value, one VIMNMX per max: the maximum is off by
    // < 2^-22 relative); 2 compares the values rounded to float (one conversion per value, one FMNMX per max: < 2^-24).
    // Measured deviation from EXACT arithmetic: tools/fast_accuracy.cpp (cube40, 2155 sweeps) and bench.py `parity`.
#ifndef LSF_EPS_MODE
#define LSF_EPS_MODE 1
#endif
#if defined(__CUDA_ARCH__)
    static LSF_HD unsigned ekey(double x) { return __funnelshift_l((unsigned)__double2loint(x), (unsigned)__double2hiint(x), 1); }
    static LSF_HD double eunkey(unsigned k) { return __hiloint2double((int)(k >> 1), (int)((k << 31) | 0x40000000u)); }   // midpoint of the key's interval
    static LSF_HD unsigned umax2(unsigned a, unsigned b) { return max(a, b); }
    static LSF_HD float efl(double x) { return __double2float_rn(fabs(x)); }
#else
    static LSF_HD unsigned ekey(double x) { unsigned long long b; memcpy(&b, &x, 8); return (unsigned)((b << 1) >> 32); }
    static LSF_HD double eunkey(unsigned k) { const unsigned long long b = ((unsigned long long)k << 31) | 0x40000000ull; double x; memcpy(&x, &b, 8); return x; }
    static LSF_HD unsigned umax2(unsigned a, unsigned b) { return a > b ? a : b; }
    static LSF_HD float efl(double x) { return (float)fabs(x); }
#endif
    // mp = max(|e1|..|e5|) (|e1|..|e4| if YQ: subs.f90:576), mm = max(|e0|..|e4|)
    template <bool YQ>
    static LSF_HD void eps_max(double e0, double e1, double e2, double e3, double e4, double e5, double &mp, double &mm)
    {
#if LSF_EPS_MODE == 1
        const unsigned mc = umax2(umax2(ekey(e1), ekey(e2)), umax2(ekey(e3), ekey(e4)));
        mp = eunkey(YQ ? mc : umax2(mc, ekey(e5)));
        mm = eunkey(umax2(mc, ekey(e0)));
#elif LSF_EPS_MODE == 2
        const float mc = fmaxf(fmaxf(efl(e1), efl(e2)), fmaxf(efl(e3), efl(e4)));
        mp = (double)(YQ ? mc : fmaxf(mc, efl(e5)));
        mm = (double)fmaxf(mc, efl(e0));
#elif !defined(LSF_NO_ABS_INT) && defined(__CUDA_ARCH__)
        // |e| through an opaque 32-bit AND on the high word (the 64-bit sign mask is turned into a
        // DADD |x| by the compiler, i.e. back onto the FP64 pipe)
        const double mc = max_nn(max_nn(dabs_i(e1), dabs_i(e2)), max_nn(dabs_i(e3), dabs_i(e4)));
        mp = YQ ? mc : max_nn(mc, dabs_i(e5));
        mm = max_nn(mc, dabs_i(e0));
#else
        const double mc = max_nn(max_nn(dabs(e1), dabs(e2)), max_nn(dabs(e3), dabs(e4)));
        mp = YQ ? mc : max_nn(mc, dabs(e5));
        mm = max_nn(mc, dabs(e0));
#endif
    }

    // 1/x for the Jiang-Shu weights: MUFU seed (upper 32 bits of x: relative error <~ 2^-20) + ONE Newton step ->
    // relative error <~ 1e-12.  The weights multiply third differences, so the one-sided derivatives move by
    // < 1e-12 * |WENO correction| (< 2e-13 even at a kink); a phi update is h times that.  LSF_RCP_NEWTON2 restores
    // the second step (round 1).
    static LSF_HD double rcp(double x)
    {
#if defined(__CUDA_ARCH__)
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        double e = fma(-x, r, 1.0);
        r = fma(r, e, r);
#if defined(LSF_RCP_NEWTON2)
        e = fma(-x, r, 1.0);
        r = fma(r, e, r);
#endif
        return r;
#else
        return 1.0 / x;
#endif
    }

    // One side of one direction.  E_k = (eps + IS_k)/3 (a common factor cancels in the weights).  With q_k = E_k^2,
    // n0 = q1 q2, x = q0 q2, P = q0 q1 and D = n0 + 6 x + 3 P the Jiang-Shu weights are w0 = n0/D, w2 = 3P/D, and the
    // WENO correction  2 w0 A + (w2 - 1/2) s  equals  r (2 n0 A + 3 P s) - s/2,  r = 1/D.  Returns r*(n0*A + P*s15),
    // s15 = 1.5 s: HALF the correction plus s/4; the caller multiplies by 4 inside its final FMA (12 dx times the
    // one-sided derivative = cen -/+ s +/- 4 * side).
    static LSF_HD double side(double E0, double E1, double E2, double A, double s15)
    {
        const double q0 = E0 * E0, q1 = E1 * E1, q2 = E2 * E2;
        const double n0 = q1 * q2, x = q0 * q2, P = q0 * q1;
        const double D = fma(6.0, x, fma(3.0, P, n0));
        return rcp(D) * fma(n0, A, P * s15);
    }

    // dminus / dplus are returned UNSCALED: 12 dx times the one-sided derivatives (godunov applies 1/(12 dx) once,
    // to |grad phi|, instead of six times here)
    template <bool YQ>
    static LSF_HD void weno_dir(const double v[7], const CellConst &cc, double &dminus, double &dplus)
    {
        (void)cc;
        const double e0 = v[1] - v[0], e1 = v[2] - v[1], e2 = v[3] - v[2];
        const double e3 = v[4] - v[3], e4 = v[5] - v[4], e5 = v[6] - v[5];
        const double am = e1 - e0, bm = e2 - e1, c = e3 - e2, bp = e4 - e3, ap = e5 - e4;
        const double tpa = ap - bp, tpb = bp - c, tmc = c - bm, tma = am - bm;
        double mp, mm;
        eps_max<YQ>(e0, e1, e2, e3, e4, e5, mp, mm);
        // everything below is (eps + IS)/3: eps/3 = (1e-6/3) max e^2 + tiny, IS/3 = (13/3) u^2 + v^2
        constexpr double k13 = 13.0 / 3.0, keps = 1.0e-6 / 3.0, tiny = 1.0e-60;
        const double epsp = fma(keps * mp, mp, tiny);
        const double epsm = fma(keps * mm, mm, tiny);
        const double kb = k13 * tpb, kc = k13 * tmc;
        double t;
        t = fma(-3.0, bp, ap); const double E0p = fma(k13 * tpa, tpa, fma(t, t, epsp));
        t = bp + c;            const double E1p = fma(t, t, fma(kb, tpb, epsp));
        t = fma(3.0, c, -bm);  const double E2p = fma(t, t, fma(kc, tmc, epsp));
        t = fma(-3.0, bm, am); const double E0m = fma(k13 * tma, tma, fma(t, t, epsm));
        t = bm + c;            const double E1m = fma(t, t, fma(kc, tmc, epsm));
        t = fma(3.0, c, -bp);  const double E2m = fma(t, t, fma(kb, tpb, epsm));
        const double s = tpb - tmc, s15 = 1.5 * s;
        const double cen = fma(7.0, e2 + e3, -(e1 + e4));
        dplus = fma(4.0, side(E0p, E1p, E2p, tpa - tpb, s15), cen - s);
        dminus = fma(-4.0, side(E0m, E1m, E2m, tma + tmc, s15), cen + s);
    }

    static LSF_HD void lo_dir(double vm, double vc, double vp, const CellConst &cc, double &dminus, double &dplus)
    {
        (void)cc;
        dminus = 12.0 * (vc - vm);          // unscaled like weno_dir: 12 dx times the difference quotient
        dplus = 12.0 * (vp - vc);
    }

    // a..f: 12 dx times the one-sided derivatives.  g[] receives the reference's gradX/Y/Z (squared upwind terms),
    // the return value is gM = sqrt(gradX+gradY+gradZ) (subs.f90:702)
    static LSF_HD double godunov(double phic, double a, double b, double c, double d, double e, double f, double g[3], const CellConst &cc)
    {
        // upwind pair per axis: phi>0 -> (max(a,0), min(b,0)) else (max(b,0), min(a,0))
        const bool pos = phic > 0.;
#if !defined(LSF_GODUNOV_SELECT)
        // phi <= 0 takes (max(b,0), min(a,0)); their squares are those of (min(-b,0), max(-a,0)): flip the signs of
        // a..f instead of swapping each pair (one XOR on the high word per value instead of two selects); the squares
        // and their maximum are bit for bit the same
        const double ax = flip_if(a, !pos), bx = flip_if(b, !pos);
        const double ay = flip_if(c, !pos), by = flip_if(d, !pos);
        const double az = flip_if(e, !pos), bz = flip_if(f, !pos);
#else
        const double ax = pos ? a : b, bx = pos ? b : a;
        const double ay = pos ? c : d, by = pos ? d : c;
        const double az = pos ? e : f, bz = pos ? f : e;
#endif
        const double x1 = pos_part(ax), x2 = neg_part(bx);
        const double y1 = pos_part(ay), y2 = neg_part(by);
        const double z1 = pos_part(az), z2 = neg_part(bz);
        const double u0 = max_nn(x1 * x1, x2 * x2), u1 = max_nn(y1 * y1, y2 * y2), u2 = max_nn(z1 * z1, z2 * z2);
        const double k2 = cc.k12 * cc.k12;
        g[0] = k2 * u0; g[1] = k2 * u1; g[2] = k2 * u2;      // only stored by the WG plane kernel; dead code elsewhere
        return cc.k12 * fsqrt((u0 + u1) + u2);
    }

    // `sens` flags an ill-conditioned update: d(sgn (1-gM))/d(gM) contains  k1 * dx^2 / (2 D)  with
    // D = phiS^2 + dx^2 gM, which blows up where the sign source AND the Godunov gradient vanish (a flat
    // extremum of phi next to the interface, e.g. an under-resolved gap).  There, 1e-13 differences between
    // this arithmetic and the reference's are amplified beyond the 1e-10 contract -- and the reference's own
    // 0/0 NaN (subs.f90:169, phiS == 0 and gM == 0 exactly) hinges on the last bit of gM.  LSF_ARITH_AUTO
    // reruns the call in EXACT arithmetic when any cell raises the flag.
    static LSF_HD double update(double phic, double phiS, double gM, const CellConst &cc, bool &sens)
    {
        // phiS / sqrt(phiS^2 + dx^2 gM): 0 * rsqrt(0) = NaN reproduces the reference's 0/0 (subs.f90:169)
        const double r = rsq(fma(phiS, phiS, cc.dx2 * gM));
        const double sgn = phiS * r;
        const double k1 = sgn * (1.0 - gM);
        const double amp = dabs(k1) * (cc.dx2 * (r * r));           // = 2 * |k1| dx^2 / (2 D)
        // threshold calibrated on the reference's two inputs (oracle run of every sweep): cube40 peaks at 400 in
        // its first two sweeps and is < 7 after 25 (FAST ends 2e-14 from the reference); twoCube10 -- which the
        // reference NaN-STOPs on -- shows 1.8e9 in sweep 0 and the zero-source/zero-gradient cell from sweep 94
        sens = !(amp <= 2000.0) || (phiS == 0.0 && gM < 1.0e-6);     // NaN amp (D == 0) is flagged too
        return fma(cc.h, k1, phic);
    }
};

// ------------------------------------------------------------------------------------------
// F32Arith: the optional single-precision mode (fp32 storage, 12 B per cell update,
// contract 1e-4 relative to the fp64 reference).  Same algebra as FastArith -- unscaled differences,
// one reciprocal per side for the Jiang-Shu weights -- with two changes forced by the fp32 exponent
// range: (1) the three E_k = eps + IS_k of a side are normalised by their sum before they are squared
// and multiplied pairwise (eps >= 1e-6 max(e^2) bounds E_k / sum(E) from below by ~1e-9, so no product
// of two squares leaves the normal range), (2) the 1.E-99 floor of subs.f90:533 becomes 1.E-30.
// max/abs/select are native single instructions in fp32, so none of FastArith's integer tricks.
struct F32Arith {
    typedef float real;
#if defined(__CUDA_ARCH__)
    // MUFU approximations (1-2 ulp), flush-to-zero: no denormal fix-up code around them
    static LSF_HD float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static LSF_HD float rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static LSF_HD float sqr(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
    static LSF_HD float rcp(float x) { return 1.0f / x; }
    static LSF_HD float rsq(float x) { return 1.0f / sqrtf(x); }
    static LSF_HD float sqr(float x) { return sqrtf(x); }
#endif
    // One side of one direction, FastArith's fused form: r (n0 A + P s15) = half the WENO correction plus s/4, with the three
    // E_k normalised by their sum first (see above: keeps the products of squares inside the fp32 range)
    static LSF_HD float side(float E0, float E1, float E2, float A, float s15)
    {
        const float rs = rcp((E0 + E1) + E2);
        const float n0e = E0 * rs, n1e = E1 * rs, n2e = E2 * rs;
        const float q0 = n0e * n0e, q1 = n1e * n1e, q2 = n2e * n2e;
        const float n0 = q1 * q2, x = q0 * q2, P = q0 * q1;
        const float D = fmaf(6.0f, x, fmaf(3.0f, P, n0));
        return rcp(D) * fmaf(n0, A, P * s15);
    }

    // dminus / dplus are returned UNSCALED like FastArith's: 12 dx times the one-sided derivatives (godunov applies 1/(12 dx) once)
    template <bool YQ>
    static LSF_HD void weno_dir(const float v[7], const CellConstT<float> &cc, float &dminus, float &dplus)
    {
        (void)cc;
        const float e0 = v[1] - v[0], e1 = v[2] - v[1], e2 = v[3] - v[2];
        const float e3 = v[4] - v[3], e4 = v[5] - v[4], e5 = v[6] - v[5];
        const float am = e1 - e0, bm = e2 - e1, c = e3 - e2, bp = e4 - e3, ap = e5 - e4;
        const float tpa = ap - bp, tpb = bp - c, tmc = c - bm, tma = am - bm;
        const float mc = fmaxf(fmaxf(fabsf(e1), fabsf(e2)), fmaxf(fabsf(e3), fabsf(e4)));
        const float mp = YQ ? mc : fmaxf(mc, fabsf(e5));          // subs.f90:576: p5 == 0 in y
        const float mm = fmaxf(mc, fabsf(e0));
        // everything below is (eps + IS)/3 (the common factor cancels in the weights): eps/3 = (1e-6/3) max e^2 + tiny,
        // IS/3 = (13/3) u^2 + v^2
        constexpr float k13 = 13.0f / 3.0f, keps = 1.0e-6f / 3.0f, tiny = 1.0e-30f;
        const float epsp = fmaf(keps * mp, mp, tiny);
        const float epsm = fmaf(keps * mm, mm, tiny);
        const float kb = k13 * tpb, kc = k13 * tmc;
        float t;
        t = fmaf(-3.0f, bp, ap); const float E0p = fmaf(k13 * tpa, tpa, fmaf(t, t, epsp));
        t = bp + c;              const float E1p = fmaf(t, t, fmaf(kb, tpb, epsp));
        t = fmaf(3.0f, c, -bm);  const float E2p = fmaf(t, t, fmaf(kc, tmc, epsp));
        t = fmaf(-3.0f, bm, am); const float E0m = fmaf(k13 * tma, tma, fmaf(t, t, epsm));
        t = bm + c;              const float E1m = fmaf(t, t, fmaf(kc, tmc, epsm));
        t = fmaf(3.0f, c, -bp);  const float E2m = fmaf(t, t, fmaf(kb, tpb, epsm));
        const float s = tpb - tmc, s15 = 1.5f * s;
        const float cen = fmaf(7.0f, e2 + e3, -(e1 + e4));
        dplus = fmaf(4.0f, side(E0p, E1p, E2p, tpa - tpb, s15), cen - s);
        dminus = fmaf(-4.0f, side(E0m, E1m, E2m, tma + tmc, s15), cen + s);
    }

    static LSF_HD void lo_dir(float vm, float vc, float vp, const CellConstT<float> &cc, float &dminus, float &dplus)
    {
        (void)cc;
        dminus = 12.0f * (vc - vm);          // unscaled like weno_dir
        dplus = 12.0f * (vp - vc);
    }

    // a..f: 12 dx times the one-sided derivatives; returns gM = |grad phi| (subs.f90:702)
    static LSF_HD float godunov(float phic, float a, float b, float c, float d, float e, float f, float g[3], const CellConstT<float> &cc)
    {
        const bool pos = phic > 0.f;
        const float x1 = fmaxf(pos ? a : b, 0.f), x2 = fminf(pos ? b : a, 0.f);
        const float y1 = fmaxf(pos ? c : d, 0.f), y2 = fminf(pos ? d : c, 0.f);
        const float z1 = fmaxf(pos ? e : f, 0.f), z2 = fminf(pos ? f : e, 0.f);
        const float u0 = fmaxf(x1 * x1, x2 * x2), u1 = fmaxf(y1 * y1, y2 * y2), u2 = fmaxf(z1 * z1, z2 * z2);
        const float k2 = cc.k12 * cc.k12;
        g[0] = k2 * u0; g[1] = k2 * u1; g[2] = k2 * u2;      // dead code in the fp32 kernels (no gradPhi outputs in this mode)
        return cc.k12 * sqr((u0 + u1) + u2);
    }

    // no conditioning guard in fp32 mode: the 1e-4 contract is far above the amplification FastArith guards against
    static LSF_HD float update(float phic, float phiS, float gM, const CellConstT<float> &cc, bool &sens)
    {
        sens = false;
        const float r = rsq(fmaf(phiS, phiS, cc.dx2 * gM));     // 0 * rsqrt(0) = NaN: the reference's 0/0 (subs.f90:169)
        const float k1 = (phiS * r) * (1.0f - gM);
        return fmaf(cc.h, k1, phic);
    }
};

// Full cell update from the three 7-point lines (physical orientation, index 3 = centre).
// hi = high-order branch condition of subs.f90:506.  Returns the new phi; g[3]/gM as in weno.
template <class AR>
LSF_HD typename AR::real reinit_cell(const typename AR::real vx[7], const typename AR::real vy[7],
                                     const typename AR::real vz[7], typename AR::real phiS, bool hi,
                                     const CellConstT<typename AR::real> &cc, typename AR::real g[3],
                                     typename AR::real &gM, bool &sens)
{
    typename AR::real a, b, c, d, e, f;
    if (hi) {
        AR::template weno_dir<false>(vx, cc, a, b);
        AR::template weno_dir<true>(vy, cc, c, d);
        AR::template weno_dir<false>(vz, cc, e, f);
    } else {
        AR::lo_dir(vx[2], vx[3], vx[4], cc, a, b);
        AR::lo_dir(vy[2], vy[3], vy[4], cc, c, d);
        AR::lo_dir(vz[2], vz[3], vz[4], cc, e, f);
    }
    gM = AR::godunov(vx[3], a, b, c, d, e, f, g, cc);
    return AR::update(vx[3], phiS, gM, cc, sens);
}

// The same update in two parts, for schedules that compute the x direction ahead of the y/z gathers (lsf_march.cuh,
// split step barrier): the operations and their order are those of reinit_cell, so ExactArith stays bit-identical.
template <class AR>
LSF_HD void reinit_dir_x(const typename AR::real vx[7], bool hi, const CellConstT<typename AR::real> &cc,
                         typename AR::real &a, typename AR::real &b)
{
    if (hi) AR::template weno_dir<false>(vx, cc, a, b);
    else AR::lo_dir(vx[2], vx[3], vx[4], cc, a, b);
}

template <class AR>
LSF_HD typename AR::real reinit_cell_rest(typename AR::real a, typename AR::real b, const typename AR::real vy[7],
                                          const typename AR::real vz[7], typename AR::real phiS, bool hi,
                                          const CellConstT<typename AR::real> &cc, typename AR::real g[3],
                                          typename AR::real &gM, bool &sens)
{
    typename AR::real c, d, e, f;
    if (hi) {
        AR::template weno_dir<true>(vy, cc, c, d);
        AR::template weno_dir<false>(vz, cc, e, f);
    } else {
        AR::lo_dir(vy[2], vy[3], vy[4], cc, c, d);
        AR::lo_dir(vz[2], vz[3], vz[4], cc, e, f);
    }
    gM = AR::godunov(vy[3], a, b, c, d, e, f, g, cc);
    return AR::update(vy[3], phiS, gM, cc, sens);
}

}  // namespace lsf
