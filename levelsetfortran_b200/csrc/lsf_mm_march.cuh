// lsf_mm_march.cuh -- one min/max-flow iteration (set3d.f90:399-431) as ONE kernel: skewed x-marching
// column tiles (the schedule of lsf_march.cuh with stencil half-width 1), out of place in global memory.
//
// Reference semantics of one iteration n:
//   pass A (Jacobi): for band cells, L = phiXX + phiYY + phiZZ of the iteration's OLD phi
//                    (secondDeriv, subs.f90:384-389; minMax curv, subs.f90:461);
//   pass B (in place, ascending i,j,k): pAve = (p + p(i-1) + p(i+1) + p(j+1) + p(j-1) + p(k+1) + p(k-1))/7.
//                    from the LIVE array (subs.f90:473-474) -> the three -1 neighbours are already
//                    updated; F = pAve < 0 ? min(L,0) : max(L,0) (subs.f90:477-481); phi += h1*F.
//   band of iteration n = abs(phi_old) < 4.1*dx (narrowBand of the previous iterate, set3d.f90:460;
//                    for n = 1 the caller's phiNB, set3d.f90:360).
// The weno call and the mixed derivatives of the reference loop feed nothing that is read.
//
// Here: A = phi_old (read only), B = phi_new (every interior cell written: band cells updated,
// the others copied), so that
//   * the Laplacian of OLD values and the Gauss-Seidel average of LIVE values are both available
//     at tile edges (an in-place kernel has lost the old value of an updated neighbour tile),
//   * phiN = phi (set3d.f90:454) costs nothing: after the iteration A IS phiN,
//   * algorithmic traffic is 16 B per grid point per iteration.
// Two shared-memory rings per tile (4 hyperplane slots each): So = old values, Sn = new values.
// -b/-c halo rows need the neighbour tiles' NEW values (from B, behind the progress flag) and their
// OLD values (from A); +b/+c halo rows need OLD values only.  Since A is never written there is no
// write-after-read constraint between tiles.  All arithmetic is explicitly rounded in the
// reference's order: bit-exact.
#pragma once
#include "lsf_march.cuh"

namespace lsf {

constexpr int MM_NSLOT = 4;
constexpr int MM_LOOK = 2;
constexpr int MM_CHUNK = 8;

template <int TB_, int TC_>
struct MmCfg {
    static constexpr int TB = TB_, TC = TC_;
    static constexpr int THREADS = TB * TC;
    static constexpr int SW = TB + 2, SH = TC + 2;
    static constexpr int PW = MM_NSLOT + 1;                 // doubles per position (+1 pad: bank spread)
    static constexpr int RP = SW * PW;
    static constexpr int NPOS = SW * SH;
    static constexpr int NHALO = 2 * (TB + TC);
    static constexpr int HR = (NHALO + THREADS - 1) / THREADS;
};
typedef MmCfg<16, 16> MmCfgDefault;

struct MmParams {
    const double *A;               // phi_old
    double *B;                     // phi_new
    const uint8_t *mask;           // band mask of this iteration, or nullptr: abs(old) < bNB
    int nx, ny, nz;
    long long sx, sxy;
    double bNB;                    // 4.1*dx (subs.f90:194)
    double dxx;                    // 1./(dx*dx) (subs.f90:384)
    double h1;
    int c_lo, c_hi, c_max;         // planes c_lo..c_hi are updated, planes 0..c_max exist (one GPU: 1, nz-1, nz; z-slab: lsf_slab.cuh)
    int ntb, ntc, ntiles, tend;
    double *partial;
    unsigned *ticket;
    const int *order;
    long long *progress;
    long long epoch;
    Ctrl *ctrl;
    // z-slab sharding (the sweep order is always ascending k: the rank below is upstream); all zero on one GPU.
    // The upstream rank streams the NEW values of its last updated plane into this rank's ghost plane of B
    // and publishes its last tile row's progress in in_progress[J]; ghost planes of A hold the OLD values
    // (bulk exchange between iterations, k_slab_exchange).
    const long long *in_progress;
    long long push_delta;          // (downstream rank's B, shifted to this rank's indexing) - B, in elements; 0: none
    long long *push_progress;
    const long long *halo_seq;
    long long halo_need[2];
};

template <class CFG>
struct MmSmem {
    double So[CFG::NPOS * CFG::PW];
    double Sn[CFG::NPOS * CFG::PW];
    double red[CFG::THREADS];
    int tile;
};

template <class CFG>
inline void mm_orient(MmParams &p, int nx, int ny, int nz, int kupd_lo = 1, int kupd_hi = -1)
{
    if (kupd_hi < 0) kupd_hi = nz - 1;
    p.nx = nx; p.ny = ny; p.nz = nz;
    p.sx = (long long)nx + 1;
    p.sxy = p.sx * ((long long)ny + 1);
    p.c_lo = kupd_lo; p.c_hi = kupd_hi; p.c_max = nz;
    p.ntb = (ny - 1 + CFG::TB - 1) / CFG::TB;
    p.ntc = (p.c_hi - p.c_lo + 1 + CFG::TC - 1) / CFG::TC;
    p.ntiles = p.ntb * p.ntc;
    p.tend = (nx - 1) - 1 + (CFG::TB - 1) + (CFG::TC - 1) + 1;
}

template <class CFG>
LSF_DEV void mm_tile(const MmParams &p, MmSmem<CFG> &sm, const int tid, const int J, const int K)
{
    typedef ExactArith X;
    constexpr int TB = CFG::TB, TC = CFG::TC, THREADS = CFG::THREADS, RP = CFG::RP, PW = CFG::PW;
    const int tb = tid % TB, tc = tid / TB;
    const int b = 1 + J * TB + tb, c = p.c_lo + K * TC + tc;
    const bool rowValid = (b <= p.ny) && (c <= p.c_max);
    const bool compValid = (b <= p.ny - 1) && (c <= p.c_hi);
    const bool pushRow = (p.push_delta != 0) && compValid && (c == p.c_hi);
    const int sig = tb + tc + 1;
    const long long rowoff = (long long)b * p.sx + (long long)c * p.sxy;
    const double *rowA = p.A + (rowValid ? rowoff : 0);
    double *rowB = p.B + (rowValid ? rowoff : 0);
    const uint8_t *rowM = p.mask ? p.mask + (rowValid ? rowoff : 0) : nullptr;
    const int pos = (tc + 1) * RP + (tb + 1) * PW;
    double *const So = sm.So + pos;
    double *const Sn = sm.Sn + pos;

    bool hvalid[CFG::HR], hlow[CFG::HR];
    int hsig[CFG::HR], hpos[CFG::HR];
    long long hoff[CFG::HR];
#pragma unroll
    for (int r = 0; r < CFG::HR; ++r) {
        const int q = tid + r * THREADS;
        hvalid[r] = false; hlow[r] = false; hsig[r] = 0; hpos[r] = 0; hoff[r] = 0;
        if (q < CFG::NHALO) {
            int htb, htc;
            if (q < 2 * TC) { htc = q % TC; if (q < TC) { htb = -1; hlow[r] = true; } else htb = TB; }
            else { const int q2 = q - 2 * TC; htb = q2 % TB; if (q2 < TB) { htc = -1; hlow[r] = true; } else htc = TC; }
            const int hb = 1 + J * TB + htb, hc = p.c_lo + K * TC + htc;
            hvalid[r] = (hb >= 0) && (hb <= p.ny) && (hc >= 0) && (hc <= p.c_max);
            hsig[r] = htb + htc + 1;
            hpos[r] = (htc + 1) * RP + (htb + 1) * PW;
            if (hvalid[r]) hoff[r] = (long long)hb * p.sx + (long long)hc * p.sxy;
        }
    }

    const long long ebase = p.epoch << 32;
    const long long *predB = (J > 0) ? p.progress + ((J - 1) + p.ntb * K) : nullptr;
    const bool predCpeer = (K == 0) && p.in_progress;
    const long long *predC = (K > 0) ? p.progress + (J + p.ntb * (K - 1)) : (predCpeer ? p.in_progress + J : nullptr);
    long long *mine = p.progress + (J + p.ntb * K);
    long long *minePeer = (K == p.ntc - 1 && p.push_progress) ? p.push_progress + J : nullptr;
    double acc = 0.;

    // Global loads are software-pipelined by one step: values requested in step t are deposited into the
    // rings in step t+1, so their latency overlaps a whole step instead of stalling the step that issued
    // them (the cell update itself is only a handful of instructions).
    //   old values : hyperplane t+LOOK+1 requested at step t, deposited at t+1, first read at t+LOOK
    //   new values of -b/-c halo rows : hyperplane t+1 requested at step t, deposited at t+1, read at t+2;
    //                the predecessor tile computed that cell at its step t+1+TB -> wait bound below
    bool q_look = false, q_do[CFG::HR], q_dn[CFG::HR];
    double q_la = 0., q_vo[CFG::HR], q_vn[CFG::HR];
#pragma unroll
    for (int r = 0; r < CFG::HR; ++r) { q_do[r] = q_dn[r] = false; q_vo[r] = q_vn[r] = 0.; }

    for (int t = -MM_LOOK - 1; t <= p.tend; ++t) {
        if (t >= 0 && (t % MM_CHUNK) == 0) {
            const long long need_b = ebase + M_BIAS + (t + MM_CHUNK + TB);
            const long long need_c = ebase + M_BIAS + (t + MM_CHUNK + TC);
            if (tid == 0 && predB) wait_ge<false>(predB, need_b, p.ctrl);
            if (tid == 32 % THREADS && predC) {
                if (predCpeer) wait_ge<true>(predC, need_c, p.ctrl); else wait_ge<false>(predC, need_c, p.ctrl);
            }
            p_sync();
        }
        const int a = 1 + t - sig;
        // ---- request the values that step t+1 will deposit ---------------------------------------
        const int al = a + MM_LOOK + 1;
        const bool ldLook = rowValid && (al >= 0) && (al <= p.nx);
        double la = 0.;
        if (ldLook) la = p_ldcg(rowA + al);
        const bool active = compValid && (a >= 1) && (a <= p.nx - 1);
        bool band = false;
        bool hdo[CFG::HR], hdn[CFG::HR];
        double hvo[CFG::HR], hvn[CFG::HR];
#pragma unroll
        for (int r = 0; r < CFG::HR; ++r) {
            hdo[r] = hdn[r] = false; hvo[r] = hvn[r] = 0.;
            if (hvalid[r]) {
                const int ao = 1 + (t + MM_LOOK + 1) - hsig[r];               // old value, hyperplane t+LOOK+1
                if (ao >= 0 && ao <= p.nx) { hvo[r] = p_ldcg(p.A + hoff[r] + ao); hdo[r] = true; }
                if (hlow[r]) {                                                // new value, hyperplane t+1
                    const int an = 1 + (t + 1) - hsig[r];
                    if (t + 1 >= 0 && an >= 0 && an <= p.nx) { hvn[r] = p_ldcg(p.B + hoff[r] + an); hdn[r] = true; }
                }
            }
        }
        // ---- cell update ------------------------------------------------------------------------
        // (cells that exist but are never updated -- a = 0 or nx -- still feed their live value, which is
        // their old one, to the new-value ring: cell a = 1 reads it as its updated -1 neighbour)
        const bool exists = rowValid && (a >= 0) && (a <= p.nx);
        double pn = 0.;
        if (exists) {
            const int s0 = t & (MM_NSLOT - 1), sm1 = (t - 1) & (MM_NSLOT - 1), sp1 = (t + 1) & (MM_NSLOT - 1);
            const double pc = So[s0];
            pn = pc;
            if (active) {
                if (rowM) band = rowM[a] != 0;
                else band = fabs(pc) < p.bNB;
            }
            if (band) {
                const double oxm = So[sm1], oxp = So[sp1];
                const double oym = So[sm1 - PW], oyp = So[sp1 + PW];
                const double ozm = So[sm1 - RP], ozp = So[sp1 + RP];
                const double m2 = X::mul(-2., pc);
                const double xx = X::mul(X::add(X::add(m2, oxp), oxm), p.dxx);        // subs.f90:387
                const double yy = X::mul(X::add(X::add(m2, oyp), oym), p.dxx);        // :388
                const double zz = X::mul(X::add(X::add(m2, ozp), ozm), p.dxx);        // :389
                const double curv = X::add(X::add(xx, yy), zz);                       // :461
                const double nxm = Sn[sm1], nym = Sn[sm1 - PW], nzm = Sn[sm1 - RP];   // live -1 neighbours
                double pAve = X::add(pc, nxm);                                         // :473, left to right
                pAve = X::add(pAve, oxp);
                pAve = X::add(pAve, oyp);
                pAve = X::add(pAve, nym);
                pAve = X::add(pAve, ozp);
                pAve = X::add(pAve, nzm);
                pAve = X::div(pAve, 7.);                                               // :474
                const double F = (pAve < 0.) ? fmin_f(curv, 0.0) : fmax_f(curv, 0.0);  // :477-481
                pn = X::add(pc, X::mul(p.h1, F));                                      // set3d.f90:426
                const double df = X::sub(pn, pc);
                acc = X::add(acc, X::mul(df, df));
            }
            if (active) {
                p_stcg(rowB + a, pn);
                if (pushRow) p_st_peer(rowB + a + p.push_delta, pn);
            }
        }
        // ---- deposits: own new value, and the values requested one step ago -------------------------
        if (exists) Sn[t & (MM_NSLOT - 1)] = pn;
        if (q_look) So[(t + MM_LOOK) & (MM_NSLOT - 1)] = q_la;
#pragma unroll
        for (int r = 0; r < CFG::HR; ++r) {
            if (q_do[r]) sm.So[hpos[r] + ((t + MM_LOOK) & (MM_NSLOT - 1))] = q_vo[r];
            if (q_dn[r]) sm.Sn[hpos[r] + (t & (MM_NSLOT - 1))] = q_vn[r];
        }
        q_look = ldLook; q_la = la;
#pragma unroll
        for (int r = 0; r < CFG::HR; ++r) { q_do[r] = hdo[r]; q_vo[r] = hvo[r]; q_dn[r] = hdn[r]; q_vn[r] = hvn[r]; }
        const bool pub = (t >= 0) && ((t % MM_CHUNK) == MM_CHUNK - 1);
        p_sync();
        if (pub && tid == 0) {
            p_fence(); p_st_release(mine, ebase + M_BIAS + t);
            if (minePeer) { p_fence_sys(); p_st_release_sys(minePeer, ebase + M_BIAS + t); }
        }
    }
    sm.red[tid] = acc;
    p_sync();
    if (tid == 0) {
        p_fence(); p_st_release(mine, ebase + M_BIAS + M_FIN);
        if (minePeer) { p_fence_sys(); p_st_release_sys(minePeer, ebase + M_BIAS + M_FIN); }
    }
    for (int wdt = THREADS / 2; wdt > 0; wdt >>= 1) {
        if (tid < wdt) sm.red[tid] = X::add(sm.red[tid], sm.red[tid + wdt]);
        p_sync();
    }
    if (tid == 0) p.partial[J + p.ntb * K] = sm.red[0];
    p_sync();
}

template <class CFG>
LSF_DEV void mm_cta(const MmParams &p, MmSmem<CFG> &sm, const int tid)
{
    if (p.ctrl->done) return;
    if (p.halo_seq) {          // z-slab: both neighbours must have refreshed this rank's ghost planes
        if (tid == 0) {
            if (p.halo_need[0]) wait_ge<true>(p.halo_seq + 0, p.halo_need[0], p.ctrl);
            if (p.halo_need[1]) wait_ge<true>(p.halo_seq + 1, p.halo_need[1], p.ctrl);
        }
        p_sync();
    }
    for (;;) {
        if (tid == 0) sm.tile = (int)p_ticket(p.ticket);
        p_sync();
        const int tk = sm.tile;
        p_sync();
        if (tk >= p.ntiles) break;
        const int jk = p.order[tk];
        mm_tile<CFG>(p, sm, tid, jk & 0xffff, jk >> 16);
    }
}

}  // namespace lsf
