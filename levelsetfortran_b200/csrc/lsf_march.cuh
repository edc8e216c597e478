// lsf_march.cuh -- the production schedule of the in-place Gauss-Seidel WENO5 sweep
// (reference loop nest subs.f90:742-852): skewed x-marching column tiles.
//
// Idea.  Any topological order of the dependence DAG  c - e_d -> c  (d = x,y,z in the sweep's
// own direction) reproduces the reference's lexicographic in-place sweep exactly (SURVEY.md 3.2;
// proven bitwise by tests).  In sweep-oriented indices (a,b,c), all ascending:
//   * the (b,c) cross-section is cut into TBxTC column tiles, one CTA per tile at a time;
//   * thread (tb,tc) owns the whole x-row (b0+tb, c0+tc) and marches along a (the contiguous
//     axis), skewed by one cell per unit of tb+tc: at step t it updates a = 1 + t - (tb+tc+3).
//     All TB*TC threads are busy every step (except the ~TB+TC ramp steps at both ends);
//   * the row's own +-3 window lives in registers (3 new values behind, 3 old ahead);
//   * y/z neighbours are exchanged through a ring of NSLOT hyperplane slots in shared memory:
//     slot(h) holds, for every row of the tile and its 3-wide halo, the cell whose step index
//     is h -- the OLD value until its owner reaches step h, the NEW one afterwards.  A thread
//     at step t therefore reads slots t-3..t-1 (new) and t+1..t+3 (old): exactly what the
//     reference's in-place loop sees.  One __syncthreads per step;
//   * halo rows on the -b/-c side are the neighbouring tiles' NEW values, read from global
//     memory (L2) behind a per-tile progress flag: the same cell is LAG=TB steps later in the
//     neighbour's frame, so a tile simply runs >= LAG+CHUNK steps behind its two predecessors.
//     +b/+c halo rows are OLD values, read 4 steps ahead; the successor tile cannot have
//     overwritten them because it runs behind this tile by the same rule;
//   * tiles are handed out through an atomic ticket in an order that is topological for
//     (J-1,K) -> (J,K) <- (J,K-1), so a waiting CTA's predecessors are always running: no
//     deadlock, no grid-wide barrier, one launch per sweep.
//
// The file is written against a tiny set of primitives (sync, cache-global load/store, fence,
// acquire/release flag access) so that tests/emu can compile the very same code for the CPU,
// with one OS thread per CUDA thread, and check it bitwise against the oracle without a GPU.
#pragma once
#include "lsf_common.cuh"

#if defined(LSF_EMU)
#include "../../tests/emu/emu_prims.h"
#else
#define LSF_DEV __device__ __forceinline__
namespace lsf {
LSF_DEV void p_sync() { __syncthreads(); }
LSF_DEV double p_ldcg(const double *p) { return __ldcg(p); }
LSF_DEV void p_stcg(double *p, double v) { __stcg(p, v); }
LSF_DEV void p_fence() { __threadfence(); }
LSF_DEV unsigned p_ticket(unsigned *ctr) { return atomicAdd(ctr, 1u); }
LSF_DEV long long p_ld_acquire(const long long *p)
{
    long long v;
    asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
LSF_DEV void p_st_release(long long *p, long long v)
{
    asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
LSF_DEV void p_sleep() { __nanosleep(40); }
}  // namespace lsf
#endif

namespace lsf {

constexpr int M_TB = 16, M_TC = 16;          // tile cross-section (oriented b x c)
constexpr int M_THREADS = M_TB * M_TC;
constexpr int M_H = 3;                       // stencil half-width
constexpr int M_SW = M_TB + 2 * M_H;         // 22
constexpr int M_SH = M_TC + 2 * M_H;         // 22
constexpr int M_NSLOT = 8;
constexpr int M_CHUNK = 8;                   // publish / wait granularity in steps
constexpr int M_LOOK = 4;                    // old values are deposited this many steps ahead
constexpr int M_NHALO = 2 * M_H * (M_TB + M_TC);   // 192 halo rows
constexpr long long M_FIN = 1LL << 30;       // "tile finished" progress value
constexpr long long M_BIAS = 1000;

struct MarchParams {
    double *phi;
    const double *phiS;
    int nx, ny, nz;
    long long sa, sb, sc, off0;    // signed strides / origin of the sweep-oriented frame
    int fa, fb, fc;                // axis flipped?
    int lo_a, hi_a, lo_b, hi_b, lo_c, hi_c;   // high-order window (subs.f90:506) in oriented indices
    int ntb, ntc, ntiles;
    int tend;                      // last step index
    CellConst cc;
    double *partial;               // per tile: sum over its cells of (new-old)^2
    unsigned *ticket;
    const int *order;              // ticket -> J | (K << 16)
    long long *progress;           // per tile (J + ntb*K)
    long long epoch;               // progress values are epoch*2^32 + step + BIAS
    const Ctrl *ctrl;
};

struct MarchSmem {
    double S[M_NSLOT][M_SH][M_SW];
    double red[M_THREADS];
    int tile;
};

// Host helper shared with the emulator: fill the orientation-dependent fields.
inline void march_orient(MarchParams &p, int nx, int ny, int nz, long long sx, long long sxy, int raster)
{
    int d[3];
    raster_dirs(raster, d);
    p.nx = nx; p.ny = ny; p.nz = nz;
    p.fa = d[0] < 0; p.fb = d[1] < 0; p.fc = d[2] < 0;
    p.sa = p.fa ? -1 : 1;
    p.sb = p.fb ? -sx : sx;
    p.sc = p.fc ? -sxy : sxy;
    p.off0 = (p.fa ? nx : 0) + (p.fb ? ny : 0) * sx + (p.fc ? nz : 0) * sxy;
    // physical window i in [4, n-5]; flipped index a = n - i -> [5, n-4]
    p.lo_a = p.fa ? 5 : 4; p.hi_a = p.fa ? nx - 4 : nx - 5;
    p.lo_b = p.fb ? 5 : 4; p.hi_b = p.fb ? ny - 4 : ny - 5;
    p.lo_c = p.fc ? 5 : 4; p.hi_c = p.fc ? nz - 4 : nz - 5;
    p.ntb = (ny - 1 + M_TB - 1) / M_TB;
    p.ntc = (nz - 1 + M_TC - 1) / M_TC;
    p.ntiles = p.ntb * p.ntc;
    p.tend = (nx - 1) - 1 + (M_TB - 1) + (M_TC - 1) + M_H;
}

// ticket order: anti-diagonals of the tile grid (J+K ascending): topological, and all tiles of
// a diagonal are mutually independent, so the resident CTAs are never blocked for long.
inline void march_fill_order(int ntb, int ntc, int *order)
{
    int n = 0;
    for (int s = 0; s <= ntb + ntc - 2; ++s)
        for (int K = 0; K < ntc; ++K) {
            const int J = s - K;
            if (J < 0 || J >= ntb) continue;
            order[n++] = J | (K << 16);
        }
}

template <class AR>
LSF_DEV void march_tile(const MarchParams &p, MarchSmem &sm, const int tid, const int J, const int K)
{
    const int tb = tid % M_TB, tc = tid / M_TB;
    const int b = 1 + J * M_TB + tb, c = 1 + K * M_TC + tc;
    const bool rowValid = (b <= p.ny) && (c <= p.nz);
    const bool compValid = (b <= p.ny - 1) && (c <= p.nz - 1);
    const bool hiBC = (b >= p.lo_b) && (b <= p.hi_b) && (c >= p.lo_c) && (c <= p.hi_c);
    const int sig = tb + tc + M_H;
    const long long rowoff = p.off0 + (long long)b * p.sb + (long long)c * p.sc;
    double *rowp = p.phi + (rowValid ? rowoff : 0);
    const double *rowS = p.phiS + (rowValid ? rowoff : 0);

    // halo duty: thread q < 192 feeds one halo row
    bool hvalid = false, hlow = false;
    int htb = 0, htc = 0, hsig = 0;
    const double *hrow = p.phi;
    if (tid < M_NHALO) {
        const int side = tid / (M_H * M_TB), r = tid % (M_H * M_TB);
        const int m = r / M_TB + 1, idx = r % M_TB;
        if (side == 0) { htb = -m; htc = idx; hlow = true; }
        else if (side == 1) { htb = M_TB - 1 + m; htc = idx; }
        else if (side == 2) { htb = idx; htc = -m; hlow = true; }
        else { htb = idx; htc = M_TC - 1 + m; }
        const int hb = 1 + J * M_TB + htb, hc = 1 + K * M_TC + htc;
        hvalid = (hb >= 0) && (hb <= p.ny) && (hc >= 0) && (hc <= p.nz);
        hsig = htb + htc + M_H;
        if (hvalid) hrow = p.phi + p.off0 + (long long)hb * p.sb + (long long)hc * p.sc;
    }

    const long long ebase = p.epoch << 32;
    const long long *predB = (J > 0) ? p.progress + ((J - 1) + p.ntb * K) : nullptr;
    const long long *predC = (K > 0) ? p.progress + (J + p.ntb * (K - 1)) : nullptr;
    long long *mine = p.progress + (J + p.ntb * K);

    double w0 = 0., w1 = 0., w2 = 0., w3 = 0., w4 = 0., w5 = 0., w6 = 0.;
    double acc = 0.;

    for (int t = -M_LOOK; t <= p.tend; ++t) {
        // ---- wait for the two predecessor tiles at chunk starts ---------------------------
        if (t >= 0 && (t % M_CHUNK) == 0) {
            const long long need_b = ebase + M_BIAS + (t + M_CHUNK - 1 + M_TB);
            const long long need_c = ebase + M_BIAS + (t + M_CHUNK - 1 + M_TC);
            if (tid == 0 && predB) { while (p_ld_acquire(predB) < need_b) p_sleep(); }
            if (tid == 32 && predC) { while (p_ld_acquire(predC) < need_c) p_sleep(); }
            p_sync();
        }
        const int a = 1 + t - sig;
        // ---- (1) issue the global loads of this step --------------------------------------
        const int a4 = a + M_LOOK;
        const bool ldLook = rowValid && (a4 >= 0) && (a4 <= p.nx);
        double la = 0.;
        if (ldLook) la = p_ldcg(rowp + a4 * p.sa);
        const bool active = compValid && (a >= 1) && (a <= p.nx - 1);
        double ps = 0.;
        if (active) ps = p_ldcg(rowS + a * p.sa);
        bool hdep = false;
        double hv = 0.;
        int hh = 0;
        if (hvalid) {
            hh = hlow ? t : t + M_LOOK;
            const int ah = 1 + hh - hsig;
            if (hh >= 0 && ah >= 0 && ah <= p.nx) { hv = p_ldcg(hrow + ah * p.sa); hdep = true; }
        }
        // ---- (2) cell update ----------------------------------------------------------------
        double pn = w3;
        if (active) {
            double vy[7], vz[7], vx[7];
            const int sy = tc + M_H, sx = tb + M_H;
#pragma unroll
            for (int m = -3; m <= 3; ++m) {
                if (m == 0) continue;
                const int sl = (t + m) & (M_NSLOT - 1);
                const double yv = sm.S[sl][sy][sx + m];
                const double zv = sm.S[sl][sy + m][sx];
                vy[p.fb ? 3 - m : 3 + m] = yv;
                vz[p.fc ? 3 - m : 3 + m] = zv;
            }
            vy[3] = w3; vz[3] = w3;
            if (p.fa) { vx[0] = w6; vx[1] = w5; vx[2] = w4; vx[3] = w3; vx[4] = w2; vx[5] = w1; vx[6] = w0; }
            else      { vx[0] = w0; vx[1] = w1; vx[2] = w2; vx[3] = w3; vx[4] = w4; vx[5] = w5; vx[6] = w6; }
            const bool hi = hiBC && (a >= p.lo_a) && (a <= p.hi_a);
            double g[3], gM;
            pn = reinit_cell<AR>(vx, vy, vz, ps, hi, p.cc, g, gM);
            const double df = pn - w3;
            acc += df * df;
            p_stcg(rowp + a * p.sa, pn);
        }
        // ---- (3) deposits into the slot ring ------------------------------------------------
        if (active) sm.S[t & (M_NSLOT - 1)][tc + M_H][tb + M_H] = pn;
        if (ldLook) sm.S[(t + M_LOOK) & (M_NSLOT - 1)][tc + M_H][tb + M_H] = la;
        if (hdep) sm.S[hh & (M_NSLOT - 1)][htc + M_H][htb + M_H] = hv;
        // ---- (4) slide the register window --------------------------------------------------
        w0 = w1; w1 = w2; w2 = pn; w3 = w4; w4 = w5; w5 = w6; w6 = la;
        // ---- (5) publish progress every CHUNK steps -----------------------------------------
        const bool pub = (t >= 0) && ((t % M_CHUNK) == M_CHUNK - 1);
        if (pub) p_fence();
        p_sync();
        if (pub && tid == 0) p_st_release(mine, ebase + M_BIAS + t);
    }
    // ---- tile done: final publish + deterministic block reduction of the RMS partial --------
    p_fence();
    sm.red[tid] = acc;
    p_sync();
    if (tid == 0) p_st_release(mine, ebase + M_BIAS + M_FIN);
    for (int wdt = M_THREADS / 2; wdt > 0; wdt >>= 1) {
        if (tid < wdt) sm.red[tid] = sm.red[tid] + sm.red[tid + wdt];
        p_sync();
    }
    if (tid == 0) p.partial[J + p.ntb * K] = sm.red[0];
    p_sync();
}

// Persistent CTA: take tickets until the tile list is exhausted.
template <class AR>
LSF_DEV void march_cta(const MarchParams &p, MarchSmem &sm, const int tid)
{
    if (p.ctrl->done) return;
    for (;;) {
        if (tid == 0) sm.tile = (int)p_ticket(p.ticket);
        p_sync();
        const int tk = sm.tile;
        p_sync();
        if (tk >= p.ntiles) break;
        const int jk = p.order[tk];
        march_tile<AR>(p, sm, tid, jk & 0xffff, jk >> 16);
    }
}

}  // namespace lsf
