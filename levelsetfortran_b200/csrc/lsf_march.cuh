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
//   * ALL 19 stencil values come from a ring of hyperplane slots in shared memory:
//     slot(h) holds, for every row of the tile and its 3-wide halo, the cell whose step index
//     is h -- the OLD value until its owner reaches step h, the NEW one afterwards.  A thread
//     at step t therefore reads slots t-3..t-1 (new) and t+1..t+3 (old) of its x/y/z neighbour
//     rows: exactly what the reference's in-place loop sees.  One __syncthreads per step.
//     The ring is stored position-major with the 8 slots written twice (S[pos][h&7] and
//     S[pos][(h&7)+8]), so the 7-slot window t-3..t+3 is contiguous and every one of the 19 loads
//     is  base(thread) + ((t-3)&7) + compile-time constant: no per-load address arithmetic;
//   * the sweep orientation (which axes run backwards) is a template parameter, so the mapping of
//     oriented stencil offsets to the physical -3..+3 stencil is resolved at compile time;
//   * halo rows on the -b/-c side are the neighbouring tiles' NEW values, read from global
//     memory (L2) behind a per-tile progress flag: the same cell is LAG=TB (TC) steps later in the
//     neighbour's frame, so a tile simply runs >= LAG+CHUNK steps behind its two predecessors.
//     +b/+c halo rows are OLD values, read 4 steps ahead; the successor tile cannot have
//     overwritten them because it runs behind this tile by the same rule;
//   * tiles are handed out through an atomic ticket in an order that is topological for
//     (J-1,K) -> (J,K) <- (J,K-1), so a waiting CTA's predecessors are always running: no
//     deadlock, no grid-wide barrier, one launch per sweep;
//   * the sum over the tile of (new-old)^2 -- the interior part of the RMS exit test,
//     subs.f90:902-914 -- is accumulated on the fly (one partial per tile, fixed order).
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
// The three global loads of a step and the twelve ring gathers are ORDERED accesses (relaxed.gpu loads / volatile shared loads in
// volatile asm), which the compiler keeps in program order: the global loads issue before the gathers, at the top of the step.
// With plain loads ptxas sinks each of them to 25-105 instructions before its first use (tools/sass_ldg_distance.py) -- and in one
// build (z-slab kernels without the library sqrt's CALL) to 7-29, which made a sweep 7x slower.  LSF_ORDERED_LD: 2 = relaxed.gpu
// (default; session 21: 34.6 fp64 / 49.5 fp32 Gcell/s at 1024^3), 1 = ld.volatile (33.7 / 45.2), 0 = plain ld.cg (33.9 / 49.1)
#ifndef LSF_ORDERED_LD
#define LSF_ORDERED_LD 2
#endif
#if LSF_ORDERED_LD == 2
LSF_DEV double p_ldcg(const double *p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
LSF_DEV float p_ldcg(const float *p) { float v; asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v; }
template <class T> LSF_DEV T p_lds(const T *p) { return *(const volatile T *)p; }
#elif LSF_ORDERED_LD == 1
LSF_DEV double p_ldcg(const double *p) { double v; asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
LSF_DEV float p_ldcg(const float *p) { float v; asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v; }
template <class T> LSF_DEV T p_lds(const T *p) { return *(const volatile T *)p; }
#else
template <class T> LSF_DEV T p_lds(const T *p) { return *p; }
LSF_DEV double p_ldcg(const double *p) { return __ldcg(p); }
LSF_DEV float p_ldcg(const float *p) { return __ldcg(p); }
#endif
LSF_DEV void p_stcg(double *p, double v) { __stcg(p, v); }
LSF_DEV void p_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L1-allocating load: only for data no other CTA writes while this tile may still hold the line (phiS; OLD values)
LSF_DEV double p_ldca(const double *p) { return __ldca(p); }
LSF_DEV float p_ldca(const float *p) { return __ldca(p); }
LSF_DEV void p_stcg(float *p, float v) { __stcg(p, v); }
// aligned VEC-element chunk, L1-bypassing (one LDG.64 / LDG.128)
template <int VEC> LSF_DEV void p_ldcg_vec(const float *p, float *out)
{
    if (VEC == 4) { const float4 v = __ldcg(reinterpret_cast<const float4 *>(p)); out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
    else if (VEC == 2) { const float2 v = __ldcg(reinterpret_cast<const float2 *>(p)); out[0] = v.x; out[1] = v.y; }
    else out[0] = __ldcg(p);
}
template <int VEC> LSF_DEV void p_ldcg_vec(const double *p, double *out)
{
    if (VEC == 2) { const double2 v = __ldcg(reinterpret_cast<const double2 *>(p)); out[0] = v.x; out[1] = v.y; }
    else out[0] = __ldcg(p);
}
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
// polling load: relaxed (LDG.STRONG.GPU only).  ld.acquire would add an L1 invalidate (CCTL.IVALL) to
// EVERY poll, which floods the LSU/MIO pipe the shared-memory loads go through; instead one acquire
// fence is executed after the poll loop has seen the value it waits for.
LSF_DEV long long p_ld_relaxed(const long long *p)
{
    long long v;
    asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
LSF_DEV void p_fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
LSF_DEV void p_sleep() { __nanosleep(100); }
// system-scope flavours for flags / data that cross NVLink (z-slab sharding, lsf_slab.cuh): a flag that a
// PEER GPU writes into this GPU's memory is polled relaxed.sys and followed by one acq_rel.sys fence; data
// pushed into a peer's ghost planes is an L1-bypassing store made visible by fence.sys + st.release.sys.
LSF_DEV long long p_ld_relaxed_sys(const long long *p)
{
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
LSF_DEV void p_fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
LSF_DEV void p_st_release_sys(long long *p, long long v)
{
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
LSF_DEV void p_st_peer(double *p, double v) { __stcg(p, v); }
LSF_DEV void p_st_peer(float *p, float v) { __stcg(p, v); }
LSF_DEV int p_ld_relaxed_i32(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
LSF_DEV bool p_cas_i32(int *p, int expect, int desired) { return atomicCAS(p, expect, desired) == expect; }
LSF_DEV void p_emu_hook(bool) {}   // CPU emulation only (tests/emu/emu_prims.h)
// split-phase CTA barrier (mbarrier in shared memory): every warp arrives once per phase (one elected lane after a
// __syncwarp, release), a waiter spins on the phase parity (acquire).  Between arrive and wait a thread may do anything
// that touches neither the slot ring nor another thread's data.
// experiments (session 16): LSF_BAR_ALLARRIVE -- every thread arrives (no __syncwarp + elected lane); LSF_BAR_TESTWAIT -- the
// waiter polls the non-blocking mbarrier.test_wait instead of the potentially suspending try_wait
LSF_DEV void p_bar_init(unsigned long long *bar, int nthreads)
{
#if defined(LSF_BAR_ALLARRIVE)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(nthreads) : "memory");
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(nthreads / 32) : "memory");
#endif
}
LSF_DEV void p_bar_arrive(unsigned long long *bar)
{
#if defined(LSF_BAR_ALLARRIVE)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
#else
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
#endif
}
LSF_DEV void p_bar_wait(unsigned long long *bar, unsigned phase)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
#if defined(LSF_BAR_TESTWAIT)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(phase & 1u) : "memory");
#else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(phase & 1u) : "memory");
#endif
    } while (!ok);
}
}  // namespace lsf
#endif

// ---- build-time switches (defaults = the production kernel; the others are the measured alternatives and the
// ---- switch-off experiments of DESIGN.md section 9, built with tools/build_variant.sh) ---------------------------
#ifndef LSF_W32
#define LSF_W32 17              // fp32 ring: floats per position (16 slots + pad); 18 makes the column-halo deposits conflict-free (-0.7 %)
#endif
#ifndef LSF_VEC32
#define LSF_VEC32 1             // cells per global vector load of a row walk in fp32; 4 (LDG.128) measured 5-18 % SLOWER
#endif
#ifndef LSF_VEC64
#define LSF_VEC64 1             // same in fp64; 2 (LDG.64) measured 40 % slower
#endif
#ifndef LSF_LD_CACHED
#define LSF_LD_CACHED 0         // bit 0: look-ahead loads through L1, bit 1: phiS loads, bit 2: +b/+c halo rows (OLD values); measured 1-6 % slower
#endif
#if defined(LSF_EXP_NOLDG) && !defined(LSF_EXP_NOLDG_MASK)
#define LSF_EXP_NOLDG_MASK 7
#endif
#ifndef LSF_EXP_NOLDG_MASK
#define LSF_EXP_NOLDG_MASK 0    // timing experiments only (results wrong): 1 = no look-ahead load, 2 = no phiS load, 4 = no halo loads
#endif
#ifndef LSF_CHUNK
#define LSF_CHUNK 8             // steps between two progress publications / predecessor waits of a tile
#endif
#ifndef LSF_ASYNC_POLL
#define LSF_ASYNC_POLL 0        // 1: the predecessor flags are read one step before the chunk start that tests them
#endif
#ifndef LSF_RING_DUP
#define LSF_RING_DUP 0          // 0 (default): every ring slot stored once, 7 wrapped slot offsets computed per step (35 KB per fp64 CTA:
#endif                          // 3 CTAs/SM leave the L1 its size); 1: stored twice, window loads need no wrap arithmetic (66 KB per CTA; round 1)
#ifndef LSF_SPREAD_DUTIES
#define LSF_SPREAD_DUTIES 1     // flag polling / progress publishing by lanes of the warps that feed no halo rows
#endif
#ifndef LSF_FOLD_FLAGS
#define LSF_FOLD_FLAGS 1        // (with LSF_SPLIT_BAR) the predecessor-flag check of a chunk is folded into the step barrier before it
#endif
#ifndef LSF_NO_ACQ_FENCE
#define LSF_NO_ACQ_FENCE 0      // experiment: no acquire fence after a flag that was already satisfied when pre-read
#endif
#ifndef LSF_SPLIT_BAR
#define LSF_SPLIT_BAR 1         // 1 (default): split step barrier -- a thread ARRIVES after its deposits, computes the x direction of its NEXT
#endif                          //    cell from a register window of its own row, and only then WAITS for the other threads' deposits
#ifndef LSF_PREFETCH
#define LSF_PREFETCH 1          // global loads issued one step ahead and carried in registers (ncu r2b: 12 % / 17 % of fp64 / fp32 warp time waited
#endif                          //    for phiS at its first use).  Bit 0: phiS, bit 1: look-ahead cell, bit 2: +b/+c halo rows, bit 3: -b/-c halo rows.
                                //    Session 14, fp64 / fp32 Gcell/s at 1024^3: 0 -> 33.5 / 45.8, 1 -> 33.9 / 49.0, 3 -> 33.6 / 8.5 (!), 7 -> 30.9 / 7.1,
                                //    15 -> 33.6 / 13.0: a look-ahead or halo load left in flight across the step's stores to the same rows is
                                //    expensive (fp32 most: 8 cells per sector), so only phiS -- read-only -- is fetched ahead
#ifndef LSF_PUB_FENCE
#define LSF_PUB_FENCE 1         // 1: a __threadfence (fence.sc.gpu) before the release store that publishes a tile's progress.  The release store
#endif                          //    alone is sufficient (all threads' stores -> step barrier -> one thread's gpu-scope release: cumulative); 0 drops the
                                //    sequentially-consistent fence from the critical path of every chunk
#ifndef LSF_PIN_LOADS
#define LSF_PIN_LOADS 2         // see the step body
#endif
#ifndef LSF_L2_AHEAD
#define LSF_L2_AHEAD 0          // > 0: every 16 steps a thread asks for the 128-byte lines of its own row (phi and phiS) that its look-ahead
#endif                          //      will reach this many cells later (prefetch.global.L2: no register, no scoreboard).  ncu r2c: 12 % of the
                                //      warp time waits for the look-ahead load at its deposit -- a first touch that comes from HBM -- yet the
                                //      prefetch is a loss (session 27, 48 / 96 cells ahead: 33.4 instead of 34.2 Gcell/s fp64, 46.9 vs 49.4 fp32)
#ifndef LSF_PREFETCH_MG
#define LSF_PREFETCH_MG 0       // the same on z-slab kernels (see march_tile)
#endif
#ifndef LSF_STEADY
#define LSF_STEADY 1            // 1 (default): second copy of the step body for the steps of interior tiles in which every range test holds
#endif
// LSF_EXP_NOSTG / LSF_EXP_NOSYNC: no global stores / no CTA barrier per step (timing experiments, results wrong)

namespace lsf {

// Per-thread reader of one grid row that is walked one cell per step: fetches the aligned VEC-element chunk when the
// walk enters it (one vector load per VEC cells instead of one scalar load per cell).  Measured slower than scalar
// loads (the skew puts the lanes of a warp in different alignment phases, so the load is still issued every step, for
// a fraction of the lanes, under a divergent branch): kept behind LSF_VEC32 / LSF_VEC64, off by default.  FA: the row
// is walked towards lower addresses.  Elements of the chunk outside the row are fetched but never used (arrays padded).
template <class real, int VEC, bool FA>
struct RowReader {
    real buf[VEC];
    bool have;
    LSF_DEV void reset() { have = false; }
    LSF_DEV real get(const real *p)
    {
        const int idx = (int)(((unsigned long long)p / sizeof(real)) & (unsigned long long)(VEC - 1));
        if (!have || idx == (FA ? VEC - 1 : 0)) { p_ldcg_vec<VEC>(p - idx, buf); have = true; }
        real v = buf[0];
#pragma unroll
        for (int q = 1; q < VEC; ++q) v = (idx == q) ? buf[q] : v;
        return v;
    }
};

// one application of "+ dx" of the boundary block, rounded once (a plain add: nothing to contract)
LSF_DEV double bc_add(double v, double dx) { return ExactArith::add(v, dx); }
LSF_DEV float bc_add(float v, float dx) { return v + dx; }
LSF_DEV int m_imax(int a, int b) { return a > b ? a : b; }
LSF_DEV int m_imin(int a, int b) { return a < b ? a : b; }

struct StepGeneric { static constexpr bool value = false; };
struct StepSteady { static constexpr bool value = true; };

constexpr int M_PFM = (LSF_PREFETCH & 8) ? 1 : 0;  // extra steps a predecessor must be ahead: -b/-c halo values are fetched one step before their use
                                              // (progress is published in the same phase, at steps = CHUNK-1+M_PFM mod CHUNK, so the lag stays minimal);
constexpr int M_H = 3;                       // stencil half-width
constexpr int M_NSLOT = 8;                   // hyperplane slots in the ring (each stored twice)
constexpr bool M_DUP = (LSF_RING_DUP != 0);
constexpr int M_SLOTW = (M_DUP ? 2 : 1) * M_NSLOT + 1;   // doubles per position: 16 (8 slots stored twice) or 8, + 1 pad (bank spread)
constexpr int M_CHUNK = LSF_CHUNK;           // publish / wait granularity in steps
constexpr int M_LOOK = 4;                    // old values are deposited this many steps ahead
constexpr long long M_FIN = 1LL << 30;       // "tile finished" progress value
constexpr long long M_BIAS = 1000;
constexpr long long M_SPIN_LIMIT = 40000000; // polls before a flag wait gives up (tens of seconds): a lost peer must not hang the GPU
constexpr int M_ERR_TIMEOUT = -4;            // = LSF_ERR_TIMEOUT (include/lsf_b200.h)

// Tile geometry.  TB x TC rows per CTA (oriented b x c), R rows per thread.
#ifndef LSF_ROWS
#define LSF_ROWS 1
#endif
template <int TB_, int TC_, int R_ = LSF_ROWS, class T_ = double>
struct MarchCfg {
    typedef T_ real;                                          // element type of phi and of the slot ring
    static constexpr int TB = TB_, TC = TC_, R = R_;          // R rows (cells per step) per thread
    static constexpr int THREADS = TB * TC / R;
    static constexpr int SW = TB + 2 * M_H, SH = TC + 2 * M_H;
    // elements per position: 16 slots + pad.  fp32 may use a different pad (LSF_W32) so that the column-halo
    // deposits (stride = row pitch) do not all fall on two banks
    static constexpr int W = sizeof(T_) == 4 ? (M_DUP ? LSF_W32 : M_SLOTW) : M_SLOTW;
    // cells per global vector load of a row walk (RowReader); 1 = scalar loads
    static constexpr int VEC = (R_ > 1) ? 1 : (sizeof(T_) == 4 ? LSF_VEC32 : LSF_VEC64);
    // pitch of one c-row of positions, in doubles; for TB = 8 a warp spans 4 c-rows and the pitch
    // is padded to 8 (mod 16) doubles so that the two rows of a half-warp hit disjoint banks
    // fp32 ring: a warp spans two c-rows of 16 positions; the 17-float position stride puts one row on 16
    // distinct banks, a pitch of 16 (mod 32) floats puts the other row on the complementary 16
    static constexpr int RP_F64 = (TB >= 16) ? SW * M_SLOTW : ((SW * M_SLOTW + 15) / 16) * 16 + 8;
    // W = 17: one row of 16 positions covers 16 banks, pitch = 16 (mod 32) puts the next row on the other 16;
    // W = 18: a row covers the 16 even banks, pitch = 1 (mod 32) puts the next row on the odd ones
    static constexpr int RP_F32 = (W % 2) ? ((SW * W + 15) / 32) * 32 + 16 : ((SW * W + 30) / 32) * 32 + 1;
    static constexpr int RP = sizeof(T_) == 4 ? RP_F32 : RP_F64;
    static constexpr int NHALO = 2 * M_H * (TB + TC);               // halo rows
    static constexpr int HR = (NHALO + THREADS - 1) / THREADS;      // halo rows fed per thread
    static constexpr int SMEM_DOUBLES = SH * RP;
};
typedef MarchCfg<16, 16> MarchCfgDefault;
typedef MarchCfg<16, 16, 1, float> MarchCfgF32;

template <class T>
struct MarchParamsT {
    T *phi;
    const T *phiS;
    int nx, ny, nz;
    long long sa, sb, sc, off0;    // signed strides / origin of the sweep-oriented frame
    int fa, fb, fc;                // axis flipped?
    int lo_a, hi_a, lo_b, hi_b, lo_c, hi_c;   // high-order window (subs.f90:506) in oriented indices
    int c_lo, c_hi, c_max;         // oriented c: planes c_lo..c_hi are updated, planes 0..c_max exist
                                   // (whole grid on one GPU: 1, nz-1, nz; a z-slab with ghost planes: lsf_slab.cu)
    int ntb, ntc, ntiles;
    int tend;                      // last step index
    int slack;                     // a tile STARTS only once its two predecessors are this many steps further ahead than the data
                                   // dependence (TB/TC + CHUNK) needs: an elastic buffer against per-chunk timing jitter, which
                                   // otherwise couples all resident tiles like a barrier per chunk (27 % of warp time, round-2 ncu)
    CellConstT<T> cc;
    double *partial;               // per tile: sum over its cells of (new-old)^2
    unsigned *ticket;
    const int *order;              // ticket -> J | (K << 16)
    int *col_next;                 // [ntb] dynamic tile scheduler (march_pick): next tile row K of every tile column, zero before the
                                   // sweep; null: static tickets in `order`
    long long *progress;           // per tile (J + ntb*K)
    long long epoch;               // progress values are epoch*2^32 + step + BIAS
    Ctrl *ctrl;
    long long *dbg;                // timing experiments only (LSF_EXP_TIMING): 6 words per tile
    // ---- z-slab sharding: the streaming halo of the Gauss-Seidel pipeline (all zero on a single GPU) ----
    // The rank upstream in c (the one owning lower oriented c) pushes the NEW values of its last three
    // updated planes straight into this rank's ghost planes (peer stores over NVLink) and publishes the
    // progress of its last tile row into in_progress[J]; tiles (J,0) of this rank wait on that flag exactly
    // as they would on a predecessor tile of their own GPU.  Ghost planes on the +c side hold the OLD values
    // (snapshot exchanged between sweeps).
    const long long *in_progress;  // [ntb], written by the upstream rank; null: no upstream rank
    long long push_delta;          // (downstream rank's phi, shifted to this rank's indexing) - phi, in elements; 0: none
    long long *push_progress;      // downstream rank's in_progress; null: no downstream rank
    const long long *halo_seq;     // [2] counters the neighbours bump when they have refreshed this rank's ghost planes
    long long halo_need[2];        // the sweep starts once halo_seq[s] >= halo_need[s] (0: no neighbour on side s)
    // Write-through on the other side: the first three updated planes are also stored into the UPSTREAM
    // rank's ghost planes.  That rank reads them as OLD values LOOK steps ahead of its own march and this
    // rank runs >= TC steps behind it, so the store can never overtake the read -- the very argument that
    // makes the in-place update of neighbouring tiles safe on one GPU.  With both pushes a rank's ghost
    // planes are always what an in-place sweep of the whole grid would hold there, and consecutive sweeps
    // need no bulk halo exchange: only per-tile completion flags (edge_*).
    long long push_up_delta;       // (upstream rank's phi, shifted) - phi; 0: none
    long long *edge_pub[2];        // where the tiles of row K = 0 / K = ntc-1 publish "done in sweep `epoch`" (the adjacent rank's memory), or null
    const long long *edge_wait;    // [ntb of the previous sweep] completion flags of the DOWNSTREAM rank's adjacent tile row, or null
    long long edge_need;           // epoch of the previous sweep (tiles of row ntc-1 wait for it before reading / overwriting that rank's planes)
    int edge_prev_fb;              // b-orientation of the previous sweep (maps this sweep's tile columns onto that sweep's)
    // ---- overlapped sweeps: ONE launch runs several consecutive sweeps (march_multi_cta) ----
    // CTAs go on to the tiles of sweep s+1 while sweep s drains.  A tile of sweep s+1 starts once the 3x3 neighbourhood
    // of tiles of sweep s covering its cells and halo has finished.  The boundary block (subs.f90:858-897) is folded
    // into the tiles: a finished tile writes the boundary points its cells are the source of into the array the NEXT
    // sweep reads boundary values from -- phi and a shell array (phiN) alternate, so no reader of the current sweep
    // ever sees them -- and adds their part of the RMS sum.
    long long shell_rd_delta;      // (array holding the boundary values this sweep READS) - phi, in elements; 0: phi itself
    long long shell_wr_delta;      // (array the boundary values for the NEXT sweep are written to) - phi
    const long long *prev_progress;// progress flags of the previous sweep of this launch; null: first sweep
    long long prev_fin;            // value a finished tile of that sweep holds
    int prev_fb, prev_fc;          // b / c orientation of that sweep
    int fold_bc;                   // 1: finished tiles apply the boundary block for the points they source
    double *partial_bc;            // per tile: boundary part of the RMS sum
};
typedef MarchParamsT<double> MarchParams;

// Spin until *flag >= need.  SYS: the flag is written by a peer GPU.  Gives up (and poisons the loop
// status) after M_SPIN_LIMIT polls or as soon as another waiter has given up.
template <bool SYS>
LSF_DEV void wait_ge(const long long *flag, long long need, Ctrl *ctrl)
{
    long long spins = 0;
    while ((SYS ? p_ld_relaxed_sys(flag) : p_ld_relaxed(flag)) < need) {
        p_sleep();
        if (((++spins) & 1023) == 0) {
            if (*(volatile int *)&ctrl->status == M_ERR_TIMEOUT) return;
            if (spins > M_SPIN_LIMIT) { *(volatile int *)&ctrl->status = M_ERR_TIMEOUT; return; }
        }
    }
    if (SYS) p_fence_sys(); else p_fence_acquire();
}

template <class CFG>
struct MarchSmem {
    typename CFG::real S[CFG::SMEM_DOUBLES];
    double red[CFG::THREADS];
    unsigned long long bar;      // split step barrier (LSF_SPLIT_BAR)
    int tile;
};

// Host helper shared with the emulator: fill the orientation-dependent fields.
// nz = last plane index of the LOCAL array.  Whole grid on one GPU: kupd_lo = 1, kupd_hi = nz-1, kbase = 0,
// NZ = nz.  z-slab: local planes kupd_lo..kupd_hi are updated, local plane k is global plane k + kbase of
// a grid 0..NZ (the high-order window of subs.f90:506 is a property of the GLOBAL index).
template <class CFG, class T>
inline void march_orient(MarchParamsT<T> &p, int nx, int ny, int nz, long long sx, long long sxy, int raster,
                         int kupd_lo = 1, int kupd_hi = -1, int kbase = 0, int NZ = -1)
{
    if (kupd_hi < 0) kupd_hi = nz - 1;
    if (NZ < 0) NZ = nz;
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
    const int wlo = 4 - kbase, whi = NZ - 5 - kbase;          // window in local plane indices
    p.lo_c = p.fc ? nz - whi : wlo; p.hi_c = p.fc ? nz - wlo : whi;
    p.c_lo = p.fc ? nz - kupd_hi : kupd_lo; p.c_hi = p.fc ? nz - kupd_lo : kupd_hi; p.c_max = nz;
    p.ntb = (ny - 1 + CFG::TB - 1) / CFG::TB;
    p.ntc = (p.c_hi - p.c_lo + 1 + CFG::TC - 1) / CFG::TC;
    p.ntiles = p.ntb * p.ntc;
    p.tend = (nx - 1) - 1 + (CFG::TB - 1) + (CFG::TC - 1) + M_H;
    p.slack = 0;
    p.col_next = nullptr;
}

// ticket order: fronts m*J + K ascending (m = 1: anti-diagonals of the tile grid).  Topological for
// (J-1,K) -> (J,K) <- (J,K-1), and all tiles of a front are mutually independent, so the resident CTAs
// are never blocked for long.  m > 1 tilts the front so that the sweep crosses the c extent of the grid
// sooner: on a z-slab this is what lets the downstream rank start early (lsf_slab.cu).
inline void march_fill_order(int ntb, int ntc, int *order, int m = 1)
{
    int n = 0;
    for (int s = 0; s <= m * (ntb - 1) + ntc - 1; ++s)
        for (int K = s % m; K < ntc && K <= s; K += m) {
            const int J = (s - K) / m;
            if (J >= ntb) continue;
            order[n++] = J | (K << 16);
        }
}

// Ring access.  DUP: slot h lives at [h&7] and [(h&7)+8], the window t-3..t+3 is the contiguous run starting at (t-3)&7.
// Single copy: slot h lives at [h&7]; the caller computes the 7 wrapped offsets of the window once per step.
template <class real> LSF_DEV void ring_put(real *pos, int h, real v)
{
    real *d = pos + (h & (M_NSLOT - 1));
    d[0] = v;
    if (M_DUP) d[M_NSLOT] = v;
}
struct RingWin {
    int o[7];                        // DUP: o[k] = ((t-3)&7) + k ; single copy: (t-3+k)&7
    LSF_DEV explicit RingWin(int t)
    {
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = M_DUP ? ((t - M_H) & (M_NSLOT - 1)) + k : ((t - M_H + k) & (M_NSLOT - 1));
    }
};

// One cell of row `Sown` at step t: gather the 19 stencil values from the ring (orientation resolved at
// compile time) and update.  HI: the caller knows the high-order branch applies (compile-time), else `hi`.
template <class AR, bool FA, bool FB, bool FC, class CFG, bool HI>
LSF_DEV typename AR::real march_cell(const typename AR::real *Sown, int t, typename AR::real ps, bool hi,
                                     const CellConstT<typename AR::real> &cc, bool &sens, typename AR::real &df2)
{
    typedef typename AR::real real;
    constexpr int W = CFG::W, RP = CFG::RP;
    const RingWin w(t);                                       // window t-3..t+3
    real vx[7], vy[7], vz[7];
#pragma unroll
    for (int m = -3; m <= 3; ++m) {
        vx[FA ? 3 - m : 3 + m] = Sown[w.o[3 + m]];
        if (m != 0) {
            vy[FB ? 3 - m : 3 + m] = Sown[m * W + w.o[3 + m]];
            vz[FC ? 3 - m : 3 + m] = Sown[m * RP + w.o[3 + m]];
        }
    }
    vy[3] = vx[3]; vz[3] = vx[3];
    real g[3], gM;
    const real pn = reinit_cell<AR>(vx, vy, vz, ps, HI ? true : hi, cc, g, gM, sens);
    const real df = pn - vx[3];
    df2 = df * df;
    return pn;
}

// SPLIT variant of march_cell: the x direction (a, b) was computed ahead from the thread's register window; only the 12
// y/z stencil values come from the ring.  phic = the cell's current value (window centre).
template <class AR, bool FB, bool FC, class CFG>
LSF_DEV typename AR::real march_cell_yz(const typename AR::real *Sown, int t, typename AR::real phic, typename AR::real a, typename AR::real b,
                                        typename AR::real ps, bool hi, const CellConstT<typename AR::real> &cc, bool &sens,
                                        typename AR::real &df2)
{
    typedef typename AR::real real;
    constexpr int W = CFG::W, RP = CFG::RP;
    const RingWin w(t);                                       // window t-3..t+3
    real vy[7], vz[7];
#pragma unroll
    for (int m = -3; m <= 3; ++m) {
        if (m != 0) {
            vy[FB ? 3 - m : 3 + m] = p_lds(Sown + m * W + w.o[3 + m]);
            vz[FC ? 3 - m : 3 + m] = p_lds(Sown + m * RP + w.o[3 + m]);
        }
    }
    vy[3] = phic; vz[3] = phic;
    real g[3], gM;
    const real pn = reinit_cell_rest<AR>(a, b, vy, vz, ps, hi, cc, g, gM, sens);
    const real df = pn - phic;
    df2 = df * df;
    return pn;
}

// MG = false compiles the z-slab hooks (peer stores, peer flags) out of the single-GPU kernel.
// OV = true compiles the overlapped-sweeps hooks in (cross-sweep tile wait, boundary values from the shell array,
// folded boundary block); single GPU only.
template <class AR, bool FA, bool FB, bool FC, class CFG, bool MG = true, bool OV = false>
LSF_DEV void march_tile(const MarchParamsT<typename AR::real> &p, MarchSmem<CFG> &sm, const int tid, const int J, const int K,
                        unsigned &bar_phase)
{
    // split step barrier + own-row register window (scalar loads, one row per thread, one sweep per launch)
    constexpr bool SPLIT = (LSF_SPLIT_BAR != 0) && CFG::R == 1 && CFG::VEC == 1 && !OV;
    // Per-chunk duties go to warps WITHOUT halo duty (halo rows are fed by threads 0..NHALO-1): the warp that polls a flag or
    // publishes (fence + release store) is the one every other warp of the CTA waits for at the next barrier.
    constexpr int TID_PB = (LSF_SPREAD_DUTIES && CFG::THREADS >= 256) ? CFG::THREADS - 96 : 0;                    // polls predB
    constexpr int TID_PC = (LSF_SPREAD_DUTIES && CFG::THREADS >= 256) ? CFG::THREADS - 64 : 32 % CFG::THREADS;    // polls predC
    constexpr int TID_PUB = (LSF_SPREAD_DUTIES && CFG::THREADS >= 256) ? CFG::THREADS - 32 : 0;                   // publishes progress
    static_assert(!(OV && MG), "overlapped sweeps are a single-GPU schedule");
    static_assert(!OV || (CFG::VEC == 1 && CFG::R == 1), "overlapped sweeps use scalar loads, one row per thread");
    typedef typename AR::real real;
    constexpr int TB = CFG::TB, TC = CFG::TC, THREADS = CFG::THREADS, RP = CFG::RP, R = CFG::R;
    constexpr int W = CFG::W;
    // thread tid owns rows tid, tid + THREADS, ... of the tile (row q: tb = q % TB, tc = q / TB); all rows of a
    // step lie on one hyperplane, so the R cells a thread updates per step are independent of each other
    bool rowValid[R], compValid[R], pushRow[R], pushUpRow[R], hiBC[R];
    long long rowShell[R];           // OV: offset to the shell array if the whole row consists of boundary points
    int sig[R];
    real *Sown[R];
    real *pOut[R];
    const real *pSgn[R];
    constexpr long long SA = FA ? -1 : 1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int q = tid + r * THREADS;
        const int tb = q % TB, tc = q / TB;
        const int b = 1 + J * TB + tb, c = p.c_lo + K * TC + tc;
        rowValid[r] = (b <= p.ny) && (c <= p.c_max);
        compValid[r] = (b <= p.ny - 1) && (c <= p.c_hi);
        pushRow[r] = MG && (p.push_delta != 0) && compValid[r] && (c > p.c_hi - M_H);
        pushUpRow[r] = MG && (p.push_up_delta != 0) && compValid[r] && (c < p.c_lo + M_H);
        hiBC[r] = (b >= p.lo_b) && (b <= p.hi_b) && (c >= p.lo_c) && (c <= p.hi_c);
        rowShell[r] = (OV && (b == p.ny || c == p.c_max)) ? p.shell_rd_delta : 0;
        sig[r] = tb + tc + M_H;
        const long long rowoff = rowValid[r] ? p.off0 + (long long)b * p.sb + (long long)c * p.sc : 0;
        Sown[r] = sm.S + (tc + M_H) * RP + (tb + M_H) * W;
        // running pointers: cell a of the own row (phi, phiS), advanced by one cell per step (a = 1 + t - sig)
        pOut[r] = p.phi + rowoff + (long long)(1 - M_LOOK - sig[r]) * SA;
        pSgn[r] = p.phiS + rowoff + (long long)(1 - M_LOOK - sig[r]) * SA;
    }

    // halo duty: thread q feeds halo rows q, q+THREADS, ... (< NHALO)
    bool hvalid[CFG::HR], hlow[CFG::HR];
    long long hShell[CFG::HR];
    int hsig[CFG::HR];
    const real *hrow[CFG::HR];
    real *hS[CFG::HR];
#pragma unroll
    for (int r = 0; r < CFG::HR; ++r) {
        const int q = tid + r * THREADS;
        hvalid[r] = false; hlow[r] = false; hsig[r] = 0; hrow[r] = p.phi; hS[r] = sm.S; hShell[r] = 0;
        if (q < CFG::NHALO) {
            int htb, htc;
            if (q < 2 * M_H * TC) {                     // -b / +b sides: 3 rows x TC each
                const int side = q / (M_H * TC), rr = q % (M_H * TC);
                const int m = rr / TC + 1, idx = rr % TC;
                htc = idx;
                if (side == 0) { htb = -m; hlow[r] = true; } else htb = TB - 1 + m;
            } else {                                    // -c / +c sides: 3 rows x TB each
                const int q2 = q - 2 * M_H * TC;
                const int side = q2 / (M_H * TB), rr = q2 % (M_H * TB);
                const int m = rr / TB + 1, idx = rr % TB;
                htb = idx;
                if (side == 0) { htc = -m; hlow[r] = true; } else htc = TC - 1 + m;
            }
            const int hb = 1 + J * TB + htb, hc = p.c_lo + K * TC + htc;
            hvalid[r] = (hb >= 0) && (hb <= p.ny) && (hc >= 0) && (hc <= p.c_max);
            hsig[r] = htb + htc + M_H;
            hS[r] = sm.S + (htc + M_H) * RP + (htb + M_H) * W;
            if (hvalid[r]) hrow[r] = p.phi + p.off0 + (long long)hb * p.sb + (long long)hc * p.sc;
            if (OV && (hb == 0 || hb == p.ny || hc == 0 || hc == p.c_max)) hShell[r] = p.shell_rd_delta;
        }
    }

    const long long ebase = p.epoch << 32;
    const long long *predB = (J > 0) ? p.progress + ((J - 1) + p.ntb * K) : nullptr;
    const bool predCpeer = MG && (K == 0) && p.in_progress;       // predecessor in c lives on the upstream rank
    const long long *predC = (K > 0) ? p.progress + (J + p.ntb * (K - 1)) : (predCpeer ? p.in_progress + J : nullptr);
    long long *mine = p.progress + (J + p.ntb * K);
    long long *minePeer = (MG && K == p.ntc - 1 && p.push_progress) ? p.push_progress + J : nullptr;

    // overlapped sweeps: every tile of the previous sweep that holds a cell this tile reads or overwrites (its rows and
    // their 3-wide halo) must be finished -- tile columns / rows are counted in that sweep's own orientation
    if (OV && p.prev_progress) {
        const int blo = m_imax(1, 1 + J * TB - M_H), bhi = m_imin(p.ny - 1, J * TB + TB + M_H);
        const int clo = m_imax(1, 1 + K * TC - M_H), chi = m_imin(p.c_max - 1, K * TC + TC + M_H);
        const bool flipb = p.prev_fb != (FB ? 1 : 0), flipc = p.prev_fc != (FC ? 1 : 0);
        const int b0 = flipb ? p.ny - bhi : blo, b1 = flipb ? p.ny - blo : bhi;
        const int c0 = flipc ? p.c_max - chi : clo, c1 = flipc ? p.c_max - clo : chi;
        const int J0 = (b0 - 1) / TB, J1 = (b1 - 1) / TB, K0 = (c0 - 1) / TC, K1 = (c1 - 1) / TC;
        const int nj = J1 - J0 + 1, nk = K1 - K0 + 1;
        if (tid < nj * nk) wait_ge<false>(p.prev_progress + ((J0 + tid % nj) + p.ntb * (K0 + tid / nj)), p.prev_fin, p.ctrl);
        p_sync();
    }

    // z-slabs: before a tile of the last row touches the downstream rank's planes (reads them as OLD values,
    // streams NEW values into its ghost planes) that rank's adjacent tiles of the PREVIOUS sweep covering the
    // same physical j range must be complete (they may have used a different b orientation).
    // (a short last tile row leaves some of the three boundary planes to the row before it: any tile whose
    // rows or +c halo reach beyond c_hi is concerned)
    if (MG && p.edge_wait && p.c_lo + (K + 1) * TC + M_H - 1 > p.c_hi) {
        if (tid < 2) {
            const int bq = 1 + J * TB + (tid ? TB - 1 : 0);
            const int be = bq < p.ny - 1 ? bq : p.ny - 1;                       // clamp to the updated range
            const int bp = (p.edge_prev_fb != (FB ? 1 : 0)) ? p.ny - be : be;    // same physical j in the previous orientation
            wait_ge<true>(p.edge_wait + (bp - 1) / TB, p.edge_need, p.ctrl);
        }
        p_sync();
    }

    // start slack (see MarchParamsT::slack): same flags, a larger head start
    if (p.slack > 0) {
        if (tid == TID_PB && predB) wait_ge<false>(predB, ebase + M_BIAS + (M_CHUNK - 1 + TB + (CFG::VEC - 1) + M_PFM + p.slack), p.ctrl);
        if (tid == TID_PC && predC) {
            const long long need = ebase + M_BIAS + (M_CHUNK - 1 + TC + (CFG::VEC - 1) + M_PFM + p.slack);
            if (predCpeer) wait_ge<true>(predC, need, p.ctrl); else wait_ge<false>(predC, need, p.ctrl);
        }
        p_sync();
    }

    real acc = 0;
    constexpr int VEC = CFG::VEC;
    RowReader<real, VEC, FA> rdLook, rdSgn, rdHalo[CFG::HR];
    rdLook.reset(); rdSgn.reset();
#pragma unroll
    for (int r = 0; r < CFG::HR; ++r) rdHalo[r].reset();
#if defined(LSF_EXP_TIMING)
    long long dbg_wait = 0, dbg_t0 = 0, dbg_c0 = 0;
    if (tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0)); dbg_c0 = clock64(); }
#endif
    // cell ah of each halo row, advanced by one cell per step
    const real *hp[CFG::HR];
#pragma unroll
    for (int r = 0; r < CFG::HR; ++r)
        hp[r] = hrow[r] + (long long)(1 - M_LOOK + (hlow[r] ? 0 : M_LOOK) - hsig[r]) * SA;

    long long preB = 0, preC = 0;           // LSF_ASYNC_POLL: flag values read ahead of the wait
    // SPLIT: xw[k] = what slot t-3+k of the own row holds when step t is computed (the thread's own new values behind,
    // its own look-ahead loads ahead), xa/xb = the x-direction one-sided derivatives of the cell of step t, computed at
    // the end of step t-1 between the arrive and the wait of that step's barrier
    real xw[7] = {0, 0, 0, 0, 0, 0, 0};
    real xa = 0, xb = 0;
    // Steady state.  On a tile whose rows and halo rows all exist and lie inside the high-order window in b and c (all but the
    // outermost ring of tiles), every range test of a step -- look-ahead load, cell active, high-order branch, halo row in range,
    // next cell's x direction -- holds for every thread during steps T0 .. T1-1 (92 % of the steps at nx = 1024).  The step body
    // is compiled a second time for that range (ST = true) with the tests folded away: straight-line code, no low-order branch.
    //   a = 1+t-sig, sig in [M_H, TB+TC-2+M_H]:  a >= lo_a for all rows  <=>  t >= lo_a + TB + TC;  a+1 <= hi_a  <=>  t <= hi_a + 1
    //   (hi_a <= nx-4 covers a+LOOK <= nx, lo_a >= 4 covers the halo rows, whose sig is at most TB+TC+4)
    int T0 = 0, T1 = 0;
    if (LSF_STEADY && R == 1 && !OV && CFG::VEC == 1) {
        const int b0 = 1 + J * TB, c0 = p.c_lo + K * TC;
        const bool interior = (b0 >= p.lo_b) && (b0 + TB - 1 <= p.hi_b) && (b0 - M_H >= 0) && (b0 + TB - 1 + M_H <= p.ny) &&
                              (c0 >= p.lo_c) && (c0 + TC - 1 <= p.hi_c) && (c0 + TC - 1 <= p.c_hi) && (c0 - M_H >= 0) &&
                              (c0 + TC - 1 + M_H <= p.c_max) && (p.lo_a >= 4) && (p.hi_a <= p.nx - 4);
        if (interior) { T0 = p.lo_a + TB + TC; T1 = p.hi_a + 2; }
        if (T1 < T0) T1 = T0;
    }
    // Prefetch (PF).  The values step t needs from global memory -- the look-ahead cell of the own row, phiS of the cell being
    // updated, one cell of the halo row this thread feeds -- are loaded during step t-1 and carried in registers, so their L2 /
    // HBM latency overlaps a whole step of arithmetic instead of the few hundred instructions between a load and its use.  All
    // three are safe a step early: phiS is read-only; look-ahead and +b/+c halo cells are OLD values that their owner overwrites
    // many steps later; -b/-c halo cells are the predecessors' NEW values, for which every flag test asks for one more step of
    // progress (M_PFM).  The first flag test of a tile sits at the end of step -1, so the -b/-c cells of step 0 are loaded in
    // step 0 itself.
    // (not on z-slabs: there the same one-step-ahead phiS load made every sweep 2.6x slower -- session 15/17, N = 2: 95 ms instead
    // of 36 ms per sweep kernel with bit-identical results; the tiles of the slab's first and last tile rows, which execute
    // system-scope fences and peer stores every chunk and pace all other tiles, do not tolerate a load in flight across them)
    constexpr bool PF = (LSF_PREFETCH != 0) && R == 1 && CFG::VEC == 1 && !OV && (!MG || LSF_PREFETCH_MG);
    real laN = 0, psN = 0, hvN[CFG::HR];
#pragma unroll
    for (int r = 0; r < CFG::HR; ++r) hvN[r] = 0;
    // loads of step tn; pOut / pSgn point `off` cells before those of step tn, hp at them
    auto prefetch = [&](const int tn, const int off, auto steady_tag) {
        constexpr bool ST = decltype(steady_tag)::value;
        const int a = 1 + tn - sig[0], a4 = a + M_LOOK;
        laN = 0; psN = 0;
        if ((LSF_PREFETCH & 2) && (ST || (rowValid[0] && (a4 >= 0) && (a4 <= p.nx)))) laN = p_ldcg(pOut[0] + (M_LOOK + off) * SA);
        if ((LSF_PREFETCH & 1) && (ST || (compValid[0] && (a >= 1) && (a <= p.nx - 1)))) psN = p_ldcg(pSgn[0] + off * SA);
#pragma unroll
        for (int r = 0; r < CFG::HR; ++r) {
            hvN[r] = 0;
            if (hvalid[r]) {
                const int hhn = hlow[r] ? tn : tn + M_LOOK;
                const int ah = 1 + hhn - hsig[r];
                if (!(LSF_PREFETCH & (hlow[r] ? 8 : 4))) continue;
                if (ST || (hhn >= 0 && ah >= 0 && ah <= p.nx && !(hlow[r] && tn == 0))) hvN[r] = p_ldcg(hp[r]);
            }
        }
    };
    if constexpr (PF) prefetch(-M_LOOK, 0, StepGeneric());
    auto body = [&](const int t, auto steady_tag) {
        constexpr bool ST = decltype(steady_tag)::value;
        // ---- wait for the two predecessor tiles at chunk starts ---------------------------
        // (SPLIT && LSF_FOLD_FLAGS: folded into the previous step's barrier instead, see below)
        if (!(SPLIT && LSF_FOLD_FLAGS) && t >= 0 && (t % M_CHUNK) == 0) {
            // (a vector load of a -b/-c halo row fetches the predecessor's cells of up to VEC-1 later steps)
            const long long need_b = ebase + M_BIAS + (t + M_CHUNK - 1 + TB + (VEC - 1) + M_PFM);
            const long long need_c = ebase + M_BIAS + (t + M_CHUNK - 1 + TC + (VEC - 1) + M_PFM);
#if defined(LSF_EXP_TIMING)
            long long tw0 = 0;
            if (tid == 0) tw0 = clock64();
#endif
            if (tid == TID_PB && predB) {
                if (LSF_ASYNC_POLL && preB >= need_b) p_fence_acquire(); else wait_ge<false>(predB, need_b, p.ctrl);
            }
            if (tid == TID_PC && predC) {
                if (LSF_ASYNC_POLL && preC >= need_c) { if (predCpeer) p_fence_sys(); else p_fence_acquire(); }
                else if (predCpeer) wait_ge<true>(predC, need_c, p.ctrl); else wait_ge<false>(predC, need_c, p.ctrl);
            }
            p_sync();
#if defined(LSF_EXP_TIMING)
            if (tid == 0) dbg_wait += clock64() - tw0;
#endif
        }
        const bool chunk_next = (t + 1 >= 0) && (((t + 1) % M_CHUNK) == 0);       // the next step starts a chunk
        if ((LSF_ASYNC_POLL || (SPLIT && LSF_FOLD_FLAGS)) && chunk_next) {          // the flags the next chunk start will test: the
            if (tid == TID_PB && predB) preB = p_ld_relaxed(predB);            // L2 round trip overlaps this step's arithmetic
            if (tid == TID_PC && predC) preC = predCpeer ? p_ld_relaxed_sys(predC) : p_ld_relaxed(predC);
        }
        p_emu_hook(tid == 0 && t == 6);
        if (LSF_L2_AHEAD > 0 && R == 1 && !OV && (t & 15) == 0) {
            const int af = 1 + t - sig[0] + M_LOOK + LSF_L2_AHEAD;            // the cell the look-ahead reaches LSF_L2_AHEAD steps from now
            if (rowValid[0] && af >= 0 && af <= p.nx) {
                p_prefetch_l2(pOut[0] + (M_LOOK + LSF_L2_AHEAD) * SA);
                p_prefetch_l2(pSgn[0] + (M_LOOK + LSF_L2_AHEAD) * SA);
            }
        }
        // ---- (1) issue the global loads of this step --------------------------------------
        bool ldLook[R], active[R], hi[R];
        real la[R], ps[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int a = 1 + t - sig[r];
            const int a4 = a + M_LOOK;
            ldLook[r] = ST ? true : (rowValid[r] && (a4 >= 0) && (a4 <= p.nx));
            la[r] = 0;
#if !(LSF_EXP_NOLDG_MASK & 1)
            if constexpr (PF && (LSF_PREFETCH & 2)) la[r] = laN;
            else if (ldLook[r]) {
                if constexpr (VEC > 1) la[r] = rdLook.get(pOut[r] + M_LOOK * SA);
                else if (OV) la[r] = p_ldcg(pOut[r] + M_LOOK * SA + ((a4 == 0 || a4 == p.nx) ? p.shell_rd_delta : rowShell[r]));
                else la[r] = (LSF_LD_CACHED & 1) ? p_ldca(pOut[r] + M_LOOK * SA) : p_ldcg(pOut[r] + M_LOOK * SA);
            }
#endif
            active[r] = ST ? true : (compValid[r] && (a >= 1) && (a <= p.nx - 1));
            ps[r] = 0;
#if !(LSF_EXP_NOLDG_MASK & 2)
            if constexpr (PF && (LSF_PREFETCH & 1)) ps[r] = psN;
            else if (active[r]) {
                if constexpr (VEC > 1) ps[r] = rdSgn.get(pSgn[r]);
                else ps[r] = (LSF_LD_CACHED & 2) ? p_ldca(pSgn[r]) : p_ldcg(pSgn[r]);
            }
#else
            if (active[r]) ps[r] = (real)0.5;
#endif
            hi[r] = ST ? true : (hiBC[r] && (a >= p.lo_a) && (a <= p.hi_a));
        }
        bool hdep[CFG::HR];
        real hv[CFG::HR];
        int hh[CFG::HR];
#pragma unroll
        for (int r = 0; r < CFG::HR; ++r) {
            hdep[r] = false; hv[r] = 0; hh[r] = 0;
            if (hvalid[r]) {
                hh[r] = hlow[r] ? t : t + M_LOOK;
                const int ah = 1 + hh[r] - hsig[r];
#if !(LSF_EXP_NOLDG_MASK & 4)
                if (ST || (hh[r] >= 0 && ah >= 0 && ah <= p.nx)) {
                    if constexpr (PF) hv[r] = (!(LSF_PREFETCH & (hlow[r] ? 8 : 4)) || (hlow[r] && !ST && t == 0)) ? p_ldcg(hp[r]) : hvN[r];
                    else if constexpr (VEC > 1) hv[r] = rdHalo[r].get(hp[r]);
                    else if (OV) hv[r] = p_ldcg(hp[r] + ((ah == 0 || ah == p.nx) ? p.shell_rd_delta : hShell[r]));
                    else hv[r] = ((LSF_LD_CACHED & 4) && !hlow[r]) ? p_ldca(hp[r]) : p_ldcg(hp[r]);
                    hdep[r] = true;
                }
#else
                if (hh[r] >= 0 && ah >= 0 && ah <= p.nx) { hv[r] = (real)hh[r]; hdep[r] = true; }
#endif
            }
            hp[r] += SA;
        }
        if constexpr (PF) prefetch(t + 1, 1, steady_tag);          // the loads of step t+1 (hp has moved on, pOut / pSgn not yet)
#if !defined(LSF_EMU)
        // A never-taken branch ends the basic block here: ptxas does not move the loads above past it, so they are issued before
        // the arithmetic of the step instead of right before their first use.  LSF_PIN_LOADS bit 0: single-GPU kernels (session
        // 25: 33.3 instead of 34.5 Gcell/s -- off), bit 1: z-slab kernels, whose steady loop otherwise has 28-72 instructions
        // between a load and its use (session 28, N = 2: 60.1 instead of 50.4 Gcell/s fp64, 72.1 instead of 67.5 fp32 -- on)
        if ((LSF_PIN_LOADS & (MG ? 2 : 1)) && p.tend < 0) asm volatile("trap;");
#endif
        // ---- (2) cell updates ---------------------------------------------------------------
        real pn[R];
        bool sens = false;
        bool fused = false;
        if (R == 2) {
            // both cells on the high-order branch (the bulk of the grid): one straight-line block, so the two
            // independent dependence chains interleave on the FP64 pipe
            if (active[0] && active[R - 1] && hi[0] && hi[R - 1]) {
                bool s0, s1;
                real d0, d1;
                pn[0] = march_cell<AR, FA, FB, FC, CFG, true>(Sown[0], t, ps[0], true, p.cc, s0, d0);
                pn[R - 1] = march_cell<AR, FA, FB, FC, CFG, true>(Sown[R - 1], t, ps[R - 1], true, p.cc, s1, d1);
                sens = s0 || s1;
                acc += d0;
                acc += d1;
                fused = true;
            }
        }
        if (!fused) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                pn[r] = 0;
                if (active[r]) {
                    bool s0;
                    real d0;
                    if constexpr (SPLIT) pn[r] = march_cell_yz<AR, FB, FC, CFG>(Sown[r], t, xw[3], xa, xb, ps[r], hi[r], p.cc, s0, d0);
                    else pn[r] = march_cell<AR, FA, FB, FC, CFG, false>(Sown[r], t, ps[r], hi[r], p.cc, s0, d0);
                    sens = sens || s0;
                    acc += d0;
                }
            }
        }
        if (sens) p.ctrl->guard = 1;
        // ---- (3) stores and deposits into the slot ring (each value to slot h&7 and its double) -------
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (active[r]) {
#if !defined(LSF_EXP_NOSTG)
                p_stcg(pOut[r], pn[r]);
#endif
                if (pushRow[r]) p_st_peer(pOut[r] + p.push_delta, pn[r]);
                if (pushUpRow[r]) p_st_peer(pOut[r] + p.push_up_delta, pn[r]);
                ring_put(Sown[r], t, pn[r]);
            }
            if (ldLook[r]) ring_put(Sown[r], t + M_LOOK, la[r]);
            pOut[r] += SA; pSgn[r] += SA;
        }
#pragma unroll
        for (int r = 0; r < CFG::HR; ++r)
            if (hdep[r]) ring_put(hS[r], hh[r], hv[r]);
        // ---- (4) publish progress every CHUNK steps -----------------------------------------
        // (all threads' stores -> CTA barrier -> one thread's gpu-scope release: cumulative, so the
        // whole tile's stores of this chunk are visible to whoever acquires the flag)
        const bool pub = (t >= M_PFM) && (((t - M_PFM) % M_CHUNK) == M_CHUNK - 1);
#if defined(LSF_EXP_NOSYNC)          // timing experiment only (results are wrong): no CTA barrier per step
        __syncwarp();
#else
        if constexpr (SPLIT) {
            if (LSF_FOLD_FLAGS && chunk_next) {
                // The predecessor flags of the chunk that starts with step t+1 were read at the top of this step (the L2 round
                // trip is long over).  The two polling threads settle them BEFORE they arrive: whoever passes this step's
                // barrier knows the flags hold -- no separate CTA barrier at the chunk start.
                const long long need_b = ebase + M_BIAS + (t + 1 + M_CHUNK - 1 + TB + (VEC - 1) + M_PFM);
                const long long need_c = ebase + M_BIAS + (t + 1 + M_CHUNK - 1 + TC + (VEC - 1) + M_PFM);
#if defined(LSF_EXP_TIMING)
                long long tw0 = 0;
                if (tid == 0) tw0 = clock64();
#endif
                if (tid == TID_PB && predB) {
                    if (preB >= need_b) { if (!LSF_NO_ACQ_FENCE) p_fence_acquire(); } else wait_ge<false>(predB, need_b, p.ctrl);
                }
                if (tid == TID_PC && predC) {
                    if (preC >= need_c) { if (predCpeer) p_fence_sys(); else if (!LSF_NO_ACQ_FENCE) p_fence_acquire(); }
                    else if (predCpeer) wait_ge<true>(predC, need_c, p.ctrl); else wait_ge<false>(predC, need_c, p.ctrl);
                }
#if defined(LSF_EXP_TIMING)
                if (tid == 0) dbg_wait += clock64() - tw0;
#endif
            }
            p_bar_arrive(&sm.bar);                 // this thread's deposits of step t are done
            // advance the own-row window to step t+1 and compute that cell's x direction: registers only
            const real c0 = active[0] ? pn[0] : xw[3];
            xw[0] = xw[1]; xw[1] = xw[2]; xw[2] = c0; xw[3] = xw[4]; xw[4] = xw[5]; xw[5] = xw[6]; xw[6] = la[0];
            const int an = 2 + t - sig[0];
            if (ST || (compValid[0] && (an >= 1) && (an <= p.nx - 1))) {
                real vxn[7];
#pragma unroll
                for (int m = -3; m <= 3; ++m) vxn[FA ? 3 - m : 3 + m] = xw[3 + m];
                reinit_dir_x<AR>(vxn, ST ? true : (hiBC[0] && (an >= p.lo_a) && (an <= p.hi_a)), p.cc, xa, xb);
            }
            p_bar_wait(&sm.bar, bar_phase++);      // every thread's deposits of step t are visible
        } else p_sync();
#endif
        if (pub && tid == TID_PUB) {
            if (LSF_PUB_FENCE) p_fence();
            p_st_release(mine, ebase + M_BIAS + t);
            if (minePeer) { p_fence_sys(); p_st_release_sys(minePeer, ebase + M_BIAS + t); }
        }
    };
    {
        int t = -M_LOOK;
        for (int pass = 0; pass < 2; ++pass) {           // ramp-in, steady state, ramp-out (one call site per body variant)
            const int tstop = (pass == 0) ? T0 : p.tend + 1;
            for (; t < tstop; ++t) body(t, StepGeneric());
            if (pass == 0)
                for (; t < T1; ++t) body(t, StepSteady());
        }
    }
    // ---- overlapped sweeps: the boundary block for the points this tile's cells are the source of ------------
    // (closed form of subs.f90:858-897, see k_reinit_bc_rms: phi(c) = phi(clamp(c)) + dx applied min(1+H,B) times,
    // B boundary axes of which H on the physical high side).  Results go to the array the NEXT sweep reads boundary
    // values from; the differences to the values THIS sweep read are the boundary part of the RMS sum.
    if (OV && p.fold_bc) {
        p_sync();                                       // the tile's cells are in global memory (L2) for all its threads
        const real *Xrd = p.phi + p.shell_rd_delta;
        real *Xwr = p.phi + p.shell_wr_delta;
        const int bA = 1 + J * TB, bB = m_imin(p.ny - 1, bA + TB - 1);
        const int cA = 1 + K * TC, cB = m_imin(p.c_max - 1, cA + TC - 1);
        double accb = 0.;
        auto point = [&](int a2, int b2, int c2, int b, int c) {        // boundary point (a2,b2,c2), source row (b,c)
            const int as = a2 < 1 ? 1 : (a2 > p.nx - 1 ? p.nx - 1 : a2);
            const int i = FA ? p.nx - a2 : a2, j = FB ? p.ny - b2 : b2, k = FC ? p.c_max - c2 : c2;
            const int Bn = (i == 0 || i == p.nx) + (j == 0 || j == p.ny) + (k == 0 || k == p.c_max);
            const int Hn = (i == p.nx) + (j == p.ny) + (k == p.c_max);
            const int m = 1 + Hn < Bn ? 1 + Hn : Bn;
            real v = p_ldcg(p.phi + p.off0 + (long long)as * SA + (long long)b * p.sb + (long long)c * p.sc);
            for (int r2 = 0; r2 < m; ++r2) v = bc_add(v, p.cc.dx);
            const long long q = p.off0 + (long long)a2 * SA + (long long)b2 * p.sb + (long long)c2 * p.sc;
            const double d = (double)v - (double)p_ldcg(Xrd + q);
            accb += d * d;
            p_stcg(Xwr + q, v);
        };
        // (1) the two ends of every computed row
        {
            const int tb = tid % TB, tc = tid / TB;
            const int b = bA + tb, c = cA + tc;
            if (b <= bB && c <= cB) { point(0, b, c, b, c); point(p.nx, b, c, b, c); }
        }
        // (2) whole boundary rows next to computed rows on the b / c faces (tiles touching a face only)
        if (bA == 1 || bB == p.ny - 1 || cA == 1 || cB == p.c_max - 1) {
            for (int c = cA; c <= cB; ++c)
                for (int b = bA; b <= bB; ++b)
                    for (int eb = 0; eb < 3; ++eb)
                        for (int ec = 0; ec < 3; ++ec) {
                            if (eb == 0 && ec == 0) continue;
                            if ((eb == 1 && b != 1) || (eb == 2 && b != p.ny - 1)) continue;
                            if ((ec == 1 && c != 1) || (ec == 2 && c != p.c_max - 1)) continue;
                            const int b2 = eb == 0 ? b : (eb == 1 ? 0 : p.ny), c2 = ec == 0 ? c : (ec == 1 ? 0 : p.c_max);
                            for (int a2 = tid; a2 <= p.nx; a2 += THREADS) point(a2, b2, c2, b, c);
                        }
        }
        sm.red[tid] = accb;
        p_sync();
        for (int wdt = THREADS / 2; wdt > 0; wdt >>= 1) {
            if (tid < wdt) sm.red[tid] = sm.red[tid] + sm.red[tid + wdt];
            p_sync();
        }
        if (tid == 0) p.partial_bc[J + p.ntb * K] = sm.red[0];
        p_sync();
    }
    // ---- tile done: final publish + deterministic block reduction of the RMS partial --------
    sm.red[tid] = (double)acc;
    p_sync();
    if (tid == 0) {
        if (LSF_PUB_FENCE) p_fence();
        p_st_release(mine, ebase + M_BIAS + M_FIN);
        if (minePeer) { p_fence_sys(); p_st_release_sys(minePeer, ebase + M_BIAS + M_FIN); }
        if (MG && K == 0 && p.edge_pub[0]) { p_fence_sys(); p_st_release_sys(p.edge_pub[0] + J, p.epoch); }
        if (MG && K == p.ntc - 1 && p.edge_pub[1]) { p_fence_sys(); p_st_release_sys(p.edge_pub[1] + J, p.epoch); }
    }
    for (int wdt = THREADS / 2; wdt > 0; wdt >>= 1) {
        if (tid < wdt) sm.red[tid] = sm.red[tid] + sm.red[tid + wdt];
        p_sync();
    }
    if (tid == 0) p.partial[J + p.ntb * K] = sm.red[0];
#if defined(LSF_EXP_TIMING)
    if (tid == 0 && p.dbg) {
        long long t1; unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long *d = p.dbg + 6 * (long long)(J + p.ntb * K);
        d[0] = dbg_t0; d[1] = t1; d[2] = dbg_wait; d[3] = clock64() - dbg_c0; d[4] = smid; d[5] = blockIdx.x;
    }
#endif
    p_sync();
}

// Dynamic tile scheduler (one lane per CTA).  A static ticket order hands a CTA a tile whether or not that tile can start, and
// a CTA holding a tile whose predecessors are not far enough ahead just spins -- which is what makes a steep ticket order (needed
// on z-slabs so that the sweep reaches the downstream rank quickly) expensive.  Here every tile column J keeps the index of its
// next unstarted tile, col_next[J]; a free CTA takes the READY tile of the lowest column -- ready: both predecessors (and, on the
// first tile row of a z-slab, the upstream rank) have published the progress the tile's first chunk needs -- by a compare-and-swap
// on the column's counter.  Lowest column first is column-major priority: the last tile row, which the downstream rank waits for,
// completes as early as the dependences allow, and no CTA ever holds a tile that cannot run.  Deadlock-free: a tile is only taken
// when its predecessors are running or done.  Returns J | (K << 16), or -1 when every column is finished.
template <class CFG, bool MG, class T>
LSF_DEV int march_pick(const MarchParamsT<T> &p, int &jlo)
{
    constexpr int TB = CFG::TB, TC = CFG::TC;
    const long long ebase = p.epoch << 32;
    const long long need_b = ebase + M_BIAS + (M_CHUNK - 1 + TB + (CFG::VEC - 1) + M_PFM + p.slack);
    const long long need_c = ebase + M_BIAS + (M_CHUNK - 1 + TC + (CFG::VEC - 1) + M_PFM + p.slack);
    long long spins = 0;
    for (;;) {
        bool all_done = true;
        for (int J = jlo; J < p.ntb; ++J) {
            const int K = p_ld_relaxed_i32(p.col_next + J);
            if (K >= p.ntc) { if (all_done) jlo = J + 1; continue; }
            all_done = false;
            if (J > 0 && p_ld_relaxed(p.progress + ((J - 1) + p.ntb * K)) < need_b) continue;
            bool ready = true;
            if (K > 0) ready = p_ld_relaxed(p.progress + (J + p.ntb * (K - 1))) >= need_c;
            else if (MG && p.in_progress) ready = p_ld_relaxed_sys(p.in_progress + J) >= need_c;
            if (ready && p_cas_i32(p.col_next + J, K, K + 1)) return J | (K << 16);
        }
        if (all_done) return -1;
        p_sleep();
        if (((++spins) & 1023) == 0) {
            if (*(volatile int *)&p.ctrl->status == M_ERR_TIMEOUT) return -1;
            if (spins > M_SPIN_LIMIT) { *(volatile int *)&p.ctrl->status = M_ERR_TIMEOUT; return -1; }
        }
    }
}

// Persistent CTA: take tickets until the tile list is exhausted.
template <class AR, bool FA, bool FB, bool FC, class CFG, bool MG = true, bool OV = false>
LSF_DEV void march_cta(const MarchParamsT<typename AR::real> &p, MarchSmem<CFG> &sm, const int tid)
{
    if (p.ctrl->done) return;
    unsigned bar_phase = 0;
    if (LSF_SPLIT_BAR) {
        if (tid == 0) p_bar_init(&sm.bar, CFG::THREADS);
        p_sync();
    }
    if (MG && p.halo_seq) {          // z-slab: both neighbours must have refreshed this rank's ghost planes
        if (tid == 0) {
            if (p.halo_need[0]) wait_ge<true>(p.halo_seq + 0, p.halo_need[0], p.ctrl);
            if (p.halo_need[1]) wait_ge<true>(p.halo_seq + 1, p.halo_need[1], p.ctrl);
        }
        p_sync();
    }
    int jlo = 0;                     // first tile column that still has unstarted tiles (march_pick)
    for (;;) {
        if (tid == 0) sm.tile = p.col_next ? march_pick<CFG, MG>(p, jlo) : (int)p_ticket(p.ticket);
        p_sync();
        const int tk = sm.tile;
        p_sync();
        if (p.col_next ? tk < 0 : tk >= p.ntiles) break;
        const int jk = p.col_next ? tk : p.order[tk];
        march_tile<AR, FA, FB, FC, CFG, MG, OV>(p, sm, tid, jk & 0xffff, jk >> 16, bar_phase);
    }
}

// Overlapped sweeps: one launch works through the tiles of `nsweeps` consecutive sweeps (ticket q -> sweep q / ntiles,
// tile order[q % ntiles]).  Tickets are handed out sweep by sweep and topologically within a sweep, so whatever a tile
// waits for has a smaller ticket and is held by a resident CTA: no deadlock.  `sweeps` is the per-sweep parameter
// array (on the GPU: a __grid_constant__ kernel parameter, i.e. constant-bank reads with a run-time index).
template <class AR, class CFG>
LSF_DEV void march_multi_cta(const MarchParamsT<typename AR::real> *sweeps, int nsweeps, unsigned *ticket, MarchSmem<CFG> &sm,
                             const int tid)
{
    if (sweeps[0].ctrl->done) return;
    unsigned bar_phase = 0;          // unused: the split barrier is a one-sweep-per-launch feature
    const int ntiles = sweeps[0].ntiles;
    for (;;) {
        if (tid == 0) sm.tile = (int)p_ticket(ticket);
        p_sync();
        const int tk = sm.tile;
        p_sync();
        if (tk >= ntiles * nsweeps) break;
        const MarchParamsT<typename AR::real> &p = sweeps[tk / ntiles];
        const int jk = p.order[tk % ntiles];
        const int J = jk & 0xffff, K = jk >> 16;
        switch ((p.fa ? 1 : 0) | (p.fb ? 2 : 0) | (p.fc ? 4 : 0)) {
        case 0: march_tile<AR, false, false, false, CFG, false, true>(p, sm, tid, J, K, bar_phase); break;
        case 1: march_tile<AR, true, false, false, CFG, false, true>(p, sm, tid, J, K, bar_phase); break;
        case 2: march_tile<AR, false, true, false, CFG, false, true>(p, sm, tid, J, K, bar_phase); break;
        case 3: march_tile<AR, true, true, false, CFG, false, true>(p, sm, tid, J, K, bar_phase); break;
        case 4: march_tile<AR, false, false, true, CFG, false, true>(p, sm, tid, J, K, bar_phase); break;
        case 5: march_tile<AR, true, false, true, CFG, false, true>(p, sm, tid, J, K, bar_phase); break;
        case 6: march_tile<AR, false, true, true, CFG, false, true>(p, sm, tid, J, K, bar_phase); break;
        default: march_tile<AR, true, true, true, CFG, false, true>(p, sm, tid, J, K, bar_phase); break;
        }
    }
}

// Run-time orientation -> compile-time orientation.
template <class AR, class CFG>
LSF_DEV void march_cta_any(const MarchParamsT<typename AR::real> &p, MarchSmem<CFG> &sm, const int tid)
{
    const int o = (p.fa ? 1 : 0) | (p.fb ? 2 : 0) | (p.fc ? 4 : 0);
    switch (o) {
    case 0: march_cta<AR, false, false, false, CFG>(p, sm, tid); break;
    case 1: march_cta<AR, true, false, false, CFG>(p, sm, tid); break;
    case 2: march_cta<AR, false, true, false, CFG>(p, sm, tid); break;
    case 3: march_cta<AR, true, true, false, CFG>(p, sm, tid); break;
    case 4: march_cta<AR, false, false, true, CFG>(p, sm, tid); break;
    case 5: march_cta<AR, true, false, true, CFG>(p, sm, tid); break;
    case 6: march_cta<AR, false, true, true, CFG>(p, sm, tid); break;
    default: march_cta<AR, true, true, true, CFG>(p, sm, tid); break;
    }
}

}  // namespace lsf
