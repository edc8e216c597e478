// lsf_march.cu -- GPU launch of the skewed x-marching column-tile sweep (lsf_march.cuh).
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "lsf_internal.cuh"
#include "lsf_march.cuh"

namespace lsf {

#ifndef LSF_TB
#define LSF_TB 16
#endif
#ifndef LSF_TC
#define LSF_TC 16
#endif
#ifndef LSF_OCC
#define LSF_OCC 2      // resident CTAs per SM the kernels are COMPILED for (register budget 65536 / (OCC * THREADS) = 128).  Round 2, session
                       // 11: with ~125 registers the compiler keeps the ring / row addresses and the thread index in registers instead of
                       // rematerialising them every step (steady loop 651 instead of 681 instructions); 2 and 3 resident CTAs deliver the
                       // same rate at 1024^3 and 2 are faster below (512^3: 23.7 vs 22.2 Gcell/s).  3 = the 80-register build of sessions 3-10
#endif
typedef MarchCfg<LSF_TB, LSF_TC> CFG;
#ifndef LSF_OCC_EXACT
#define LSF_OCC_EXACT 2   // ExactArith (IEEE division / sqrt sequences) does not fit the 80-register budget of 3 CTAs per SM
#endif
template <class AR> struct MarchOcc { static constexpr int v = LSF_OCC; };
template <> struct MarchOcc<ExactArith> { static constexpr int v = LSF_OCC_EXACT; };

// MG: z-slab variant (peer stores / peer flags compiled in)
template <class AR, bool FA, bool FB, bool FC, bool MG>
__global__ void __launch_bounds__(CFG::THREADS, MarchOcc<AR>::v)
k_reinit_march(const MarchParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MarchSmem<CFG> &sm = *reinterpret_cast<MarchSmem<CFG> *>(smem_raw);
    march_cta<AR, FA, FB, FC, CFG, MG>(p, sm, threadIdx.x);
}

typedef void (*MarchKernel)(const MarchParams);

// ---- fp32 mode (F32Arith, float ring): single GPU -------------------------------------------------
#ifndef LSF_OCC32
#define LSF_OCC32 3    // resident CTAs per SM the fp32 kernels are compiled for (85 registers).  Session 12, 1024^3 / 512^3 Gcell/s:
                       // 4 CTAs (64 regs) 45.2 / 21.9, 3 CTAs 45.8 / 25.7, 2 CTAs (128 regs) 44.0 / 29.3 -- below 2048 tiles a sweep
                       // cannot feed more than 2 CTAs per SM (the dependence chain between tiles binds), so those run with 2
#endif
typedef MarchCfg<LSF_TB, LSF_TC, LSF_ROWS, float> CFG32;
typedef MarchParamsT<float> MarchParamsF;

template <bool FA, bool FB, bool FC, bool MG>
__global__ void __launch_bounds__(CFG32::THREADS, LSF_OCC32)
k_reinit_march_f32(const MarchParamsF p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MarchSmem<CFG32> &sm = *reinterpret_cast<MarchSmem<CFG32> *>(smem_raw);
    march_cta<F32Arith, FA, FB, FC, CFG32, MG>(p, sm, threadIdx.x);
}

typedef void (*MarchKernelF)(const MarchParamsF);
template <bool MG>
static MarchKernelF march_kernel_f32_mg(int fa, int fb, int fc)
{
    switch ((fa ? 1 : 0) | (fb ? 2 : 0) | (fc ? 4 : 0)) {
    case 0: return k_reinit_march_f32<false, false, false, MG>;
    case 1: return k_reinit_march_f32<true, false, false, MG>;
    case 2: return k_reinit_march_f32<false, true, false, MG>;
    case 3: return k_reinit_march_f32<true, true, false, MG>;
    case 4: return k_reinit_march_f32<false, false, true, MG>;
    case 5: return k_reinit_march_f32<true, false, true, MG>;
    case 6: return k_reinit_march_f32<false, true, true, MG>;
    default: return k_reinit_march_f32<true, true, true, MG>;
    }
}
static MarchKernelF march_kernel_f32(int fa, int fb, int fc, bool mg = false)
{
    return mg ? march_kernel_f32_mg<true>(fa, fb, fc) : march_kernel_f32_mg<false>(fa, fb, fc);
}

template <class AR, bool MG>
static MarchKernel march_kernel_mg(int fa, int fb, int fc)
{
    switch ((fa ? 1 : 0) | (fb ? 2 : 0) | (fc ? 4 : 0)) {
    case 0: return k_reinit_march<AR, false, false, false, MG>;
    case 1: return k_reinit_march<AR, true, false, false, MG>;
    case 2: return k_reinit_march<AR, false, true, false, MG>;
    case 3: return k_reinit_march<AR, true, true, false, MG>;
    case 4: return k_reinit_march<AR, false, false, true, MG>;
    case 5: return k_reinit_march<AR, true, false, true, MG>;
    case 6: return k_reinit_march<AR, false, true, true, MG>;
    default: return k_reinit_march<AR, true, true, true, MG>;
    }
}

template <class AR>
static MarchKernel march_kernel(int fa, int fb, int fc, bool mg = false)
{
    return mg ? march_kernel_mg<AR, true>(fa, fb, fc) : march_kernel_mg<AR, false>(fa, fb, fc);
}

struct MarchHost {
    int *d_order = nullptr;
    int ntb = 0, ntc = 0, m = 0;
};
static MarchHost MH;   // order table of the most recent grid shape

const int *march_order() { return MH.d_order; }

int march_ntiles(const Grid *g)
{
    return ((g->dm.ny - 1 + CFG::TB - 1) / CFG::TB) * ((g->sg.kupd_hi - g->sg.kupd_lo + 1 + CFG::TC - 1) / CFG::TC);
}

// Tilt of the ticket fronts (march_fill_order).  One GPU: anti-diagonals.  z-slabs: the downstream rank can
// only start once this rank's sweep has crossed the slab in c, so the fronts are tilted as far as the
// LAG between successive tiles of a column allows without starving the resident CTAs.
// Dynamic tile scheduler (march_pick, lsf_march.cuh) unless LSF_STATIC_TICKETS=1: the per-column counters are zeroed before the sweep.
static int *march_colnext_for_sweep(Grid *g, int ntb)
{
    // Opt-in (LSF_STATIC_TICKETS=0).  Measured, round 2: one GPU 1024^3 31.2 static / 30.2 dynamic, 512^3 22.1 / 20.7 (fp32 512^3:
    // 17.9 / 22.1); two z-slabs 53.5 (static, tilt 4) / 53.0 (dynamic): lowest-column priority starts every tile right at the heels
    // of its predecessors, which costs what the steep static order costs, and the lag at the END of a sweep is a full column chain
    // either way.  Correct (emulation + 72 GPU tests + the 2-GPU parity worker ran with it), not faster: the static order stays.
    static const int env = getenv("LSF_STATIC_TICKETS") ? atoi(getenv("LSF_STATIC_TICKETS")) : 1;
    if (env != 0) return nullptr;
    cudaMemsetAsync(g->march_colnext, 0, sizeof(int) * (size_t)ntb, G.stream);
    return g->march_colnext;
}

// start slack of a tile over its predecessors, in steps (MarchParamsT::slack): LSF_SLACK overrides
static int march_slack(const Grid *g)
{
    static const int env = getenv("LSF_SLACK") ? atoi(getenv("LSF_SLACK")) : -1;
    if (env >= 0) return env;
    (void)g;
    return 0;      // measured (round 2, session 4): 8 .. 64 steps of slack are monotonically slower at 1024^3 and 512^3
}

static int march_order_tilt(const Grid *g)
{
    if (const char *e = getenv("LSF_ORDER_TILT")) { const int m = atoi(e); if (m >= 1 && m <= 64) return m; }
    // round 2 (session 7, one GPU, 1024^3): tilt 1 / 2 / 4 / 8 cost 0 / 1 / 3.3 / 7.6 % at 2 CTAs/SM and 0 / 1 / 5 / 16 % at 3 CTAs/SM:
    // the tilt caps the number of tiles that can run concurrently at (tile length / lag) x (ntc / m) = 44 x 64/m.  With 2 CTAs/SM
    // (296 CTAs) the first ~1060 steps hand out the tickets of the J = 0 column chain fast enough for m = 4.
    // z-slabs: the downstream rank lags by (ntc-1) fronts = (ntc-1) ntb/m tickets (measured L = 7.5 ms at m = 4, 4.8 ms at m = 8,
    // 1024^3 per rank) and every k flip costs (P-1) L per sweep cycle: 8 T(m) + 2 (P-1) L(m) is smallest at m = 4 for P <= 3, m = 8 beyond
    if (!sharded(g)) return 1;
    return g->sg.nranks <= 3 ? 4 : 8;
}

template <class T>
static void march_orient_grid(MarchParamsT<T> &p, const Grid *g, int raster)
{
    const SlabGeom &sg = g->sg;
    march_orient<CFG>(p, g->dm.nx, g->dm.ny, g->dm.nz, g->dm.sx, g->dm.sxy, raster, sg.kupd_lo, sg.kupd_hi, sg.kbase, sg.NZ);
}

int march_prepare(Grid *g)
{
    MarchParams p;
    march_orient_grid(p, g, 1);
    if (p.ntiles > 65536) return set_error(LSF_ERR_ARG, "march: more than 65536 column tiles");
    if (sharded(g) && p.ntb > SLAB_MAX_NTB) return set_error(LSF_ERR_ARG, "march: more than %d tile columns on a sharded grid", SLAB_MAX_NTB);
    const int tilt = march_order_tilt(g);
    if (!g->march_ticket) LSF_CUDA(cudaMalloc(&g->march_ticket, sizeof(unsigned)));
    if (g->march_colnext_cap < p.ntb) {
        cudaFree(g->march_colnext);
        g->march_colnext = nullptr; g->march_colnext_cap = 0;
        LSF_CUDA(cudaMalloc(&g->march_colnext, sizeof(int) * (size_t)p.ntb));
        g->march_colnext_cap = p.ntb;
    }
    if (g->march_tiles_cap < p.ntiles) {
        cudaFree(g->march_progress);
        g->march_progress = nullptr;
        LSF_CUDA(cudaMalloc(&g->march_progress, sizeof(long long) * (size_t)p.ntiles));
        LSF_CUDA(cudaMemsetAsync(g->march_progress, 0, sizeof(long long) * (size_t)p.ntiles, G.stream));
        g->march_tiles_cap = p.ntiles;
        g->march_epoch = 0;
    }
    if (MH.ntb != p.ntb || MH.ntc != p.ntc || MH.m != tilt || !MH.d_order) {
        cudaFree(MH.d_order);
        MH.d_order = nullptr;
        std::vector<int> order(p.ntiles);
        march_fill_order(p.ntb, p.ntc, order.data(), tilt);
        LSF_CUDA(cudaMalloc(&MH.d_order, sizeof(int) * (size_t)p.ntiles));
        LSF_CUDA(cudaMemcpy(MH.d_order, order.data(), sizeof(int) * (size_t)p.ntiles, cudaMemcpyHostToDevice));
        MH.ntb = p.ntb; MH.ntc = p.ntc; MH.m = tilt;
    }
    static bool attr_done = false;
    if (!attr_done) {
        // Shared-memory carve-out: left to the driver unless LSF_CARVEOUT_MAX is defined.  Measured (round 2, session 1):
        // forcing the maximum carve-out halves the sweep rate of BOTH the fp64 and the fp32 kernel -- the L1 that is left is too
        // small to hold the in-flight global loads of 16-24 warps.
#if defined(LSF_CARVEOUT_MAX)
        const int carve = (int)cudaSharedmemCarveoutMaxShared;
#else
        const int carve = (int)cudaSharedmemCarveoutDefault;
#endif
        for (int o = 0; o < 16; ++o) {
            LSF_CUDA(cudaFuncSetAttribute(march_kernel<FastArith>(o & 1, o & 2, o & 4, o & 8), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem<CFG>)));
            LSF_CUDA(cudaFuncSetAttribute(march_kernel<ExactArith>(o & 1, o & 2, o & 4, o & 8), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem<CFG>)));
            LSF_CUDA(cudaFuncSetAttribute(march_kernel<FastArith>(o & 1, o & 2, o & 4, o & 8), cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            LSF_CUDA(cudaFuncSetAttribute(march_kernel<ExactArith>(o & 1, o & 2, o & 4, o & 8), cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        }
        for (int o = 0; o < 16; ++o) {
            LSF_CUDA(cudaFuncSetAttribute(march_kernel_f32(o & 1, o & 2, o & 4, o & 8), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem<CFG32>)));
            LSF_CUDA(cudaFuncSetAttribute(march_kernel_f32(o & 1, o & 2, o & 4, o & 8), cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        }
        attr_done = true;
    }
    return LSF_OK;
}

// z-slabs: the streaming-halo / write-through / completion-flag fields of the sweep parameters (lsf_march.cuh).
// phi: this rank's field inside its shared allocation (fp64 or fp32).
template <class T>
static void march_fill_slab(MarchParamsT<T> &p, Grid *g, T *phi)
{
    // the Gauss-Seidel pipeline along k (lsf_slab.cuh): upstream = the rank owning lower oriented c
    const SlabGeom &sg = g->sg;
    const int up = p.fc ? sg.rank + 1 : sg.rank - 1, down = p.fc ? sg.rank - 1 : sg.rank + 1;
    const int side_up = p.fc ? 1 : 0, side_down = p.fc ? 0 : 1;     // which of THIS rank's sides that neighbour is on
    if (up >= 0 && up < sg.nranks) {
        SlabGeom ug;
        slab_geom(sg.NZ, sg.nranks, up, ug);
        p.in_progress = g->sync->in_progress;
        p.push_up_delta = (peer_ptr(g, up, phi) + (long long)(sg.kbase - ug.kbase) * g->dm.sxy) - phi;
        p.edge_pub[0] = peer_ptr(g, up, g->sync)->edge_done[1 - side_up];        // address computation only
    }
    if (down >= 0 && down < sg.nranks) {
        SlabGeom dg;
        slab_geom(sg.NZ, sg.nranks, down, dg);
        p.push_delta = (peer_ptr(g, down, phi) + (long long)(sg.kbase - dg.kbase) * g->dm.sxy) - phi;
        p.push_progress = peer_ptr(g, down, g->sync)->in_progress;
        p.edge_pub[1] = peer_ptr(g, down, g->sync)->edge_done[1 - side_down];
        if (g->prev_sweep_valid) {
            p.edge_wait = g->sync->edge_done[side_down];
            p.edge_need = g->prev_sweep_epoch;
            p.edge_prev_fb = g->prev_sweep_fb;
        }
    }
    // the first sweep after a bulk exchange: both neighbours' planes must have landed in the ghost planes
    p.halo_seq = g->sync->halo_seq;
    if (sg.rank > 0) p.halo_need[0] = g->phase;
    if (sg.rank < sg.nranks - 1) p.halo_need[1] = g->phase;
    g->prev_sweep_valid = true; g->prev_sweep_epoch = p.epoch; g->prev_sweep_fb = p.fb;
}

// fp32 grid (lsf_f32.cu): same schedule, float phi / phiS and a float slot ring
void launch_reinit_sweep_march_f32(Grid *g, int raster, const CellConst &cc)
{
    MarchParamsF p;
    memset(&p, 0, sizeof(p));
    march_orient_grid(p, g, raster);
    p.phi = g->phi_f; p.phiS = g->phiS_f; p.slack = march_slack(g);
    p.cc.dx = (float)cc.dx; p.cc.inv_dx = (float)cc.inv_dx; p.cc.k12 = (float)cc.k12; p.cc.dx2 = (float)cc.dx2; p.cc.h = (float)cc.h;
    p.partial = g->partial; p.ticket = g->march_ticket; p.order = MH.d_order;
    p.progress = g->march_progress; p.epoch = ++g->march_epoch; p.ctrl = g->ctrl;
    if (sharded(g)) march_fill_slab(p, g, g->phi_f);
    cudaMemsetAsync(g->march_ticket, 0, sizeof(unsigned), G.stream);
    p.col_next = march_colnext_for_sweep(g, p.ntb);
    static const int occ_env32 = getenv("LSF_OCC32_RUN") ? atoi(getenv("LSF_OCC32_RUN")) : 0;   // experiments: fewer resident CTAs
    const int occ = occ_env32 > 0 ? occ_env32 : ((sharded(g) || p.ntiles < 2048) ? 2 : LSF_OCC32);   // z-slabs: see launch_reinit_sweep_march
    const int ncta = p.ntiles < occ * G.num_sms ? p.ntiles : occ * G.num_sms;
    march_kernel_f32(p.fa, p.fb, p.fc, sharded(g))<<<ncta, CFG32::THREADS, sizeof(MarchSmem<CFG32>), G.stream>>>(p);
    G.n_launch++;
}

// ---- overlapped sweeps (march_multi_cta, lsf_march.cuh): one launch = OV_BATCH consecutive sweeps -----------------
struct MultiParams {
    MarchParams p[OV_BATCH];
};

template <class AR>
__global__ void __launch_bounds__(CFG::THREADS, MarchOcc<AR>::v)
k_reinit_march_multi(const __grid_constant__ MultiParams mp, int nsweeps, unsigned *ticket)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MarchSmem<CFG> &sm = *reinterpret_cast<MarchSmem<CFG> *>(smem_raw);
    march_multi_cta<AR, CFG>(mp.p, nsweeps, ticket, sm, threadIdx.x);
}

#if defined(LSF_EXP_PDL)
// Second packaging of the same schedule (LSF_OVERLAP_PDL=1): ONE KERNEL PER SWEEP as in the production path -- compile-
// time orientation, by-value parameters -- with the overlapped-sweeps hooks, chained by programmatic dependent launch:
// every CTA executes griddepcontrol.launch_dependents as soon as it is resident, so the next sweep's kernel is admitted
// when all CTAs of this one hold their SM slots and its CTAs move in as these run out of tickets.  No
// griddepcontrol.wait: the tile flags carry the data dependences.  Compiled only with -DLSF_EXP_PDL (tools/build_variant.sh) until it has
// passed the GPU parity tests: a wrong ordering assumption would corrupt data silently.
template <class AR, bool FA, bool FB, bool FC>
__global__ void __launch_bounds__(CFG::THREADS, MarchOcc<AR>::v)
k_reinit_march_ov(const MarchParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MarchSmem<CFG> &sm = *reinterpret_cast<MarchSmem<CFG> *>(smem_raw);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    march_cta<AR, FA, FB, FC, CFG, false, true>(p, sm, threadIdx.x);
    // completion must be transitive along the chain (the batch's loop-control kernels follow the LAST sweep's kernel in the
    // stream and read every sweep's RMS partials): do not retire before the previous sweep's grid has
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <class AR>
static MarchKernel march_kernel_ov(int fa, int fb, int fc)
{
    switch ((fa ? 1 : 0) | (fb ? 2 : 0) | (fc ? 4 : 0)) {
    case 0: return k_reinit_march_ov<AR, false, false, false>;
    case 1: return k_reinit_march_ov<AR, true, false, false>;
    case 2: return k_reinit_march_ov<AR, false, true, false>;
    case 3: return k_reinit_march_ov<AR, true, true, false>;
    case 4: return k_reinit_march_ov<AR, false, false, true>;
    case 5: return k_reinit_march_ov<AR, true, false, true>;
    case 6: return k_reinit_march_ov<AR, false, true, true>;
    default: return k_reinit_march_ov<AR, true, true, true>;
    }
}

#endif   // LSF_EXP_PDL

// Sweeps n_first .. n_first + nsweeps - 1 of the reinit loop in one launch, then their RMS / EXIT / NaN tests in order
// (k_finalize per sweep: sweeps after an exit leave the history alone; phi then holds the state after the WHOLE batch and
// the caller rolls back).  The shell array is phiN.  Returns LSF_OK or an allocation error.
int launch_reinit_sweeps_overlapped(Grid *g, int n_first, int nsweeps, const CellConst &cc, double tol)
{
    MarchParams p0;
    march_orient_grid(p0, g, 1);
    const int ntiles = p0.ntiles;
    if (g->ov_tiles_cap < ntiles) {
        cudaFree(g->ov_progress); cudaFree(g->ov_partial);
        g->ov_progress = nullptr; g->ov_partial = nullptr; g->ov_tiles_cap = 0;
        LSF_CUDA(cudaMalloc(&g->ov_progress, sizeof(long long) * (size_t)OV_BATCH * ntiles));
        LSF_CUDA(cudaMemsetAsync(g->ov_progress, 0, sizeof(long long) * (size_t)OV_BATCH * ntiles, G.stream));
        LSF_CUDA(cudaMalloc(&g->ov_partial, sizeof(double) * (size_t)OV_BATCH * 2 * ntiles));
        g->ov_tiles_cap = ntiles;
    }
    static bool attr_done = false;
    if (!attr_done) {
        LSF_CUDA(cudaFuncSetAttribute(k_reinit_march_multi<FastArith>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem<CFG>)));
        LSF_CUDA(cudaFuncSetAttribute(k_reinit_march_multi<ExactArith>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem<CFG>)));
        attr_done = true;
    }
    MultiParams MP;
    memset(&MP, 0, sizeof(MP));
    MarchParams *P = MP.p;
    const long long epoch = ++g->march_epoch;
    for (int s = 0; s < nsweeps; ++s) {
        MarchParams &p = P[s];
        memset(&p, 0, sizeof(p));
        march_orient_grid(p, g, (n_first + s) % 8 + 1);                  // subs.f90:740,855
        p.phi = g->phi; p.phiS = g->phiS; p.cc = cc;
        p.partial = g->ov_partial + (size_t)s * 2 * ntiles;
        p.partial_bc = p.partial + ntiles;
        p.ticket = g->march_ticket; p.order = MH.d_order;
        p.progress = g->ov_progress + (size_t)s * ntiles; p.epoch = epoch; p.ctrl = g->ctrl;
        p.shell_rd_delta = (s & 1) ? g->phiN - g->phi : 0;
        p.shell_wr_delta = ((s + 1) & 1) ? g->phiN - g->phi : 0;
        p.fold_bc = 1;
        if (s > 0) {
            p.prev_progress = g->ov_progress + (size_t)(s - 1) * ntiles;
            p.prev_fin = (epoch << 32) + M_BIAS + M_FIN;
            p.prev_fb = P[s - 1].fb; p.prev_fc = P[s - 1].fc;
        }
    }
    const int occ_ov = (G.arith_run == LSF_ARITH_EXACT) ? MarchOcc<ExactArith>::v : MarchOcc<FastArith>::v;
    const int ncta = ntiles < occ_ov * G.num_sms ? ntiles : occ_ov * G.num_sms;
#if defined(LSF_EXP_PDL)
    static const bool pdl = getenv("LSF_OVERLAP_PDL") != nullptr && atoi(getenv("LSF_OVERLAP_PDL")) != 0;
#else
    const bool pdl = false;
#endif
    if (pdl) {
#if defined(LSF_EXP_PDL)
        // one ticket counter per sweep slot (a memset between two kernels would break their adjacency in the stream)
        static unsigned *tickets = nullptr;
        if (!tickets) LSF_CUDA(cudaMalloc(&tickets, sizeof(unsigned) * OV_BATCH));
        LSF_CUDA(cudaMemsetAsync(tickets, 0, sizeof(unsigned) * OV_BATCH, G.stream));
        static bool ov_attr_done = false;
        if (!ov_attr_done) {
            for (int o = 0; o < 8; ++o) {
                LSF_CUDA(cudaFuncSetAttribute(march_kernel_ov<FastArith>(o & 1, o & 2, o & 4), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem<CFG>)));
                LSF_CUDA(cudaFuncSetAttribute(march_kernel_ov<ExactArith>(o & 1, o & 2, o & 4), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem<CFG>)));
            }
            ov_attr_done = true;
        }
        for (int s = 0; s < nsweeps; ++s) {
            P[s].ticket = tickets + s;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(CFG::THREADS);
            cfg.dynamicSmemBytes = sizeof(MarchSmem<CFG>); cfg.stream = G.stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = s > 0 ? 1 : 0;            // the first sweep of a batch is an ordinary launch
            MarchKernel kern = (G.arith_run == LSF_ARITH_EXACT) ? march_kernel_ov<ExactArith>(P[s].fa, P[s].fb, P[s].fc)
                                                            : march_kernel_ov<FastArith>(P[s].fa, P[s].fb, P[s].fc);
            LSF_CUDA(cudaLaunchKernelEx(&cfg, kern, P[s]));
            G.n_launch++;
        }
#endif
    } else {
    cudaMemsetAsync(g->march_ticket, 0, sizeof(unsigned), G.stream);
    if (G.arith_run == LSF_ARITH_EXACT)
        k_reinit_march_multi<ExactArith><<<ncta, CFG::THREADS, sizeof(MarchSmem<CFG>), G.stream>>>(MP, nsweeps, g->march_ticket);
    else
        k_reinit_march_multi<FastArith><<<ncta, CFG::THREADS, sizeof(MarchSmem<CFG>), G.stream>>>(MP, nsweeps, g->march_ticket);
    G.n_launch++;
    }
    if (nsweeps & 1) launch_copy_boundary(g, g->phi, g->phiN);            // the last boundary block went to the shell array
    for (int s = 0; s < nsweeps; ++s)
        launch_finalize(g, 2 * ntiles, 0, tol, g->ov_partial + (size_t)s * 2 * ntiles);   // subs.f90:914-926, in sweep order
    g->prev_sweep_valid = false;
    return LSF_OK;
}

void launch_reinit_sweep_march(Grid *g, int raster, const CellConst &cc)
{
    MarchParams p;
    march_orient_grid(p, g, raster);
    p.phi = g->phi; p.phiS = g->phiS; p.cc = cc; p.slack = march_slack(g);
    p.partial = g->partial; p.ticket = g->march_ticket; p.order = MH.d_order;
    p.progress = g->march_progress; p.epoch = ++g->march_epoch; p.ctrl = g->ctrl;
    p.dbg = nullptr;
    p.in_progress = nullptr; p.push_delta = 0; p.push_progress = nullptr; p.halo_seq = nullptr;
    p.halo_need[0] = p.halo_need[1] = 0;
    p.push_up_delta = 0; p.edge_pub[0] = p.edge_pub[1] = nullptr; p.edge_wait = nullptr; p.edge_need = 0; p.edge_prev_fb = 0;
    if (sharded(g)) march_fill_slab(p, g, g->phi);
#if defined(LSF_EXP_TIMING)
    {   // experiment: dump per-tile timing of this sweep to $LSF_TIMING_DUMP after the launch (synchronous)
        static long long *d_dbg = nullptr; static int cap = 0;
        if (cap < p.ntiles) { cudaFree(d_dbg); cudaMalloc(&d_dbg, sizeof(long long) * 6 * p.ntiles); cap = p.ntiles; }
        p.dbg = d_dbg;
    }
#endif
    cudaMemsetAsync(g->march_ticket, 0, sizeof(unsigned), G.stream);
    p.col_next = march_colnext_for_sweep(g, p.ntb);
    // resident CTAs per SM at RUN time: a sweep over few tiles is bound by the dependence chain between tiles, not by throughput,
    // and then fewer CTAs per SM (each progressing faster) finish sooner -- measured at 512^3: 2 CTAs/SM 22.1, 3 CTAs/SM 19.0 Gcell/s
    static const int occ_env = getenv("LSF_OCC_RUN") ? atoi(getenv("LSF_OCC_RUN")) : 0;
    int occ = (G.arith_run == LSF_ARITH_EXACT) ? MarchOcc<ExactArith>::v : MarchOcc<FastArith>::v;
    if (occ_env > 0) occ = occ_env < occ ? occ_env : occ;
    else if ((p.ntiles < 2048 || sharded(g)) && occ > 2) occ = 2;      // z-slabs: the tilted ticket order cannot feed more CTAs
    const int ncta = p.ntiles < occ * G.num_sms ? p.ntiles : occ * G.num_sms;
    MarchKernel kern = (G.arith_run == LSF_ARITH_EXACT) ? march_kernel<ExactArith>(p.fa, p.fb, p.fc, sharded(g))
                                                    : march_kernel<FastArith>(p.fa, p.fb, p.fc, sharded(g));
    kern<<<ncta, CFG::THREADS, sizeof(MarchSmem<CFG>), G.stream>>>(p);
    G.n_launch++;
#if defined(LSF_EXP_TIMING)
    if (const char *path = getenv("LSF_TIMING_DUMP")) {
        cudaStreamSynchronize(G.stream);
        std::vector<long long> h(6 * (size_t)p.ntiles);
        cudaMemcpy(h.data(), p.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        FILE *f = fopen(path, "w");
        if (f) {
            fprintf(f, "# raster %d ntb %d ntc %d : tile J K start_ns end_ns wait_cycles total_cycles smid cta\n", raster, p.ntb, p.ntc);
            long long t0 = h[0];
            for (int q = 0; q < p.ntiles; ++q) if (h[6 * q] < t0) t0 = h[6 * q];
            for (int q = 0; q < p.ntiles; ++q)
                fprintf(f, "%d %d %lld %lld %lld %lld %lld %lld\n", q % p.ntb, q / p.ntb, h[6 * q] - t0, h[6 * q + 1] - t0, h[6 * q + 2], h[6 * q + 3], h[6 * q + 4], h[6 * q + 5]);
            fclose(f);
        }
    }
#endif
}

}  // namespace lsf
