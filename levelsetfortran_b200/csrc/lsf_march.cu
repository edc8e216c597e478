// lsf_march.cu -- GPU launch of the skewed x-marching column-tile sweep (lsf_march.cuh).
#include <vector>

#include "lsf_internal.cuh"
#include "lsf_march.cuh"

namespace lsf {

template <class AR>
__global__ void __launch_bounds__(M_THREADS, 2)
k_reinit_march(const MarchParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MarchSmem &sm = *reinterpret_cast<MarchSmem *>(smem_raw);
    march_cta<AR>(p, sm, threadIdx.x);
}

struct MarchHost {
    int *d_order = nullptr;
    int ntb = 0, ntc = 0;
};
static MarchHost MH;   // order table of the most recent grid shape

int march_prepare(Grid *g)
{
    MarchParams p;
    march_orient(p, g->dm.nx, g->dm.ny, g->dm.nz, g->dm.sx, g->dm.sxy, 1);
    if (p.ntiles > 65536) return set_error(LSF_ERR_ARG, "march: more than 65536 column tiles");
    if (!g->march_ticket) LSF_CUDA(cudaMalloc(&g->march_ticket, sizeof(unsigned)));
    if (g->march_tiles_cap < p.ntiles) {
        cudaFree(g->march_progress);
        g->march_progress = nullptr;
        LSF_CUDA(cudaMalloc(&g->march_progress, sizeof(long long) * (size_t)p.ntiles));
        LSF_CUDA(cudaMemsetAsync(g->march_progress, 0, sizeof(long long) * (size_t)p.ntiles, G.stream));
        g->march_tiles_cap = p.ntiles;
        g->march_epoch = 0;
    }
    if (MH.ntb != p.ntb || MH.ntc != p.ntc || !MH.d_order) {
        cudaFree(MH.d_order);
        MH.d_order = nullptr;
        std::vector<int> order(p.ntiles);
        march_fill_order(p.ntb, p.ntc, order.data());
        LSF_CUDA(cudaMalloc(&MH.d_order, sizeof(int) * (size_t)p.ntiles));
        LSF_CUDA(cudaMemcpy(MH.d_order, order.data(), sizeof(int) * (size_t)p.ntiles, cudaMemcpyHostToDevice));
        MH.ntb = p.ntb; MH.ntc = p.ntc;
    }
    static bool attr_done = false;
    if (!attr_done) {
        LSF_CUDA(cudaFuncSetAttribute(k_reinit_march<FastArith>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem)));
        LSF_CUDA(cudaFuncSetAttribute(k_reinit_march<ExactArith>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem)));
        attr_done = true;
    }
    return LSF_OK;
}

void launch_reinit_sweep_march(Grid *g, int raster, const CellConst &cc)
{
    MarchParams p;
    march_orient(p, g->dm.nx, g->dm.ny, g->dm.nz, g->dm.sx, g->dm.sxy, raster);
    p.phi = g->phi; p.phiS = g->phiS; p.cc = cc;
    p.partial = g->partial; p.ticket = g->march_ticket; p.order = MH.d_order;
    p.progress = g->march_progress; p.epoch = ++g->march_epoch; p.ctrl = g->ctrl;
    cudaMemsetAsync(g->march_ticket, 0, sizeof(unsigned), G.stream);
    const int ncta = p.ntiles < 2 * G.num_sms ? p.ntiles : 2 * G.num_sms;
    if (G.arith == LSF_ARITH_EXACT) k_reinit_march<ExactArith><<<ncta, M_THREADS, sizeof(MarchSmem), G.stream>>>(p);
    else k_reinit_march<FastArith><<<ncta, M_THREADS, sizeof(MarchSmem), G.stream>>>(p);
    G.n_launch++;
}

}  // namespace lsf
