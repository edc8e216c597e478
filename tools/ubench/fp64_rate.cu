// Micro-benchmark: FP64 pipe throughput on B200 (warp-instructions per clock per SM) for DFMA / DADD / DMUL
// and a mix, as a function of resident warps per SM and ILP.  Timing experiment for DESIGN.md section 4.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            else if (OP == 1) x[i] = x[i] + a;
            else if (OP == 2) x[i] = x[i] * a;
            else { x[i] = fma(x[i], a, b); x[i] = x[i] + b; x[i] = x[i] * a; }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int OP, int ILP>
void run(const char *name, int warps_per_sm, int sms, double clk_ghz)
{
    double *d;
    cudaMalloc(&d, 8);
    int threads = 256, blocks_per_sm = warps_per_sm * 32 / threads;
    if (blocks_per_sm < 1) { blocks_per_sm = 1; threads = warps_per_sm * 32; }
    int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP, ILP><<<sms * blocks_per_sm, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<OP, ILP><<<sms * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ninstr_per_warp = (double)iters * ILP * (OP == 3 ? 3 : 1);
    double warp_instr_per_sm = ninstr_per_warp * warps_per_sm;
    double cycles = ms * 1e-3 * clk_ghz * 1e9;
    printf("%-5s ILP=%d warps/SM=%2d : %.3f ms  -> %.3f warp-instr/clk/SM (= %.1f lanes/clk/SM), %.2f T lane-ops/s chip\n", name, ILP,
           warps_per_sm, ms, warp_instr_per_sm / cycles, 32 * warp_instr_per_sm / cycles,
           warp_instr_per_sm * 32 * sms / (ms * 1e-3) / 1e12);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double clk = p.clockRate * 1e-6;
    printf("%s, %d SMs, clockRate %.3f GHz (cycles computed at this clock)\n", p.name, sms, clk);
    for (int w : {4, 8, 16, 32, 64}) {
        run<0, 1>("DFMA", w, sms, clk);
        run<0, 4>("DFMA", w, sms, clk);
        run<1, 4>("DADD", w, sms, clk);
        run<2, 4>("DMUL", w, sms, clk);
        run<3, 4>("MIX", w, sms, clk);
    }
    return 0;
}
