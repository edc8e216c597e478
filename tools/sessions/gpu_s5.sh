#!/bin/bash
# GPU session 5: per-tile timing dump of one sweep (LSF_EXP_TIMING build): where do tiles wait?
mkdir -p gpurun_out
export LSF_LIB_PATH=$PWD/variants/timing.so
LSF_TIMING_DUMP=$PWD/gpurun_out/s5_tiles_1024.txt timeout 300 python bench.py --grid 1024 --steps 1 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32 > gpurun_out/s5_bench_1024.json 2> gpurun_out/s5_bench_1024.err
LSF_TIMING_DUMP=$PWD/gpurun_out/s5_tiles_512.txt timeout 300 python bench.py --grid 512 --steps 1 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32 > gpurun_out/s5_bench_512.json 2> gpurun_out/s5_bench_512.err
LSF_OCC_RUN=2 LSF_TIMING_DUMP=$PWD/gpurun_out/s5_tiles_1024_occ2.txt timeout 300 python bench.py --grid 1024 --steps 1 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32 > gpurun_out/s5_bench_1024_occ2.json 2> gpurun_out/s5_bench_1024_occ2.err
wc -l gpurun_out/s5_tiles_*.txt; tail -2 gpurun_out/s5_bench_1024.err
