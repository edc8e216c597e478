#!/bin/bash
# GPU session 30: ncu capture of the K2' stage kernel (k_rk_stage) at 1024^3
mkdir -p gpurun_out
NB="--grid 1024 --steps 1 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_rk_stage -s 4 -c 1 -f -o gpurun_out/r2l_rk_stage_1024 python bench.py $NB > gpurun_out/s30_ncu_rk.log 2>&1; tail -1 gpurun_out/s30_ncu_rk.log | cut -c1-200
