#!/bin/bash
# GPU session 26: ncu captures of the final sweep kernels of the round (r2c)
mkdir -p gpurun_out
NB="--grid 1024 --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_reinit_march -s 10 -c 1 -f -o gpurun_out/r2c_march_1024 python bench.py $NB --no-f32 > gpurun_out/s26_ncu_full.log 2>&1; tail -1 gpurun_out/s26_ncu_full.log
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_reinit_march_f32 -s 10 -c 1 -f -o gpurun_out/r2c_march_f32_1024 python bench.py $NB > gpurun_out/s26_ncu_full32.log 2>&1; tail -1 gpurun_out/s26_ncu_full32.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches_1024.csv python bench.py --grid 1024 --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32 > gpurun_out/s26_ncu_launch.log 2>&1; tail -1 gpurun_out/s26_ncu_launch.log
