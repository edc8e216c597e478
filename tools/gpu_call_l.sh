#!/bin/bash
# ncu --set full of the kernels that have no capture yet: active-list min/max, sign search, boundary block, node projection
mkdir -p gpurun_out
T=r1u
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_mml_iter|k_sign_search_tiled|k_reinit_bc_rms|k_advect_nodes|k_mml_settle' \
    -c 12 -o gpurun_out/${T}_other_kernels_1024 -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-f32 --minmax-iters 3 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${T}_ncu.log; ls -la gpurun_out/${T}_*
