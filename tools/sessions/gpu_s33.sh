#!/bin/bash
# GPU session 33: K2' stage kernel with the in-plane neighbours through shared memory (k_rk_stage_tile)
mkdir -p gpurun_out
B="--grid 1024 --steps 1 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32"
for v in rkt4 rkt3 rkt3z32; do
export LSF_LIB_PATH=$PWD/variants/$v.so
T=$(timeout 300 python -m pytest tests/test_gpu_rk.py -x -q 2>&1 | tail -1)
timeout 300 python bench.py $B > gpurun_out/s33_$v.json 2> gpurun_out/s33_$v.err
python -c "
import json
d=json.loads(open('gpurun_out/s33_$v.json').read().strip().splitlines()[-1]); r=d['rk3_mode']
print('$v [$T] rk3 %.2f Gcell-stage/s, %.2f ms per RK step, HBM frac %.3f' % (r['value'], r['ms_per_rk_step'], r['roofline']['frac']))" || tail -3 gpurun_out/s33_$v.err
done
