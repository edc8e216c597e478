#!/bin/bash
mkdir -p gpurun_out
LSF_F32_IO_LOG=1 timeout 100 python bench.py --f32 --steps 1 --warmup 0 --no-cpu --grid 1024 > gpurun_out/r1y_f32_e2e.json 2> gpurun_out/r1y_f32_e2e.err
grep "lsf f32" gpurun_out/r1y_f32_e2e.err | tail -8
python -c "import json; d=json.load(open('gpurun_out/r1y_f32_e2e.json')); print('e2e', d['e2e'], 'value', d['value'])"
