#!/bin/bash
# GPU session 34 (2 GPUs): per-sweep timeline of the strong-scaling companion (ONE 1024^3 grid on 2 slabs)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
LSF_SWEEP_LOG=1 timeout 300 $RUN bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-f32 --no-rk3 --no-config3 > gpurun_out/s34_bench.json 2> gpurun_out/s34_timeline.txt
grep -c "lsf sweep" gpurun_out/s34_timeline.txt
python -c "
import json
d=json.loads(open('gpurun_out/s34_bench.json').read().strip().splitlines()[-1]); s=d['strong']
print('weak', d['value'], 'strong', s['value'], s['reinit_ms'])"
