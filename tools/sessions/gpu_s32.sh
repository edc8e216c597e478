#!/bin/bash
# GPU session 32 (8 GPUs): bench line of the final state incl. the K2' weak-scaling companion
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613"
timeout 600 $RUN bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu > gpurun_out/s32_bench_n8.json 2> gpurun_out/s32_bench_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/s32_bench_n8.json').read().strip().splitlines()[-1])
print('N=8 value %.2f ms/step %.2f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('strong', d['strong']['value'], d['strong']['digest_after_reinit'], d['strong']['digest_after_minmax'])
print('rk3', d.get('rk3_mode')); print('fp32', (d.get('fp32_mode') or {}).get('value'))" || tail -5 gpurun_out/s32_bench_n8.err
