#!/bin/bash
# GPU session 18 (2 GPUs): z-slab kernels without the one-step-ahead phiS load
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
timeout 600 $RUN bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/s18_main.json 2> gpurun_out/s18_main.err
python -c "
import json
d=json.loads(open('gpurun_out/s18_main.json').read().strip().splitlines()[-1])
print('main N=2 value %.2f ms/step %.2f' % (d['value'], d['ms_per_step']))
print('strong', d['strong']['value'], d['strong']['reinit_ms'], d['strong']['digest_after_reinit'], d['strong']['digest_after_minmax'])
print('fp32', (d.get('fp32_mode') or {}).get('value'))" || tail -3 gpurun_out/s18_main.err
