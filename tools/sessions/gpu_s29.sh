#!/bin/bash
# GPU session 29 (2 GPUs): K2' (Jacobi / TVD-RK3) on z-slabs -- parity worker incl. the new checks, bench line with the rk3 companion
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
timeout 900 $RUN tests/mgpu/worker.py > gpurun_out/s29_worker.txt 2>&1
grep -a "MGPU_OK\|Error\|error" gpurun_out/s29_worker.txt | tail -4
timeout 600 $RUN bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu > gpurun_out/s29_bench.json 2> gpurun_out/s29_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/s29_bench.json').read().strip().splitlines()[-1])
print('N=2 value %.2f ms/step %.2f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('rk3', d.get('rk3_mode'))
print('fp32', (d.get('fp32_mode') or {}).get('value'))" || tail -5 gpurun_out/s29_bench.err
