#!/bin/bash
# round-1 session j, GPU call D (4 GPUs): sharded parity at 2 and 4 ranks + weak-scaling bench lines N = 2, 4
mkdir -p gpurun_out
T=r1m
nvidia-smi -L > gpurun_out/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/${T}_tests_multi.txt 2>&1
tail -4 gpurun_out/${T}_tests_multi.txt
for N in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --steps 3 --warmup 3 --no-cpu 2>gpurun_out/${T}_bench_n$N.err | grep '^{' > gpurun_out/${T}_bench_n$N.json
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_n$N.json')); print('N=$N value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2) if d['e2e'] else None, 'mm', d['minmax_flow']['ms_per_iteration'] if d.get('minmax_flow') else None)" || tail -5 gpurun_out/${T}_bench_n$N.err
done
