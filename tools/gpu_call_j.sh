#!/bin/bash
# sharded fp32 grids on 2 GPUs: parity section of the worker + one weak-scaling bench line
mkdir -p gpurun_out
T=r1s
LSF_MGPU_ONLY=f32 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    tests/mgpu/worker.py > gpurun_out/${T}_worker_f32.txt 2>&1
echo "worker rc=$?"; grep -E "MGPU_OK|Error|error|assert" gpurun_out/${T}_worker_f32.txt | head -8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --f32 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/${T}_bench_f32_n2.err | grep '^{' > gpurun_out/${T}_bench_f32_n2.json
python -c "import json; d=json.load(open('gpurun_out/${T}_bench_f32_n2.json')); print('N=2 f32 value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2) if d['e2e'] else None)" || tail -5 gpurun_out/${T}_bench_f32_n2.err
