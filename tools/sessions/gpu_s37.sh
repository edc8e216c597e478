#!/bin/bash
# GPU session 37 (4 GPUs): ticket tilt 4 instead of 8 at P = 4 with the final z-slab kernels
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613"
B="--gpus 4 --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-f32 --no-rk3 --no-config3"
for m in 4 8; do
  LSF_ORDER_TILT=$m timeout 200 $RUN bench.py $B > gpurun_out/s37_tilt$m.json 2> gpurun_out/s37_tilt$m.err
  python -c "
import json
d=json.loads(open('gpurun_out/s37_tilt$m.json').read().strip().splitlines()[-1])
print('tilt $m N=4 value %.2f ms/step %.2f' % (d['value'], d['ms_per_step']))" || tail -3 gpurun_out/s37_tilt$m.err
done
