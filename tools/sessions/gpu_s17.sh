#!/bin/bash
# GPU session 17 (2 GPUs): bisect of the z-slab slowdown over the builds of sessions 11-15, then the parity worker
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
B="--gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-f32 --no-config3"
for v in main nosteady eps1 occ2eps1 pf0 nopf; do
  if [ $v = main ]; then unset LSF_LIB_PATH; else export LSF_LIB_PATH=$PWD/variants/$v.so; fi
  timeout 300 $RUN bench.py $B > gpurun_out/s17_$v.json 2> gpurun_out/s17_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/s17_$v.json').read().strip().splitlines()[-1])
print('$v N=2 value %.2f ms/step %.2f' % (d['value'], d['ms_per_step']))" || tail -3 gpurun_out/s17_$v.err
done
unset LSF_LIB_PATH
timeout 900 $RUN tests/mgpu/worker.py > gpurun_out/s17_worker.txt 2>&1
grep -a "MGPU_OK\|Error\|error" gpurun_out/s17_worker.txt | tail -3
