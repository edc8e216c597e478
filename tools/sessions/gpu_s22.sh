#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
B="--gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
for v in main mglib; do
  if [ $v = main ]; then unset LSF_LIB_PATH; else export LSF_LIB_PATH=$PWD/variants/$v.so; fi
  timeout 300 $RUN bench.py $B > gpurun_out/s22_$v.json 2> gpurun_out/s22_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/s22_$v.json').read().strip().splitlines()[-1])
print('$v N=2 value %.2f ms/step %.2f fp32 %s' % (d['value'], d['ms_per_step'], (d.get('fp32_mode') or {}).get('value')))" || tail -3 gpurun_out/s22_$v.err
done
