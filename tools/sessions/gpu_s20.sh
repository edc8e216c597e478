#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
B="--gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-f32 --no-config3"
for v in libsqrt main; do
  if [ $v = main ]; then unset LSF_LIB_PATH; else export LSF_LIB_PATH=$PWD/variants/$v.so; fi
  LSF_SWEEP_LOG=1 timeout 300 $RUN bench.py $B > gpurun_out/s20_$v.json 2> gpurun_out/s20_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/s20_$v.json').read().strip().splitlines()[-1])
print('$v N=2 value %.2f ms/step %.2f' % (d['value'], d['ms_per_step']))" || tail -3 gpurun_out/s20_$v.err
  grep "lsf sweep" gpurun_out/s20_$v.err | tail -4
done
