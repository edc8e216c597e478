#!/bin/bash
# experiment: ticket-front tilt (LSF_ORDER_TILT) vs throughput, N GPUs.  usage: tools/exp_tilt.sh N GRID "tilts"
N=$1; GRID=$2; shift 2
for M in $@; do
  if [ "$N" = "1" ]; then
    LSF_ORDER_TILT=$M timeout 300 python bench.py --grid $GRID --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 2>/dev/null | grep '^{' > gpurun_out/tilt_n${N}_g${GRID}_m${M}.json
  else
    LSF_ORDER_TILT=$M timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --grid $GRID --steps 2 --warmup 1 --no-cpu --no-e2e 2>/dev/null | grep '^{' > gpurun_out/tilt_n${N}_g${GRID}_m${M}.json
  fi
  python -c "import sys,json; d=json.load(open('gpurun_out/tilt_n${N}_g${GRID}_m${M}.json')); print('N=$N grid=$GRID tilt=$M value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'launch_ms', round(d['roofline']['launch_ms'],2))"
done
