#!/bin/bash
# experiment: per-sweep kernel times of every rank (LSF_SWEEP_LOG).  usage: tools/exp_sweeplog.sh N GRID TILT TAG
N=$1; GRID=$2; M=$3; TAG=$4
if [ "$N" = "1" ]; then
  LSF_SWEEP_LOG=1 LSF_ORDER_TILT=$M timeout 300 python bench.py --grid $GRID --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 > gpurun_out/sl_${TAG}.out 2> gpurun_out/sl_${TAG}.err
else
  LSF_SWEEP_LOG=1 LSF_ORDER_TILT=$M timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --grid $GRID --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/sl_${TAG}.out 2> gpurun_out/sl_${TAG}.err
fi
grep '^{' gpurun_out/sl_${TAG}.out | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$TAG value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'launch_ms', round(d['roofline']['launch_ms'],2))"
grep "lsf sweep" gpurun_out/sl_${TAG}.err | tail -$((16*N)) | sort -k4,4 -s | awk '{print $4, $5, $7, $10}' | tr '\n' ';' | sed 's/;dev/\ndev/g'; echo
