#!/bin/bash
# GPU session 8: dynamic tile scheduler (march_pick) vs static tickets, single GPU
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); f=d.get('fp32_mode') or {}
print('$2 value=%.2f launch_ms=%.3f fp32=%s' % (d['value'], d['roofline']['launch_ms'], f.get('value')))" || tail -3 ${1%.json}.err; }
python -m pytest tests -m gpu -x -q > gpurun_out/s8_tests.txt 2>&1; tail -2 gpurun_out/s8_tests.txt
for n in 1024 512 256; do
  timeout 400 python bench.py --grid $n $B > gpurun_out/s8_dyn_$n.json 2> gpurun_out/s8_dyn_$n.err; show gpurun_out/s8_dyn_$n.json "dynamic $n"
  LSF_STATIC_TICKETS=1 timeout 400 python bench.py --grid $n $B > gpurun_out/s8_static_$n.json 2> gpurun_out/s8_static_$n.err; show gpurun_out/s8_static_$n.json "static $n"
done
LSF_OCC_RUN=3 timeout 400 python bench.py --grid 512 $B --no-f32 > gpurun_out/s8_dyn_512_occ3.json 2> gpurun_out/s8_dyn_512_occ3.err; show gpurun_out/s8_dyn_512_occ3.json "dynamic 512 occ3"
LSF_OCC_RUN=2 timeout 400 python bench.py --grid 1024 $B --no-f32 > gpurun_out/s8_dyn_1024_occ2.json 2> gpurun_out/s8_dyn_1024_occ2.err; show gpurun_out/s8_dyn_1024_occ2.json "dynamic 1024 occ2"
