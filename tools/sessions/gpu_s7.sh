#!/bin/bash
# GPU session 7: cost of the tilted ticket order (what a z-slab rank runs with) x resident CTAs per SM, single GPU
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 value=%.2f launch_ms=%.3f' % (d['value'], d['roofline']['launch_ms']))" || tail -3 ${1%.json}.err; }
for T in 1 2 4 8; do for O in 2 3; do
  LSF_ORDER_TILT=$T LSF_OCC_RUN=$O timeout 300 python bench.py --grid 1024 $B > gpurun_out/s7_tilt${T}_occ${O}.json 2> gpurun_out/s7_tilt${T}_occ${O}.err; show gpurun_out/s7_tilt${T}_occ${O}.json "tilt=$T occ=$O 1024"
done; done
