#!/bin/bash
# GPU session 16: (a) the z-slab regression of session 15 (N=2: 95 ms per sweep) reproduced on ONE GPU through the dynamic scheduler / tilted tickets,
# bisected over the builds of sessions 11-15; (b) barrier variants
mkdir -p gpurun_out
B="--steps 3 --warmup 2 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 value=%.2f launch_ms=%.3f' % (d['value'], d['roofline']['launch_ms']))" || tail -3 ${1%.json}.err; }
run() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --grid 1024 $B > gpurun_out/s16_$name.json 2> gpurun_out/s16_$name.err; show gpurun_out/s16_$name.json "$name"
}
run main_static LSF_X=0
run main_dyn LSF_STATIC_TICKETS=0
run main_tilt4 LSF_ORDER_TILT=4
run main_dyn_tilt4 LSF_STATIC_TICKETS=0 LSF_ORDER_TILT=4
for v in nosteady eps0 occ2 occ2eps1 pf0 nopf; do
  run ${v}_dyn LSF_LIB_PATH=$PWD/variants/$v.so LSF_STATIC_TICKETS=0
done
for v in allarrive testwait both; do
  run ${v} LSF_LIB_PATH=$PWD/variants/$v.so
  env LSF_LIB_PATH=$PWD/variants/$v.so timeout 300 python bench.py --grid 512 $B > gpurun_out/s16_${v}_512.json 2> gpurun_out/s16_${v}_512.err; show gpurun_out/s16_${v}_512.json "$v 512"
done
