#!/bin/bash
# progress-flag polling: chunk length and pre-read of the flags
mkdir -p gpurun_out
T=r1v
run() {  # name lib extra-args...
  local name=$1 lib=$2; shift 2
  if [ "$lib" != default ]; then export LSF_LIB_PATH=$PWD/variants/$lib.so; else unset LSF_LIB_PATH; fi
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-f32 "$@" 2>gpurun_out/${T}_$name.err | grep '^{' > gpurun_out/${T}_$name.json
  python -c "import json; d=json.load(open('gpurun_out/${T}_$name.json')); print('EXP $name value', round(d['value'],2), 'launch_ms', round(d['roofline']['launch_ms'],3))" 2>/dev/null || { echo "EXP $name FAILED"; tail -3 gpurun_out/${T}_$name.err; }
  unset LSF_LIB_PATH
}
run f64_default default
for v in async chunk16 chunk16async chunk32async; do run f64_$v $v; done
run f32_default default --f32
run f32_chunk16async chunk16async --f32
run f32_chunk32async chunk32async --f32
