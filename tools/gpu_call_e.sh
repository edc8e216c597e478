#!/bin/bash
# round-1 session j, GPU call E: software-pipelined global loads (LSF_PREFETCH=1, default) vs the old order (pf0)
mkdir -p gpurun_out
T=r1n
timeout 900 python -m pytest tests/test_gpu_f32.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/${T}_tests.txt 2>&1
tail -4 gpurun_out/${T}_tests.txt
run() {  # name lib extra-args...
  local name=$1 lib=$2; shift 2
  if [ "$lib" != default ]; then export LSF_LIB_PATH=$PWD/variants/$lib.so; else unset LSF_LIB_PATH; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-f32 "$@" 2>gpurun_out/${T}_$name.err | grep '^{' > gpurun_out/${T}_$name.json
  python -c "import json; d=json.load(open('gpurun_out/${T}_$name.json')); print('EXP $name value', round(d['value'],2), 'launch_ms', round(d['roofline']['launch_ms'],3))" 2>/dev/null || { echo "EXP $name FAILED"; tail -3 gpurun_out/${T}_$name.err; }
  unset LSF_LIB_PATH
}
run f32_pf1 default --f32
run f32_pf0 pf0 --f32
run f64_pf1 default
run f64_pf0 pf0
run f32_pf1_512 default --f32 --grid 512
run f32_pf0_512 pf0 --f32 --grid 512
run f64_pf1_512 default --grid 512
run f64_pf0_512 pf0 --grid 512
