#!/bin/bash
# round-1 session j, GPU call B: re-test (gradPhi replay, |e| on the ALU pipe, fp32 fixes) + where-does-the-time-go experiments
mkdir -p gpurun_out
T=r1k
timeout 900 python -m pytest tests/test_gpu_f32.py tests/test_gpu_parity.py tests/test_gpu_nodes.py -q -m gpu > gpurun_out/${T}_tests.txt 2>&1
tail -4 gpurun_out/${T}_tests.txt
run() {  # name lib extra-args...
  local name=$1 lib=$2; shift 2
  if [ "$lib" != default ]; then export LSF_LIB_PATH=$PWD/variants/$lib.so; else unset LSF_LIB_PATH; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-f32 "$@" 2>gpurun_out/${T}_$name.err | grep '^{' > gpurun_out/${T}_$name.json
  python -c "import json; d=json.load(open('gpurun_out/${T}_$name.json')); print('EXP $name value', round(d['value'],2), 'launch_ms', round(d['roofline']['launch_ms'],3))" 2>/dev/null || { echo "EXP $name FAILED"; tail -3 gpurun_out/${T}_$name.err; }
  unset LSF_LIB_PATH
}
run f64_default default
run f32_default default --f32
LSF_OCC32_RUN=1 run f32_occ1 default --f32
for v in w18 nosync noldg nostg t16x8; do run f32_$v $v --f32; done
for v in nosync noldg nostg t16x8; do run f64_$v $v; done
run f32_512 default --f32 --grid 512
run f32_512_w18 w18 --f32 --grid 512
