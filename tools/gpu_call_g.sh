#!/bin/bash
# L1-allocating loads for the data that is safe to cache (phiS, OLD values)
mkdir -p gpurun_out
T=r1p
run() {  # name lib extra-args...
  local name=$1 lib=$2; shift 2
  if [ "$lib" != default ]; then export LSF_LIB_PATH=$PWD/variants/$lib.so; else unset LSF_LIB_PATH; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-f32 "$@" 2>gpurun_out/${T}_$name.err | grep '^{' > gpurun_out/${T}_$name.json
  python -c "import json; d=json.load(open('gpurun_out/${T}_$name.json')); print('EXP $name value', round(d['value'],2), 'launch_ms', round(d['roofline']['launch_ms'],3))" 2>/dev/null || { echo "EXP $name FAILED"; tail -3 gpurun_out/${T}_$name.err; }
  unset LSF_LIB_PATH
}
for v in ca2 ca3 ca7 ca7occ2; do run f32_$v $v --f32; done
for v in ca2 ca3 ca7; do run f64_$v $v; done
LSF_LIB_PATH=$PWD/variants/ca7.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py -q -m gpu -k "reinit or f32" > gpurun_out/${T}_tests_ca7.txt 2>&1
tail -3 gpurun_out/${T}_tests_ca7.txt
