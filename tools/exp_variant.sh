#!/bin/bash
# experiment: run parity + bench with an alternative build of the library (LSF_LIB_PATH).  usage: tools/exp_variant.sh LIB GRID...
LIB=$1; shift
export LSF_LIB_PATH=$PWD/$LIB
T=$(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "reinit" 2>&1 | tail -1)
for G in $@; do
  timeout 300 python bench.py --grid $G --steps 3 --warmup 2 --no-cpu --no-e2e --minmax-iters 0 2>/dev/null | grep '^{' > gpurun_out/var_$(basename $LIB .so)_$G.json
  python -c "import json; d=json.load(open('gpurun_out/var_$(basename $LIB .so)_$G.json')); print('$LIB grid=$G parity=[$T] value', round(d['value'],2), 'launch_ms', round(d['roofline']['launch_ms'],3), 'fp64', round(d['roofline']['fp64_pipe_frac'],3))"
done
