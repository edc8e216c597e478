#!/bin/bash
# GPU session 11: steady-state step body, eps-max modes, 2 CTAs/SM builds (126 registers), doubled ring -- one GPU
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); f=d.get('fp32_mode') or {}; p=(d.get('parity') or {}).get('full_grid') or {}
print('$2 value=%.2f launch_ms=%.3f fp32=%s fast_vs_exact=%s' % (d['value'], d['roofline']['launch_ms'], f.get('value'), p.get('max_abs_fast_vs_exact')))" || tail -3 ${1%.json}.err; }
python -m pytest tests -m gpu -x -q > gpurun_out/s11_tests.txt 2>&1; tail -2 gpurun_out/s11_tests.txt
timeout 400 python bench.py --grid 1024 $B > gpurun_out/s11_main_1024.json 2> gpurun_out/s11_main_1024.err; show gpurun_out/s11_main_1024.json "main(steady) 1024"
timeout 400 python bench.py --grid 512 $B --no-f32 > gpurun_out/s11_main_512.json 2> gpurun_out/s11_main_512.err; show gpurun_out/s11_main_512.json "main(steady) 512"
for v in nosteady eps1 eps2 occ2 occ2dup occ2eps1 occ2eps2 occ2dupeps2; do
  export LSF_LIB_PATH=$PWD/variants/$v.so
  T=$(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -x -q -k "reinit or march or scale" 2>&1 | tail -1)
  echo "$v parity: $T"
  timeout 400 python bench.py --grid 1024 $B > gpurun_out/s11_${v}_1024.json 2> gpurun_out/s11_${v}_1024.err; show gpurun_out/s11_${v}_1024.json "$v 1024"
  timeout 400 python bench.py --grid 512 $B --no-f32 > gpurun_out/s11_${v}_512.json 2> gpurun_out/s11_${v}_512.err; show gpurun_out/s11_${v}_512.json "$v 512"
done
unset LSF_LIB_PATH
