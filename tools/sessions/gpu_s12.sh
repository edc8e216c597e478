#!/bin/bash
# GPU session 12: new default build (steady body, eps keys, 2 CTAs/SM, fused F32Arith) -- tests, bench, fp32 occupancy variants, ncu
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); f=d.get('fp32_mode') or {}; p=(d.get('parity') or {}).get('full_grid') or {}
print('$2 value=%.2f launch_ms=%.3f fp32=%s fast_vs_exact=%s' % (d['value'], d['roofline']['launch_ms'], f.get('value'), p.get('max_abs_fast_vs_exact')))" || tail -3 ${1%.json}.err; }
python -m pytest tests -m gpu -x -q > gpurun_out/s12_tests.txt 2>&1; tail -2 gpurun_out/s12_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s12_smoke.txt 2>&1; tail -1 gpurun_out/s12_smoke.txt
timeout 400 python bench.py --grid 1024 $B > gpurun_out/s12_main_1024.json 2> gpurun_out/s12_main_1024.err; show gpurun_out/s12_main_1024.json "main 1024"
timeout 400 python bench.py --grid 512 $B > gpurun_out/s12_main_512.json 2> gpurun_out/s12_main_512.err; show gpurun_out/s12_main_512.json "main 512"
for v in f32occ3 f32occ2; do
  export LSF_LIB_PATH=$PWD/variants/$v.so
  T=$(timeout 600 python -m pytest tests/test_gpu_f32.py -x -q 2>&1 | tail -1)
  echo "$v f32 tests: $T"
  timeout 400 python bench.py --grid 1024 $B > gpurun_out/s12_${v}_1024.json 2> gpurun_out/s12_${v}_1024.err; show gpurun_out/s12_${v}_1024.json "$v 1024"
  timeout 400 python bench.py --grid 512 $B > gpurun_out/s12_${v}_512.json 2> gpurun_out/s12_${v}_512.err; show gpurun_out/s12_${v}_512.json "$v 512"
done
unset LSF_LIB_PATH
NB="--grid 1024 --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_reinit_march -s 10 -c 1 -f -o gpurun_out/r2b_march_1024 python bench.py $NB --no-f32 > gpurun_out/s12_ncu_full.log 2>&1; tail -2 gpurun_out/s12_ncu_full.log
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_reinit_march_f32 -s 10 -c 1 -f -o gpurun_out/r2b_march_f32_1024 python bench.py $NB > gpurun_out/s12_ncu_full32.log 2>&1; tail -2 gpurun_out/s12_ncu_full32.log
