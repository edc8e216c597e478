#!/bin/bash
# GPU session 3 (round 2): new default build (single-copy ring, split barrier, 3 CTAs/SM with run-time selection), variants, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.txt 2>&1
tail -4 gpurun_out/s3_tests.txt
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); f=d.get('fp32_mode') or {}
print('$2 value=%.2f launch_ms=%.3f frac=%.4f fp32=%s' % (d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], f.get('value')))" || tail -3 ${1%.json}.err; }
timeout 400 python bench.py --grid 1024 $B > gpurun_out/s3_default_1024.json 2> gpurun_out/s3_default_1024.err; show gpurun_out/s3_default_1024.json "default 1024"
timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s3_default_512.json 2> gpurun_out/s3_default_512.err; show gpurun_out/s3_default_512.json "default 512 (auto occ)"
LSF_OCC_RUN=3 timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s3_occ3_512.json 2> gpurun_out/s3_occ3_512.err; show gpurun_out/s3_occ3_512.json "occ_run=3 512"
LSF_OCC_RUN=1 timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s3_occ1_512.json 2> gpurun_out/s3_occ1_512.err; show gpurun_out/s3_occ1_512.json "occ_run=1 512"
LSF_OCC_RUN=2 timeout 400 python bench.py --grid 1024 $B --no-f32 > gpurun_out/s3_occ2_1024.json 2> gpurun_out/s3_occ2_1024.err; show gpurun_out/s3_occ2_1024.json "occ_run=2 1024"
timeout 300 python bench.py --grid 256 $B --no-f32 > gpurun_out/s3_default_256.json 2> gpurun_out/s3_default_256.err; show gpurun_out/s3_default_256.json "default 256"
for v in chunk4 chunk16 dsetp t16x8 nosplit; do
  export LSF_LIB_PATH=$PWD/variants/$v.so
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "fast_mode_within or one_sweep_vs_oracle or march_equals_plane" > gpurun_out/s3_${v}_parity.txt 2>&1
  echo "$v parity: $(tail -1 gpurun_out/s3_${v}_parity.txt)"
  timeout 400 python bench.py --grid 1024 $B > gpurun_out/s3_${v}_1024.json 2> gpurun_out/s3_${v}_1024.err; show gpurun_out/s3_${v}_1024.json "$v 1024"
  timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s3_${v}_512.json 2> gpurun_out/s3_${v}_512.err; show gpurun_out/s3_${v}_512.json "$v 512"
done
unset LSF_LIB_PATH
# ncu: launch list of a short bench run, then one full-set capture of the fp64 and of the fp32 sweep kernel at 1024^3
NB="--grid 1024 --steps 1 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_1024.csv python bench.py $NB > gpurun_out/s3_ncu_launch.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_reinit_march -s 10 -c 1 -f -o gpurun_out/r2a_march_1024 python bench.py $NB --no-f32 > gpurun_out/s3_ncu_full.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_reinit_march_f32 -s 10 -c 1 -f -o gpurun_out/r2a_march_f32_1024 python bench.py $NB > gpurun_out/s3_ncu_full32.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
