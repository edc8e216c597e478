#!/bin/bash
# GPU session 4 (round 2): start slack of the tile schedule (LSF_SLACK, steps) x resident CTAs per SM
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); f=d.get('fp32_mode') or {}
print('$2 value=%.2f launch_ms=%.3f frac=%.4f fp32=%s' % (d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], f.get('value')))" || tail -3 ${1%.json}.err; }
python -m pytest tests/test_gpu_parity.py tests/test_gpu_rk.py -x -q > gpurun_out/s4_tests.txt 2>&1; tail -2 gpurun_out/s4_tests.txt
for S in 0 8 16 24 32 48 64; do
  LSF_SLACK=$S timeout 400 python bench.py --grid 1024 $B > gpurun_out/s4_slack${S}_1024.json 2> gpurun_out/s4_slack${S}_1024.err; show gpurun_out/s4_slack${S}_1024.json "slack=$S 1024"
done
for S in 0 8 16 32; do
  LSF_SLACK=$S timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s4_slack${S}_512.json 2> gpurun_out/s4_slack${S}_512.err; show gpurun_out/s4_slack${S}_512.json "slack=$S 512"
  LSF_SLACK=$S LSF_OCC_RUN=2 timeout 300 python bench.py --grid 1024 $B --no-f32 > gpurun_out/s4_slack${S}_occ2_1024.json 2> gpurun_out/s4_slack${S}_occ2_1024.err; show gpurun_out/s4_slack${S}_occ2_1024.json "slack=$S occ_run=2 1024"
done
LSF_SLACK=16 LSF_OCC_RUN=3 timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s4_slack16_occ3_512.json 2> gpurun_out/s4_slack16_occ3_512.err; show gpurun_out/s4_slack16_occ3_512.json "slack=16 occ_run=3 512"
LSF_SLACK=32 LSF_OCC_RUN=3 timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s4_slack32_occ3_512.json 2> gpurun_out/s4_slack32_occ3_512.err; show gpurun_out/s4_slack32_occ3_512.json "slack=32 occ_run=3 512"
