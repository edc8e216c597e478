#!/bin/bash
# GPU session 36: loads carried across the step (LSF_PREFETCH 3 / 15) with a POLLING barrier wait (mbarrier.test_wait) instead of try_wait
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-rk3"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); f=d.get('fp32_mode') or {}
print('$2 value=%.2f launch_ms=%.3f fp32=%s' % (d['value'], d['roofline']['launch_ms'], f.get('value')))" || tail -3 ${1%.json}.err; }
for v in pf15tw pf3tw; do
export LSF_LIB_PATH=$PWD/variants/$v.so
T=$(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py -x -q -k "reinit or march or f32" 2>&1 | tail -1); echo "$v tests: $T"
timeout 400 python bench.py --grid 1024 $B > gpurun_out/s36_${v}_1024.json 2> gpurun_out/s36_${v}_1024.err; show gpurun_out/s36_${v}_1024.json "$v 1024"
done
