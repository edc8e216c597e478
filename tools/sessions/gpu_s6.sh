#!/bin/bash
# GPU session 6: predecessor-flag check folded into the step barrier; acquire-fence experiment; chunk 4
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); f=d.get('fp32_mode') or {}
print('$2 value=%.2f launch_ms=%.3f frac=%.4f fp32=%s' % (d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], f.get('value')))" || tail -3 ${1%.json}.err; }
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -x -q > gpurun_out/s6_tests.txt 2>&1; tail -2 gpurun_out/s6_tests.txt
timeout 400 python bench.py --grid 1024 $B > gpurun_out/s6_default_1024.json 2> gpurun_out/s6_default_1024.err; show gpurun_out/s6_default_1024.json "fold (default) 1024"
timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s6_default_512.json 2> gpurun_out/s6_default_512.err; show gpurun_out/s6_default_512.json "fold (default) 512"
for v in nofold fold_nofence fold_nofence_c4 fold_c4; do
  export LSF_LIB_PATH=$PWD/variants/$v.so
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -q -x > gpurun_out/s6_${v}_parity.txt 2>&1
  echo "$v parity: $(tail -1 gpurun_out/s6_${v}_parity.txt)"
  timeout 400 python bench.py --grid 1024 $B > gpurun_out/s6_${v}_1024.json 2> gpurun_out/s6_${v}_1024.err; show gpurun_out/s6_${v}_1024.json "$v 1024"
  timeout 300 python bench.py --grid 512 $B --no-f32 > gpurun_out/s6_${v}_512.json 2> gpurun_out/s6_${v}_512.err; show gpurun_out/s6_${v}_512.json "$v 512"
done
