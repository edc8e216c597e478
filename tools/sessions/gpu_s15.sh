#!/bin/bash
# GPU session 15: phiS fetched a step ahead + root / reciprocal root without the library's special-case path
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); f=d.get('fp32_mode') or {}
print('$2 value=%.2f launch_ms=%.3f fp32=%s' % (d['value'], d['roofline']['launch_ms'], f.get('value')))" || tail -3 ${1%.json}.err; }
python -m pytest tests -m gpu -x -q > gpurun_out/s15_tests.txt 2>&1; tail -2 gpurun_out/s15_tests.txt
for n in 1024 512 256; do
timeout 400 python bench.py --grid $n $B > gpurun_out/s15_main_$n.json 2> gpurun_out/s15_main_$n.err; show gpurun_out/s15_main_$n.json "main $n"
done
