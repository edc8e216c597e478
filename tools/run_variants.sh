#!/bin/bash
# tuning helper: bench every liblsf variant in variants/ (built with -DLSF_TB/-DLSF_TC/-DLSF_OCC)
mkdir -p gpurun_out
for so in variants/*.so; do
  export LSF_LIB_PATH=$PWD/$so
  ok=$(timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "exact_mode_bitwise or march_equals_plane" 2>&1 | tail -1)
  for n in 512 1024; do
    line=$(timeout 300 python bench.py --grid $n --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1)
    echo "$so grid=$n parity=[$ok] $(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value=%.2f launch_ms=%.3f frac=%.4f'%(d['value'], d['roofline']['launch_ms'], d['roofline']['frac']))" 2>/dev/null || echo "FAILED: $line")"
  done
done
