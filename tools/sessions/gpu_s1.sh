#!/bin/bash
# GPU session 1 (round 2): full GPU test suite on the new FAST arithmetic, then the occupancy / arithmetic variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s1_tests.txt 2>&1
tail -5 gpurun_out/s1_tests.txt
for v in occ2 occ3 occ3_split occ2_split occ3_dsetp occ3_n2; do
  export LSF_LIB_PATH=$PWD/variants/$v.so
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "cube40_reinit_full_parity or fast_mode_within or one_sweep_vs_oracle" > gpurun_out/s1_${v}_parity.txt 2>&1
  echo "$v parity: $(tail -1 gpurun_out/s1_${v}_parity.txt)"
  timeout 400 python bench.py --grid 1024 --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 > gpurun_out/s1_${v}_1024.json 2> gpurun_out/s1_${v}_1024.err
  timeout 300 python bench.py --grid 512 --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-f32 > gpurun_out/s1_${v}_512.json 2> gpurun_out/s1_${v}_512.err
  python - <<PY
import json
for n in (1024, 512):
    try:
        d = json.loads(open("gpurun_out/s1_${v}_%d.json" % n).read().strip().splitlines()[-1])
        f = d.get("fp32_mode") or {}
        print("${v} grid=%d value=%.2f launch_ms=%.3f frac=%.4f fp32=%s" % (n, d["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"], f.get("value")))
    except Exception as e:
        print("${v} grid=%d FAILED %s" % (n, e))
PY
done
