#!/bin/bash
# last validation of round 1: node-projection tests (changed upload path), smoke, default bench line
mkdir -p gpurun_out
T=r1w
timeout 200 python -m pytest tests/test_gpu_nodes.py tests/test_gpu_f32.py -q -m gpu > gpurun_out/${T}_tests.txt 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${T}_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.txt
timeout 300 python bench.py > gpurun_out/${T}_bench_1024.json 2> gpurun_out/${T}_bench_1024.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1w_bench_1024.json').read().strip().splitlines()[-1])
print('BENCH value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],4),'fp32',round(d['fp32_mode']['value'],2),'fp32 traffic',d['fp32_mode']['roofline']['traffic'],'nodes ms',d['node_projection'].get('ms'))
PY
