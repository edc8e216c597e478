#!/bin/bash
# final validation of the round's state on one GPU: smoke, the whole -m gpu suite, the default bench line, launch list
mkdir -p gpurun_out
T=r1t
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.txt
timeout 600 python bench.py > gpurun_out/${T}_bench_1024.json 2> gpurun_out/${T}_bench_1024.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${T}_bench_1024.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1t_bench_1024.json').read().strip().splitlines()[-1])
print('BENCH value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],4),'launch_ms',round(d['roofline']['launch_ms'],2),'cpu',d['cpu_baseline']['value'])
print('FP32',round(d['fp32_mode']['value'],2), 'MM',d['minmax_flow']['ms_per_iteration'], 'NODES',d['node_projection'])
print('clocks',d['clocks'])
PY
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>&1; tail -c 300 gpurun_out/${T}_bench_reference.json
