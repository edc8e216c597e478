#!/bin/bash
# final single-GPU state of the round: what the driver runs (GPU tests, smoke, bench with its flags, reference arm), plus the ncu launch list
TAG=${1:-r2_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.txt 2>&1; tail -3 gpurun_out/${TAG}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_1024.json 2> gpurun_out/${TAG}_bench_1024.err
tail -c 300 gpurun_out/${TAG}_bench_1024.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_1024.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'fp64frac', d['roofline']['fp64_pipe_frac'], 'launch_ms', d['roofline']['launch_ms'])
print('e2e', d['e2e']['value'], d['e2e'].get('as_fortran_driver'))
print('parity', d['parity']); print('strong', d['strong']['value'], d['strong']['digest_after_reinit'], d['strong']['digest_after_minmax'])
print('config3', d['config3']['value'], 'fp32', d['fp32_mode']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], 'mm', d['minmax_flow']['value'], 'nodes', d['node_projection'])
print('clocks', d['clocks'], 'launches', d['gpu_launches'])"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>&1; tail -c 300 gpurun_out/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_1024.csv python bench.py --grid 1024 --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-config3 --no-f32 > gpurun_out/${TAG}_ncu_launch.log 2>&1
