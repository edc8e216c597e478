#!/bin/bash
# round-1 session j, GPU call A: new tests first, full suite, bench line (fp64 + fp32 companion), experiments, ncu
mkdir -p gpurun_out
T=r1j
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_f32.py tests/test_gpu_nodes.py -q -m gpu > gpurun_out/${T}_tests_new.txt 2>&1
tail -15 gpurun_out/${T}_tests_new.txt
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_f32.py --deselect tests/test_gpu_nodes.py > gpurun_out/${T}_tests_rest.txt 2>&1
tail -5 gpurun_out/${T}_tests_rest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_1024.json 2> gpurun_out/${T}_bench_1024.err
tail -c 600 gpurun_out/${T}_bench_1024.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r1j_bench_1024.json').read().strip().splitlines()[-1])
    print('BENCH value',d['value'],'e2e',d['e2e']['value'] if d['e2e'] else None,'frac',d['roofline']['frac'],'launch_ms',d['roofline']['launch_ms'])
    print('FP32',json.dumps(d.get('fp32_mode')))
    print('MM',d['minmax_flow']['ms_per_iteration'] if d.get('minmax_flow') else None)
except Exception as e: print('bench parse failed',e)
PY
# A/B: default vs absint variant (fp64 kernel only)
for v in default absint; do
  if [ $v != default ]; then export LSF_LIB_PATH=$PWD/variants/$v.so; else unset LSF_LIB_PATH; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-f32 2>/dev/null | grep '^{' > gpurun_out/${T}_var_$v.json
  python -c "import json; d=json.load(open('gpurun_out/${T}_var_$v.json')); print('VAR $v value', round(d['value'],2), 'launch_ms', round(d['roofline']['launch_ms'],3))"
done
unset LSF_LIB_PATH
# fp32 main line + resident-CTA sensitivity
for occ in 4 3 2; do
  LSF_OCC32_RUN=$occ timeout 300 python bench.py --f32 --steps 3 --warmup 3 --no-cpu $( [ $occ != 4 ] && echo --no-e2e ) 2>/dev/null | grep '^{' > gpurun_out/${T}_f32_occ$occ.json
  python -c "import json; d=json.load(open('gpurun_out/${T}_f32_occ$occ.json')); print('F32 occ $occ value', round(d['value'],2), 'launch_ms', round(d['roofline']['launch_ms'],3), 'e2e', d['e2e'])"
done
timeout 300 python bench.py --f32 --grid 512 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | grep '^{' > gpurun_out/${T}_f32_512.json
# ncu: launch list of a short fp64 bench (with fp32 companion) and full capture of one f32 sweep kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_1024.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --minmax-iters 8 > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reinit_march_f32 -s 9 -c 1 -o gpurun_out/${T}_march_f32_1024 -f \
    python bench.py --f32 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
