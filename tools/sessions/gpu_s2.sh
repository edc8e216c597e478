#!/bin/bash
# GPU session 2 (round 2): tests, sweep-kernel variants (occupancy / ring copy / split barrier), PDL overlap, full bench, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.txt 2>&1
tail -4 gpurun_out/s2_tests.txt
for v in occ2 occ2_split occ2_nodup occ3_nodup occ3_nodup_split occ3_dup; do
  export LSF_LIB_PATH=$PWD/variants/$v.so
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "cube40_reinit_full_parity or fast_mode_within or one_sweep_vs_oracle" > gpurun_out/s2_${v}_parity.txt 2>&1
  echo "$v parity: $(tail -1 gpurun_out/s2_${v}_parity.txt)"
  timeout 400 python bench.py --grid 1024 --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-config3 > gpurun_out/s2_${v}_1024.json 2> gpurun_out/s2_${v}_1024.err
  timeout 300 python bench.py --grid 512 --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-f32 --no-config3 > gpurun_out/s2_${v}_512.json 2> gpurun_out/s2_${v}_512.err
  python - <<PY
import json
for n in (1024, 512):
    try:
        d = json.loads(open("gpurun_out/s2_${v}_%d.json" % n).read().strip().splitlines()[-1])
        f = d.get("fp32_mode") or {}
        print("${v} grid=%d value=%.2f launch_ms=%.3f frac=%.4f fp32=%s" % (n, d["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"], f.get("value")))
    except Exception as e:
        print("${v} grid=%d FAILED %s" % (n, e))
PY
done
# overlapped sweeps, PDL packaging (build flag LSF_EXP_PDL): parity tests, then 512^3 / 1024^3 with and without
export LSF_LIB_PATH=$PWD/variants/occ2_pdl.so
LSF_TEST_OVERLAP=1 LSF_OVERLAP_PDL=1 timeout 600 python -m pytest tests/test_gpu_overlap.py -x -q > gpurun_out/s2_pdl_tests.txt 2>&1
echo "pdl tests: $(tail -1 gpurun_out/s2_pdl_tests.txt)"
for n in 512 1024; do
  LSF_OVERLAP_PDL=1 timeout 300 python bench.py --grid $n --steps 3 --warmup 3 --no-cpu --no-e2e --minmax-iters 0 --no-f32 --no-config3 --overlap > gpurun_out/s2_pdl_$n.json 2> gpurun_out/s2_pdl_$n.err
  python -c "
import json
d=json.loads(open('gpurun_out/s2_pdl_$n.json').read().strip().splitlines()[-1]); print('pdl overlap grid=$n value=%.2f ms_per_step=%.2f' % (d['value'], d['ms_per_step']))" || tail -3 gpurun_out/s2_pdl_$n.err
done
unset LSF_LIB_PATH
# the full default bench line (all companions), then the reference arm at 1 step
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s2_bench_full.json 2> gpurun_out/s2_bench_full.err
tail -c 600 gpurun_out/s2_bench_full.err
python -c "
import json
d=json.loads(open('gpurun_out/s2_bench_full.json').read().strip().splitlines()[-1])
print('FULL value', d['value'], 'e2e', d['e2e'], 'parity', d['parity'], 'strong', d['strong'], 'config3', d['config3'], 'cpu', d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s2_bench_reference.json 2>&1
tail -c 400 gpurun_out/s2_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_smoke.txt 2>&1; tail -2 gpurun_out/s2_smoke.txt
