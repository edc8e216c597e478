#!/bin/bash
# multi-GPU session: usage tools/gpu_multi.sh N TAG   (run under `gpurun --gpus N`)
N=$1; TAG=$2
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/${TAG}_gpus.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
# (1) parity worker: sharded == single GPU, bit for bit (fp64 + fp32, tolerance exits, sign search, band, min/max)
timeout 900 $RUN tests/mgpu/worker.py > gpurun_out/${TAG}_worker.txt 2>&1
grep -a "MGPU_OK\|Error\|error" gpurun_out/${TAG}_worker.txt | tail -3
# (2) the bench line the driver will produce at this N (weak scaling + strong companion with digests + fp32 companion)
timeout 1200 $RUN bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('N=$N value %.2f ms/step %.2f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('strong', d['strong']['value'], d['strong']['reinit_ms'], d['strong']['digest_after_reinit'], d['strong']['digest_after_minmax'])
print('fp32', (d.get('fp32_mode') or {}).get('value'))" || tail -5 gpurun_out/${TAG}_bench.err
# (3) per-rank per-sweep timeline of two steps
LSF_SWEEP_LOG=1 timeout 600 $RUN bench.py --gpus $N --steps 2 --warmup 1 --no-cpu --no-e2e --minmax-iters 0 --no-f32 > gpurun_out/${TAG}_timeline.json 2> gpurun_out/${TAG}_timeline.txt
grep -c "lsf sweep" gpurun_out/${TAG}_timeline.txt
