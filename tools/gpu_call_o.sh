#!/bin/bash
# full multi-GPU parity worker (fp32 + fp64 sections) at 2 GPUs on the final state of round 1
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 tests/mgpu/worker.py > gpurun_out/r1x_worker_n2.txt 2>&1
echo "worker rc=$?"; grep -E "MGPU_OK|Error|assert" gpurun_out/r1x_worker_n2.txt | head -5
