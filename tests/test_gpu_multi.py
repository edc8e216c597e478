"""Multi-GPU parity (z-slab sharding, include/lsf_b200.h): launches tests/mgpu/worker.py with one rank per
GPU.  Needs >= 2 GPUs on the box (`gpurun --gpus 2 -- python -m pytest tests -m gpu`); skipped otherwise."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_path_is_bit_identical_to_single_gpu(lsf, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    port = 29500 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu", "worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "MGPU_OK" in out.stdout, out.stdout[-4000:]
