import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from conftest import dist_field
from levelsetfortran_b200 import DeviceGrid, ShardedGrid, _lib
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
_lib.check(_lib.lib().lsf_init(local))
DX = 0.05
shape = (44, 38, 18 * world + 7)
p0 = dist_field(shape, seed=21)
for iters in (1, 2, 3, 8, 9):
    G = DeviceGrid(shape[0] - 1, shape[1] - 1, shape[2] - 1)
    G.upload(p0); rc, n, h = G.minMaxFlow(iters, DX, 1.0e-4, tol=0.0); ref = G.download(); G.close()
    SG = ShardedGrid(shape[0] - 1, shape[1] - 1, shape[2] - 1)
    SG.upload(np.asfortranarray(p0[:, :, SG.k0:SG.k1]))
    rc, n2, h2 = SG.minMaxFlow(iters, DX, 1.0e-4, tol=0.0)
    got = SG.download()
    d = got != ref[:, :, SG.k0:SG.k1]
    planes = np.unique(np.argwhere(d)[:, 2]) + SG.k0
    unchanged = np.array_equal(got, p0[:, :, SG.k0:SG.k1])
    print(f"iters {iters} rank {rank} k0 {SG.k0} k1 {SG.k1}: ndiff {d.sum()} planes {planes.tolist()} hist {h[-1]:.6e} vs {h2[-1]:.6e} n {n} {n2} got==p0 {unchanged}", flush=True)
    SG.close()
    dist.barrier()
dist.destroy_process_group()
