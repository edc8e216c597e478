"""Host-side logic of the z-slab sharding (include/lsf_b200.h: lsf_slab_range; levelsetfortran_b200.ShardedGrid's
set-up protocol) on CPU: two `gloo` ranks.  No GPU: the compute entry points are not called."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nz, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from levelsetfortran_b200 import _lib, slab_range
        k0, k1 = slab_range(nz, world, rank)
        # every rank all-gathers a fixed-size opaque handle exactly as ShardedGrid.__init__ does
        mine = torch.tensor([rank * 16 + (i % 16) for i in range(_lib.IPC_HANDLE_BYTES)], dtype=torch.uint8)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        blob = b"".join(bytes(t.numpy().tobytes()) for t in every)
        assert len(blob) == world * _lib.IPC_HANDLE_BYTES and blob[rank * _lib.IPC_HANDLE_BYTES] == rank * 16
        # scatter a global field into slabs and gather it back through the ranks
        rng = np.random.default_rng(7)
        full = np.asfortranarray(rng.standard_normal((5, 4, nz + 1)))
        slab = np.asfortranarray(full[:, :, k0:k1])
        assert slab.flags.f_contiguous and slab.tobytes(order="F") == full.tobytes(order="F")[k0 * 160:k1 * 160]   # contiguous range
        rngs = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(rngs, torch.tensor([k0, k1]))
        q.put((rank, [tuple(int(v) for v in r) for r in rngs]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nz", [15, 40, 1023])
def test_slab_partition_two_gloo_ranks(lsf, nz):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nz, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    out = dict(q.get(timeout=10) for _ in range(2))
    assert out[0] == out[1]
    (a0, a1), (b0, b1) = out[0]
    assert a0 == 0 and a1 == b0 and b1 == nz + 1 and a1 - a0 >= 8 and b1 - b0 >= 8


def test_slab_range_properties(lsf):
    from levelsetfortran_b200 import _lib, slab_range
    for nz, world in [(1023, 8), (8191, 8), (100, 3), (63, 8), (2047, 2)]:
        prev = 0
        for r in range(world):
            k0, k1 = slab_range(nz, world, r)
            assert k0 == prev and k1 - k0 >= 8
            prev = k1
        assert prev == nz + 1
    with pytest.raises(_lib.LsfError):
        slab_range(20, 4, 0)        # fewer than 8 planes per rank
    with pytest.raises(_lib.LsfError):
        slab_range(100, 2, 2)       # rank out of range
