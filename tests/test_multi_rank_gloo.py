"""world_size-2 gloo test of the N>1 plumbing bench.py uses: every rank owns a contiguous slice
of the batch (no data-path collective), times are MAX-reduced, outputs gather back in image order.
The CUDA encoder is replaced by the CPU oracle here (test infrastructure) -- what is under test is
the sharding / reduction / ordering logic, which is device independent."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from image_webp_b200 import shard, synth
    b, e = shard.shard_range(n, rank, world)
    digests = []
    pix = 0
    for i in range(b, e):
        img = synth.photo_like(48 + 16 * (i % 3), 32 + 16 * (i % 2), i)
        rc, data, _ = O.encode(img, 75, 4)
        assert rc == 0
        digests.append(hashlib.sha256(data).digest())
        pix += img.shape[0] * img.shape[1]
    t = torch.tensor([0.25 * (rank + 1), float(pix)], dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    if rank == 0:
        ret["max_time"] = float(tmax[0])
        ret["pixels"] = float(tsum[1])
        ret["digests"] = shard.gather_in_order(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_gather_in_order():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from image_webp_b200 import synth
    O.lib()  # build before forking
    n, world = 7, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    ref, pix = [], 0
    for i in range(n):
        img = synth.photo_like(48 + 16 * (i % 3), 32 + 16 * (i % 2), i)
        ref.append(hashlib.sha256(O.encode(img, 75, 4)[1]).digest())
        pix += img.shape[0] * img.shape[1]
    assert ret["digests"] == ref            # image order preserved across ranks
    assert ret["max_time"] == 0.5           # MAX over ranks
    assert ret["pixels"] == float(pix)      # whole-job units
