"""CPU, world_size 2 over gloo: the z-batch sharding and the single all-reduce of
(mom2, count) (rewriting_b200/dist.py).  The accumulate step is replaced by a CPU matmul —
the collective plumbing, batch partition and cache writing are what is under test."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _cpu_accumulate(r2mom, sample):
    if r2mom.mom2 is None:
        r2mom.mom2 = torch.zeros(sample.shape[1], sample.shape[1])
    r2mom.mom2 += sample.t() @ sample
    r2mom.count += sample.shape[0]


def _worker(rank, world, port, tmpdir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from rewriting_b200 import dist as rdist
        torch.manual_seed(0)
        data = torch.randn(50, 16)          # 5 batches of 10 -> ranks get 3 and 2 batches
        proj = torch.randn(16, 128)
        compute = lambda z: z @ proj
        cache = os.path.join(tmpdir, 'r2m.npz')
        r2m = rdist.sharded_second_moment(compute, data, batch_size=10, cachefile=cache,
                                          accumulate=_cpu_accumulate)
        full = compute(data)
        ref = full.t() @ full
        assert r2m.count == 50, r2m.count
        assert torch.allclose(r2m.mom2, ref, rtol=1e-5, atol=1e-3)
        lo, hi = rdist.shard_range(103)
        sizes = [rdist.shard_range(103, r, world) for r in range(world)]
        assert sizes[0][0] == 0 and sizes[-1][1] == 103
        assert all(a[1] == b[0] for a, b in zip(sizes, sizes[1:]))
        assert (lo, hi) == sizes[rank]
        dist.barrier()
        if rank == 0:
            dat = np.load(cache, allow_pickle=True)
            assert int(dat['count']) == 50
        torch.save(r2m.mom2, os.path.join(tmpdir, 'mom2_%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_sharded_second_moment_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(str(tmp_path / 'mom2_0.pt'))
    b = torch.load(str(tmp_path / 'mom2_1.pt'))
    assert torch.equal(a, b)                 # every rank ends with the identical matrix
