"""Data-parallel sharding of the z batch: one process per GPU under torchrun,
`torch.distributed` for the plumbing.

* `sharded_second_moment`: rank r takes batches {j : j % R == r} of the z dataset, runs the
  context model + tensor-core accumulator locally, then ONE all-reduce(sum) of
  mom2[C,C] fp32 (1 MiB for C=512) and of the int64 count.  Every rank ends with the
  identical matrix; rank 0 writes the `r2m.npz` cache in the reference's format.
  (The reference has no collective on this path — it is single-GPU, tally.py:424-443.)
* `shard_range`: contiguous seed/batch ranges for sharded sampling (config 5) — no exchange.

With the `gloo` backend (CPU tests) tensors are staged through host memory for the reduce.
"""
import torch
import torch.distributed as dist

from .utils import pbar, runningstats, tally


def world_size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def shard_range(n_items, r=None, R=None):
    """[lo, hi) of n_items for rank r of R, contiguous, sizes differing by at most one."""
    r = rank() if r is None else r
    R = world_size() if R is None else R
    base, extra = divmod(n_items, R)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def allreduce_moment_(mom2, count):
    """In-place sum of (mom2, count) over ranks; returns the global count."""
    if world_size() == 1:
        return count
    backend = dist.get_backend()
    cnt = torch.tensor([count], dtype=torch.int64,
                       device=mom2.device if backend == 'nccl' else 'cpu')
    if backend == 'nccl':
        dist.all_reduce(mom2, op=dist.ReduceOp.SUM)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    else:
        host = mom2.detach().cpu()
        dist.all_reduce(host, op=dist.ReduceOp.SUM)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        mom2.copy_(host)
    return int(cnt.item())


def sharded_second_moment(compute, dataset, sample_size=None, batch_size=10, cachefile=None,
                          device=None, accumulate=None):
    """Sharded `tally.tally_second_moment`.  `compute(zbatch)` returns [N,C] samples or
    `ops.KeyPlanes`; `accumulate(r2mom, sample)` can be overridden (CPU tests)."""
    from . import ops
    args = dict(sample_size=sample_size)
    cached = tally.load_cached_state(cachefile, args)
    if cached is not None:
        return runningstats.RunningSecondMoment(state=cached)
    loader = tally.batches(dataset, sample_size, batch_size)
    r2mom = runningstats.RunningSecondMoment()
    R, r = world_size(), rank()
    for j, batch in enumerate(pbar(loader)):
        if j % R != r:
            continue
        sample = tally.call_compute(compute, batch)
        if accumulate is not None:
            accumulate(r2mom, sample)
        elif isinstance(sample, ops.KeyPlanes):
            r2mom.add_planes(sample.hi, sample.lo, sample.B * sample.H * sample.W)
        else:
            r2mom.add(sample)
    if r2mom.mom2 is None:      # a rank that received no batch still joins the collective
        probe = tally.call_compute(compute, next(iter(loader)))
        C = probe.C if isinstance(probe, ops.KeyPlanes) else probe.shape[1]
        dev = device if device is not None else (
            probe.hi.device if isinstance(probe, ops.KeyPlanes) else probe.device)
        r2mom.mom2 = torch.zeros(C, C, dtype=torch.float32, device=dev)
    r2mom.count = allreduce_moment_(r2mom.mom2, r2mom.count)
    r2mom.to_('cpu')
    if r == 0:
        tally.save_cached_state(cachefile, r2mom, args)
    return r2mom
