"""Streaming statistics on the rewrite path (API of the reference's `utils/runningstats.py`,
hot-path subset).

`RunningSecondMoment` keeps `mom2 = sum_n a_n a_n^T` and `count`; `moment()` is the key
second moment C = E[k k^T] that the rewriter whitens with (ganrewrite.py:83-96).  The
reference accumulates with `mom2.addbmm_(a[:,:,None], a[:,None,:])`, i.e. one rank-1 batched
GEMM per sample row (runningstats.py:1086-1097,1181-1190); here `add` is one tensor-core
col-GEMM over bf16 hi/lo planes (`rw_second_moment_accum`, fp32 accumulate, fixed reduction
order).  The on-disk state (`state_dict`: constructor, count, mom2) is byte-compatible with
the reference's `r2m.npz` caches.

`RunningMean` is provided for the erase path's `unit_rs.npz` (ganrewrite.py:541-552).
Quantile / top-k / IoU statistics of the reference serve the UI search and dissection
tools and are out of scope.
"""
import torch

from .. import ops


class RunningSecondMoment(object):
    """Uncentered second moment of a stream of [N, C] batches."""

    def __init__(self, state=None):
        if state is not None:
            self.set_state_dict(resolve_state_dict(state))
            return
        self.count = 0
        self.mom2 = None

    def add(self, a):
        if len(a.shape) == 1:
            a = a[None, :]
        if self.count == 0:
            self.mom2 = a.new_zeros(a.shape[1], a.shape[1])
        self.count += a.shape[0]
        if a.shape[0] == 0:
            return
        if not a.is_cuda:
            raise RuntimeError('RunningSecondMoment.add needs CUDA data (no CPU fallback); got '
                               + str(a.device))
        if a.shape[1] % 128 != 0:
            raise RuntimeError('RunningSecondMoment.add: channel count %d is not a multiple of '
                               '128 (tensor-core tile)' % a.shape[1])
        ops.second_moment_accum(self.mom2, a.detach())

    def add_planes(self, hi, lo, count):
        """Fast path: accumulate from bf16 hi/lo planes that a producer kernel already wrote
        (zero rows contribute nothing); `count` = number of real sample rows."""
        if self.count == 0 and self.mom2 is None:
            self.mom2 = torch.zeros(hi.shape[1], hi.shape[1], dtype=torch.float32,
                                    device=hi.device)
        self.count += count
        ops.second_moment_accum_planes(self.mom2, hi, lo)

    def cpu_(self):
        self.mom2 = self.mom2.cpu()

    def cuda_(self):
        self.mom2 = self.mom2.cuda()

    def to_(self, device):
        self.mom2 = self.mom2.to(device)

    def moment(self):
        return self.mom2 / self.count

    def state_dict(self):
        return dict(constructor=self.__module__ + '.' + self.__class__.__name__ + '()',
                    count=self.count, mom2=self.mom2.cpu().numpy())

    def set_state_dict(self, dic):
        self.count = dic['count'].item()
        self.mom2 = torch.from_numpy(dic['mom2'])


class RunningMean(object):
    """Streaming per-unit mean of [N, C] batches (Chan-style merge like the reference)."""

    def __init__(self, state=None):
        if state is not None:
            self.set_state_dict(resolve_state_dict(state))
            return
        self.count = 0
        self.batchcount = 0
        self._mean = None

    def add(self, a):
        if len(a.shape) == 1:
            a = a[None, :]
        if len(a.shape) > 2:
            a = a.permute(0, *range(2, a.dim()), 1).reshape(-1, a.shape[1])
        batch_count = a.shape[0]
        batch_mean = a.sum(0) / batch_count
        self.batchcount += 1
        if self._mean is None:
            self.count, self._mean = batch_count, batch_mean
            return
        self.count += batch_count
        self._mean = self._mean + (batch_mean - self._mean) * (batch_count / self.count)

    def size(self):
        return self.count

    def mean(self):
        return self._mean

    def to_(self, device):
        self._mean = self._mean.to(device)

    def cpu_(self):
        self._mean = self._mean.cpu()

    def cuda_(self):
        self._mean = self._mean.cuda()

    def state_dict(self):
        return dict(constructor=self.__module__ + '.' + self.__class__.__name__ + '()',
                    count=self.count, batchcount=self.batchcount,
                    mean=self._mean.cpu().numpy())

    def set_state_dict(self, dic):
        self.count = dic['count'].item()
        self.batchcount = dic['batchcount'].item()
        self._mean = torch.from_numpy(dic['mean'])


def resolve_state_dict(s):
    """Accepts a dict / NpzFile, or the path of an .npz file."""
    import numpy
    if isinstance(s, str):
        return numpy.load(s, allow_pickle=True)
    return s
