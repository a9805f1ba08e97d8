"""`FixedSubsetSampler` — the one sampler `tally.make_loader` needs (reference:
utils/sampler.py, used at tally.py:640-642)."""
from torch.utils.data.sampler import Sampler


class FixedSubsetSampler(Sampler):
    """Yields the given indices in order, optionally windowed to [start, end)."""

    def __init__(self, samples, start=None, end=None):
        self.samples = list(samples)[start:end]

    def __iter__(self):
        return iter(self.samples)

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, key):
        return self.samples[key]

    def dereference(self, indices):
        return [self.samples[i] for i in indices]
