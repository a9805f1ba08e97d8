"""Deterministic latent samples (API of the reference's `utils/zdataset.py`).

`standard_z_sample(size, depth, seed)` must be bit-identical to the reference
(zdataset.py:37-51): numpy `RandomState(seed).standard_normal(size*depth)` reshaped and cast
to fp32 — row i is the same vector whatever `size` is.  The draw stays on the host (numpy's
Mersenne-Twister stream is the contract); tensors move to the GPU in batches.
"""
import numpy
import torch
from torch.utils.data import TensorDataset


def standard_z_sample(size, depth, seed=1, device=None):
    rng = numpy.random.RandomState(seed)
    z = torch.from_numpy(rng.standard_normal(size * depth).reshape(size, depth)).float()
    return z if device is None else z.to(device)


def standard_y_sample(size, num_classes, seed=1, device=None):
    rng = numpy.random.RandomState(seed)
    y = torch.from_numpy(rng.randint(num_classes, size=size)).long()
    return y if device is None else y.to(device)


def z_sample_for_model(model, size=100, seed=1):
    """z batch shaped for `model`: uses `model.input_shape` if present, else the input width
    of the first Conv/ConvTranspose/Linear layer (conv models get [N,C,1,1])."""
    if hasattr(model, 'input_shape'):
        shape = tuple(model.input_shape[1:])
        return standard_z_sample(size, model.input_shape[1], seed=seed).view((size,) + shape)
    kinds = (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)
    first = next(m for m in model.modules() if isinstance(m, kinds))
    if isinstance(first, torch.nn.Linear):
        return standard_z_sample(size, first.in_features, seed=seed)
    return standard_z_sample(size, first.in_channels, seed=seed)[:, :, None, None]


def z_dataset_for_model(model, size=100, seed=1, indices=None):
    if indices is None:
        return TensorDataset(z_sample_for_model(model, size, seed))
    indices = torch.as_tensor(indices, dtype=torch.int64, device='cpu')
    zs = z_sample_for_model(model, indices.max().item() + 1, seed)
    return TensorDataset(zs[indices])
