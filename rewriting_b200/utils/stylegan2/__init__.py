"""StyleGAN2 (sequential form) on the rewriting_b200 kernels.

`load_seq_stylegan(category, truncation=1.0, **kw)` mirrors the reference entry point
(utils/stylegan2/__init__.py:39-47 there): build a `SeqStyleGAN2` of the right resolution for
`category`, load rosinality-format weights (`{'g_ema': ..., 'latent_avg': ...}`), move it to the
GPU.  Weight files keep the reference's names (`stylegan2_<category>-<hash>.pt`); they are read
from `path=`, from the directory in $REWRITING_B200_WEIGHTS, or — last resort, needs network —
fetched from the reference's server.
"""
import os

import torch

from .models import SeqStyleGAN2, DataBag  # noqa: F401

WEIGHT_URLS = 'http://rewriting.csail.mit.edu/data/models/'

# category -> content hash of the published checkpoint
_HASHES = {
    'bedroom': '6fa55a6e', 'car': '3659b4b6', 'cat': 'd8dc98b2', 'church': 'e8ca9fd0',
    'faces': '2858cc2e', 'horse': '499b5380', 'kitchen': 'b3a526e9', 'places': 'a3b72d71',
}
FILENAMES = {cat: 'stylegan2_%s-%s.pt' % (cat, h) for cat, h in _HASHES.items()}


class _Sizes(dict):
    """output resolution per category (256 unless listed)"""

    def __missing__(self, key):
        return 256


sizes = _Sizes(faces=1024, car=512)


def _checkpoint_source(category, path):
    if path is not None:
        return path, True
    root = os.environ.get('REWRITING_B200_WEIGHTS')
    if root:
        return os.path.join(root, FILENAMES[category]), True
    return WEIGHT_URLS + FILENAMES[category], False


def load_state_dict(category, path=None):
    src, local = _checkpoint_source(category, path)
    if local:
        return torch.load(src, map_location='cpu')
    return torch.hub.load_state_dict_from_url(src, map_location='cpu')


def load_seq_stylegan(category, truncation=1.0, path=None, **kwargs):
    """The nn.Sequential StyleGAN2 for `category`, weights loaded, on the GPU.
    `real_truncation=True` (keyword, not in the reference) makes `truncation` really apply;
    the default reproduces the reference, whose 0-dim `latent_avg` buffer silently disables it."""
    ckpt = load_state_dict(category, path=path)
    net = SeqStyleGAN2(sizes[category], style_dim=512, n_mlp=8, truncation=truncation, **kwargs)
    net.load_state_dict(ckpt['g_ema'], latent_avg=ckpt['latent_avg'])
    return net.cuda()
