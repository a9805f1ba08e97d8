"""StyleGAN2 (sequential form) on the rewriting_b200 kernels.

Mirror of the reference's `utils/stylegan2/__init__.py:39-47`: `load_seq_stylegan(category,
truncation, **kw)` builds a `SeqStyleGAN2` and loads rosinality-format weights.  The
reference downloads from rewriting.csail.mit.edu; pass `path=` (or set
REWRITING_B200_WEIGHTS to a directory holding the same file names) to load from disk — the
download is attempted only if neither is given.
"""
import os
from collections import defaultdict

import torch

from .models import SeqStyleGAN2, DataBag

WEIGHT_URLS = 'http://rewriting.csail.mit.edu/data/models/'
sizes = defaultdict(lambda: 256, faces=1024, car=512)

FILENAMES = dict(
    bedroom='stylegan2_bedroom-6fa55a6e.pt',
    car='stylegan2_car-3659b4b6.pt',
    cat='stylegan2_cat-d8dc98b2.pt',
    church='stylegan2_church-e8ca9fd0.pt',
    faces='stylegan2_faces-2858cc2e.pt',
    horse='stylegan2_horse-499b5380.pt',
    kitchen='stylegan2_kitchen-b3a526e9.pt',
    places='stylegan2_places-a3b72d71.pt',
)


def load_state_dict(category, path=None):
    fn = FILENAMES[category]
    if path is None and os.environ.get('REWRITING_B200_WEIGHTS'):
        path = os.path.join(os.environ['REWRITING_B200_WEIGHTS'], fn)
    if path is not None:
        return torch.load(path, map_location='cpu')
    return torch.hub.load_state_dict_from_url(WEIGHT_URLS + fn, map_location='cpu')


def load_seq_stylegan(category, truncation=1.0, path=None, **kwargs):
    """Loads the nn.Sequential StyleGAN2 for `category` and puts it on the GPU."""
    state_dict = load_state_dict(category, path=path)
    g = SeqStyleGAN2(sizes[category], style_dim=512, n_mlp=8, truncation=truncation, **kwargs)
    g.load_state_dict(state_dict['g_ema'], latent_avg=state_dict['latent_avg'])
    g.cuda()
    return g
