"""Image <-> tensor range conversions used on the rewrite path (API of the reference's
`utils/renormalize.py`): decoding the UI's data-URL PNG masks (`from_url`) and the
'zc' [-1,1] / 'pt' [0,1] / 'byte' [0,255] conversions.

Mask semantics to preserve: masks are RGBA PNGs, converted to RGB, and
bilinearly resized with PIL to the feature resolution; the RED channel in [0,1] is the
per-pixel weight the rewriter uses (`from_url(mask, target='pt', size=...)[0]`).
"""
import base64
import io
import re

import numpy
import PIL.Image
import torch

OFFSET_SCALE = dict(
    pt=([0.0, 0.0, 0.0], [1.0, 1.0, 1.0]),
    zc=([0.5, 0.5, 0.5], [0.5, 0.5, 0.5]),
    imagenet=([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
    byte=([0.0, 0.0, 0.0], [1.0 / 255, 1.0 / 255, 1.0 / 255]),
)


class Renormalizer(object):
    """x_new = (x_old*oldscale + oldoffset - newoffset) / newscale per channel."""

    def __init__(self, oldoffset, oldscale, newoffset, newscale, tobyte=False):
        oo, os_, no, ns = [numpy.array(v, dtype=numpy.float64)
                           for v in (oldoffset, oldscale, newoffset, newscale)]
        # computed in float64 on the host, cast to the data dtype on use (as the reference)
        self.mul = torch.from_numpy(os_ / ns)
        self.add = torch.from_numpy((oo - no) / ns)
        self.tobyte = tobyte
        self.mean, self.std = newoffset, newscale

    def __call__(self, data):
        shape = [1] * data.dim()
        shape[-3] = -1
        mul = self.mul.to(data.device, data.dtype).view(shape)
        add = self.add.to(data.device, data.dtype).view(shape)
        out = data.mul(mul).add_(add)
        if self.tobyte:
            out = out.clamp(0, 255).byte()
        return out


def renormalizer(source='zc', target='zc'):
    oldoffset, oldscale = OFFSET_SCALE[source] if isinstance(source, str) else OFFSET_SCALE['pt']
    newoffset, newscale = target if isinstance(target, tuple) else OFFSET_SCALE[target]
    return Renormalizer(oldoffset, oldscale, newoffset, newscale, tobyte=(target == 'byte'))


def as_tensor(data, source='zc', target='zc'):
    return renormalizer(source=source, target=target)(data)


def as_image(data, source='zc', target='byte'):
    assert len(data.shape) == 3
    arr = renormalizer(source=source, target=target)(data).permute(1, 2, 0).cpu().numpy()
    return PIL.Image.fromarray(arr)


def as_url(data, source='zc', size=None):
    img = data if isinstance(data, PIL.Image.Image) else as_image(data, source)
    if size is not None:
        img = img.resize(size, resample=PIL.Image.BILINEAR)
    buf = io.BytesIO()
    img.save(buf, format='png')
    return 'data:image/png;base64,%s' % base64.b64encode(buf.getvalue()).decode('utf-8')


def _pil_to_pt(im):
    """HWC uint8 PIL image -> CHW float in [0,1] (what torchvision's to_tensor does)."""
    arr = numpy.asarray(im, dtype=numpy.uint8)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    return torch.from_numpy(arr.copy()).permute(2, 0, 1).float().div(255)


def from_image(im, target='zc', size=None):
    # the reference tests `im.format != 'RGB'` (always true) and converts: RGBA masks lose
    # alpha and keep their colour channels.
    im = im.convert('RGB')
    if size is not None:
        im = im.resize(tuple(size), resample=PIL.Image.BILINEAR)
    return renormalizer(source='pt', target=target)(_pil_to_pt(im))


def from_url(url, target='zc', size=None):
    raw = base64.b64decode(re.sub('^data:image/.+;base64,', '', url))
    im = PIL.Image.open(io.BytesIO(raw))
    if target == 'image' and size is None:
        return im
    return from_image(im, target, size=size)
