"""Overlay rendering for the rewriter's UI hooks: the part of the reference's
`utils/imgviz.py` that `ganrewrite.render_image(key=, level=)`, `render_image_batch` and
`render_object(box=)` call (imgviz.py:83-122,170-196,309-330 with `utils/upsample.py:124-155`
for the activation upsampling).  Heat-map colouring (matplotlib colour maps), top-k image grids
and the segmentation visualisers of the reference are outside the rewriting path.

`ImageVisualizer(size).masked_image(imagedata, activations, level=..)` thresholds the
bilinearly upsampled activation map at `level`, keeps the image inside the mask, dims it
outside (`outside_bright`) and draws a border of `thickness` pixels around the mask.
Pure torch + PIL on whatever device the data is on; results are byte images on the host.
"""
import PIL.Image
import torch

from . import renormalize


def upsample_grid(data_shape, target_shape, dtype=torch.float, device=None):
    """grid_sample grid that stretches `data_shape` features over `target_shape` pixels with
    feature centres kept at pixel-block centres (upsample.py:124-155, default scale/offset)."""
    axes = []
    for ts, ds in zip(target_shape, data_shape):
        scale = float(ts) / ds
        offset = 0.5 * scale - 0.5
        axes.append((torch.arange(ts, dtype=dtype, device=device) - offset)
                    * (2 / (scale * max(1, ds - 1))) - 1)
    ty, tx = axes
    return torch.stack((tx[None, :].expand(*target_shape), ty[:, None].expand(*target_shape)),
                       2)[None]


def upsample(data, target_shape, mode='bilinear'):
    """[N,C,h,w] -> [N,C,H,W] (upsample.py:30-43: grid_sample, zeros padding, align_corners)."""
    grid = upsample_grid(data.shape[2:], tuple(target_shape), dtype=data.dtype, device=data.device)
    return torch.nn.functional.grid_sample(data, grid.expand(data.shape[0], -1, -1, -1), mode=mode,
                                           padding_mode='zeros', align_corners=True)


def border_from_mask(mask, thickness=1, outside=True):
    """Pixels within `thickness` of a change of `mask` (8-neighbourhood); with `outside`, only
    those not in the mask (imgviz.py:309-330)."""
    a = mask
    out = torch.zeros_like(a)
    for it in range(thickness):
        h = a[:-1, :] != a[1:, :]
        v = a[:, :-1] != a[:, 1:]
        d = a[:-1, :-1] != a[1:, 1:]
        u = a[1:, :-1] != a[:-1, 1:]
        out[:-1, :-1] |= d
        out[1:, 1:] |= d
        out[1:, :-1] |= u
        out[:-1, 1:] |= u
        out[:-1, :] |= h
        out[1:, :] |= h
        out[:, :-1] |= v
        out[:, 1:] |= v
        if it > 0:
            out |= a
        a = out
    if outside:
        out &= ~mask
    return out


class ImageVisualizer(object):
    def __init__(self, size, renormalizer=None, level=None, percent_level=None):
        if isinstance(size, int):
            size = (size, size)
        self.size = tuple(int(s) for s in size)
        self.renormalizer = renormalizer
        self.level = level
        self.percent_level = percent_level

    def pytorch_image(self, imagedata):
        """[-1,1] image tensor -> float byte-range tensor [3, H, W] at the visualised size."""
        if len(imagedata.shape) == 4:
            imagedata = imagedata[0]
        renorm = self.renormalizer or renormalize.renormalizer('zc', 'byte')
        return torch.nn.functional.interpolate(renorm(imagedata).float()[None, ...],
                                               size=self.size)[0]

    def image(self, imagedata):
        return PIL.Image.fromarray(self.pytorch_image(imagedata).permute(1, 2, 0).byte()
                                   .cpu().numpy())

    def level_for(self, activations, unit, percent_level=None):
        if unit is not None and self.level is not None:
            if hasattr(unit, '__len__'):
                unit = unit[1]
            return self.level[unit].item()
        s, _ = activations.reshape(-1).sort()
        if percent_level is None:
            percent_level = self.percent_level or 0.95
        return s[int(len(s) * percent_level)]

    def pytorch_mask(self, activations, unit, level=None, percent_level=None):
        a = activations if unit is None else activations[unit]
        if level is None:
            level = self.level_for(activations, unit, percent_level=percent_level)
        return upsample(a[None, None, ...].float(), self.size)[0, 0] > level

    def pytorch_masked_image(self, imagedata, activations=None, unit=None, level=None,
                             percent_level=None, thickness=1, mask=None, border_color=None,
                             outside_bright=0.5, inside_color=None):
        scaled = self.pytorch_image(imagedata).float().cpu()
        if mask is None:
            mask = self.pytorch_mask(activations, unit, level=level,
                                     percent_level=percent_level).cpu()
        border = border_from_mask(mask, thickness)
        inside = (mask & (~border)).float()
        outside = (~mask & (~border)).float()
        border = border.float()
        if border_color is None:
            border_color = [255.0, 255.0, 0]                     # yellow
        border_color = torch.tensor(border_color, dtype=border.dtype)[:, None, None]
        body = scaled
        if inside_color is not None:
            body = torch.tensor(inside_color, dtype=border.dtype)[:, None, None]
        return (body * inside + border_color * border +
                outside_bright * scaled * outside).clamp(0, 255).byte()

    def masked_image(self, imagedata, activations=None, unit=None, level=None,
                     percent_level=None, **kwargs):
        img = self.pytorch_masked_image(imagedata, activations=activations, unit=unit, level=level,
                                        percent_level=percent_level, **kwargs)
        return PIL.Image.fromarray(img.permute(1, 2, 0).cpu().numpy())
