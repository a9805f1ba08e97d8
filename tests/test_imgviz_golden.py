"""CPU: the overlay renderer behind render_image(key, level) / render_object(box) against
byte images produced by the live reference's ImageVisualizer (oracle/make_golden_imgviz.py)."""
import os

import numpy as np
import torch

from rewriting_b200.utils import imgviz

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'imgviz_overlay.npz')


def _inputs():
    g = torch.Generator().manual_seed(5)
    img = torch.randn(1, 3, 64, 64, generator=g).clamp(-1, 1)
    g = torch.Generator().manual_seed(6)
    return img, torch.randn(8, 8, generator=g)


def test_masked_image_matches_reference_bytes():
    gold = np.load(GOLD)
    img, heat = _inputs()
    iv = imgviz.ImageVisualizer(img.shape[2:])
    np.testing.assert_array_equal(np.asarray(iv.masked_image(img, heat, level=0.3)), gold['heat_level'])
    np.testing.assert_array_equal(
        np.asarray(iv.masked_image(img, heat, level=-0.2, thickness=2, outside_bright=0.25)),
        gold['heat_thick'])
    low = torch.zeros(8, 8)
    low[2:5, 1:6] = 1
    np.testing.assert_array_equal(
        np.asarray(iv.masked_image(img, activations=low, level=0.0, border_color=[255, 0, 0],
                                   thickness=3)), gold['box'])
    mask = torch.zeros(64, 64, dtype=torch.bool)
    mask[10:30, 20:50] = True
    np.testing.assert_array_equal(np.asarray(iv.masked_image(img, mask=mask)), gold['mask'])
    np.testing.assert_array_equal(np.asarray(iv.masked_image(img, heat, percent_level=0.8)),
                                  gold['percent'])


def test_border_is_outside_the_mask_and_thickness_grows():
    mask = torch.zeros(20, 20, dtype=torch.bool)
    mask[5:12, 6:14] = True
    b1 = imgviz.border_from_mask(mask, 1)
    b3 = imgviz.border_from_mask(mask, 3)
    assert not (b1 & mask).any() and not (b3 & mask).any()
    assert b1.sum() < b3.sum() and (b1 & ~b3).sum() == 0
    assert imgviz.border_from_mask(mask, 1, outside=False).sum() > b1.sum()
