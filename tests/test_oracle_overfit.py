"""CPU: the oracle's restatement of `all_weights_insert` (oracle/sg2_oracle.py, reference
rewrite/ganrewrite.py:300-331) and of the selection that feeds it (`rgb_from_selection` /
`rgbpaste_from_selection`, :522-539) against what the UNMODIFIED live reference produced
(tests/golden/overfit3.npz, oracle/make_golden_overfit.py)."""
import os

import numpy as np
import torch

from oracle import sg2_oracle as orc
from conftest import GOLD


def test_oracle_all_weights_insert_matches_live_reference(seeded_sd, z40, edit_request):
    from rewriting_b200.rewrite.ganrewrite import (positive_bounding_box, centered_location,
                                                   paste_clip_at_center)
    from rewriting_b200.synthetic import seeded_vgg16
    from rewriting_b200.utils import renormalize
    g = np.load(os.path.join(GOLD, 'overfit3.npz'))
    names = [str(n) for n in g['names']]
    o_imgnum, o_mask = edit_request['object']
    p_imgnum, p_mask = edit_request['paste']
    with torch.no_grad():
        x_obj = orc.generator_forward(seeded_sd, z40[o_imgnum][None])
        unchanged = orc.generator_forward(seeded_sd, z40[p_imgnum][None])
    area = renormalize.from_url(o_mask, target='pt', size=(256, 256))[0]
    t, l, b, r = positive_bounding_box(area)
    p_area = renormalize.from_url(p_mask, target='pt', size=(256, 256))[0]
    changed, bounds = paste_clip_at_center(unchanged, x_obj[:, :, t:b, l:r], centered_location(p_area),
                                           area[t:b, l:r])
    losses, grad0 = [], {}
    niter = 2                                  # two Adam steps (the golden holds three losses)
    trained = orc.all_weights_insert(seeded_sd, names, z40[p_imgnum][None], changed, bounds,
                                     seeded_vgg16().features, niter, lr=float(g['lr']),
                                     record_loss=losses, record_grad0=grad0)
    np.testing.assert_allclose(losses, g['losses'][:niter], rtol=2e-5)
    norms = np.array([float(grad0[k].norm()) for k in names])
    np.testing.assert_allclose(norms, g['grad0_norms'], rtol=2e-4)
    for i, k in enumerate(str(n) for n in g['kept']):
        want = torch.from_numpy(g['grad0_%d' % i])
        assert float((grad0[k] - want).norm() / want.norm()) < 2e-4, k
    assert set(trained) == set(names)
    assert all(not torch.equal(trained[k], seeded_sd[k]) for k in names)     # every tensor moved
