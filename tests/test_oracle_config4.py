"""CPU: the oracle against the BASELINE config-4 goldens that oracle/make_golden_config4.py
recorded from the live reference replaying notebooks/masks/stylegan/horse/hat_on_horse_ears.json
(zds = 1000, layer 8, rank 1, piter 10, lr 0.05)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sg2_oracle as orc
from conftest import GOLD


@pytest.fixture(scope='module')
def c4():
    return dict(np.load(os.path.join(GOLD, 'config4_hat.npz')))


@pytest.fixture(scope='module')
def hat_request():
    with open(os.path.join(GOLD, 'hat_on_horse_ears.json')) as f:
        return json.load(f)


def test_request_is_the_shipped_fixture(hat_request, c4):
    assert hat_request['object'][0] == 441 and hat_request['paste'][0] == 854
    assert [k[0] for k in hat_request['key']] == [354, 956, 309, 926]
    assert int(c4['n_z']) == 1000 and int(c4['layer']) == 8
    assert c4['goal_in_fmap'].shape == (1, 512, 8, 9) == c4['goal_out_fmap'].shape


def test_direction_from_golden_c(seeded_sd, c4, hat_request):
    from rewriting_b200.utils import renormalize, zdataset
    z = zdataset.standard_z_sample(1000, 512, seed=1)
    zca = orc.zca_from_cov(torch.from_numpy(c4['C']))
    obs, wts = [], []
    for imgnum, mask in hat_request['key']:
        with torch.no_grad():
            k = orc.generator_forward(seeded_sd, z[imgnum][None], upto_key_layer=8)
        obs.append(k.permute(0, 2, 3, 1).reshape(-1, 512))
        wts.append(renormalize.from_url(mask, target='pt', size=(32, 32))[0].view(-1)[:, None])
    d = orc.multi_key_zca(obs, wts, zca, rank=1)
    assert (d - torch.from_numpy(c4['d'])).abs().max().item() < 1e-4


def test_insert_50_iterations_match_reference(seeded_sd, c4):
    W0 = seeded_sd['layer8.sconv.mconv.dconv.weight']
    d = torch.from_numpy(c4['d'])
    losses = []
    W = orc.insert_loop(W0, torch.from_numpy(c4['goal_in_fmap']),
                        torch.from_numpy(c4['goal_in_style']),
                        torch.from_numpy(c4['goal_out_fmap']),
                        seeded_sd['layer8.sconv.noise.weight'],
                        seeded_sd['layer8.sconv.activate.bias'], d, 50, piter=10, lr=0.05,
                        record_loss=losses)
    lam = torch.einsum('goiyx,i->goyx', W - W0, d[0])[0]
    np.testing.assert_allclose(lam.numpy(), c4['lam50'], atol=1e-5, rtol=0)
    np.testing.assert_allclose(np.array(losses), c4['loss50'], rtol=1e-5)
    # after the final projection the edit lies exactly in span(d)
    resid = (W - W0)[0] - torch.einsum('oyx,i->oiyx', lam, d[0])
    assert resid.abs().max().item() < 1e-5


def test_2001_iteration_statistics_of_the_reference(c4):
    """what the reference itself achieves over the full horizon — the bars the
    GPU test holds the fused loop to."""
    assert float(c4['rel_fro_ref32_vs_fp64']) < 2e-2
    assert float(c4['sigma_ratio_ref32']) < 1e-6
    assert abs(float(c4['final_loss_ref32']) - float(c4['final_loss_fp64'])) < \
        1e-2 * float(c4['final_loss_fp64'])
    assert c4['lam2001_fp64'].shape == (512, 3, 3)
