"""CPU: the ProgGAN oracle (oracle/proggan_oracle.py) and the host side of
rewriting_b200.utils.proggan against the goldens oracle/make_golden_proggan.py recorded from the
live reference and from the kitchen fixtures the reference ships."""
import os

import numpy as np
import pytest
import torch

from oracle import proggan_oracle as ppo
from conftest import GOLD


@pytest.fixture(scope='module')
def pg():
    return dict(np.load(os.path.join(GOLD, 'proggan64.npz')))


@pytest.fixture(scope='module')
def kitchen():
    return dict(np.load(os.path.join(GOLD, 'proggan_kitchen_layer6.npz'), allow_pickle=True))


@pytest.fixture(scope='module')
def pg_model():
    from rewriting_b200.utils import proggan
    return ppo.seeded_state_dict(lambda: proggan.ProgressiveGenerator(resolution=64))


def test_module_tree_and_state_dict_keys(pg_model):
    names = list(pg_model._modules)
    assert names == ['layer%d' % i for i in range(1, 11)] + ['output_64x64']
    assert list(pg_model.layer3._modules) == ['norm', 'up', 'conv', 'wscale', 'relu']
    assert list(pg_model.layer6._modules) == ['norm', 'conv', 'wscale', 'relu']
    sd = pg_model.state_dict()
    assert sd['layer6.conv.weight'].shape == (512, 512, 3, 3) and sd['layer1.conv.weight'].shape[2] == 4
    from rewriting_b200.utils import proggan
    assert proggan.sizes_from_state_dict(sd) == [512, 512, 512, 512, 256, 128]
    again = proggan.from_state_dict({'state_dict': sd})
    assert all(torch.equal(v, sd[k]) for k, v in again.state_dict().items())
    old = {}
    for i in range(10):
        old['features.%d.conv.weight' % i] = sd['layer%d.conv.weight' % (i + 1)]
        old['features.%d.wscale.b' % i] = sd['layer%d.wscale.b' % (i + 1)]
    old['output.conv.weight'] = sd['output_64x64.conv.weight']
    old['output.wscale.b'] = sd['output_64x64.wscale.b']
    conv = proggan.from_old_pt_dict(old)
    assert all(torch.equal(v, sd[k]) for k, v in conv.state_dict().items())
    # the rewriter's split addresses the same leaves as in the reference
    from rewriting_b200.utils import nethook
    ctx = nethook.subsequence(pg_model, upto_layer='layer6.conv', share_weights=True)
    tgt = nethook.subsequence(pg_model, first_layer='layer6.conv', last_layer='layer6.conv',
                              share_weights=True)
    assert [n for n, _ in tgt.named_parameters()] == ['layer6.conv.weight']
    assert list(ctx._modules)[-1] == 'layer6' and list(ctx.layer6._modules) == ['norm']


def test_oracle_forward_and_statistics_match_golden(pg_model, pg):
    from rewriting_b200.utils import zdataset
    sd = pg_model.state_dict()
    z = zdataset.z_sample_for_model(pg_model, 40, seed=1)
    assert z.shape == (40, 512, 1, 1)
    with torch.no_grad():
        pix = ppo.generator_forward(sd, z[:2])
        keys = torch.cat([ppo.generator_forward(sd, z[i:i + 10], upto_key_layer=6)
                          for i in range(0, 40, 10)])
    np.testing.assert_allclose(pix[:, :, ::2, ::2].numpy(), pg['pixels_sub'], atol=1e-5, rtol=0)
    np.testing.assert_allclose(keys[0, ::16].numpy(), pg['key_sub'], atol=1e-5, rtol=0)
    flat = keys.permute(0, 2, 3, 1).reshape(-1, 512).double()
    C = (flat.t() @ flat / flat.shape[0])
    Cg = torch.from_numpy(pg['C']).double()
    assert ((C - Cg).norm() / Cg.norm()).item() < 1e-5
    # rendering from the raw layer-6 conv output == the whole generator
    with torch.no_grad():
        v = torch.nn.functional.conv2d(keys[:2], sd['layer6.conv.weight'], padding=1)
        again = ppo.generator_forward(sd, None, from_layer_output=(6, v))
    assert (again - pix).abs().max().item() < 1e-5


def test_oracle_insert_matches_reference(pg_model, pg):
    W0 = pg_model.state_dict()['layer6.conv.weight']
    d = torch.from_numpy(pg['d'])
    losses = []
    W = ppo.insert_loop(W0, torch.from_numpy(pg['goal_in']), torch.from_numpy(pg['goal_out']), d,
                        int(pg['niter']), piter=10, lr=0.05, record_loss=losses)
    lam = torch.einsum('oiyx,i->oyx', W - W0, d[0])
    np.testing.assert_allclose(lam.numpy(), pg['lam'], atol=1e-5, rtol=0)
    np.testing.assert_allclose(np.array(losses), pg['losses'], rtol=1e-5)


def test_kitchen_fixtures_known_answers(kitchen):
    """the real-data fixtures of the reference: cache format of r2m.npz and the rank-one edit of
    the paper's reflection example (reflection_switched_layer6.npz)"""
    from rewriting_b200.utils import runningstats
    assert str(kitchen['constructor']).endswith('runningstats.RunningSecondMoment()')
    r = runningstats.RunningSecondMoment(state=kitchen)
    assert r.count == 256000 and r.mom2.shape == (512, 512)          # 1000 z x 16 x 16 keys
    C = r.moment()
    assert torch.allclose(C, C.t(), atol=1e-4 * float(C.abs().max())) and float(C.diag().min()) > 0
    st = r.state_dict()
    assert set(st) == {'constructor', 'count', 'mom2'} and st['count'] == 256000
    assert float(kitchen['sigma_ratio']) < 1e-6                         # delta W is rank one
    assert abs(float(np.linalg.norm(kitchen['d'])) - 1.0) < 1e-5
    # the oracle's edit of the REAL weights along the REAL direction reproduces its golden
    d = torch.from_numpy(kitchen['d'])[None]
    W0 = torch.from_numpy(kitchen['W_unopt_sub'])
    W = ppo.insert_loop(W0, torch.from_numpy(kitchen['key_crop']), torch.from_numpy(kitchen['target']),
                        d, 20, piter=10, lr=0.05)
    lam = torch.einsum('oiyx,i->oyx', W - W0, d[0])
    np.testing.assert_allclose(lam.numpy(), kitchen['lam20'], atol=2e-5, rtol=0)
