"""CPU: the oracle (oracle/sg2_oracle.py) against the committed golden vectors that
oracle/make_golden.py produced from the live reference."""
import json
import os

import numpy as np
import torch

from oracle import sg2_oracle as orc
from conftest import GOLD


def test_seeded_weights_match_reference_checksums(seeded_sd):
    with open(os.path.join(GOLD, 'weights_checksum.json')) as f:
        chk = json.load(f)
    assert set(chk) == set(seeded_sd)
    assert len(chk) == 136
    for k, (s, a) in chk.items():
        v = seeded_sd[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, a), k


def test_package_synthetic_weights_are_the_golden_weights(seeded_sd):
    """bench.py / smoke() build their model with rewriting_b200.synthetic (nothing under oracle/
    is imported by the product arm); it must be the model the goldens were generated with."""
    from rewriting_b200.synthetic import seeded_generator
    sd = seeded_generator().state_dict()
    assert list(sd) == list(seeded_sd)
    for k in sd:
        assert torch.equal(sd[k], seeded_sd[k]), k


def test_generator_pixels_match_golden(seeded_sd, z40, golden):
    rec = {}
    with torch.no_grad():
        pix = orc.generator_forward(seeded_sd, z40[:2], record=rec)
    assert pix.shape == (2, 3, 256, 256)
    np.testing.assert_allclose(pix[:, :, ::8, ::8].numpy(), golden['pixels_sub'], atol=1e-5, rtol=0)
    np.testing.assert_allclose(rec['layer8']['k'][0, ::16, ::4, ::4].numpy(), golden['l8_key_sub'],
                               atol=1e-6, rtol=0)
    np.testing.assert_allclose(rec['layer8']['y'][0, ::16, ::4, ::4].numpy(), golden['l8_out_sub'],
                               atol=1e-6, rtol=0)
    np.testing.assert_allclose(rec['layer9']['y'][0, ::16, ::8, ::8].numpy(), golden['l9_out_sub'],
                               atol=1e-6, rtol=0)


def test_second_moment_zca_direction_match_golden(seeded_sd, z40, golden, edit_request):
    from rewriting_b200.utils import renormalize
    with torch.no_grad():
        keys = [orc.generator_forward(seeded_sd, z40[i:i + 10], upto_key_layer=8)
                for i in range(0, 40, 10)]
    mom2, count = orc.second_moment(keys)
    assert count == int(golden['count']) == 40 * 32 * 32
    C = mom2 / count
    np.testing.assert_allclose(C[::8, ::8].numpy(), golden['C_sub'], rtol=0,
                               atol=2e-5 * float(golden['C_diag'].max()))
    assert abs(float(C.trace()) - float(golden['C_trace'])) < 1e-4 * float(golden['C_trace'])
    zca = orc.zca_from_cov(C)
    np.testing.assert_allclose(zca[::8, ::8].numpy(), golden['zca_sub'], rtol=0,
                               atol=2e-3 * float(np.abs(golden['zca_sub']).max()))
    obs, wts = [], []
    for imgnum, mask in edit_request['key']:
        with torch.no_grad():
            k = orc.generator_forward(seeded_sd, z40[imgnum][None], upto_key_layer=8)
        obs.append(k.permute(0, 2, 3, 1).reshape(-1, 512))
        wts.append(renormalize.from_url(mask, target='pt', size=(32, 32))[0].view(-1)[:, None])
    d = orc.multi_key_zca(obs, wts, zca, rank=1)
    assert d.shape == (1, 512)
    assert abs(float(d.norm()) - 1.0) < 1e-5
    # d is ill-conditioned in C (few samples): compare as directions
    cos = float((d[0] * torch.from_numpy(golden['d'][0])).sum())
    assert cos > 1 - 1e-4


def test_insert_loop_matches_golden(seeded_sd, golden):
    W0 = seeded_sd['layer8.sconv.mconv.dconv.weight']
    losses = []
    W = orc.insert_loop(W0, torch.from_numpy(golden['goal_in_fmap']),
                        torch.from_numpy(golden['goal_in_style']),
                        torch.from_numpy(golden['goal_out_fmap']),
                        seeded_sd['layer8.sconv.noise.weight'],
                        seeded_sd['layer8.sconv.activate.bias'],
                        torch.from_numpy(golden['d']), int(golden['niter']), piter=10, lr=0.05,
                        record_loss=losses)
    delta = (W - W0)[0, ::37, ::41].numpy()
    np.testing.assert_allclose(delta, golden['W_delta_sub'], atol=1e-5, rtol=0)
    np.testing.assert_allclose(np.array(losses), golden['losses'], rtol=1e-5)
    # the edit is rank one in the (out*taps) x in matricisation, row space = d
    dW = (W - W0)[0].permute(0, 2, 3, 1).reshape(-1, 512)
    s = torch.linalg.svdvals(dW.double())
    assert float(s[1] / s[0]) < 1e-5


def test_upfirdn2d_and_projection_properties():
    torch.manual_seed(3)
    x = torch.randn(2, 3, 8, 8)
    k = orc.make_kernel([1, 3, 3, 1]) * 4
    up = orc.upfirdn2d(x, k, up=2, pad=(2, 1))
    assert up.shape == (2, 3, 16, 16)
    # a constant image stays constant under the normalised 2x upsampler (interior)
    c = orc.upfirdn2d(torch.ones(1, 1, 8, 8), k, up=2, pad=(2, 1))
    assert torch.allclose(c[:, :, 2:-2, 2:-2], torch.ones_like(c[:, :, 2:-2, 2:-2]), atol=1e-6)
    W = torch.randn(1, 16, 32, 3, 3)
    q, _ = torch.linalg.qr(torch.randn(32, 2))
    d = q.t()
    P = orc.projected_conv(W, d)
    # idempotent, and the residual is orthogonal to d
    assert torch.allclose(orc.projected_conv(P, d), P, atol=1e-5)
    assert torch.einsum('goiyx,di->godyx', W - P, d).abs().max() < 1e-5
