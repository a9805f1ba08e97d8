"""CPU: the oracle's restatement of the UI-search / erase statistics (oracle/sg2_oracle.py)
against golden vectors produced by the live reference (oracle/make_golden_search.py).  The GPU
tests compare the CUDA path with these same oracle functions (tests/test_gpu_parity.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sg2_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def sgold():
    return dict(np.load(os.path.join(GOLD, 'search_erase.npz')))


@pytest.fixture(scope='module')
def keys40(seeded_sd, z40):
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    with torch.no_grad():
        return [orc.generator_forward(seeded_sd, z40[i:i + 10], upto_key_layer=8)
                for i in range(0, 40, 10)]


@pytest.fixture(scope='module')
def observed(seeded_sd, z40, edit_request):
    """keys of the selected images, each computed ALONE: the reference regenerates the noise
    table for the batch at hand (models.py:542-545), so an image's key depends on its batch
     — get_z(imgnum) is a batch of one, the tallies use batches of 10."""
    from rewriting_b200.utils import renormalize
    obs, wts = [], []
    for imgnum, mask in edit_request['key']:
        with torch.no_grad():
            k = orc.generator_forward(seeded_sd, z40[imgnum][None], upto_key_layer=8)
        obs.append(orc.flat_keys(k))
        wts.append(renormalize.from_url(mask, target='pt', size=(32, 32))[0].view(-1)[:, None])
    return obs, wts


def test_ranking_for_key_matches_reference(keys40, golden, sgold):
    d = torch.from_numpy(golden['d'])[0]
    sel, resp = orc.ranking_for_key(torch.cat(keys40), d, k=6)
    np.testing.assert_array_equal(sel.numpy(), sgold['rank_sel'])
    assert resp.numel() == int(sgold['rank_count'])
    # the reference's numbers come from its randomised sketch (r = 4096): rank-space agreement
    qs = torch.tensor([0.01, 0.5, 0.99, 0.999])
    ranks = (resp[:, None] <= torch.from_numpy(sgold['rank_quantiles'])[None, :]).float().mean(0)
    assert (ranks - qs).abs().max().item() < 3e-3
    exact = torch.quantile(resp.double(), qs.double())
    ref_q = torch.from_numpy(sgold['rank_quantiles']).double()
    assert (exact - ref_q).abs().max().item() < 0.02 * ref_q.abs().max().item()


def test_square_scales_and_normdissect_units_match_reference(keys40, observed, sgold):
    rs = orc.square_scales_for_units(keys40)
    np.testing.assert_allclose(rs.numpy(), sgold['unit_rs'], rtol=2e-5)
    obs, wts = observed
    units = orc.normdissect_units(obs, wts, rs, 30)
    np.testing.assert_array_equal(units.numpy(), sgold['normdissect_units'])


def test_gandissect_units_match_reference(keys40, observed, sgold):
    flat = orc.flat_keys(torch.cat(keys40))
    sorted_units = torch.sort(flat.t().contiguous(), dim=1)[0]
    obs, wts = observed
    units = orc.gandissect_units(obs, wts, sorted_units, 3)
    # exact quantiles vs the reference's sketch: the winning units agree
    assert units.tolist() == sgold['gandissect_units'].tolist()
    # and the product's exact RunningQuantile gives the same scores as the oracle's rank formula
    from rewriting_b200.utils import runningstats
    rq = runningstats.RunningQuantile()
    for kb in keys40:
        rq.add(orc.flat_keys(kb))
    all_obs, all_w = torch.cat(obs), torch.cat(wts)
    logscore = -torch.log(1.0 - rq.normalize(all_obs.permute(1, 0))).permute(1, 0)
    mean_logscore = (logscore * all_w).sum(0) / all_w.sum()
    assert mean_logscore.sort(descending=True)[1][:3].tolist() == units.tolist()


def test_rank2_direction_matches_reference(keys40, observed, golden, sgold, seeded_sd):
    mom2, count = orc.second_moment(keys40)
    zca = orc.zca_from_cov(mom2 / count)
    obs, wts = observed
    d2 = orc.multi_key_zca(obs, wts, zca, rank=2)
    ref = torch.from_numpy(sgold['d_rank2'])
    # C from 40 z is ill-conditioned: compare as directions, row by row
    for r in range(2):
        assert abs(float((d2[r] * ref[r]).sum())) > 1 - 1e-4
    assert (d2 @ d2.t() - torch.eye(2)).abs().max().item() < 1e-5


def test_erase_goal_matches_reference(keys40, observed, sgold, seeded_sd, z40, edit_request):
    """erase_from_selection (ganrewrite.py:473-496, tight_paste off): goal_in is the untouched
    key, goal_out the target model's output with the normdissect units zeroed."""
    rs = orc.square_scales_for_units(keys40)
    obs, wts = observed
    units = orc.normdissect_units(obs, wts, rs, 30)
    imgnum = edit_request['paste'][0]
    with torch.no_grad():
        k = orc.generator_forward(seeded_sd, z40[imgnum][None], upto_key_layer=8)
    np.testing.assert_allclose(k[:, ::8, ::2, ::2].numpy(), sgold['erase_goal_in_sub'],
                               rtol=1e-5, atol=1e-6)
    assert abs(float(k.norm()) - float(sgold['erase_goal_in_fro'])) < 1e-4 * float(sgold['erase_goal_in_fro'])
    p = orc._layer_params(seeded_sd, 'layer8')
    with torch.no_grad():
        style = orc.modulate(orc.mapping(seeded_sd, z40[imgnum][None]), p['mod_w'], p['mod_b'])
        without = k.clone()
        without[:, units] = 0.0
        out = orc.target_forward(without, style, p['weight'], p['noise_w'], p['bias'])
    np.testing.assert_allclose(out[:, ::8, ::2, ::2].numpy(), sgold['erase_goal_out_sub'],
                               rtol=1e-4, atol=1e-5)
    assert abs(float(out.norm()) - float(sgold['erase_goal_out_fro'])) < 1e-4 * float(sgold['erase_goal_out_fro'])
