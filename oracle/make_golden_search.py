"""TEST INFRASTRUCTURE ONLY — golden vectors for the UI-search and erase paths of the rewriter
(ganrewrite.py:375-400,453-496,541-594), produced by the live reference in the authoring
container (needs /root/reference; ~15 min on 8 cores — the reference's 512-unit quantile
sketch and its per-unit numpy.interp loop dominate):

    python oracle/make_golden_search.py     ->  tests/golden/search_erase.npz

Same seeded model / z / masks as oracle/make_golden.py.  Stored: ranking_for_key of the golden
direction d (image indexes, quantiles of the responses), the unit square scales, the
normdissect units, the gandissect (quantile-scored) units, the rank-2 zca directions, and the
erase goal crops.  tests/test_oracle_search.py pins the CPU oracle to these numbers.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')

from oracle import sg2_oracle as orc          # noqa: E402
from oracle.ref_shim import load_reference    # noqa: E402

QS = [0.01, 0.5, 0.99, 0.999]


def main():
    torch.set_num_threads(os.cpu_count())
    ref = load_reference()
    model = orc.seeded_state_dict(
        lambda: ref.models.SeqStyleGAN2(256, style_dim=512, n_mlp=8, mconv='seq')).eval()
    z = ref.zdataset.standard_z_sample(40, 512, seed=1)
    zds = torch.utils.data.TensorDataset(z)
    with open(os.path.join(GOLD, 'edit_request.json')) as f:
        request = json.load(f)
    gold = np.load(os.path.join(GOLD, 'sg2_layer8.npz'))
    d = torch.from_numpy(gold['d'])
    gw = ref.ganrewrite.SeqStyleGanRewriter(model, zds, 8, cachedir=None)
    out = {}
    torch.manual_seed(11)                       # the reference's quantile sketch draws random bits
    sel, rq = gw.ranking_for_key(d[0], k=6)
    out['rank_sel'] = sel.numpy()
    out['rank_quantiles'] = rq.quantiles(QS)[0].numpy()
    out['rank_count'] = rq.size()
    out['unit_rs'] = gw.square_scales_for_units().numpy()
    pairs = [tuple(p) for p in request['key']]
    out['normdissect_units'] = gw.normdissect_units(pairs, 30).numpy()
    torch.manual_seed(12)
    gd = gw.multi_key_from_selection(pairs, rank=3, key_method='gandissect')
    out['gandissect_units'] = gd.argmax(dim=1).numpy()
    with torch.no_grad():
        d2 = gw.multi_key_from_selection(pairs, rank=2)
    out['d_rank2'] = d2.numpy()
    with torch.no_grad():
        gin, gout = gw.erase_from_selection(request['paste'][0], request['paste'][1], pairs, 30)
    gi, go = gin.fmap.numpy(), gout.fmap.numpy()          # [1,512,32,32] (tight_paste is off)
    out['erase_goal_in_sub'], out['erase_goal_out_sub'] = gi[:, ::8, ::2, ::2], go[:, ::8, ::2, ::2]
    out['erase_goal_in_fro'] = np.linalg.norm(gi.ravel())
    out['erase_goal_out_fro'] = np.linalg.norm(go.ravel())
    np.savez_compressed(os.path.join(GOLD, 'search_erase.npz'), **out)
    print({k: getattr(v, 'shape', v) for k, v in out.items()})


if __name__ == '__main__':
    main()
