"""Bring-up probe: C over 1000 z (config 4 golden) from the fast path, fused vs unfused up-convs."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from rewriting_b200.rewrite import ganrewrite
from rewriting_b200.utils import zdataset
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
c4 = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'config4_hat.npz')))
zds = torch.utils.data.TensorDataset(zdataset.standard_z_sample(1000, 512, seed=1))
for bs in (None, 10, 250):
    gw = ganrewrite.SeqStyleGanRewriter(model, zds, 8) if bs is None else gw
    C = gw.collect_2nd_moment(batch_size=bs).double().cpu() if bs else gw.c_matrix.double().cpu()
    Cg = torch.from_numpy(c4['C']).double()
    print('batch', bs, 'fused', os.environ.get('RW_UP_FUSED', '1'), 'rel-Fro vs reference %.3g' % ((C - Cg).norm() / Cg.norm()).item(),
          'max diag rel %.3g' % ((C.diag() - Cg.diag()).abs() / Cg.diag()).max().item(), flush=True)
