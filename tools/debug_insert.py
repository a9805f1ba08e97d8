"""Bring-up probe: fused insert loop vs the CPU oracle (torch.optim.Adam) from identical state —
how many of the 2.4 M weights differ by more than 1e-4 after n iterations, and by how much."""
import copy
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle import sg2_oracle as orc  # noqa: E402
import bench  # noqa: E402
from rewriting_b200.rewrite import ganrewrite  # noqa: E402
from rewriting_b200.utils import zdataset  # noqa: E402

dev = torch.device('cuda', 0)
model = bench.build_model(dev)
golden = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'sg2_layer8.npz')))
z = zdataset.standard_z_sample(10, 512, seed=1)
zds = torch.utils.data.TensorDataset(z)
sd = {k: v.cpu() for k, v in model.state_dict().items()}
for niter in (1, 2, 3, 11, 30):
    gw = ganrewrite.SeqStyleGanRewriter(model, zds, 8)
    bag = gw.context_model(gw.get_z(0))
    gin = type(bag)(bag, fmap=torch.from_numpy(golden['goal_in_fmap']).cuda(),
                    style=torch.from_numpy(golden['goal_in_style']).cuda())
    gout = type(bag)(bag, fmap=torch.from_numpy(golden['goal_out_fmap']).cuda())
    W0 = gw.target_weights().detach().clone().cpu()
    gw.insert(gin, gout, torch.from_numpy(golden['d']).cuda(), niter=niter, piter=10, lr=0.05)
    W = gw.target_weights().detach().cpu()
    W_orc = orc.insert_loop(W0, torch.from_numpy(golden['goal_in_fmap']),
                            torch.from_numpy(golden['goal_in_style']),
                            torch.from_numpy(golden['goal_out_fmap']),
                            sd['layer8.sconv.noise.weight'], sd['layer8.sconv.activate.bias'],
                            torch.from_numpy(golden['d']), niter, piter=10, lr=0.05)
    diff = (W - W_orc).abs()
    lam = torch.einsum('goiyx,i->goyx', W - W0, torch.from_numpy(golden['d'])[0])
    lam_o = torch.einsum('goiyx,i->goyx', W_orc - W0, torch.from_numpy(golden['d'])[0])
    dl = (lam - lam_o).abs()
    print('niter %d: max|dW| %.3g; max diff %.3g; entries > 1e-4: %d of %d; rel-Fro %.3g; lambda: max diff '
          '%.3g, entries > 1e-4: %d of %d' % (
              niter, (W_orc - W0).abs().max(), diff.max(), int((diff > 1e-4).sum()), diff.numel(),
              ((W - W_orc).norm() / (W_orc - W0).norm()), dl.max(), int((dl > 1e-4).sum()), dl.numel()),
          flush=True)
