"""Bring-up probe: key planes of layer 8 from the fast path at batch 40 vs the CPU oracle."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import sg2_oracle as orc  # noqa: E402
import bench  # noqa: E402
from rewriting_b200 import fastpath  # noqa: E402
from rewriting_b200.utils import zdataset  # noqa: E402

dev = torch.device('cuda', 0)
model = bench.build_model(dev)
sd = {k: v.cpu() for k, v in model.state_dict().items()}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 40
z = zdataset.standard_z_sample(B, 512, seed=1)
with torch.no_grad():
    planes = fastpath.forward(model, z.to(dev), upto_key_layer=8)
    got = (planes.hi.float() + planes.lo.float()).view(B, 33, 33, 512)[:, :32, :32].permute(0, 3, 1, 2).cpu()
    for i in (0, 1, 15, 16, 17, 31, 32, 33, B - 1):
        if i >= B:
            continue
        want = orc.generator_forward(sd, z[i:i + 1], upto_key_layer=8)
        print('img %d: max err %.3g (max %.3g)' % (i, (got[i] - want[0]).abs().max().item(),
                                                    want.abs().max().item()), flush=True)
