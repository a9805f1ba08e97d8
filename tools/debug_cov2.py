"""Bring-up probe: per-batch key second moment from the fast path vs the CPU oracle (fp32 and the
same in fp64), to see where the 1.6e-4 rel-Frobenius of C over 1000 z comes from."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import sg2_oracle as orc
from rewriting_b200 import fastpath
from rewriting_b200.utils import zdataset
torch.set_num_threads(32)
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
sd = {k: v.cpu() for k, v in model.state_dict().items()}
sd64 = {k: v.double() for k, v in sd.items()}
z = zdataset.standard_z_sample(1000, 512, seed=1)
for j in (0, 1, 37, 85):
    zb = z[10 * j:10 * j + 10]
    with torch.no_grad():
        k32 = orc.generator_forward(sd, zb, upto_key_layer=8)
        k64 = orc.generator_forward(sd64, zb.double(), upto_key_layer=8)
        planes = fastpath.forward(model, zb.to(dev), upto_key_layer=8)
    kg = (planes.hi.float() + planes.lo.float()).view(10, 33, 33, 512)[:, :32, :32].permute(0, 3, 1, 2).cpu().double()
    def mom(k):
        f = k.permute(0, 2, 3, 1).reshape(-1, 512).double()
        return f.t() @ f / f.shape[0]
    C64, C32, Cg = mom(k64), mom(k32.double()), mom(kg)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    print('batch %d: keys gpu-vs-fp64 rel %.3g (max abs %.3g of %.3g), cpu32-vs-fp64 rel %.3g | C: gpu-vs-fp64 %.3g, '
          'cpu32-vs-fp64 %.3g, gpu-vs-cpu32 %.3g' % (j, rel(kg, k64), (kg - k64).abs().max(), k64.abs().max(),
                                                    rel(k32.double(), k64), rel(Cg, C64), rel(C32, C64), rel(Cg, C32)),
          flush=True)
    per_img = [(kg[i] - k64[i]).norm().item() / k64[i].norm().item() for i in range(10)]
    print('   per-image key rel error:', ' '.join('%.2g' % v for v in per_img), flush=True)
