"""Prints how close apply_overfit is to tests/golden/overfit3.npz (losses, first-iteration gradient
norms of all parameter tensors, element-wise updates): the numbers behind the tolerances of
tests/test_gpu_overfit.py."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import json          # noqa: E402
import numpy as np   # noqa: E402
import torch         # noqa: E402


def main():
    from test_gpu_overfit import _run
    from oracle import sg2_oracle as orc
    from rewriting_b200.utils.stylegan2 import SeqStyleGAN2
    from rewriting_b200.utils import zdataset
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'overfit3.npz'))
    model = orc.seeded_state_dict(lambda: SeqStyleGAN2(256, style_dim=512, n_mlp=8, mconv='seq')).eval()
    z40 = zdataset.standard_z_sample(40, 512, seed=1)
    with open(os.path.join(ROOT, 'tests', 'golden', 'edit_request.json')) as f:
        req = json.load(f)
    losses, grad0, upd = _run(model, z40, req, int(g['niter']), float(g['lr']))
    names = [str(n) for n in g['names']]
    print('losses', losses, 'golden', g['losses'].tolist())
    print('rel', (np.array(losses) / g['losses'] - 1).tolist())
    norms = np.array([float(grad0[k].norm()) for k in names])
    rel = np.abs(norms / g['grad0_norms'] - 1)
    print('grad norm rel error: max %.2e (%s), median %.2e' % (rel.max(), names[int(rel.argmax())], np.median(rel)))
    for i, k in enumerate(str(n) for n in g['kept']):
        want = torch.from_numpy(g['grad0_%d' % i])
        d = (upd[k] - torch.from_numpy(g['upd_%d' % i])).abs()
        print('%-40s grad rel-fro %.2e | update: max %.2e, median %.2e, frac<1e-4 %.4f' % (
            k, float((grad0[k] - want).norm() / want.norm()), float(d.max()), float(d.median()),
            float((d < 1e-4).float().mean())))
    for lname in ('layer3', 'layer8', 'layer13', 'layer14'):
        want = torch.from_numpy(g['grad0_w_%s' % lname])
        got = grad0['%s.sconv.mconv.dconv.weight' % lname][0, ::37, ::41]
        print(lname, 'dconv.weight grad sample rel-fro %.2e' % float((got - want).norm() / want.norm()))
    sums = np.array([float(upd[k].abs().sum()) for k in names])
    r = np.abs(sums / g['abs_update_sums'] - 1)
    print('update sums rel error: max %.2e (%s)' % (r.max(), names[int(r.argmax())]))
    # throughput of the all-weights baseline (batch 1: forward + backward of the whole generator,
    # VGG-16 features of the crop, Adam over 30 M parameters)
    import copy
    import time
    from rewriting_b200.rewrite import ganrewrite
    from rewriting_b200.synthetic import seeded_vgg16
    gw = ganrewrite.SeqStyleGanRewriter(copy.deepcopy(model).cuda().eval(),
                                        torch.utils.data.TensorDataset(z40), 8)
    vgg = seeded_vgg16()
    gw.apply_overfit(req, niter=3, lr=1e-4, feature_net=vgg)            # warm-up
    for mode, n in ((False, 50), (True, 50), (True, 500)):
        torch.cuda.synchronize()
        t0 = time.time()
        gw.apply_overfit(req, niter=n, lr=1e-5, feature_net=vgg, use_graph=mode)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print('apply_overfit (%s): %d iterations in %.2f s = %.1f it/s (incl. target rendering, VGG '
              'setup%s)' % ('one CUDA graph per iteration' if mode else 'eager', n, dt, n / dt,
                            ', 3 eager warm-up iterations and the capture' if mode else ''))


if __name__ == '__main__':
    main()
