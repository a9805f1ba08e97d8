"""TEST INFRASTRUCTURE ONLY — the paper's all-weights baseline (`apply_overfit` ->
`all_weights_insert`, reference rewrite/ganrewrite.py:171-181, 300-331) run through the UNMODIFIED
live reference (oracle/ref_shim.py) on the request of tests/golden/edit_request.json, 40 z, layer 8,
for a few Adam iterations over ALL generator parameters.  The pretrained VGG-16 of the perceptual
term cannot be downloaded here: `torchvision.models.vgg16` is replaced, for this run only, by
`rewriting_b200.synthetic.seeded_vgg16()` (same architecture, seeded random weights), which the GPU
test rebuilds identically.  Authoring container only (~3 min):

    python oracle/make_golden_overfit.py

Recorded in tests/golden/overfit3.npz: the loss of every iteration (and of the same run started from
parameters perturbed by 1e-6 relative: the trajectory's own sensitivity), the pasted target crop bounds,
the update (after - before) of a few whole parameter tensors and sum|update| of every parameter,
and the FIRST iteration's gradients (norm of every tensor, whole small tensors, samples of four conv
weights).
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

N_Z = 40
LAYER = 8
NITER = 3
LR = 0.01


def main():
    torch.set_num_threads(os.cpu_count())
    ref = load_reference()
    import torchvision
    from rewriting_b200.synthetic import seeded_vgg16
    stand_in = seeded_vgg16()                      # built with the real constructor, then swapped in
    torchvision.models.vgg16 = lambda *a, **k: stand_in
    ref_model = orc.seeded_state_dict(
        lambda: ref.models.SeqStyleGAN2(256, style_dim=512, n_mlp=8, mconv='seq')).eval()
    z = ref.zdataset.standard_z_sample(N_Z, 512, seed=1)
    zds = torch.utils.data.TensorDataset(z)
    with open(os.path.join(GOLD, 'edit_request.json')) as f:
        request = json.load(f)
    gw = ref.ganrewrite.SeqStyleGanRewriter(ref_model, zds, LAYER, cachedir=None)
    before = {k: v.detach().clone() for k, v in gw.model.named_parameters()}
    losses = []
    grad0 = {}

    def callback(it, loss):
        losses.append(float(loss))
        if it == 0:       # .grad still holds the first iteration's gradient (zeroed at the next one)
            for k, p_ in gw.model.named_parameters():
                grad0[k] = p_.grad.detach().clone()
    gw.apply_overfit(request, niter=NITER, lr=LR, update_callback=callback)
    after = dict(gw.model.named_parameters())
    # how sharp is this trajectory?  The same run from parameters perturbed by 1e-6 relative
    # (the scale of fp32 rounding differences between two implementations of the forward pass)
    pert_model = orc.seeded_state_dict(
        lambda: ref.models.SeqStyleGAN2(256, style_dim=512, n_mlp=8, mconv='seq')).eval()
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for p_ in pert_model.parameters():
            p_.mul_(1 + 1e-6 * torch.randn(p_.shape, generator=gen))
    gw2 = ref.ganrewrite.SeqStyleGanRewriter(pert_model, zds, LAYER, cachedir=None)
    losses_pert = []
    gw2.apply_overfit(request, niter=NITER, lr=LR,
                      update_callback=lambda it, loss: losses_pert.append(float(loss)))
    names = sorted(before)
    upd = {k: (after[k].detach() - before[k]) for k in names}
    sums = np.array([float(upd[k].abs().sum()) for k in names])
    small = [k for k in names if upd[k].numel() <= 4096 and float(upd[k].abs().sum()) > 0][:12]
    out = {'losses': np.array(losses), 'losses_perturbed_1e-6': np.array(losses_pert), 'niter': NITER, 'lr': LR, 'names': np.array(names),
           'abs_update_sums': sums, 'kept': np.array(small)}
    for i, k in enumerate(small):
        out['upd_%d' % i] = upd[k].numpy()
    # first-iteration gradients: Frobenius norm of every tensor, whole small tensors, samples of
    # the big convolution weights (the direct check of the full backward pass)
    out['grad0_norms'] = np.array([float(grad0[k].norm()) for k in names])
    for i, k in enumerate(small):
        out['grad0_%d' % i] = grad0[k].numpy()
    for lname in ('layer3', 'layer8', 'layer13', 'layer14'):
        out['grad0_w_%s' % lname] = grad0['%s.sconv.mconv.dconv.weight' % lname][0, ::37, ::41].numpy()
    # one large tensor, subsampled: the target layer's conv weight
    wkey = 'layer8.sconv.mconv.dconv.weight'
    out['upd_w8_sub'] = upd[wkey][0, ::37, ::41].numpy()
    np.savez_compressed(os.path.join(GOLD, 'overfit3.npz'), **out)
    print('losses', losses, 'from parameters perturbed by 1e-6:', losses_pert)
    print('parameters', len(names), 'updated', int((sums > 0).sum()), 'kept', small)


if __name__ == '__main__':
    main()
