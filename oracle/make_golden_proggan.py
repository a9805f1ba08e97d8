"""TEST INFRASTRUCTURE ONLY — SURVEY.md §8 f-3: pins oracle/proggan_oracle.py against the live
reference (`utils/proggan.py` + `ProgressiveGanRewriter`, run unmodified through
oracle/ref_shim.py) and digests the two real-data fixtures the reference ships
(notebooks/masks/reflections/progan-kitchen/{r2m.npz, reflection_switched_layer6.npz}) into small
committed goldens.  Authoring container only (~2 min):

    python oracle/make_golden_proggan.py

tests/golden/proggan64.npz   seeded ProgressiveGenerator(resolution=64): pixels, layer-6 key
    second moment C over 40 z, ZCA, d (4 context keys, rank 1), goal crops, edited layer6.conv
    weight after 11 insert iterations (Lambda = delta W . d) and the loss trajectory
tests/golden/proggan_kitchen_layer6.npz   from the shipped kitchen fixtures: the r2m.npz
    statistics (reference cache format: constructor / count / mom2 / sample_size), the rank-one
    edit the paper's reflection example stores (direction d, Lambda, sigma2/sigma1 of
    opt_layer6 - unopt_layer6), 128 output channels of the real trained layer-6 weights, and a
    20-iteration oracle edit of those real weights along the real direction on a seeded crop
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')
KITCHEN = '/root/reference/notebooks/masks/reflections/progan-kitchen'

from oracle import proggan_oracle as ppo       # noqa: E402
from oracle.make_golden import make_request    # noqa: E402
from oracle.ref_shim import load_reference     # noqa: E402

N_Z, LAYER, NITER = 40, 6, 11


def main():
    torch.set_num_threads(os.cpu_count())
    ref = load_reference()
    import utils.proggan as rp                  # the reference's module (sys.path set by the shim)
    model = ppo.seeded_state_dict(lambda: rp.ProgressiveGenerator(resolution=64))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    from rewriting_b200.utils import proggan as myp
    mine = ppo.seeded_state_dict(lambda: myp.ProgressiveGenerator(resolution=64))
    assert list(mine.state_dict()) == list(sd)
    for k in sd:
        assert torch.equal(mine.state_dict()[k], sd[k]), k
    print('seeded init identical (%d tensors)' % len(sd))
    z = ref.zdataset.z_sample_for_model(model, N_Z, seed=1)          # [N, 512, 1, 1]
    assert z.shape == (N_Z, 512, 1, 1)
    with torch.no_grad():
        pix_ref = model(z[:2])
        pix_orc = ppo.generator_forward(sd, z[:2])
    print('pixels ref vs oracle', (pix_ref - pix_orc).abs().max().item(), 'range',
          pix_ref.min().item(), pix_ref.max().item())
    assert (pix_ref - pix_orc).abs().max().item() < 1e-5

    zds = torch.utils.data.TensorDataset(z)
    gw = ref.ganrewrite.ProgressiveGanRewriter(model, zds, LAYER, cachedir=None)
    C = gw.c_matrix.clone()
    with torch.no_grad():
        keys = torch.cat([ppo.generator_forward(sd, z[i:i + 10], upto_key_layer=LAYER)
                          for i in range(0, N_Z, 10)])
        flat = keys.permute(0, 2, 3, 1).reshape(-1, 512)
        C_orc = flat.t() @ flat / flat.shape[0]
    print('C ref vs oracle rel', ((C - C_orc).norm() / C.norm()).item(), 'k_shape', tuple(gw.k_shape))
    request = make_request()
    with torch.no_grad():
        obj_acts, _, obj_area, obj_bounds = gw.object_from_selection(*request['object'])
        goal_in, goal_out, _, paste_bounds = gw.paste_from_selection(
            request['paste'][0], request['paste'][1], obj_acts, obj_area)
        d = gw.multi_key_from_selection(request['key'], rank=1)
    print('crop', tuple(goal_in.shape), tuple(goal_out.shape), obj_bounds, paste_bounds)
    W0 = gw.target_weights().detach().clone()
    losses = []
    gw.insert(goal_in, goal_out, d, niter=NITER, piter=10, lr=0.05,
              update_callback=lambda it, loss: losses.append(float(loss)))
    W = gw.target_weights().detach().clone()
    lo = []
    W_orc = ppo.insert_loop(W0, goal_in, goal_out, d, NITER, piter=10, lr=0.05, record_loss=lo)
    print('edited W: ref vs oracle', (W - W_orc).abs().max().item(), 'max|dW|',
          (W - W0).abs().max().item())
    assert (W - W_orc).abs().max().item() < 1e-5
    lam = torch.einsum('oiyx,i->oyx', W - W0, d[0])
    np.savez_compressed(
        os.path.join(GOLD, 'proggan64.npz'), n_z=N_Z, layer=LAYER, niter=NITER,
        pixels_sub=pix_ref[:, :, ::2, ::2].numpy(), C=C.numpy(), zca_sub=gw.zca_matrix[::8, ::8].numpy(),
        d=d.numpy(), goal_in=goal_in.numpy(), goal_out=goal_out.numpy(),
        obj_bounds=np.array(obj_bounds), paste_bounds=np.array(paste_bounds),
        lam=lam.numpy(), losses=np.array(losses), W_delta_fro=(W - W0).norm().numpy(),
        key_sub=keys[0, ::16].numpy())

    # ---- the shipped real-data fixtures ------------------------------------------------------
    r2m = np.load(os.path.join(KITCHEN, 'r2m.npz'), allow_pickle=True)
    fx = np.load(os.path.join(KITCHEN, 'reflection_switched_layer6.npz'))
    W_un = torch.from_numpy(fx['unopt_layer6'])
    W_op = torch.from_numpy(fx['opt_layer6'])
    dW = (W_op - W_un).permute(0, 2, 3, 1).reshape(-1, 512).double()
    u, s, vh = torch.linalg.svd(dW, full_matrices=False)
    d_fix = vh[0].float()
    lam_fix = torch.einsum('oiyx,i->oyx', W_op - W_un, d_fix)
    print('kitchen: sigma2/sigma1 %.3g, max|dW| %.3g, ||d||=%.6f' % (s[1] / s[0], dW.abs().max(), d_fix.norm()))
    # oracle edit of REAL weights (128 output channels) along the REAL direction, seeded crop
    g = torch.Generator().manual_seed(5)
    W_sub = W_un[:128].contiguous()
    kc = torch.randn(1, 512, 6, 7, generator=g)
    kc = kc / torch.sqrt(torch.mean(kc ** 2, dim=1, keepdim=True) + 1e-8)      # pixel-normed keys
    tgt = torch.nn.functional.conv2d(kc, W_sub, padding=1) * 1.3 + 0.2
    l20 = []
    W20 = ppo.insert_loop(W_sub, kc, tgt, d_fix[None], 20, piter=10, lr=0.05, record_loss=l20)
    lam20 = torch.einsum('oiyx,i->oyx', W20 - W_sub, d_fix)
    np.savez_compressed(
        os.path.join(GOLD, 'proggan_kitchen_layer6.npz'),
        constructor=r2m['constructor'], count=r2m['count'], mom2=r2m['mom2'],
        sample_size=r2m['sample_size'],
        d=d_fix.numpy(), lam=lam_fix.numpy(), sigma_ratio=(s[1] / s[0]).item(),
        dW_fro=dW.norm().item(), W_unopt_sub=W_sub.numpy(), key_crop=kc.numpy(), target=tgt.numpy(),
        lam20=lam20.numpy(), loss20=np.array(l20))
    print('wrote goldens')


if __name__ == '__main__':
    main()
