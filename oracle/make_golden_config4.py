"""TEST INFRASTRUCTURE ONLY — BASELINE.json config 4 on its own fixture and horizon: replays
`notebooks/masks/stylegan/horse/hat_on_horse_ears.json` (object img 441, paste img 854, context
keys 354/956/309/926) through the UNMODIFIED live reference (oracle/ref_shim.py) with
zds = 1000 z, layer 8, rank 1, piter 10, lr 0.05, and writes tests/golden/config4_hat.npz plus a
copy of the request (mask data-URLs are inputs).  Authoring container only
(~6 min on 8 cores):

    python oracle/make_golden_config4.py

Recorded (reference ganrewrite.py:135-169, 254-298, 333-374):
  * C over the 1000 z (full matrix, the GPU test feeds it back so that d is compared on the
    identical C), d, the goal_in / goal_out crops and their bounds
  * the edit after 50 iterations, as Lambda50 = (W - W0) . d  ([Cout,3,3]; after the final
    projection W - W0 = Lambda d^T exactly) — the short-horizon 1e-4 check
  * the full 2001 iterations: reference fp32 Lambda and loss trajectory, and the fp64 anchor
    (the oracle's insert loop in float64 from the same W0 / goal / d) — SURVEY.md §7(ii):
    rel-Frobenius(fp32 vs fp64), final losses, sigma2/sigma1 of the reference's delta W.
"""
import json
import os
import shutil
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')
REQUEST = '/root/reference/notebooks/masks/stylegan/horse/hat_on_horse_ears.json'

from oracle import sg2_oracle as orc          # noqa: E402
from oracle.ref_shim import load_reference    # noqa: E402

N_Z = 1000
LAYER = 8


def lam_of(W, W0, d):
    return torch.einsum('goiyx,i->goyx', (W - W0).double(), d[0].double())[0]


def c64_anchor(n=200):
    """tests/golden/c64_200z.npz: E[kk^T] over the first `n` z with the oracle's keys accumulated
    in float64.  The reference accumulates `mom2` in fp32 (runningstats.py:1086-1097): measured
    here, its own rounding error grows to 1.4e-5 rel-Frobenius after 200 z and ~1.5e-4 after the
    1000 z of config 4 — the reference's C is a looser target than the exact statistic."""
    from rewriting_b200.synthetic import seeded_generator
    from rewriting_b200.utils import zdataset
    sd = seeded_generator().state_dict()
    z = zdataset.standard_z_sample(1000, 512, seed=1)
    acc = torch.zeros(512, 512, dtype=torch.float64)
    rows = 0
    for j in range(n // 10):
        with torch.no_grad():
            k = orc.generator_forward(sd, z[10 * j:10 * j + 10], upto_key_layer=8)
        f = k.permute(0, 2, 3, 1).reshape(-1, 512).double()
        acc += f.t() @ f
        rows += f.shape[0]
    np.savez_compressed(os.path.join(GOLD, 'c64_200z.npz'), n_z=n, count=rows,
                        C64=(acc / rows).numpy().astype(np.float32))
    print('wrote c64_200z.npz (%d rows)' % rows)


def main():
    torch.set_num_threads(os.cpu_count())
    if '--c64-only' in sys.argv:
        return c64_anchor()
    ref = load_reference()
    ref_model = orc.seeded_state_dict(
        lambda: ref.models.SeqStyleGAN2(256, style_dim=512, n_mlp=8, mconv='seq')).eval()
    sd = {k: v.clone() for k, v in ref_model.state_dict().items()}
    z = ref.zdataset.standard_z_sample(N_Z, 512, seed=1)
    zds = torch.utils.data.TensorDataset(z)
    with open(REQUEST) as f:
        request = json.load(f)
    t0 = time.time()
    gw = ref.ganrewrite.SeqStyleGanRewriter(ref_model, zds, LAYER, cachedir=None)
    print('rewriter (C over %d z): %.1f s' % (N_Z, time.time() - t0), flush=True)
    C = gw.c_matrix.clone()
    with torch.no_grad():
        obj_acts, _, obj_area, obj_bounds = gw.object_from_selection(*request['object'])
        goal_in, goal_out, _, paste_bounds = gw.paste_from_selection(
            request['paste'][0], request['paste'][1], obj_acts, obj_area)
        d = gw.multi_key_from_selection(request['key'], rank=1)
    print('crop', tuple(goal_in.fmap.shape), tuple(goal_out.fmap.shape), 'bounds', obj_bounds,
          paste_bounds, flush=True)
    W0 = gw.target_weights().detach().clone()

    def run_ref(niter):
        with torch.no_grad():
            gw.target_weights()[...] = W0
        losses = []
        t = time.time()
        gw.insert(goal_in, goal_out, d, niter=niter, piter=10, lr=0.05,
                  update_callback=lambda it, loss: losses.append(float(loss)))
        print('reference insert %d its: %.1f s' % (niter, time.time() - t), flush=True)
        return gw.target_weights().detach().clone(), np.array(losses)

    W50, loss50 = run_ref(50)
    lam50 = lam_of(W50, W0, d)
    res50 = ((W50 - W0).double() - torch.einsum('oyx,i->oiyx', lam50, d[0].double())[None]).abs().max()
    print('50 its: max|dW| %.3g, out-of-span residual %.3g' % ((W50 - W0).abs().max(), res50))

    W2k, loss2k = run_ref(2001)
    lam2k = lam_of(W2k, W0, d)
    dW = (W2k - W0)[0].permute(0, 2, 3, 1).reshape(-1, 512).double()
    sv = torch.linalg.svdvals(dW)
    print('2001 its: max|dW| %.3g sigma2/sigma1 %.3g final loss %.6f' % (
        dW.abs().max(), sv[1] / sv[0], loss2k[-1]), flush=True)

    # fp64 anchor: the oracle's loop (== the reference bit-for-bit in fp32, make_golden.py)
    tp = dict(noise_w=sd['layer8.sconv.noise.weight'], bias=sd['layer8.sconv.activate.bias'])
    l64 = []
    t = time.time()
    W64 = orc.insert_loop(W0.double(), goal_in.fmap.double(), goal_in.style.double(),
                          goal_out.fmap.double(), tp['noise_w'].double(), tp['bias'].double(),
                          d.double(), 2001, piter=10, lr=0.05, record_loss=l64)
    print('oracle fp64 2001 its: %.1f s' % (time.time() - t), flush=True)
    lam64 = torch.einsum('goiyx,i->goyx', W64 - W0.double(), d[0].double())[0]
    rel = ((lam2k - lam64).norm() / lam64.norm()).item()
    print('2001 its: reference fp32 vs fp64 anchor rel-Frobenius %.3g; final loss %.6f vs %.6f' % (
        rel, loss2k[-1], l64[-1]))
    l32 = []
    W32 = orc.insert_loop(W0, goal_in.fmap, goal_in.style, goal_out.fmap, tp['noise_w'],
                          tp['bias'], d, 50, piter=10, lr=0.05, record_loss=l32)
    print('oracle fp32 vs reference after 50 its: max|d| %.3g' % (W32 - W50).abs().max())

    os.makedirs(GOLD, exist_ok=True)
    shutil.copyfile(REQUEST, os.path.join(GOLD, 'hat_on_horse_ears.json'))
    np.savez_compressed(
        os.path.join(GOLD, 'config4_hat.npz'),
        n_z=N_Z, layer=LAYER, C=C.numpy(), d=d.numpy(),
        goal_in_fmap=goal_in.fmap.numpy(), goal_in_style=goal_in.style.numpy(),
        goal_out_fmap=goal_out.fmap.numpy(),
        obj_bounds=np.array(obj_bounds), paste_bounds=np.array(paste_bounds),
        lam50=lam50.float().numpy(), loss50=loss50,
        lam2001_ref32=lam2k.float().numpy(), lam2001_fp64=lam64.float().numpy(),
        loss2001_ref32=loss2k[::10], loss2001_fp64=np.array(l64)[::10],
        final_loss_ref32=loss2k[-1], final_loss_fp64=l64[-1],
        rel_fro_ref32_vs_fp64=rel, sigma_ratio_ref32=(sv[1] / sv[0]).item(),
        max_abs_dW_2001=dW.abs().max().item(),
    )
    print('wrote', os.path.join(GOLD, 'config4_hat.npz'))
    c64_anchor()


if __name__ == '__main__':
    main()
