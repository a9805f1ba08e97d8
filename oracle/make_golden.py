"""TEST INFRASTRUCTURE ONLY — pins the CPU oracle against the live reference and writes the
committed fixtures under tests/golden/.  Runs in the authoring container only (needs the
read-only reference at /root/reference; ~2 min on 8 cores):

    python oracle/make_golden.py

What is pinned (reference executed unmodified through oracle/ref_shim.py, seeded synthetic
weights per SURVEY.md §8d; the masks are synthetic RGBA data-URLs drawn here, same format as
the UI's edit requests):
  * SeqStyleGAN2(256) forward, B=2                     -> pixels (oracle must match bit-exactly)
  * SeqStyleGanRewriter(layer 8) over 40 z             -> C = E[kk^T], ZCA
  * multi_key_from_selection (4 context keys, rank 1)  -> d
  * object/paste selection                             -> goal_in / goal_out crops
  * insert(niter=11, piter=10, lr=0.05)                -> edited W (sampled) + loss trajectory
and that a model built by THIS repo's SeqStyleGAN2 constructor under the same seed has
identical parameters (so the GPU box can rebuild the weights without shipping 120 MB).
"""
import base64
import io
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
NITER = 11


def ellipse_mask_url(cx, cy, rx, ry, size=256):
    from PIL import Image, ImageDraw
    im = Image.new('RGBA', (size, size), (0, 0, 0, 0))
    ImageDraw.Draw(im).ellipse([cx - rx, cy - ry, cx + rx, cy + ry], fill=(255, 255, 255, 255))
    buf = io.BytesIO()
    im.save(buf, format='png')
    return 'data:image/png;base64,' + base64.b64encode(buf.getvalue()).decode('ascii')


def make_request():
    return {
        'object': [3, ellipse_mask_url(120, 96, 30, 26)],
        'paste': [7, ellipse_mask_url(150, 140, 24, 20)],
        'key': [[5, ellipse_mask_url(100, 100, 22, 18)],
                [11, ellipse_mask_url(160, 90, 18, 22)],
                [17, ellipse_mask_url(90, 170, 20, 20)],
                [23, ellipse_mask_url(180, 180, 16, 24)]],
    }


def checksum(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    ref = load_reference()

    ref_model = orc.seeded_state_dict(
        lambda: ref.models.SeqStyleGAN2(256, style_dim=512, n_mlp=8, mconv='seq'))
    ref_model.eval()
    sd = {k: v.clone() for k, v in ref_model.state_dict().items()}

    # this repo's constructor under the same seed must give the same parameters
    from rewriting_b200.utils.stylegan2 import SeqStyleGAN2 as MySeq
    mine = orc.seeded_state_dict(lambda: MySeq(256, style_dim=512, n_mlp=8, mconv='seq'))
    my_sd = mine.state_dict()
    assert list(my_sd.keys()) == list(sd.keys()), 'state_dict keys differ'
    for k in sd:
        assert torch.equal(my_sd[k], sd[k]), 'seeded init differs at ' + k
    print('seeded init: identical (%d tensors)' % len(sd))

    z = ref.zdataset.standard_z_sample(N_Z, 512, seed=1)
    from rewriting_b200.utils import zdataset as myz
    assert torch.equal(myz.standard_z_sample(N_Z, 512, seed=1), z)

    # ---- generator forward --------------------------------------------------------------
    with torch.no_grad():
        pix_ref = ref_model(z[:2])
        rec = {}
        pix_orc = orc.generator_forward(sd, z[:2], record=rec)
    d_pix = (pix_ref - pix_orc).abs().max().item()
    print('pixels: ref vs oracle max|d| = %.3g  (range %.2f..%.2f)' % (
        d_pix, pix_ref.min().item(), pix_ref.max().item()))
    assert d_pix <= 1e-5
    with torch.no_grad():
        pix64 = orc.generator_forward({k: v.double() for k, v in sd.items()}, z[:2].double())
    print('pixels: fp32 vs fp64 oracle max|d| = %.3g' % (pix_ref.double() - pix64).abs().max())

    # ---- rewriter ----------------------------------------------------------------------
    zds = torch.utils.data.TensorDataset(z)
    gw = ref.ganrewrite.SeqStyleGanRewriter(ref_model, zds, LAYER, cachedir=None)
    C_ref = gw.c_matrix.clone()
    with torch.no_grad():
        keys = [orc.generator_forward(sd, z[i:i + 10], upto_key_layer=LAYER)
                for i in range(0, N_Z, 10)]
        mom2, count = orc.second_moment(keys)
    C_orc = mom2 / count
    rel = ((C_ref - C_orc).norm() / C_ref.norm()).item()
    print('C: ref vs oracle rel-Frobenius = %.3g (count %d)' % (rel, count))
    assert rel < 1e-5
    zca_ref = gw.zca_matrix.clone()
    zca_orc = orc.zca_from_cov(C_ref)
    assert (zca_ref - zca_orc).abs().max().item() < 1e-4 * zca_ref.abs().max().item()

    request = make_request()
    with open(os.path.join(GOLD, 'edit_request.json'), 'w') as f:
        json.dump(request, f)

    with torch.no_grad():
        obj_acts, _, obj_area, obj_bounds = gw.object_from_selection(*request['object'])
        goal_in, goal_out, _, paste_bounds = gw.paste_from_selection(
            request['paste'][0], request['paste'][1], obj_acts, obj_area)
        d_ref = gw.multi_key_from_selection(request['key'], rank=1)
    # oracle direction
    obs, wts = [], []
    for imgnum, mask in request['key']:
        with torch.no_grad():
            k = orc.generator_forward(sd, z[imgnum][None], upto_key_layer=LAYER)
        obs.append(k.permute(0, 2, 3, 1).reshape(-1, 512))
        area = ref.renormalize.from_url(mask, target='pt', size=(32, 32))[0]
        wts.append(area.view(-1)[:, None])
    d_orc = orc.multi_key_zca(obs, wts, zca_ref, rank=1)
    dd = (d_ref - d_orc).abs().max().item()
    print('d: ref vs oracle max|d| = %.3g' % dd)
    assert dd < 1e-4

    W0 = gw.target_weights().detach().clone()
    losses_ref = []
    gw.insert(goal_in, goal_out, d_ref, niter=NITER, piter=10, lr=0.05,
              update_callback=lambda it, loss: losses_ref.append(float(loss)))
    W_ref = gw.target_weights().detach().clone()

    tp = dict(noise_w=sd['layer8.sconv.noise.weight'], bias=sd['layer8.sconv.activate.bias'])
    losses_orc = []
    W_orc = orc.insert_loop(W0, goal_in.fmap, goal_in.style, goal_out.fmap, tp['noise_w'],
                            tp['bias'], d_ref, NITER, piter=10, lr=0.05, record_loss=losses_orc)
    dW = (W_ref - W_orc).abs().max().item()
    print('edited W (11 its): ref vs oracle max|d| = %.3g ; max|W-W0| = %.3g' % (
        dW, (W_ref - W0).abs().max().item()))
    assert dW < 1e-4
    print('loss ref', losses_ref[:3], '... oracle', losses_orc[:3])

    # ---- write fixtures ------------------------------------------------------------------
    np.savez_compressed(
        os.path.join(GOLD, 'sg2_layer8.npz'),
        n_z=N_Z, layer=LAYER, niter=NITER,
        pixels_sub=pix_ref[:, :, ::8, ::8].numpy(),
        pixels_absmax=pix_ref.abs().max().numpy(),
        l8_key_sub=rec['layer8']['k'][0, ::16, ::4, ::4].numpy(),
        l8_out_sub=rec['layer8']['y'][0, ::16, ::4, ::4].numpy(),
        l9_out_sub=rec['layer9']['y'][0, ::16, ::8, ::8].numpy(),
        C_sub=C_ref[::8, ::8].numpy(), C_trace=C_ref.trace().numpy(), C_fro=C_ref.norm().numpy(),
        C_diag=C_ref.diag().numpy(), count=count,
        zca_sub=zca_ref[::8, ::8].numpy(),
        d=d_ref.numpy(),
        goal_in_fmap=goal_in.fmap.numpy(), goal_in_style=goal_in.style.numpy(),
        goal_out_fmap=goal_out.fmap.numpy(),
        obj_bounds=np.array(obj_bounds), paste_bounds=np.array(paste_bounds),
        W_delta_sub=(W_ref - W0)[0, ::37, ::41].numpy(),
        W_delta_fro=(W_ref - W0).norm().numpy(),
        losses=np.array(losses_ref, dtype=np.float64),
    )
    with open(os.path.join(GOLD, 'weights_checksum.json'), 'w') as f:
        json.dump(checksum(sd), f)
    print('wrote fixtures to', GOLD)


if __name__ == '__main__':
    main()
