"""GPU (B200): the CUDA path against the CPU oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star): pixels within 1e-3 fp32 on identical z/seed, edited W
within 1e-4 (over <= 50 iterations), C rel-Frobenius <= 1e-5, d max-abs <= 1e-4.
Conv operands are 3-term split bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulate).
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sg2_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cuda_model(seeded_model):
    import copy
    return copy.deepcopy(seeded_model).cuda().eval()


# ------------------------------------------------------------------------------------------
# kernel level
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,C,H,W', [(2, 64, 4, 4), (3, 128, 5, 7), (1, 512, 32, 32), (2, 128, 40, 33)])
def test_prep_keys_layout_and_split(B, C, H, W):
    from rewriting_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device='cuda')
    s = torch.randn(B, C, device='cuda')
    planes, k = ops.prep_keys(x, s, want_k=True)
    k_ref = s[:, :, None, None] * x
    assert torch.equal(k, k_ref)                                  # fp32 product, same rounding
    hi = planes.hi.float().view(B, H + 1, W + 1, C)
    lo = planes.lo.float().view(B, H + 1, W + 1, C)
    assert hi[:, H].abs().max() == 0 and hi[:, :, W].abs().max() == 0    # zero pad row / col
    assert lo[:, H].abs().max() == 0 and lo[:, :, W].abs().max() == 0
    rec = (hi + lo)[:, :H, :W].permute(0, 3, 1, 2)
    err = (rec - k_ref).abs().max().item()
    assert err <= 2 ** -16 * k_ref.abs().max().item()
    assert torch.equal(planes.hi.view(B, H + 1, W + 1, C)[:, :H, :W].permute(0, 3, 1, 2),
                       k_ref.to(torch.bfloat16))


@pytest.mark.parametrize('B,Cin,Cout,H,W,act', [
    (2, 128, 128, 8, 8, True), (3, 64, 256, 5, 9, False), (1, 512, 512, 32, 32, True),
    (8, 512, 512, 4, 4, True), (2, 256, 128, 19, 33, True)])
def test_styled_conv_forward_vs_oracle(B, Cin, Cout, H, W, act):
    from rewriting_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(B, Cin, H, W)
    style = torch.randn(B, Cin) * 0.5 + 1.0
    weight = torch.randn(1, Cout, Cin, 3, 3)
    nw = torch.tensor([0.37])
    bias = torch.randn(Cout)
    k = style[:, :, None, None] * x
    if act:
        ref = orc.target_forward(k, style, weight, nw, bias, True)
    else:
        ref = orc.demod_conv(k, style, weight, upsample=False)
    wp = torch.nn.Parameter(weight.cuda())
    with torch.no_grad():
        y = ops.styled_conv(x.cuda(), style.cuda(), wp, torch.nn.Parameter(nw.cuda()),
                            torch.nn.Parameter(bias.cuda()), upsample=False, demodulate=True,
                            with_noise=act, with_act=act)
    err = (y.cpu() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize('B,Cin,Cout,H,W', [(2, 128, 128, 4, 4), (1, 512, 256, 16, 16),
                                             (3, 64, 128, 7, 5)])
def test_styled_conv_upsample_vs_oracle(B, Cin, Cout, H, W):
    from rewriting_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(B, Cin, H, W)
    style = torch.randn(B, Cin) * 0.5 + 1.0
    weight = torch.randn(1, Cout, Cin, 3, 3)
    nw, bias = torch.tensor([0.37]), torch.randn(Cout)
    k = style[:, :, None, None] * x
    t = orc.demod_conv(k, style, weight, upsample=True)
    kern = orc.make_kernel([1, 3, 3, 1]) * 4
    tb = orc.upfirdn2d(t, kern, pad=(1, 1))
    n = orc.noise_table(B, 4 * H * W).view(B, 1, 2 * H, 2 * W)
    ref = orc.fused_leaky_relu(tb + nw * n, bias)
    with torch.no_grad():
        y = ops.styled_conv(x.cuda(), style.cuda(), torch.nn.Parameter(weight.cuda()),
                            torch.nn.Parameter(nw.cuda()), torch.nn.Parameter(bias.cuda()),
                            upsample=True, blur_kernel=kern.cuda(), demodulate=True)
    assert y.shape == ref.shape
    err = (y.cpu() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


def test_styled_conv_backward_vs_oracle_autograd():
    from rewriting_b200 import ops
    torch.manual_seed(3)
    B, Cin, Cout, H, W = 2, 128, 128, 6, 7
    x = torch.randn(B, Cin, H, W, requires_grad=True)
    style = (torch.randn(B, Cin) * 0.5 + 1.0).requires_grad_(True)
    weight = torch.randn(1, Cout, Cin, 3, 3, requires_grad=True)
    nw = torch.tensor([0.37], requires_grad=True)
    bias = torch.randn(Cout, requires_grad=True)
    gy = torch.randn(B, Cout, H, W)
    ref = orc.target_forward(style[:, :, None, None] * x, style, weight, nw, bias, True)
    ref.backward(gy)
    xc = x.detach().cuda().requires_grad_(True)
    sc = style.detach().cuda().requires_grad_(True)
    wc = torch.nn.Parameter(weight.detach().cuda())
    nc = torch.nn.Parameter(nw.detach().cuda())
    bc = torch.nn.Parameter(bias.detach().cuda())
    y = ops.styled_conv(xc, sc, wc, nc, bc, upsample=False, demodulate=True)
    y.backward(gy.cuda())
    for name, got, want in [('x', xc.grad, x.grad), ('style', sc.grad, style.grad),
                            ('weight', wc.grad, weight.grad), ('noise_w', nc.grad, nw.grad),
                            ('bias', bc.grad, bias.grad)]:
        err = (got.cpu() - want).abs().max().item()
        assert err < 3e-4 * max(1.0, want.abs().max().item()), (name, err)


def test_styled_conv_upsample_backward_vs_oracle_autograd():
    from rewriting_b200 import ops
    torch.manual_seed(7)
    B, Cin, Cout, H, W = 2, 128, 128, 5, 6
    x = torch.randn(B, Cin, H, W, requires_grad=True)
    style = (torch.randn(B, Cin) * 0.5 + 1.0).requires_grad_(True)
    weight = torch.randn(1, Cout, Cin, 3, 3, requires_grad=True)
    nw = torch.tensor([0.37], requires_grad=True)
    bias = torch.randn(Cout, requires_grad=True)
    kern = orc.make_kernel([1, 3, 3, 1]) * 4
    gy = torch.randn(B, Cout, 2 * H, 2 * W)
    t = orc.demod_conv(style[:, :, None, None] * x, style, weight, upsample=True)
    tb = orc.upfirdn2d(t, kern, pad=(1, 1))
    n = orc.noise_table(B, 4 * H * W).view(B, 1, 2 * H, 2 * W)
    ref = orc.fused_leaky_relu(tb + nw * n, bias)
    ref.backward(gy)
    xc = x.detach().cuda().requires_grad_(True)
    sc = style.detach().cuda().requires_grad_(True)
    wc = torch.nn.Parameter(weight.detach().cuda())
    nc = torch.nn.Parameter(nw.detach().cuda())
    bc = torch.nn.Parameter(bias.detach().cuda())
    y = ops.styled_conv(xc, sc, wc, nc, bc, upsample=True, blur_kernel=kern.cuda(), demodulate=True)
    assert (y.detach().cpu() - ref.detach()).abs().max().item() < 2e-4 * ref.abs().max().item()
    y.backward(gy.cuda())
    for name, got, want in [('x', xc.grad, x.grad), ('style', sc.grad, style.grad),
                            ('weight', wc.grad, weight.grad), ('noise_w', nc.grad, nw.grad),
                            ('bias', bc.grad, bias.grad)]:
        err = (got.cpu() - want).abs().max().item()
        assert err < 3e-4 * max(1.0, want.abs().max().item()), (name, err)


def test_operator_level_ops_vs_oracle():
    from rewriting_b200.utils.stylegan2 import op
    torch.manual_seed(4)
    x = torch.randn(3, 16, 9, 11)
    b = torch.randn(16)
    y = op.fused_leaky_relu(x.cuda(), b.cuda())
    assert torch.allclose(y.cpu(), orc.fused_leaky_relu(x, b), atol=1e-6)
    lin = torch.randn(5, 16)
    assert torch.allclose(op.fused_leaky_relu(lin.cuda(), b.cuda()).cpu(),
                          orc.fused_leaky_relu(lin, b), atol=1e-6)
    # backward gates on the saved output
    xc = x.cuda().requires_grad_(True)
    bc = b.cuda().requires_grad_(True)
    op.fused_leaky_relu(xc, bc).sum().backward()
    xr = x.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    orc.fused_leaky_relu(xr, br).sum().backward()
    assert torch.allclose(xc.grad.cpu(), xr.grad, atol=1e-6)
    assert torch.allclose(bc.grad.cpu(), br.grad, atol=1e-4)
    k = orc.make_kernel([1, 3, 3, 1])
    for up, down, pad in [(2, 1, (2, 1)), (1, 1, (1, 1)), (1, 2, (1, 1)), (1, 1, (2, 2)),
                          (2, 1, (-1, 2))]:
        kk = k * (up * up)
        ref = orc.upfirdn2d(x, kk, up=up, down=down, pad=pad)
        got = op.upfirdn2d(x.cuda(), kk.cuda(), up=up, down=down, pad=pad)
        assert got.shape == ref.shape, (up, down, pad)
        assert torch.allclose(got.cpu(), ref, atol=1e-5), (up, down, pad)
    xg = x.cuda().requires_grad_(True)
    op.upfirdn2d(xg, (k * 4).cuda(), up=2, pad=(2, 1)).pow(2).sum().backward()
    xr = x.clone().requires_grad_(True)
    orc.upfirdn2d(xr, k * 4, up=2, pad=(2, 1)).pow(2).sum().backward()
    assert torch.allclose(xg.grad.cpu(), xr.grad, atol=1e-4)
    with pytest.raises(RuntimeError):
        op.fused_leaky_relu(x, b)            # CPU tensors are rejected like the reference


def test_second_moment_kernel_vs_oracle():
    from rewriting_b200.utils import runningstats
    torch.manual_seed(5)
    r = runningstats.RunningSecondMoment()
    mom_ref = torch.zeros(512, 512, dtype=torch.float64)
    total = 0
    for n in (10240, 37, 64, 5000):                 # ragged batch sizes
        a = torch.randn(n, 512) * torch.linspace(0.1, 3, 512)
        r.add(a.cuda())
        mom_ref += a.double().t() @ a.double()
        total += n
    r.add(torch.zeros(0, 512).cuda())               # empty batch
    assert r.count == total
    got = r.mom2.double().cpu()
    rel = ((got - mom_ref).norm() / mom_ref.norm()).item()
    assert rel < 1e-5, rel
    assert torch.equal(r.mom2, r.mom2.t())           # exactly symmetric (mirrored upper triangle)
    # linearity: accumulating a twice equals 2x
    r2 = runningstats.RunningSecondMoment()
    a = torch.randn(4096, 128).cuda()
    r2.add(a)
    once = r2.mom2.clone()
    r2.add(a)
    assert torch.allclose(r2.mom2, 2 * once, rtol=1e-6, atol=1e-3)


def test_projected_conv_vs_oracle():
    from rewriting_b200.rewrite import ganrewrite
    torch.manual_seed(6)
    W = torch.randn(1, 512, 512, 3, 3)
    q, _ = torch.linalg.qr(torch.randn(512, 3))
    d = q.t().contiguous()
    ref = orc.projected_conv(W, d)
    got = ganrewrite.projected_conv(W.cuda(), d.cuda())
    assert torch.allclose(got.cpu(), ref, atol=2e-5)
    ortho = ganrewrite.projected_conv(W.cuda(), d.cuda(), base=W.cuda(), sign=-1.0)
    assert torch.allclose(ortho.cpu(), W - ref, atol=2e-5)
    assert torch.allclose(ganrewrite.projected_conv(W[0].cuda(), d.cuda()).cpu(), ref[0], atol=2e-5)


# ------------------------------------------------------------------------------------------
# model level
# ------------------------------------------------------------------------------------------
def test_generator_pixels_vs_golden_and_oracle(cuda_model, seeded_sd, z40, golden):
    with torch.no_grad():
        pix = cuda_model(z40[:2].cuda()).cpu()
    assert pix.shape == (2, 3, 256, 256)
    err_g = np.abs(pix[:, :, ::8, ::8].numpy() - golden['pixels_sub']).max()
    assert err_g < 1e-3, err_g                       # north_star: pixels within 1e-3
    with torch.no_grad():
        ref = orc.generator_forward(seeded_sd, z40[:2])
    err = (pix - ref).abs().max().item()
    print('PIXEL_ERR max|d| = %.3e (pixel absmax %.1f)' % (err, ref.abs().max().item()))
    assert err < 1e-3, err
    # batch-size independence: noise row i depends only on (i, H*W), so image 0 of a batch of 2
    # equals the singleton batch
    with torch.no_grad():
        one = cuda_model(z40[0:1].cuda()).cpu()
    assert (one[0] - pix[0]).abs().max().item() < 1e-3


def test_generation_fast_path_equals_layer_path(cuda_model, z40):
    """model(z) (fused producers, no fp32 feature maps) == child-by-child execution, and the
    fast path's key planes == planes of the context model's key tensor."""
    from rewriting_b200 import fastpath, ops
    from rewriting_b200.utils import nethook
    z = z40[:3].cuda()
    with torch.no_grad():
        assert fastpath.eligible(cuda_model, z)
        fast = cuda_model(z)
        slow = torch.nn.Sequential.forward(cuda_model, z)
    assert fast.shape == slow.shape == (3, 3, 256, 256)
    # (the fast path computes all modulation linears in one kernel with its own summation
    #  order, so the two paths agree to fp32 round-off propagated through 14 layers)
    assert (fast - slow).abs().max().item() < 5e-4
    for layer in (8, 9, 4):
        ctx = nethook.subsequence(cuda_model, upto_layer='layer%d.sconv.mconv.dconv' % layer,
                                  share_weights=True)
        with torch.no_grad():
            kp = fastpath.forward(cuda_model, z, upto_key_layer=layer)
            ref_planes, _ = ops.prep_keys(ctx(z).fmap, None)
        assert (kp.B, kp.C, kp.H, kp.W) == (ref_planes.B, ref_planes.C, ref_planes.H, ref_planes.W)
        a = kp.hi.float() + kp.lo.float()
        b = ref_planes.hi.float() + ref_planes.lo.float()
        assert (a - b).abs().max().item() < 2e-4 * max(1.0, b.abs().max().item())
    # not eligible with autograd on or when hooked -> falls back transparently
    zz = z.clone().requires_grad_(True)
    assert not fastpath.eligible(cuda_model, zz)
    with nethook.InstrumentedModel(cuda_model) as inst, torch.no_grad():
        inst.retain_layer('layer4', detach=False)
        assert not fastpath.eligible(cuda_model, z)
        hooked = inst(z)
    assert (hooked - slow).abs().max().item() < 1e-4


def test_fused_layers_equal_leaf_by_leaf_execution(cuda_model, z40):
    """The nethook-split execution (context | target | rendering, leaves one by one) must give
    the same image as the fused whole-layer path."""
    from rewriting_b200.utils import nethook
    first, last = 'layer8.sconv.mconv.dconv', 'layer8.sconv.activate'
    ctx = nethook.subsequence(cuda_model, upto_layer=first, share_weights=True)
    tgt = nethook.subsequence(cuda_model, first_layer=first, last_layer=last, share_weights=True)
    rnd = nethook.subsequence(cuda_model, after_layer=last, share_weights=True)
    z = z40[:3].cuda()
    with torch.no_grad():
        whole = torch.nn.Sequential.forward(cuda_model, z)      # layer path (not the fast path)
        split = rnd(tgt(ctx(z)))
    assert (whole - split).abs().max().item() < 2e-4
    # hooks force the per-child path and still see the layer output
    with nethook.InstrumentedModel(cuda_model) as inst, torch.no_grad():
        inst.retain_layer('layer8.sconv.mconv.adain', detach=False)
        hooked = inst(z)
        key = inst.retained_layer('layer8.sconv.mconv.adain')
    assert key.fmap.shape == (3, 512, 32, 32)
    assert (hooked - whole).abs().max().item() < 2e-4
    assert torch.allclose(key.fmap, ctx(z).fmap, atol=1e-5)


def test_rewriter_statistics_direction_and_edit_vs_golden(cuda_model, z40, golden, edit_request,
                                                          seeded_sd):
    from rewriting_b200.rewrite import ganrewrite
    zds = torch.utils.data.TensorDataset(z40)
    gw = ganrewrite.SeqStyleGanRewriter(cuda_model, zds, 8)
    assert tuple(gw.k_shape) == (1, 512, 32, 32) and tuple(gw.v_shape) == (1, 512, 32, 32)
    C = gw.c_matrix.cpu()
    sub_err = np.abs(C[::8, ::8].numpy() - golden['C_sub']).max()
    assert sub_err < 2e-5 * float(golden['C_diag'].max()), sub_err
    assert abs(float(C.trace()) - float(golden['C_trace'])) < 1e-5 * float(golden['C_trace'])
    # keys come from 7 tensor-core conv layers (~5e-6 relative each): per-entry bound 2e-4,
    # the accumulator itself is held to rel-Frobenius 1e-5 in test_second_moment_kernel_vs_oracle
    np.testing.assert_allclose(C.diag().numpy(), golden['C_diag'], rtol=2e-4)
    # direction.  With only 40 z the matrix C is ill-conditioned: the reference's own fp32 and
    # fp64 pipelines differ by 3.8e-4 max-abs on this unit vector, so the
    # end-to-end d is compared as a direction, and the 1e-4 bound is applied to the key algebra
    # itself: the GPU result against the oracle evaluated on the SAME C and the same keys.
    d = gw.multi_key_from_selection(edit_request['key'], rank=1).cpu()
    d_gold = torch.from_numpy(golden['d'])
    assert float((d[0] * d_gold[0]).sum()) > 1 - 1e-5
    assert (d - d_gold).abs().max().item() < 2e-3
    from rewriting_b200.utils import renormalize
    zca_cpu = orc.zca_from_cov(C)
    obs, wts = [], []
    for imgnum, mask in edit_request['key']:
        with torch.no_grad():
            k = gw.context_model(gw.get_z(imgnum)).fmap.cpu()
        obs.append(k.permute(0, 2, 3, 1).reshape(-1, 512))
        wts.append(renormalize.from_url(mask, target='pt', size=(32, 32))[0].view(-1)[:, None])
    d_same_c = orc.multi_key_zca(obs, wts, zca_cpu, rank=1)
    assert (d - d_same_c).abs().max().item() < 1e-4
    # goal crops
    obj_acts, _, obj_area, ob = gw.object_from_selection(*edit_request['object'])
    goal_in, goal_out, _, pb = gw.paste_from_selection(edit_request['paste'][0],
                                                       edit_request['paste'][1], obj_acts, obj_area)
    assert tuple(ob) == tuple(golden['obj_bounds']) and tuple(pb) == tuple(golden['paste_bounds'])
    # keys / values after 7-8 chained tensor-core layers: same 1e-3 bound as the pixels
    assert (goal_in.fmap.cpu() - torch.from_numpy(golden['goal_in_fmap'])).abs().max() < 1e-3
    assert (goal_out.fmap.cpu() - torch.from_numpy(golden['goal_out_fmap'])).abs().max() < 1e-3
    # the edit: identical state and direction as the reference run, 11 iterations
    gin = type(goal_in)(goal_in, fmap=torch.from_numpy(golden['goal_in_fmap']).cuda(),
                        style=torch.from_numpy(golden['goal_in_style']).cuda())
    gout = type(goal_out)(goal_out, fmap=torch.from_numpy(golden['goal_out_fmap']).cuda())
    W0 = gw.target_weights().detach().clone()
    losses = []
    gw.insert(gin, gout, torch.from_numpy(golden['d']).cuda(), niter=int(golden['niter']),
              piter=10, lr=0.05, update_callback=lambda it, loss: losses.append(float(loss)))
    W = gw.target_weights().detach()
    delta = (W - W0).cpu()
    err = np.abs(delta[0, ::37, ::41].numpy() - golden['W_delta_sub']).max()
    assert err < 1e-4, err                                    # edited W within 1e-4
    assert abs(float(delta.norm()) - float(golden['W_delta_fro'])) < 1e-3 * float(golden['W_delta_fro'])
    np.testing.assert_allclose(np.array(losses), golden['losses'], rtol=2e-4)
    s = torch.linalg.svdvals(delta[0].permute(0, 2, 3, 1).reshape(-1, 512).double())
    assert float(s[1] / s[0]) < 1e-5                           # rank one, as the paper requires
    # the edit is visible to the full model (shared parameters) and to a re-render
    assert gw.model.layer8.sconv.mconv.dconv.weight is gw.target_weights()
    with torch.no_grad():
        img = gw.sample_image_from_latent(z40[7:8].cuda())
    assert torch.isfinite(img).all()


def test_fused_insert_equals_autograd_insert_and_oracle(cuda_model, z40, golden):
    """Same state, same d: fused one-kernel loop == autograd loop on the conv kernels == CPU
    oracle, over 30 iterations (short horizon)."""
    from rewriting_b200.rewrite import ganrewrite
    zds = torch.utils.data.TensorDataset(z40[:10])
    results = {}
    for mode in ('fused', 'autograd'):
        gw = ganrewrite.SeqStyleGanRewriter(cuda_model, zds, 8, fused_insert=(mode == 'fused'))
        bag = gw.context_model(gw.get_z(0))
        gin = type(bag)(bag, fmap=torch.from_numpy(golden['goal_in_fmap']).cuda(),
                        style=torch.from_numpy(golden['goal_in_style']).cuda())
        gout = type(bag)(bag, fmap=torch.from_numpy(golden['goal_out_fmap']).cuda())
        W0 = gw.target_weights().detach().clone().cpu()
        gw.insert(gin, gout, torch.from_numpy(golden['d']).cuda(), niter=30, piter=10, lr=0.05)
        results[mode] = gw.target_weights().detach().cpu()
    sd = cuda_model.state_dict()
    W_orc = orc.insert_loop(W0, torch.from_numpy(golden['goal_in_fmap']),
                            torch.from_numpy(golden['goal_in_style']),
                            torch.from_numpy(golden['goal_out_fmap']),
                            sd['layer8.sconv.noise.weight'].cpu(),
                            sd['layer8.sconv.activate.bias'].cpu(),
                            torch.from_numpy(golden['d']), 30, piter=10, lr=0.05)
    assert (results['fused'] - W_orc).abs().max().item() < 1e-4
    assert (W_orc - W0).abs().max().item() > 0.05             # the loop really moved W
    # The autograd path computes dW on the tensor cores (3-term split bf16, ~1e-6 relative).
    # Adam's first steps move every weight by lr*sign(dW): the handful of the 2.4 M entries whose
    # gradient cancels to below that error can flip sign, i.e. differ by O(lr) — the same
    # sensitivity the reference shows between its own fp32 and fp64 runs.  So
    # this path is held to a distributional bound, the fused path to the 1e-4 max-abs bound.
    diff = (results['autograd'] - W_orc).abs()
    assert (diff > 1e-3).float().mean().item() < 5e-2
    rel = ((results['autograd'] - W_orc).norm() / (W_orc - W0).norm()).item()
    assert rel < 5e-2, rel


# ------------------------------------------------------------------------------------------
# edit variants on the same kernels
# ------------------------------------------------------------------------------------------
def _goal_bags(gw, golden):
    bag = gw.context_model(gw.get_z(0))
    gin = type(bag)(bag, fmap=torch.from_numpy(golden['goal_in_fmap']).cuda(),
                    style=torch.from_numpy(golden['goal_in_style']).cuda())
    gout = type(bag)(bag, fmap=torch.from_numpy(golden['goal_out_fmap']).cuda())
    return gin, gout


def test_insert_variants_rank2_gradient_projection_and_tiny_target(cuda_model, z40, golden):
    from rewriting_b200.rewrite import ganrewrite
    zds = torch.utils.data.TensorDataset(z40[:10])
    sd = cuda_model.state_dict()
    nw = sd['layer8.sconv.noise.weight'].cpu()
    bias = sd['layer8.sconv.activate.bias'].cpu()
    torch.manual_seed(11)
    q, _ = torch.linalg.qr(torch.randn(512, 2))
    d2 = q.t().contiguous()
    k = torch.from_numpy(golden['goal_in_fmap'])
    st = torch.from_numpy(golden['goal_in_style'])
    tgt = torch.from_numpy(golden['goal_out_fmap'])
    # rank 2, with and without gradient projection
    for lrg in (False, True):
        gw = ganrewrite.SeqStyleGanRewriter(cuda_model, zds, 8, low_rank_gradient=lrg)
        gin, gout = _goal_bags(gw, golden)
        W0 = gw.target_weights().detach().clone().cpu()
        gw.insert(gin, gout, d2.cuda(), niter=12, piter=5, lr=0.05)
        W = gw.target_weights().detach().cpu()
        W_orc = orc.insert_loop(W0, k, st, tgt, nw, bias, d2, 12, piter=5, lr=0.05,
                                low_rank_gradient=lrg)
        assert (W - W_orc).abs().max().item() < 1e-4, lrg
        s = torch.linalg.svdvals((W - W0)[0].permute(0, 2, 3, 1).reshape(-1, 512).double())
        assert float(s[2] / s[0]) < 1e-5                      # rank <= 2
    # SeqTiny: the target model is the dconv leaf alone (no noise / activation)
    gw = ganrewrite.SeqTinyStyleGanRewriter(cuda_model, zds, 8)
    gin, gout = _goal_bags(gw, golden)
    assert gw._fused_plan(gin, gout, d2.cuda()) is not None
    W0 = gw.target_weights().detach().clone().cpu()
    gw.insert(gin, gout, d2[:1].cuda(), niter=12, piter=5, lr=0.05)
    W_orc = orc.insert_loop(W0, k, st, tgt, nw, bias, d2[:1], 12, piter=5, lr=0.05,
                            with_noise_act=False)
    assert (gw.target_weights().detach().cpu() - W_orc).abs().max().item() < 1e-4


def test_zero_linear_insert_and_erase_run_on_the_kernels(cuda_model, z40, golden, edit_request):
    from rewriting_b200.rewrite import ganrewrite
    zds = torch.utils.data.TensorDataset(z40)
    d = torch.from_numpy(golden['d']).cuda()
    # zero(): the component of W along d becomes `amount` times that of an all-ones weight
    gw = ganrewrite.SeqStyleGanRewriter(cuda_model, zds, 8)
    W0 = gw.target_weights().detach().clone()
    gw.zero(d, amount=0.25)
    W = gw.target_weights().detach()
    want = W0 - orc.projected_conv(W0.cpu(), d.cpu()).cuda() + \
        0.25 * orc.projected_conv(torch.ones_like(W0).cpu(), d.cpu()).cuda()
    assert (W - want).abs().max().item() < 2e-5
    # linear_insert: optimises Lambda; the edit stays rank one and the loss goes down
    gw = ganrewrite.SeqStyleGanRewriter(cuda_model, zds, 8, use_linear_insert=True)
    gin, gout = _goal_bags(gw, golden)
    W0 = gw.target_weights().detach().clone()
    losses = []
    gw.insert(gin, gout, d, niter=15, lr=0.05, update_callback=lambda it, l: losses.append(float(l)))
    dW = (gw.target_weights().detach() - W0)[0].permute(0, 2, 3, 1).reshape(-1, 512).double().cpu()
    s = torch.linalg.svdvals(dW)
    assert float(s[1] / s[0]) < 1e-5 and float(s[0]) > 0
    assert losses[-1] < losses[0]
    assert isinstance(gw.target_weights(), torch.nn.Parameter)          # parameter restored
    # apply_erase: end to end on the request (normdissect units + insert); weights must move
    gw = ganrewrite.SeqStyleGanRewriter(cuda_model, zds, 8)
    W0 = gw.target_weights().detach().clone()
    gw.apply_erase(dict(paste=edit_request['paste'], key=edit_request['key']), rank=1, drank=30,
                   niter=11, piter=10)
    moved = (gw.target_weights().detach() - W0).abs().max().item()
    assert 1e-3 < moved < 1.0 and torch.isfinite(gw.target_weights()).all()


def test_bulk_sampling_matches_reference_seed_rule(cuda_model, seeded_sd):
    """config 5: batch j is generated from zdataset seed batch*j (utils/get_samples.py:121-124)."""
    from rewriting_b200 import sampling
    from rewriting_b200.utils import zdataset
    imgs, idx = sampling.get_samples(cuda_model, nimgs=4, batch=2)       # 4//2+1 = 3 batches
    assert idx == [0, 1, 2] and imgs.shape == (6, 3, 256, 256) and imgs.dtype == torch.float32
    z1 = zdataset.standard_z_sample(2, 512, seed=2)
    assert torch.equal(sampling.z_for_batch(1, 2), z1)
    with torch.no_grad():
        ref = orc.generator_forward(seeded_sd, z1)
    assert (imgs[2:4] - ref).abs().max().item() < 1e-3
    # uint8 NHWC straight from the last ToRGB combine: exactly clamp(x*127.5+127.5).byte() of
    # this path's own fp32 image ...
    u8, _ = sampling.get_samples(cuda_model, nimgs=2, batch=2, out_dtype=torch.uint8,
                                 reference_count=False)
    want = (imgs[0:2] * 127.5 + 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
    assert u8.shape == (2, 256, 256, 3) and u8.dtype == torch.uint8
    assert torch.equal(u8, want)
    # ... and of the oracle's image up to the 1e-3 pixel tolerance (a byte flips where the
    # fp32 value sits on an integer boundary)
    with torch.no_grad():
        ref0 = orc.generator_forward(seeded_sd, sampling.z_for_batch(0, 2))
    ref_u8 = (ref0 * 127.5 + 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
    d = (u8.int() - ref_u8.int()).abs()
    assert d.max().item() <= 1 and (d > 0).float().mean().item() < 0.02
    # several reference batches per replay (noise rows repeat with the reference batch size)
    # give the same images as one reference batch per replay
    g1, _ = sampling.get_samples(cuda_model, nimgs=4, batch=2, group=1)
    assert torch.equal(g1, imgs)


def test_sample_loop_batch1_seed_imgnum_and_writer(cuda_model, seeded_sd, tmp_path):
    """metrics/sample.py:19-37: image n = G(z_sample(1, seed=n+offset)) run as a batch of one;
    32 per replay here with a period-1 noise table.  Writer: PNG per image / one npz."""
    import numpy as np
    from PIL import Image
    from rewriting_b200 import sampling
    from rewriting_b200.utils import zdataset
    nums = [0, 1, 2, 5, 7]
    u8, mine = sampling.sample_images(cuda_model, nums, offset=1000007, group=3)
    assert mine == nums and u8.shape == (5, 256, 256, 3) and u8.dtype == torch.uint8
    for i in (1, 4):
        z = zdataset.standard_z_sample(1, 512, seed=nums[i] + 1000007)
        with torch.no_grad():
            ref = orc.generator_forward(seeded_sd, z)
        ref_u8 = (ref * 127.5 + 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)[0]
        d = (u8[i].int() - ref_u8.int()).abs()
        assert d.max().item() <= 1 and (d > 0).float().mean().item() < 0.02, i
    w = sampling.ImageWriter(str(tmp_path / 'png'), fmt='png', workers=2)
    w.add(u8, mine)
    w.join()
    back = np.array(Image.open(str(tmp_path / 'png' / '5.png')))
    assert np.array_equal(back, u8[3].numpy())
    w = sampling.ImageWriter(str(tmp_path / 'npz'), fmt='npz')
    w.add(u8[:2], mine[:2])
    w.add(u8[2:], mine[2:])
    w.join()
    dat = np.load(str(tmp_path / 'npz' / 'images.npz'))
    assert np.array_equal(dat['images'], u8.numpy()) and list(dat['imgnums']) == nums


def test_ui_search_ranking_and_unit_quantiles_vs_oracle(cuda_model, z40, golden, seeded_sd):
    """ranking_for_key / quantiles_for_units / key_method='gandissect' (ganrewrite.py:375-400,
    554-594) on the device against the same statistics computed from the CPU oracle's keys."""
    from rewriting_b200.rewrite import ganrewrite
    zds = torch.utils.data.TensorDataset(z40)
    gw = ganrewrite.SeqStyleGanRewriter(cuda_model, zds, 8)
    d = torch.from_numpy(golden['d'])[0]
    sel, rq = gw.ranking_for_key(d.cuda(), k=6)
    with torch.no_grad():
        keys = torch.cat([orc.generator_forward(seeded_sd, z40[i:i + 10], upto_key_layer=8)
                          for i in range(0, 40, 10)])                # [40,512,32,32] oracle keys
    heat = (keys * d[None, :, None, None]).sum(1).reshape(40, -1)
    want_sel = heat.max(1)[0].topk(6)[1]
    assert sel.shape == (6,) and sorted(sel.tolist()) == sorted(want_sel.tolist())
    assert rq.size() == 40 * 1024
    qs = [0.01, 0.5, 0.99, 0.999]
    want_q = torch.quantile(heat.reshape(-1).double(), torch.tensor(qs, dtype=torch.float64))
    got_q = rq.quantiles(qs)[0].double().cpu()
    assert (got_q - want_q).abs().max().item() < 2e-3 * max(1.0, want_q.abs().max().item())
    urq = gw.quantiles_for_units()
    assert urq.depth == 512 and urq.size() == 40 * 1024
    flat = keys.permute(0, 2, 3, 1).reshape(-1, 512)
    # rank-space check (a unit whose two middle samples lie far apart has no well-defined value)
    got_med = urq.median().cpu()
    rank = (flat <= got_med[None, :] + 1e-3).float().mean(0)
    rank_lo = (flat <= got_med[None, :] - 1e-3).float().mean(0)
    assert (rank >= 0.5 - 1e-3).all() and (rank_lo <= 0.5 + 1e-3).all()
    one_hot = gw.multi_key_from_selection([(5, golden_mask())], rank=2, key_method='gandissect')
    assert one_hot.shape == (2, 512) and one_hot.sum().item() == 2 and (one_hot.sum(1) == 1).all()


def golden_mask():
    import base64, io
    from PIL import Image, ImageDraw
    im = Image.new('RGBA', (256, 256), (0, 0, 0, 0))
    ImageDraw.Draw(im).ellipse([70, 80, 130, 120], fill=(255, 255, 255, 255))
    buf = io.BytesIO()
    im.save(buf, format='png')
    return 'data:image/png;base64,' + base64.b64encode(buf.getvalue()).decode('ascii')
