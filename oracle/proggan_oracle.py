"""TEST INFRASTRUCTURE ONLY — CPU restatement of the ProgGAN side of the rewrite path
(SURVEY.md §8 f-3): the progressive generator (reference utils/proggan.py:63-199) and the
`ProgressiveGanRewriter` edit on `layerN.conv` (rewrite/ganrewrite.py:25-96, 254-298), i.e. a
plain 3x3 convolution target with no modulation.  Pinned to the live reference by
oracle/make_golden_proggan.py; only tests/ may import it.

A state_dict here is the reference's: `layer<i>.conv.weight` [Cout,Cin,k,k], `layer<i>.wscale.b`
[Cout], `output_<R>x<R>.conv.weight` [3,Cin,1,1], `output_<R>x<R>.wscale.b` [3]."""
import itertools
import math

import torch
import torch.nn.functional as F


def pixel_norm(x):
    """PixelNormLayer (proggan.py:128-134)."""
    return x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def wscale(x, b, fan_in, gain):
    """WScaleLayer (proggan.py:143-155): x * gain/sqrt(fan_in) + b."""
    return x * (gain / math.sqrt(fan_in)) + b.view(1, -1, 1, 1)


def num_layers(sd):
    for i in itertools.count():
        if 'layer%d.conv.weight' % (i + 1) not in sd:
            return i


def block(sd, i, x, upto_conv_input=False):
    """layer i (1-based): NormConvBlock / NormUpscaleConvBlock (proggan.py:158-181)."""
    w = sd['layer%d.conv.weight' % i]
    k = w.shape[2]
    x = pixel_norm(x)
    if i > 2 and i % 2 == 1:                               # layers 3, 5, 7, ...: 2x nearest first
        x = F.interpolate(x, scale_factor=2, mode='nearest')
    if upto_conv_input:
        return x
    x = F.conv2d(x, w, padding=3 if i == 1 else 1)
    x = wscale(x, sd['layer%d.wscale.b' % i], w.shape[1], math.sqrt(2) / k)
    return F.leaky_relu(x, 0.2)


def generator_forward(sd, z, upto_key_layer=None, from_layer_output=None, tanh=True):
    """z [B, Z] -> image [B,3,R,R]; `upto_key_layer=N` returns the input of `layerN.conv` (the
    rewriter's key: after layerN.norm and layerN.up); `from_layer_output=(N, v)` resumes after
    `layerN.conv` with its raw output v (the rewriter's rendering model)."""
    n = num_layers(sd)
    start = 1
    if from_layer_output is not None:
        N, x = from_layer_output
        w = sd['layer%d.conv.weight' % N]
        x = F.leaky_relu(wscale(x, sd['layer%d.wscale.b' % N], w.shape[1],
                                math.sqrt(2) / w.shape[2]), 0.2)
        start = N + 1
    else:
        x = z.view(z.shape[0], z.shape[1], 1, 1)
    for i in range(start, n + 1):
        if upto_key_layer == i:
            return block(sd, i, x, upto_conv_input=True)
        x = block(sd, i, x)
    res = 4 * (2 ** (n // 2 - 1))
    w = sd['output_%dx%d.conv.weight' % (res, res)]
    x = F.conv2d(pixel_norm(x), w)
    x = wscale(x, sd['output_%dx%d.wscale.b' % (res, res)], w.shape[1], 1.0)
    return F.hardtanh(x) if tanh else x


def projected_conv(weight, direction):
    """ganrewrite.py:806-813 for a 4-D conv weight."""
    cos = torch.einsum('oiyx,di->odyx', weight, direction)
    return torch.einsum('odyx,di->oiyx', cos, direction)


def insert_loop(weight, k, target, d, niter, piter=10, lr=0.05, low_rank_insert=True,
                record_loss=None):
    """ProgressiveGanRewriter.insert (ganrewrite.py:254-298) for the plain `layerN.conv` target:
    L1(v*, conv(k*, W)), Adam, every `piter` steps W <- W_ortho + P_d(W)."""
    weight = weight.clone().requires_grad_(True)
    with torch.no_grad():
        ortho = weight - projected_conv(weight, d)
    opt = torch.optim.Adam([weight], lr=lr)
    for it in range(niter):
        loss = F.l1_loss(target, F.conv2d(k, weight, padding=1))
        opt.zero_grad()
        loss.backward()
        opt.step()
        if record_loss is not None:
            record_loss.append(float(loss))
        if low_rank_insert and (it % piter == 0 or it == niter - 1):
            with torch.no_grad():
                weight[...] = ortho + projected_conv(weight, d)
    return weight.detach()


def seeded_state_dict(model_ctor, seed=0):
    """Synthetic ProgGAN weights: seeded default init, `wscale.b` ~ N(0, 0.25) (WScaleLayer draws
    b ~ N(0,1) itself; smaller biases keep the activations O(1) through 10 blocks)."""
    torch.manual_seed(seed)
    model = model_ctor()
    g = torch.Generator().manual_seed(seed + 777)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith('conv.weight'):
                p.copy_(torch.randn(p.shape, generator=g))     # He-style: wscale supplies the gain
                if name.startswith('output'):
                    p.mul_(0.3)                                # keep the hardtanh mostly unsaturated
            elif name.endswith('wscale.b'):
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
    return model.eval()
