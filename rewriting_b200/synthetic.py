"""Synthetic generator weights for benchmarks and smoke runs (no checkpoint can be downloaded on
the GPU box): seeded random init of `SeqStyleGAN2(256, style_dim=512, n_mlp=8, mconv='seq')`,
then every `*.noise.weight` = 0.37 and every `*.activate.bias` ~ N(0, 1) — both are zero at
init (reference models.py:538, op/fused_act.py:77), which would leave the noise / bias terms of
the fused epilogues unexercised.  The test suite pins this recipe to the weights
the golden vectors were generated with (tests/golden/weights_checksum.json)."""
import torch


def seeded_generator(size=256, seed=0, noise_weight=0.37, **kwargs):
    from .utils.stylegan2 import SeqStyleGAN2
    kwargs.setdefault('style_dim', 512)
    kwargs.setdefault('n_mlp', 8)
    kwargs.setdefault('mconv', 'seq')
    torch.manual_seed(seed)
    model = SeqStyleGAN2(size, **kwargs)
    g = torch.Generator().manual_seed(seed + 12345)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith('noise.weight'):
                p.fill_(noise_weight)
            elif name.endswith('activate.bias'):
                p.copy_(torch.randn(p.shape, generator=g))
    return model.eval()


def seeded_vgg16(seed=20200701):
    """torchvision VGG-16 with seeded random weights: the stand-in for the pretrained perceptual
    network of `all_weights_insert` (reference ganrewrite.py:303-304) wherever the ImageNet
    weights cannot be downloaded — tests, goldens (oracle/make_golden_overfit.py) and benches.
    The global RNG state is left untouched."""
    import torchvision
    state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        vgg = torchvision.models.vgg16(weights=None)
    finally:
        torch.random.set_rng_state(state)
    return vgg.eval()
