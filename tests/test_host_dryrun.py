"""CPU: the host side of the CUDA path, dry-run.  `_cabi.call` is replaced by a checker that
validates every launch against the C-ABI prototype table (symbol exists in librw_b200.so,
argument count, ctypes convertibility) without touching a GPU; outputs stay uninitialised, so
only shapes / control flow / the launch sequence are asserted here.  Numerics are the GPU tests'
job (tests/test_gpu_*.py)."""
import ctypes

import pytest
import torch

from rewriting_b200 import _cabi


@pytest.fixture
def dry(monkeypatch):
    from rewriting_b200 import fastpath, ops
    lib = _cabi.load()
    calls = []

    def fake_call(name, *args):
        res, argtypes = _cabi.SIGNATURES[name]
        assert hasattr(lib, name), name
        assert len(args) == len(argtypes), (name, len(args), len(argtypes))
        for a, t in zip(args, argtypes):
            if a is None:
                assert t in (_cabi.c_p,) or issubclass(t, ctypes._Pointer), (name, t)
                continue
            if isinstance(a, ctypes.Array):
                assert t is _cabi.c_p, (name, t)
                continue
            t.from_param(a)          # raises if the python value cannot become this C type
        calls.append(name)

    def f32c(t):
        if t is None:
            return None
        assert t.dtype == torch.float32
        return t.contiguous()

    monkeypatch.setattr(_cabi, 'call', fake_call)
    monkeypatch.setattr(ops, '_f32c', f32c)
    monkeypatch.setattr(ops, '_stream', lambda: None)
    monkeypatch.setattr(fastpath, '_stream', lambda: None)
    monkeypatch.setattr(lib, 'rw_gram_workspace_bytes', lambda *a: 1 << 20, raising=False)
    return calls


def test_fastpath_launch_sequence(dry, seeded_model):
    from rewriting_b200 import fastpath
    z = torch.randn(2, 512)
    with torch.no_grad():
        img = fastpath.forward(seeded_model, z)
    assert img.shape == (2, 3, 256, 256)
    assert dry.count('rw_pixel_norm') == 1 and dry.count('rw_equal_linear') == 8
    assert dry.count('rw_styles') == 1 and dry.count('rw_demod_multi') == 1
    # every upsampling layer is ONE launch (conv_transpose + blur + activation + next planes)
    assert dry.count('rw_modconv_fwd_fused') == 7 and dry.count('rw_modconv_up_fused') == 6
    assert 'rw_modconv_up_fwd_cl' not in dry and 'rw_blur_up_fused' not in dry
    assert dry.count('rw_rgb_combine') == 7
    del dry[:]
    with torch.no_grad():
        u8 = fastpath.forward(seeded_model, z, out_u8=True)
    assert u8.shape == (2, 256, 256, 3) and u8.dtype == torch.uint8
    assert dry.count('rw_rgb_combine') == 6 and dry.count('rw_rgb_combine_u8') == 1
    del dry[:]
    with torch.no_grad():
        planes = fastpath.forward(seeded_model, z, upto_key_layer=8)
    assert (planes.B, planes.C, planes.H, planes.W) == (2, 512, 32, 32)
    assert planes.hi.shape == (2 * 33 * 33, 512)
    assert dry.count('rw_modconv_fwd_fused') == 3 and dry.count('rw_modconv_up_fused') == 3
    assert 'rw_rgb_combine' not in dry      # key collection skips every ToRGB


@pytest.mark.parametrize('up,demod,noise,act,premod', [
    (False, True, True, True, False), (True, True, True, True, False),
    (False, False, False, False, False), (True, True, False, False, True),
    (False, True, True, True, True)])
def test_styled_conv_autograd_launch_sequence(dry, up, demod, noise, act, premod):
    from rewriting_b200 import ops
    B, Cin, Cout, H, W = 2, 128, 64, 4, 5
    x = torch.randn(B, Cin, H, W, requires_grad=not premod)
    style = torch.randn(B, Cin, requires_grad=True)
    w = torch.nn.Parameter(torch.randn(1, Cout, Cin, 3, 3))
    nw = torch.nn.Parameter(torch.tensor([0.3]))
    bias = torch.nn.Parameter(torch.randn(Cout))
    kern = torch.ones(4, 4)
    y = ops.styled_conv(x, style, w, nw, bias, upsample=up, blur_kernel=kern if up else None,
                        demodulate=demod, with_noise=noise, with_act=act, pre_modulated=premod)
    Ho, Wo = (2 * H, 2 * W) if up else (H, W)
    assert y.shape == (B, Cout, Ho, Wo)
    del dry[:]
    y.backward(torch.randn_like(y))
    assert dry[0] == 'rw_act_grad_reduce'
    assert ('rw_blur_adj_phase_keys' in dry) == up and ('rw_prep_keys' in dry) == (not up)
    assert ('rw_dgrad_finish' in dry) == (not premod)
    assert 'rw_wgrad_finish' in dry
    assert w.grad.shape == w.shape and style.grad is not None or not demod
    if not premod:
        assert x.grad.shape == x.shape and style.grad.shape == style.shape
    if noise:
        assert nw.grad.shape == nw.shape
    if act:
        assert bias.grad.shape == bias.shape
    if premod and not x.requires_grad:
        # key detached (the rewriter's insert path): no dgrad GEMM at all
        assert 'rw_modconv_fwd' not in dry and 'rw_modconv_up_dgrad' not in dry


def test_weight_only_backward_skips_dgrad(dry):
    from rewriting_b200 import ops
    x = torch.randn(1, 128, 4, 4)
    style = torch.randn(1, 128)
    w = torch.nn.Parameter(torch.randn(1, 128, 128, 3, 3))
    y = ops.styled_conv(x, style, w, torch.tensor([0.1]), torch.zeros(128), pre_modulated=True)
    del dry[:]
    y.sum().backward()
    assert 'rw_conv_wgrad' in dry and 'rw_wgrad_finish' in dry
    assert 'rw_modconv_fwd' not in dry and 'rw_dgrad_finish' not in dry
