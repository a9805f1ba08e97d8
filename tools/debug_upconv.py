"""Bring-up probe for csrc/upconv_tc.cu: dumps the raw tap products of the tensor-core stage and
compares them (and the final planes) with a CPU einsum."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import sg2_oracle as orc  # noqa: E402
from rewriting_b200 import _cabi, ops  # noqa: E402


def run(B, Cin, Cout, H):
    torch.manual_seed(1)
    dev = 'cuda'
    W = H
    x = torch.randn(B, Cin, H, W)
    style = torch.ones(B, Cin)
    weight = torch.randn(1, Cout, Cin, 3, 3)
    kern = orc.make_kernel([1, 3, 3, 1]) * 4
    scale = 1 / math.sqrt(Cin * 9)
    P = torch.einsum('biyx,oiuv->byxuvo', x.double(), scale * weight[0].double()).reshape(B, H, W, 9, Cout)
    planes, _ = ops.prep_keys(x.to(dev), style.to(dev))
    wp = torch.nn.Parameter(weight.to(dev))
    u_hi, u_lo, _ = ops.weight_planes(wp, 'upf')
    Ho, Wo = 2 * H, 2 * W
    noise = torch.zeros(B, Ho * Wo, device=dev)
    ones = torch.ones(B, Cout, device=dev)
    nw = torch.zeros(1, device=dev)
    bias = torch.zeros(Cout, device=dev)
    rows_o = B * (Ho + 1) * (Wo + 1)
    nh = torch.zeros((rows_o, Cout), dtype=torch.bfloat16, device=dev)
    nl = torch.zeros_like(nh)
    taps = torch.full((B, H, W, 9, Cout), float('nan'), device=dev)
    _cabi.call('rw_debug_upconv_taps', ops._p(planes.hi), ops._p(planes.lo), ops._p(u_hi),
               ops._p(u_lo), ops._p(ones), ops._p(kern.to(dev)), ops._p(noise), noise.stride(0),
               ops._p(nw), ops._p(bias), ops._p(nh), ops._p(nl), B, Cin, Cout, H, W, ops._p(taps),
               ops._stream())
    torch.cuda.synchronize()
    got = taps.cpu().double()
    err = (got - P).abs()
    print('B=%d Cin=%d Cout=%d H=%d: taps max err %.3g (max |P| %.3g) nan %d' % (
        B, Cin, Cout, H, err[torch.isfinite(err)].max().item() if torch.isfinite(err).any() else -1,
        P.abs().max().item(), int((~torch.isfinite(got)).sum())))
    if err[torch.isfinite(err)].max() > 1e-3:
        bad = (err > 1e-3).nonzero()
        print('  first bad', bad[:5].tolist())
        b, y, xx, t, o = bad[0].tolist()
        print('  got', got[b, y, xx, :, o].tolist())
        print('  want', P[b, y, xx, :, o].tolist())
        # is it a permutation?  search each got value in want
        flatP = P[b, y].reshape(-1)
        for t2 in range(9):
            v = got[b, y, xx, t2, o]
            j = (flatP - v).abs().argmin().item()
            print('   got tap %d = %.5f closest want at (x,tap,o)=%s diff %.2g' % (
                t2, v, (j // (9 * Cout), (j // Cout) % 9, j % Cout), (flatP[j] - v).abs().item()))
    # final output, leaky-ReLU of the blurred conv_transpose
    t_ref = torch.nn.functional.conv_transpose2d(x.double(), (scale * weight[0].double()).transpose(0, 1), stride=2)
    want = orc.fused_leaky_relu(orc.upfirdn2d(t_ref, kern.double(), pad=(1, 1)), torch.zeros(Cout).double())
    out = (nh.float() + nl.float()).cpu().view(B, Ho + 1, Wo + 1, Cout)[:, :Ho, :Wo].permute(0, 3, 1, 2).double()
    print('  planes max err %.3g (max %.3g)' % ((out - want).abs().max().item(), want.abs().max().item()))


if __name__ == '__main__':
    for cfg in [(2, 64, 16, 4), (2, 64, 32, 128), (3, 128, 32, 32)]:
        run(*cfg)
