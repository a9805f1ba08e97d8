"""Where does the conv error come from?  rowgemm on the GPU vs (i) exact fp64 of the fp32
operands [total error] and (ii) exact fp64 of the 3-term split products [accumulation error
of the tensor core only]."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from rewriting_b200 import _cabi


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def split(x):
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    _cabi.call('rw_split_rows', P(x), x.numel(), P(hi), P(lo), None)
    return hi, lo


def run(rows, K, N, positive=False):
    torch.manual_seed(0)
    a = torch.randn(rows, K, device='cuda')
    w = torch.randn(N, K, device='cuda') / K ** 0.5
    if positive:
        a, w = a.abs(), w.abs()
    ahi, alo = split(a)
    whi, wlo = split(w)
    out = torch.empty(rows, N, device='cuda')
    _cabi.call('rw_debug_rowgemm', P(ahi), P(alo), P(whi), P(wlo), rows, K, N, P(out), None)
    torch.cuda.synchronize()
    A, Wt = a.double(), w.double()
    exact = A @ Wt.t()
    AH, AL, WH, WL = ahi.double(), alo.double(), whi.double(), wlo.double()
    three = AH @ WH.t() + AL @ WH.t() + AH @ WL.t()
    o = out.double()
    rms = exact.pow(2).mean().sqrt().item()
    res = dict(rows=rows, K=K, N=N, positive=positive, rms=rms,
               total_max=(o - exact).abs().max().item() / rms,
               total_rms=(o - exact).pow(2).mean().sqrt().item() / rms,
               split_only_rms=(three - exact).pow(2).mean().sqrt().item() / rms,
               accum_max=(o - three).abs().max().item() / rms,
               accum_rms=(o - three).pow(2).mean().sqrt().item() / rms,
               accum_bias=((o - three) * three.sign()).mean().item() / rms)
    # fp32 reference accumulate (what the reference's cuDNN/oneDNN fp32 conv would give)
    f32 = (a @ w.t()).double()
    res['fp32_matmul_rms'] = (f32 - exact).pow(2).mean().sqrt().item() / rms
    return res


if __name__ == '__main__':
    outs = []
    for K in (64, 512, 4608):
        for pos in (False, True):
            r = run(2048, K, 128, pos)
            outs.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(outs, open('gpurun_out/precision_probe.json', 'w'), indent=1)
