"""Config 2 of BASELINE.json: every StyledConv shape of the 256^2 generator at batch 32,
fused forward and forward+backward (dX, dstyle, dW, dbias, dnoise), CUDA-event timed.
Algorithmic FLOPs: fwd = 2*B*Cin*Cout*9*H_in^2 (up) or *H^2 (non-up); bwd = 2x fwd."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from rewriting_b200 import ops

SHAPES = [('layer2', 512, 512, 4, 0), ('layer3', 512, 512, 4, 1), ('layer4', 512, 512, 8, 0),
          ('layer5', 512, 512, 8, 1), ('layer6', 512, 512, 16, 0), ('layer7', 512, 512, 16, 1),
          ('layer8', 512, 512, 32, 0), ('layer9', 512, 512, 32, 1), ('layer10', 512, 512, 64, 0),
          ('layer11', 512, 256, 64, 1), ('layer12', 256, 256, 128, 0), ('layer13', 256, 128, 128, 1),
          ('layer14', 128, 128, 256, 0)]


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main(B=32, only=None, quiet=False, save=True):
    torch.manual_seed(0)
    kern = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :])
    kern = (kern / kern.sum() * 4).cuda()
    out, tot_f, tot_fb, fl_f = [], 0.0, 0.0, 0.0
    for name, cin, cout, h, up in SHAPES:
        if only and name not in only:
            continue
        x = torch.randn(B, cin, h, h, device='cuda', requires_grad=True)
        style = (torch.randn(B, cin, device='cuda') * 0.5 + 1).requires_grad_(True)
        w = torch.nn.Parameter(torch.randn(1, cout, cin, 3, 3, device='cuda'))
        nw = torch.nn.Parameter(torch.full((1,), 0.37, device='cuda'))
        bias = torch.nn.Parameter(torch.randn(cout, device='cuda'))
        ho = 2 * h if up else h
        gy = torch.randn(B, cout, ho, ho, device='cuda')

        def fwd():
            with torch.no_grad():
                return ops.styled_conv(x, style, w, nw, bias, upsample=bool(up), blur_kernel=kern)

        def fwdbwd():
            for t in (x, style, w, nw, bias):
                t.grad = None
            y = ops.styled_conv(x, style, w, nw, bias, upsample=bool(up), blur_kernel=kern)
            y.backward(gy)
        iters = 5 if h >= 64 else 20
        flops = 2.0 * B * cin * cout * 9 * h * h
        tf = timeit(fwd, iters)
        tfb = timeit(fwdbwd, iters)
        rec = dict(layer=name, Cin=cin, Cout=cout, H_in=h, up=up, fwd_ms=tf, fwdbwd_ms=tfb,
                   fwd_TFLOPs=flops / tf / 1e9, fwdbwd_TFLOPs=3 * flops / tfb / 1e9)
        out.append(rec)
        tot_f += tf; tot_fb += tfb; fl_f += flops
        if not quiet:
            print(json.dumps(rec), flush=True)
        del x, style, w, gy
        torch.cuda.empty_cache()
    summary = dict(batch=B, total_fwd_ms=tot_f, total_fwdbwd_ms=tot_fb,
                   fwd_TFLOPs=fl_f / tot_f / 1e9, fwdbwd_TFLOPs=3 * fl_f / tot_fb / 1e9,
                   note='layer-level op (fp32 NCHW in/out, prep + conv_tc + epilogue kernels), '
                        'algorithmic FLOPs (1x); operands 3-term split bf16')
    if not quiet:
        print(json.dumps(summary))
    if save:
        os.makedirs('gpurun_out', exist_ok=True)
        json.dump(dict(layers=out, summary=summary), open('gpurun_out/modconv_fwdbwd.json', 'w'),
                  indent=1)
    return dict(layers=out, summary=summary)


if __name__ == '__main__':
    main(only=set(sys.argv[1:]) or None)
