"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (davidbau/rewriting, mounted
read-only at /root/reference) so that it runs on CPU in the authoring container.

Only `oracle/make_golden.py` uses this, to pin the CPU oracle (`oracle/sg2_oracle.py`) and to
produce the committed fixtures under tests/golden/.  /root/reference does not exist on the GPU
box; nothing in tests/, bench.py or the product imports this module there.

The shim is monkey-patches only (no reference file is touched or copied) — SURVEY.md §8c:
  1. `torch.utils.cpp_extension.load` is stubbed so the two CUDA extensions are not JIT-built;
     `fused_leaky_relu` becomes its definition `leaky_relu(x + b) * scale`
     (fused_bias_act_kernel.cu:27-47) and `upfirdn2d` is routed to the reference's OWN
     `upfirdn2d_native` (op/upfirdn2d.py:152-186, dead code there because `F` is undefined);
  2. `.cuda()` is the identity;
  3. `torch.symeig` / `torch.lstsq` (removed in torch 2) map to torch.linalg.eigh(UPLO='U') /
     torch.linalg.lstsq;
  4. `matplotlib` is stubbed (utils/imgviz.py imports it; it is never called on this path).
"""
import sys
import types

import torch
import torch.nn.functional as F

REFERENCE_ROOT = '/root/reference'


def load_reference():
    """Returns a namespace with the reference modules: models, ganrewrite, nethook, tally,
    runningstats, zdataset, renormalize."""
    import torch.utils.cpp_extension as cpp_ext

    class _Dummy(object):
        pass
    real_load = cpp_ext.load
    cpp_ext.load = lambda *a, **k: _Dummy()
    for name in ('matplotlib', 'matplotlib.cm', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        import utils.stylegan2.models as models          # noqa: E402
        import utils.stylegan2.op.fused_act as fused_act  # noqa: E402
        upf = sys.modules['utils.stylegan2.op.upfirdn2d']
        from utils import nethook, tally, runningstats, zdataset, renormalize  # noqa: E402
        from rewrite import ganrewrite                   # noqa: E402
    finally:
        cpp_ext.load = real_load

    def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
        shape = [1, -1] + [1] * (input.dim() - 2)
        return F.leaky_relu(input + bias.view(*shape), negative_slope) * scale

    upf.F = F

    def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
        b, c, h, w = input.shape
        out = upf.upfirdn2d_native(input.reshape(-1, h, w, 1), kernel, up, up, down, down,
                                   pad[0], pad[1], pad[0], pad[1])
        return out.view(b, c, out.shape[1], out.shape[2])

    fused_act.fused_leaky_relu = fused_leaky_relu
    models.op.fused_leaky_relu = fused_leaky_relu
    models.op.upfirdn2d = upfirdn2d
    models.op.FusedLeakyReLU.forward = lambda self, x: fused_leaky_relu(
        x, self.bias, self.negative_slope, self.scale)

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    # removed / raising in torch 2.x: always override
    def symeig(a, eigenvectors=False, upper=True):
        return tuple(torch.linalg.eigh(a, UPLO='U' if upper else 'L'))
    torch.symeig = symeig

    def lstsq(b, a):
        return torch.linalg.lstsq(a, b).solution, None
    torch.lstsq = lstsq
    torch.qr = lambda a: tuple(torch.linalg.qr(a))

    ns = types.SimpleNamespace(models=models, ganrewrite=ganrewrite, nethook=nethook, tally=tally,
                               runningstats=runningstats, zdataset=zdataset,
                               renormalize=renormalize)
    return ns
