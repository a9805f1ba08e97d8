"""`fused_leaky_relu(input, bias, negative_slope=0.2, scale=sqrt(2))` and the raw
`fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)` entry point.

Reference: utils/stylegan2/op/fused_act.py:19-86 (autograd wrapper) and
fused_bias_act_kernel.cu:19-98.  The forward/backward arithmetic is the C-ABI kernel
`rw_fused_bias_act`; first-order autograd is provided (the reference additionally
implements double-backward, which only GAN-training regularisers need — not the
rewriting path — and is not provided here).
"""
import torch
from torch import nn
from torch.autograd import Function

from .... import ops


class _FusedModule(object):
    """Stand-in for the reference's pybind11 module object `fused`."""

    @staticmethod
    def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
        if not input.is_cuda:
            raise RuntimeError('input must be a CUDA tensor')
        if bias is not None and bias.numel() and not bias.is_cuda:
            raise RuntimeError('bias must be a CUDA tensor')
        return ops.fused_bias_act_raw(input, bias, refer, act, grad, alpha, scale)


fused = _FusedModule()


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = fused.fused_bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        ctx.bias_shape = None if bias is None else bias.shape
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        # gate on the sign of the saved OUTPUT (fused_bias_act_kernel.cu:43)
        grad_input = fused.fused_bias_act(grad_output, None, out, 3, 1,
                                          ctx.negative_slope, ctx.scale)
        grad_bias = None
        if ctx.bias_shape is not None and ctx.needs_input_grad[1]:
            dims = [0] + list(range(2, grad_input.ndim))
            grad_bias = grad_input.sum(dims).reshape(ctx.bias_shape)
        return grad_input, grad_bias, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
