"""`upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))`: upsample by zero insertion, pad,
FIR-filter with the flipped kernel, decimate — per (batch, channel) plane.

Reference: utils/stylegan2/op/upfirdn2d.py:18-149 and upfirdn2d_kernel.cu:52-272.  The
arithmetic is the C-ABI kernel `rw_upfirdn2d`; the backward pass is the same op with the
flipped kernel, swapped up/down factors and the adjoint padding, exactly as the reference
derives it (upfirdn2d.py:112-117).
"""
import torch
from torch.autograd import Function

from .... import ops


class _UpFirDnModule(object):
    """Stand-in for the reference's pybind11 module object `upfirdn2d_op`."""

    @staticmethod
    def upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
        if not input.is_cuda or not kernel.is_cuda:
            raise RuntimeError('input must be a CUDA tensor')
        return ops.upfirdn2d_raw(input, kernel, up_x, up_y, down_x, down_y,
                                 pad_x0, pad_x1, pad_y0, pad_y1)


upfirdn2d_op = _UpFirDnModule()


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        batch, channel, in_h, in_w = input.shape
        out = upfirdn2d_op.upfirdn2d(input.reshape(-1, in_h, in_w, 1), kernel, up_x, up_y,
                                     down_x, down_y, px0, px1, py0, py1)
        out_h, out_w = out.shape[1], out.shape[2]
        ctx.save_for_backward(torch.flip(kernel, [0, 1]))
        ctx.geom = (up, down, input.shape, (out_h, out_w))
        # adjoint padding
        ctx.g_pad = (kw - px0 - 1, in_w * up_x - out_w * down_x + px0 - up_x + 1,
                     kh - py0 - 1, in_h * up_y - out_h * down_y + py0 - up_y + 1)
        return out.view(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        flipped, = ctx.saved_tensors
        (up_x, up_y), (down_x, down_y), in_shape, (out_h, out_w) = ctx.geom
        gx0, gx1, gy0, gy1 = ctx.g_pad
        g = upfirdn2d_op.upfirdn2d(grad_output.reshape(-1, out_h, out_w, 1), flipped,
                                   down_x, down_y, up_x, up_y, gx0, gx1, gy0, gy1)
        return g.view(in_shape), None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down),
                           (pad[0], pad[1], pad[0], pad[1]))
