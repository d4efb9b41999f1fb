"""``upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))`` — same signature and semantics as the reference wrapper
(model/stylegan/op/upfirdn2d.py:149-165): ``up``/``down`` int or (x, y); ``pad`` (p0, p1) -> (p0, p1, p0, p1) or
(x0, x1, y0, y1).  CUDA tensors only — the reference routes CPU tensors to ``upfirdn2d_native``; this library has no CPU path
and raises instead.

Differentiable like the reference op (upfirdn2d.py:20-146): the gradient of an upfirdn2d is the upfirdn2d with ``up`` and ``down``
swapped, the flipped kernel and the complementary padding; the second derivative is the forward op again."""
from collections import abc

import torch
from torch.autograd import Function

from .. import ops


class UpFirDn2dBackward(Function):
    """op/upfirdn2d.py:20-87"""

    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        grad_input = ops.upfirdn2d_planar(grad_output.reshape(in_size[0], in_size[1], out_size[0], out_size[1]), grad_kernel, down, up, g_pad)
        grad_input = grad_input.view(in_size[0], in_size[1], in_size[2], in_size[3])
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad, ctx.in_size, ctx.out_size = up, down, pad, in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gradgrad_out = ops.upfirdn2d_planar(gradgrad_input.reshape(ctx.in_size), kernel, ctx.up, ctx.down, ctx.pad)
        return gradgrad_out.view(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0], ctx.out_size[1]), None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    """op/upfirdn2d.py:90-146"""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        _, _, in_h, in_w = input.shape
        ctx.in_size = tuple(input.shape)
        out = ops.upfirdn2d_planar(input, kernel, up, down, pad)
        ctx.out_size = (out.shape[2], out.shape[3])
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        ctx.up, ctx.down, ctx.pad = up, down, pad
        g_pad_x0 = kernel_w - pad_x0 - 1
        g_pad_y0 = kernel_h - pad_y0 - 1
        g_pad_x1 = in_w * up_x - out.shape[3] * down_x + pad_x0 - up_x + 1
        g_pad_y1 = in_h * up_y - out.shape[2] * down_y + pad_y0 - up_y + 1
        ctx.g_pad = (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = None
        if ctx.needs_input_grad[0]:
            grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad, ctx.in_size,
                                                 ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not isinstance(up, abc.Iterable):
        up = (up, up)
    if not isinstance(down, abc.Iterable):
        down = (down, down)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    up, down, pad = tuple(int(u) for u in up), tuple(int(d) for d in down), tuple(int(p) for p in pad)
    if torch.is_grad_enabled() and input.requires_grad:
        return UpFirDn2d.apply(input, kernel, up, down, pad)
    return ops.upfirdn2d_planar(input, kernel, up, down, pad)
