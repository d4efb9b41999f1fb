"""``fused_leaky_relu`` / ``FusedLeakyReLU`` — same signatures as model/stylegan/op/fused_act.py:87-119.
``y = leaky_relu(x + bias[c], negative_slope) * scale`` with bias broadcast on dim 1 (any rank >= 2).

Differentiable like the reference op (forward, backward and double backward: fused_act.py:20-84): the two backward passes run
``vt_fused_bias_act_grad_f32`` (the op's ``grad=1`` mode, keyed on the sign of the forward OUTPUT) and the bias gradient a
deterministic per-channel reduction.  Without ``requires_grad`` the autograd machinery is bypassed."""
import torch
from torch import nn
from torch.autograd import Function

from .. import ops


class FusedLeakyReLUFunctionBackward(Function):
    """op/fused_act.py:20-53"""

    @staticmethod
    def forward(ctx, grad_output, out, bias, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        grad_input = ops.fused_bias_act_grad(grad_output, out, negative_slope, scale)
        grad_bias = ops.channel_sum(grad_input) if bias else grad_output.new_empty(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gb = gradgrad_bias if gradgrad_bias is not None and gradgrad_bias.numel() else None
        gradgrad_out = ops.fused_bias_act_grad(gradgrad_input, out, ctx.negative_slope, ctx.scale, bias=gb)
        return gradgrad_out, None, None, None, None


class FusedLeakyReLUFunction(Function):
    """op/fused_act.py:56-84"""

    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        ctx.bias = bias is not None
        out = ops.fused_bias_act(input, bias, float(negative_slope), float(scale))
        ctx.save_for_backward(out)
        ctx.negative_slope = float(negative_slope)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.bias, ctx.negative_slope, ctx.scale)
        if not ctx.bias:
            grad_bias = None
        return grad_input, grad_bias, None, None


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        if bias:
            self.bias = nn.Parameter(torch.zeros(channel))
        else:
            self.bias = None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    if torch.is_grad_enabled() and (input.requires_grad or (bias is not None and bias.requires_grad)):
        return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
    return ops.fused_bias_act(input, bias, float(negative_slope), float(scale))
