"""``fused_leaky_relu`` / ``FusedLeakyReLU`` — same signatures as model/stylegan/op/fused_act.py:87-119.
``y = leaky_relu(x + bias[c], negative_slope) * scale`` with bias broadcast on dim 1 (any rank >= 2)."""
import torch
from torch import nn

from .. import ops


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
    return ops.fused_bias_act(input, bias, float(negative_slope), float(scale))
