"""DualStyleGAN pieces used by VToonify (model/dualstylegan.py:6-76) with identical state_dict keys."""
import math

import torch
from torch import nn

from . import ops
from .stylegan import ConvLayer, EqualLinear, Generator, PixelNorm


class Linear(nn.Module):
    """nn.Linear replacement (same ``weight``/``bias`` keys and default init) running on vt_linear_f32.
    ``act``: 0 none, 2 LeakyReLU(0.2) fused."""

    def __init__(self, in_features, out_features, act=0):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(in_features)
        nn.init.uniform_(self.bias, -bound, bound)
        self.act = act

    def forward(self, input, act=None):
        return ops.linear(input, self.weight, self.bias, 1.0, 1.0, self.act if act is None else act)


class AdaptiveInstanceNorm(nn.Module):
    """model/dualstylegan.py:6-21: InstanceNorm2d(affine=False) then gamma * x + beta, [gamma|beta] = Linear(style)."""

    def __init__(self, fin, style_dim=512):
        super().__init__()
        self.style = Linear(style_dim, fin * 2)
        self.style.bias.data[:fin] = 1
        self.style.bias.data[fin:] = 0

    def gamma_beta(self, style, B):
        """[B, 2C] rows (gamma | beta) = Linear(style); style-only, so computed once per style inside a style scope
        (a shared [1, D] style is broadcast to the batch)"""
        def make():
            gb = self.style(style)
            return gb if gb.shape[0] == B else gb.expand(B, -1).contiguous()
        return ops.style_cached(self, "gb", make, extra=(B, self.style.weight._version))

    def forward_nhwc(self, x, style, x2=None):
        gb = self.gamma_beta(style, x.shape[0])
        stats = ops.instnorm_stats(x, x2)
        return ops.adain_apply(x, stats, gb, x2)

    def affine(self, x, style, stats=None):
        """The same AdaIN as a [B, C, 2] (scale, shift) table for a convolution that applies it to its input on the fly.
        ``stats``: (mean, rstd) of ``x`` when the kernel that produced ``x`` already delivered them."""
        return ops.adain_affine(ops.instnorm_stats(x) if stats is None else stats, self.gamma_beta(style, x.shape[0]))

    def forward(self, input, style):
        return ops.nhwc_as_nchw_view(self.forward_nhwc(ops.to_nhwc(input), style))


class AdaResBlock(nn.Module):
    """model/dualstylegan.py:24-45 (ModRes): x + w * conv2(AdaIN(conv(AdaIN(x, s)), s))."""

    def __init__(self, fin, style_dim=512, dilation=1):
        super().__init__()
        self.conv = ConvLayer(fin, fin, 3, dilation=dilation)
        self.conv2 = ConvLayer(fin, fin, 3, dilation=dilation)
        self.norm = AdaptiveInstanceNorm(fin, style_dim)
        self.norm2 = AdaptiveInstanceNorm(fin, style_dim)
        self.conv[0].weight.data *= 0.01
        self.conv2[0].weight.data *= 0.01

    def forward_nhwc(self, x, s, w=1, x_stats=None):
        """``x_stats``: instance-norm statistics of ``x`` from the kernel that produced it (else a statistics pass runs)."""
        if w == 0:
            return x
        if ops.affine_fusable():
            # the normalised tensors are never written: each conv applies its AdaIN (scale, shift per sample and channel) to
            # the pixels it stages, padding stays zero as in the reference (which zero-pads the normalised tensor); the
            # statistics of conv's output come out of conv's own epilogue
            out, st = self.conv.forward_nhwc(x, src_affine=self.norm.affine(x, s, x_stats), want_stats=True)
            return self.conv2.forward_nhwc(out, src_affine=self.norm2.affine(out, s, st), res=x, alpha=float(w), beta=1.0)
        out = self.conv.forward_nhwc(self.norm.forward_nhwc(x, s))
        return self.conv2.forward_nhwc(self.norm2.forward_nhwc(out, s), res=x, alpha=float(w), beta=1.0)

    def forward(self, x, s, w=1):
        return ops.nhwc_as_nchw_view(self.forward_nhwc(ops.to_nhwc(x), s, w))


class DualStyleGAN(ops.WeightsEpochMixin, nn.Module):
    """model/dualstylegan.py:47-76 constructor (parameters / keys); VToonify only uses ``.style``, ``.res[7:]``
    and ``.generator`` at inference (model/vtoonify.py:214-224, 279-283)."""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, twoRes=True, res_index=6):
        super().__init__()
        layers = [PixelNorm()]
        for _ in range(n_mlp - 6):
            layers.append(EqualLinear(512, 512, lr_mul=0.01, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        self.generator = Generator(size, style_dim, n_mlp, channel_multiplier)
        self.res = nn.ModuleList()
        self.res_index = res_index // 2 * 2
        self.res.append(AdaResBlock(self.generator.channels[2 ** 2]))
        for i in range(3, self.generator.log_size + 1):
            out_channel = self.generator.channels[2 ** i]
            if i < 3 + self.res_index // 2:
                self.res.append(AdaResBlock(out_channel))
                self.res.append(AdaResBlock(out_channel))
            else:
                for _ in range(2):
                    self.res.append(EqualLinear(512, 512))
                    self.res[-1].weight.data = torch.eye(512) * 512.0 ** 0.5 + torch.randn(512, 512) * 0.01
        self.res.append(EqualLinear(512, 512))
        self.res[-1].weight.data = torch.eye(512) * 512.0 ** 0.5 + torch.randn(512, 512) * 0.01
        self.size = self.generator.size
        self.style_dim = self.generator.style_dim
        self.log_size = self.generator.log_size
        self.num_layers = self.generator.num_layers
        self.n_latent = self.generator.n_latent
        self.channels = self.generator.channels
