"""pSp style encoder ``GradualStyleEncoder`` (IR-SE-50 + FPN + 18 style heads) with the reference's constructor,
forward and state_dict keys (model/encoder/encoders/psp_encoders.py:11-116, helpers.py:56-119), on the library's kernels.
It runs once per video (style_transfer.py:138-150); it is here for coverage of the path, not for throughput.

Inference form of each piece:
  Conv2d -> BatchNorm2d         : BN folded into the conv weights / bias (exact)
  BatchNorm2d -> Conv2d(pad 1)  : BN applied first by the AdaIN-apply kernel (the conv zero-pads the *normalised* tensor,
                                  so the affine cannot be folded into the weights at the border)
  PReLU                          : per-channel slope in the conv epilogue
  SEModule                       : plane mean (deterministic stats kernel) -> two tiny linears -> gate fused with the
                                  residual add; MaxPool2d(1, stride) shortcut == strided read in the same kernel
  _upsample_add                  : bilinear (align_corners=True) + add kernel
"""
import math
from argparse import Namespace

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import ACT_LRELU
from .stylegan import EqualLinear
from .vtoonify import Conv2d


class BatchNorm2d(nn.Module):
    """Parameter holder with nn.BatchNorm2d's keys; eval-mode affine a*x + c."""

    def __init__(self, c, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.eps = eps

    def affine(self):
        a = self.weight.detach() / torch.sqrt(self.running_var + self.eps)
        return a, self.bias.detach() - self.running_mean * a

    def apply_nhwc(self, x):
        B, _, _, C = x.shape
        stats = torch.stack([self.running_mean, torch.rsqrt(self.running_var + self.eps)], dim=1)   # [C, 2]
        gb = torch.cat([self.weight.detach(), self.bias.detach()])
        return ops.adain_apply(x, stats.unsqueeze(0).expand(B, C, 2).contiguous(), gb.unsqueeze(0).expand(B, 2 * C).contiguous())


class PReLU(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.full((c,), 0.25))


class _FoldedConv:
    """conv (no bias) followed by BatchNorm: cached folded weights in the kernels' layout + bias."""

    def __init__(self):
        self.key, self.w, self.b = None, None, None

    def get(self, conv, bn, cin_pad):
        key = (conv.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version,
               bn.running_var._version, cin_pad, ops.get_precision(), conv.weight.device)
        if key != self.key:
            a, c = bn.affine()
            self.w = ops.prep_weights((conv.weight.detach() * a.view(-1, 1, 1, 1)).contiguous(), cin_pad=cin_pad)
            self.b = c.contiguous()
            self.key = key
        return self.w, self.b


def _conv_nhwc(x, w, k, stride, pad, **epi):
    B, H, W, _ = x.shape
    Ho, Wo = ops.conv_out_size(H, k, stride, pad, 1), ops.conv_out_size(W, k, stride, pad, 1)
    return ops.conv2d_nhwc([x], w, ops.conv_taps(k, pad), stride, Ho, Wo, **epi)


class SEModule(nn.Module):
    """helpers.py:56-72"""

    def __init__(self, channels, reduction):
        super().__init__()
        self.fc1 = Conv2d(channels, channels // reduction, 1, 1, 0, bias=False)
        self.fc2 = Conv2d(channels // reduction, channels, 1, 1, 0, bias=False)

    def gate(self, x):
        mean = ops.instnorm_stats(x)[:, :, 0].contiguous()                    # AdaptiveAvgPool2d(1)
        h = ops.linear(mean, self.fc1.weight.flatten(1), None, act=3)         # 1x1 conv on a 1x1 map + ReLU
        return ops.linear(h, self.fc2.weight.flatten(1), None, act=4)         # + Sigmoid


class bottleneck_IR_SE(nn.Module):
    """helpers.py:97-119"""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        self.stride = stride
        if in_channel == depth:
            self.shortcut_layer = nn.Identity()          # MaxPool2d(1, stride): no parameters
        else:
            self.shortcut_layer = nn.Sequential(Conv2d(in_channel, depth, 1, stride, 0, bias=False), BatchNorm2d(depth))
        self.res_layer = nn.Sequential(BatchNorm2d(in_channel), Conv2d(in_channel, depth, 3, 1, 1, bias=False), PReLU(depth),
                                       Conv2d(depth, depth, 3, stride, 1, bias=False), BatchNorm2d(depth), SEModule(depth, 16))
        self._sc, self._c2 = _FoldedConv(), _FoldedConv()

    def forward_nhwc(self, x):
        bn1, conv1, prelu, conv2, bn2, se = self.res_layer
        C = x.shape[3]
        if isinstance(self.shortcut_layer, nn.Identity):
            sc, sc_stride = x, self.stride
        else:
            w, b = self._sc.get(self.shortcut_layer[0], self.shortcut_layer[1], C)
            sc, sc_stride = _conv_nhwc(x, w, 1, self.stride, 0, bias=b), 1
        t = bn1.apply_nhwc(x)
        t = _conv_nhwc(t, conv1._w.get(conv1.weight, 1.0, C), 3, 1, 1, act=ACT_LRELU, gain=1.0, slope_vec=prelu.weight)
        w2, b2 = self._c2.get(conv2, bn2, t.shape[3])
        t = _conv_nhwc(t, w2, 3, self.stride, 1, bias=b2)
        return ops.gate_shortcut_add(t, se.gate(t), sc, sc_stride)


def get_blocks(num_layers):
    """helpers.py:29-53: (in_channel, depth, stride) of every unit."""
    units = {50: [3, 4, 14, 3], 100: [3, 13, 30, 3], 152: [3, 8, 36, 3]}[num_layers]
    cfg, in_c = [], 64
    for depth, n in zip([64, 128, 256, 512], units):
        cfg.append((in_c, depth, 2))
        cfg += [(depth, depth, 1)] * (n - 1)
        in_c = depth
    return cfg


class GradualStyleBlock(nn.Module):
    """psp_encoders.py:11-32: stride-2 convs + LeakyReLU(0.01) down to 1x1, then EqualLinear."""

    def __init__(self, in_c, out_c, spatial):
        super().__init__()
        self.out_c, self.spatial = out_c, spatial
        mods = []
        for i in range(int(np.log2(spatial))):
            mods += [Conv2d(in_c if i == 0 else out_c, out_c, 3, 2, 1), nn.Identity()]     # nn.LeakyReLU() placeholder
        self.convs = nn.Sequential(*mods)
        self.linear = EqualLinear(out_c, out_c, lr_mul=1)

    def forward_nhwc(self, x):
        for m in self.convs:
            if isinstance(m, Conv2d):
                x = m.forward_nhwc(x, act=ACT_LRELU, slope=0.01, gain=1.0)
        return self.linear(x.reshape(-1, self.out_c))


class GradualStyleEncoder(nn.Module):
    """psp_encoders.py:35-116"""

    def __init__(self, num_layers, mode='ir', opts=None):
        super().__init__()
        assert num_layers in [50, 100, 152], 'num_layers should be 50,100, or 152'
        assert mode in ['ir', 'ir_se'], 'mode should be ir or ir_se'
        if mode != 'ir_se':
            raise NotImplementedError("VToonify loads the IR-SE-50 encoder (util.py:149)")
        opts = opts or Namespace(input_nc=3, n_styles=18)
        self.input_layer = nn.Sequential(Conv2d(opts.input_nc, 64, 3, 1, 1, bias=False), BatchNorm2d(64), PReLU(64))
        self.body = nn.Sequential(*[bottleneck_IR_SE(i, d, s) for i, d, s in get_blocks(num_layers)])
        self.styles = nn.ModuleList()
        self.style_count = opts.n_styles
        self.coarse_ind, self.middle_ind = 3, 7
        for i in range(self.style_count):
            self.styles.append(GradualStyleBlock(512, 512, 16 if i < self.coarse_ind else (32 if i < self.middle_ind else 64)))
        self.latlayer1 = Conv2d(256, 512, 1, 1, 0)
        self.latlayer2 = Conv2d(128, 512, 1, 1, 0)
        self._in = _FoldedConv()

    def forward(self, x):
        xn = ops.to_nhwc(x, ops._pad32(x.shape[1]))
        w, b = self._in.get(self.input_layer[0], self.input_layer[1], xn.shape[3])
        h = _conv_nhwc(xn, w, 3, 1, 1, bias=b, act=ACT_LRELU, gain=1.0, slope_vec=self.input_layer[2].weight)
        c1 = c2 = c3 = None
        for i, blk in enumerate(self.body):
            h = blk.forward_nhwc(h)
            if i == 6:
                c1 = h
            elif i == 20:
                c2 = h
            elif i == 23:
                c3 = h
        latents = [self.styles[j].forward_nhwc(c3) for j in range(self.coarse_ind)]
        p2 = ops.bilinear_add(c3, self.latlayer1.forward_nhwc(c2))
        latents += [self.styles[j].forward_nhwc(p2) for j in range(self.coarse_ind, self.middle_ind)]
        p1 = ops.bilinear_add(p2, self.latlayer2.forward_nhwc(c1))
        latents += [self.styles[j].forward_nhwc(p1) for j in range(self.middle_ind, self.style_count)]
        return torch.stack(latents, dim=1)


def load_psp_standalone(checkpoint_path, device='cuda'):
    """util.py:143-161: build the encoder from a pSp checkpoint and add ``latent_avg`` to its output."""
    ckpt = torch.load(checkpoint_path, map_location='cpu')
    opts = ckpt['opts']
    if 'output_size' not in opts:
        opts['output_size'] = 1024
    opts['n_styles'] = int(math.log(opts['output_size'], 2)) * 2 - 2
    psp = GradualStyleEncoder(50, 'ir_se', Namespace(**opts))
    psp.load_state_dict({k.replace('encoder.', ''): v for k, v in ckpt['state_dict'].items() if k.startswith('encoder.')})
    psp.eval().to(device)
    latent_avg = ckpt['latent_avg'].to(device)
    psp.register_forward_hook(lambda m, i, o: o + latent_avg.repeat(o.shape[0], 1, 1))
    return psp
