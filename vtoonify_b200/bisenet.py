"""BiSeNet face parsing as the frame loop uses it (model/bisenet/model.py:17-259, resnet.py:20-83, style_transfer.py:171-174),
same constructor and state_dict keys (191) as the reference, running on the package's kernels:

* every conv + BatchNorm (+ ReLU) is one tensor-core convolution with the BN folded into weights / bias and ReLU in the epilogue;
* the stride-2 7x7 stem runs as a 4x4 stride-1 convolution on a space-to-depth tensor that the frame-preparation kernel writes
  directly (including the frame loop's 2x bilinear up-sampling and the factor 2), so the 49-tap 3-channel stem becomes a
  16-tap 12(->32)-channel one;
* attention branches (global mean -> 1x1 conv -> BN -> sigmoid) are two tiny dense launches on the [B, C] means;
* the final ``align_corners=True`` up-sampling and the frame loop's nearest x0.5 are one read-out kernel that only evaluates
  the pixels that survive.

The host-side algebra below (BN folding, stem re-indexing) is checked on the CPU against torch (tests/test_host_logic.py), the
oracle restatement is pinned to the reference (tests/golden/bisenet.npz) and tests/test_gpu_bisenet.py checks the CUDA path
against both.
"""
import torch
from torch import nn

from . import ops
from ._lib import ACT_LRELU, ACT_NONE
from .psp import BatchNorm2d, _FoldedConv
from .vtoonify import Conv2d


# ---------------------------------------------------------------------------------------------- pure host-side algebra
def s2d_stem_weight(w7: torch.Tensor) -> torch.Tensor:
    """[Cout, 3, 7, 7] stride-2 pad-3 kernel -> [Cout, 12, 4, 4] stride-1 kernel over the space-to-depth tensor
    Z[y, x, (py*2+px)*3 + c] = X[2y+py, 2x+px, c]:  tap u = k-3 in [-3, 3] splits into parity p = u & 1 and offset
    d = (u - p) / 2 in [-2, 1];  W4[n, (py, px, c), dy+2, dx+2] = W7[n, c, ky, kx]."""
    cout = w7.shape[0]
    w4 = torch.zeros((cout, 12, 4, 4), dtype=w7.dtype, device=w7.device)
    for ky in range(7):
        uy = ky - 3
        py = uy & 1
        dy = (uy - py) // 2
        for kx in range(7):
            ux = kx - 3
            px = ux & 1
            dx = (ux - px) // 2
            q = (py * 2 + px) * 3
            w4[:, q:q + 3, dy + 2, dx + 2] = w7[:, :, ky, kx]
    return w4


S2D_TAPS = [(ky - 2, kx - 2, ky * 4 + kx) for ky in range(4) for kx in range(4)]   # (dy, dx, weight slab)


def fold_bn(weight: torch.Tensor, bn: BatchNorm2d):
    """conv (no bias) followed by eval-mode BatchNorm == conv with weight * a[n] plus bias c[n]."""
    a, c = bn.affine()
    return (weight.detach() * a.view(-1, *([1] * (weight.dim() - 1)))).contiguous(), c.contiguous()


# ---------------------------------------------------------------------------------------------- modules
def _relu(x):
    return ops.fused_bias_act(x, None, 0.0, 1.0)


def _conv(x, w, k, stride, pad, **epi):
    B, H, W, _ = x.shape
    Ho, Wo = ops.conv_out_size(H, k, stride, pad, 1), ops.conv_out_size(W, k, stride, pad, 1)
    return ops.conv2d_nhwc([x], w, ops.conv_taps(k, pad), stride, Ho, Wo, **epi)


class ConvBNReLU(nn.Module):
    """model/bisenet/model.py:17-34"""

    def __init__(self, in_chan, out_chan, ks=3, stride=1, padding=1):
        super().__init__()
        self.conv = Conv2d(in_chan, out_chan, ks, stride, padding, bias=False)
        self.bn = BatchNorm2d(out_chan)
        self.ks, self.stride, self.padding = ks, stride, padding
        self._f = _FoldedConv()

    def forward_nhwc(self, x, x2=None):
        cin = x.shape[3] + (0 if x2 is None else x2.shape[3])
        w, b = self._f.get(self.conv, self.bn, cin)
        B, H, W, _ = x.shape
        Ho = ops.conv_out_size(H, self.ks, self.stride, self.padding, 1)
        Wo = ops.conv_out_size(W, self.ks, self.stride, self.padding, 1)
        return ops.conv2d_nhwc([x] if x2 is None else [x, x2], w, ops.conv_taps(self.ks, self.padding), self.stride, Ho, Wo,
                               bias=b, act=ACT_LRELU, slope=0.0, gain=1.0)

    def forward_vector(self, v):
        """the same layer on a [B, C] vector standing for a 1x1 map (kernel centre only)"""
        w, b = fold_bn(self.conv.weight, self.bn)
        c = self.ks // 2
        return ops.linear(v, w[:, :, c, c].contiguous(), b, act=3)


class BasicBlock(nn.Module):
    """model/bisenet/resnet.py:20-47"""

    def __init__(self, in_chan, out_chan, stride=1):
        super().__init__()
        self.conv1 = Conv2d(in_chan, out_chan, 3, stride, 1, bias=False)
        self.bn1 = BatchNorm2d(out_chan)
        self.conv2 = Conv2d(out_chan, out_chan, 3, 1, 1, bias=False)
        self.bn2 = BatchNorm2d(out_chan)
        self.downsample = None
        if in_chan != out_chan or stride != 1:
            self.downsample = nn.Sequential(Conv2d(in_chan, out_chan, 1, stride, 0, bias=False), BatchNorm2d(out_chan))
        self.stride = stride
        self._f1, self._f2, self._fd = _FoldedConv(), _FoldedConv(), _FoldedConv()

    def forward_nhwc(self, x):
        C = x.shape[3]
        w1, b1 = self._f1.get(self.conv1, self.bn1, C)
        r = _conv(x, w1, 3, self.stride, 1, bias=b1, act=ACT_LRELU, slope=0.0, gain=1.0)
        sc = x
        if self.downsample is not None:
            wd, bd = self._fd.get(self.downsample[0], self.downsample[1], C)
            sc = _conv(x, wd, 1, self.stride, 0, bias=bd)
        w2, b2 = self._f2.get(self.conv2, self.bn2, r.shape[3])
        # relu(shortcut + bn2(conv2 r)): the add rides in the conv epilogue, the ReLU is one elementwise pass
        return _relu(_conv(r, w2, 3, 1, 1, bias=b2, act=ACT_NONE, res=sc, alpha=1.0, beta=1.0))


def _layer(in_chan, out_chan, bnum, stride=1):
    return nn.Sequential(BasicBlock(in_chan, out_chan, stride), *[BasicBlock(out_chan, out_chan, 1) for _ in range(bnum - 1)])


class Resnet18(nn.Module):
    """model/bisenet/resnet.py:57-83 (no download: weights come from load_state_dict)"""

    def __init__(self):
        super().__init__()
        self.conv1 = Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = BatchNorm2d(64)
        self.layer1 = _layer(64, 64, 2, 1)
        self.layer2 = _layer(64, 128, 2, 2)
        self.layer3 = _layer(128, 256, 2, 2)
        self.layer4 = _layer(256, 512, 2, 2)
        self._stem = None

    def stem_weights(self):
        key = (self.conv1.weight._version, self.bn1.weight._version, self.bn1.running_var._version, ops.get_precision(),
               self.conv1.weight.device)
        if self._stem is None or self._stem[0] != key:
            w7, b = fold_bn(self.conv1.weight, self.bn1)
            self._stem = (key, ops.prep_weights(s2d_stem_weight(w7), cin_pad=32), b)
        return self._stem[1], self._stem[2]

    def forward_s2d(self, z):
        """z: space-to-depth input [B, H/2, W/2, 32] (ops.frame_s2d) -> feat8, feat16, feat32 (NHWC)"""
        w, b = self.stem_weights()
        B, H, W, _ = z.shape
        x = ops.conv2d_nhwc([z], w, S2D_TAPS, 1, H, W, bias=b, act=ACT_LRELU, slope=0.0, gain=1.0)
        x = ops.maxpool3x3s2(x)
        for blk in self.layer1:
            x = blk.forward_nhwc(x)
        feat8 = x
        for blk in self.layer2:
            feat8 = blk.forward_nhwc(feat8)
        feat16 = feat8
        for blk in self.layer3:
            feat16 = blk.forward_nhwc(feat16)
        feat32 = feat16
        for blk in self.layer4:
            feat32 = blk.forward_nhwc(feat32)
        return feat8, feat16, feat32


def _global_mean(x):
    return ops.instnorm_stats(x)[:, :, 0].contiguous()       # F.avg_pool2d(x, x.size()[2:]) as a [B, C] vector


def _scale_shift(x, scale, shift):
    """x[b,y,x,c] * scale[b,c] + shift[b,c] through the AdaIN kernel with (mean, rstd) = (0, 1)"""
    B, _, _, C = x.shape
    stats = torch.zeros((B, C, 2), device=x.device, dtype=torch.float32)
    stats[:, :, 1] = 1.0
    return ops.adain_apply(x, stats, torch.cat([scale, shift], dim=1))


class AttentionRefinementModule(nn.Module):
    """model/bisenet/model.py:63-84"""

    def __init__(self, in_chan, out_chan):
        super().__init__()
        self.conv = ConvBNReLU(in_chan, out_chan, 3, 1, 1)
        self.conv_atten = Conv2d(out_chan, out_chan, 1, 1, 0, bias=False)
        self.bn_atten = BatchNorm2d(out_chan)

    def forward_nhwc(self, x, add_vec=None, add_map=None):
        """feat * sigmoid(bn(conv1x1(mean feat))) + (per-(b,c) vector | same-size map)"""
        feat = self.conv.forward_nhwc(x)
        w, b = fold_bn(self.conv_atten.weight, self.bn_atten)
        att = ops.linear(_global_mean(feat), w.flatten(1), b, act=4)
        if add_map is not None:
            return ops.gate_shortcut_add(feat, att, add_map)
        return _scale_shift(feat, att, add_vec if add_vec is not None else torch.zeros_like(att))


class ContextPath(nn.Module):
    """model/bisenet/model.py:87-121"""

    def __init__(self):
        super().__init__()
        self.resnet = Resnet18()
        self.arm16 = AttentionRefinementModule(256, 128)
        self.arm32 = AttentionRefinementModule(512, 128)
        self.conv_head32 = ConvBNReLU(128, 128, 3, 1, 1)
        self.conv_head16 = ConvBNReLU(128, 128, 3, 1, 1)
        self.conv_avg = ConvBNReLU(512, 128, 1, 1, 0)

    def forward_s2d(self, z):
        feat8, feat16, feat32 = self.resnet.forward_s2d(z)
        avg = self.conv_avg.forward_vector(_global_mean(feat32))                  # [B,128]; nearest up-sampling of a 1x1 map
        f32 = self.arm32.forward_nhwc(feat32, add_vec=avg)
        f32_up = self.conv_head32.forward_nhwc(ops.resize_nearest(f32, feat16.shape[1], feat16.shape[2]))
        f16 = self.arm16.forward_nhwc(feat16, add_map=f32_up)
        f16_up = self.conv_head16.forward_nhwc(ops.resize_nearest(f16, feat8.shape[1], feat8.shape[2]))
        return feat8, f16_up, f32_up


class FeatureFusionModule(nn.Module):
    """model/bisenet/model.py:187-213"""

    def __init__(self, in_chan, out_chan):
        super().__init__()
        self.convblk = ConvBNReLU(in_chan, out_chan, 1, 1, 0)
        self.conv1 = Conv2d(out_chan, out_chan // 4, 1, 1, 0, bias=False)
        self.conv2 = Conv2d(out_chan // 4, out_chan, 1, 1, 0, bias=False)

    def forward_nhwc(self, fsp, fcp):
        feat = self.convblk.forward_nhwc(fsp, fcp)                                 # virtual concat
        h = ops.linear(_global_mean(feat), self.conv1.weight.flatten(1), None, act=3)
        att = ops.linear(h, self.conv2.weight.flatten(1), None, act=4)
        return ops.gate_shortcut_add(feat, att, feat)                              # feat * atten + feat


class BiSeNetOutput(nn.Module):
    """model/bisenet/model.py:36-60; the 1x1 classifier is padded to 32 output channels for the tensor-core kernel"""

    def __init__(self, in_chan, mid_chan, n_classes):
        super().__init__()
        self.conv = ConvBNReLU(in_chan, mid_chan, 3, 1, 1)
        self.conv_out = Conv2d(mid_chan, n_classes, 1, 1, 0, bias=False)
        self.n_classes = n_classes
        self._w = None

    def logits_nhwc(self, x):
        x = self.conv.forward_nhwc(x)
        key = (self.conv_out.weight._version, ops.get_precision(), self.conv_out.weight.device)
        if self._w is None or self._w[0] != key:
            w = self.conv_out.weight.detach()
            npad = ops._pad32(w.shape[0])
            wp = torch.zeros((npad, w.shape[1], 1, 1), device=w.device, dtype=torch.float32)
            wp[:w.shape[0]] = w
            self._w = (key, ops.prep_weights(wp, cin_pad=x.shape[3]))
        return _conv(x, self._w[1], 1, 1, 0)                                       # [B, h, w, 32], first n_classes valid


class BiSeNet(nn.Module):
    """model/bisenet/model.py:216-259.  ``forward`` keeps the reference signature (three planar logit maps at input size);
    ``parsing_for_frames`` is the fused form of style_transfer.py:171-172."""

    def __init__(self, n_classes):
        super().__init__()
        self.cp = ContextPath()
        self.ffm = FeatureFusionModule(256, 256)
        self.conv_out = BiSeNetOutput(256, 256, n_classes)
        self.conv_out16 = BiSeNetOutput(128, 64, n_classes)
        self.conv_out32 = BiSeNetOutput(128, 64, n_classes)
        self.n_classes = n_classes

    def _features(self, z):
        feat_res8, feat_cp8, feat_cp16 = self.cp.forward_s2d(z)
        return self.ffm.forward_nhwc(feat_res8, feat_cp8), feat_cp8, feat_cp16

    def forward(self, x):
        H, W = x.shape[2:]
        fuse, cp8, cp16 = self._features(ops.frame_s2d(x, upsample2=False))
        return tuple(ops.logits_readout(head.logits_nhwc(f), self.n_classes, H, W)
                     for head, f in ((self.conv_out, fuse), (self.conv_out16, cp8), (self.conv_out32, cp16)))

    def parsing_for_frames(self, frames, scale=1.0, out=None):
        """frames [B,3,H,W] in [-1,1] -> ``scale * x_p`` [B,n_classes,H,W] with
        x_p = F.interpolate(self(2 * F.interpolate(frames, scale_factor=2, mode='bilinear'))[0], scale_factor=0.5)
        (style_transfer.py:171-172).  ``out``: optional channel slice of the network input (``inputs[:, 3:]``) to fill in place."""
        H, W = frames.shape[2:]
        fuse, _, _ = self._features(ops.frame_s2d(frames, upsample2=True))
        return ops.logits_readout(self.conv_out.logits_nhwc(fuse), self.n_classes, 2 * H, 2 * W, step=2, scale=scale, out=out)
