"""``conv2d_gradfix.conv2d`` / ``conv_transpose2d`` — same call signatures as model/stylegan/op/conv2d_gradfix.py:22-75.
In the reference these forward to cuDNN (F.conv2d / F.conv_transpose2d) on any modern torch; here they run the
library's NHWC convolution kernels (tcgen05 when shapes allow, fp32 FFMA otherwise).  Forward only.

Supported: groups == 1, square kernels up to 3x3; conv_transpose2d: stride 2, padding 0, 3x3 (the only form the
reference uses, model/stylegan/model.py:236-238, 281-283)."""
import contextlib

import torch

from .. import ops

enabled = True
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    if groups != 1:
        raise NotImplementedError("vtoonify_b200 conv2d: groups != 1 is not part of the inference hot path")
    s, p, d = _pair(stride), _pair(padding), _pair(dilation)
    if s[0] != s[1] or p[0] != p[1] or d[0] != d[1]:
        raise NotImplementedError("vtoonify_b200 conv2d: anisotropic stride/padding/dilation")
    Cout, Cin, kh, kw = weight.shape
    if kh != kw or kh * kw > 9:
        raise NotImplementedError("vtoonify_b200 conv2d: kernels up to 3x3 only")
    B, C, H, W = input.shape
    x = ops.to_nhwc(input, ops._pad32(C) if C % 4 else None)
    w = ops.prep_weights(weight, cin_pad=x.shape[3])
    Ho = ops.conv_out_size(H, kh, s[0], p[0], d[0])
    Wo = ops.conv_out_size(W, kw, s[0], p[0], d[0])
    if Cout <= 4:
        if s[0] != 1:
            raise NotImplementedError("vtoonify_b200 conv2d: strided conv with Cout <= 4")
        return ops.smalln_conv(x, w, ops.conv_taps(kh, p[0], d[0]), Cout, B, H, W, bias=bias)
    y = ops.conv2d_nhwc([x], w, ops.conv_taps(kh, p[0], d[0]), s[0], Ho, Wo, bias=bias)
    return ops.nhwc_as_nchw_view(y)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if groups != 1 or _pair(stride) != (2, 2) or _pair(padding) != (0, 0) or _pair(output_padding) != (0, 0) \
            or _pair(dilation) != (1, 1) or tuple(weight.shape[2:]) != (3, 3):
        raise NotImplementedError("vtoonify_b200 conv_transpose2d: only stride=2, padding=0, 3x3, groups=1")
    x = ops.to_nhwc(input)
    # F.conv_transpose2d weight is [Cin, Cout, kh, kw]
    w = ops.prep_weights(weight.transpose(0, 1).contiguous(), cin_pad=x.shape[3])
    y = ops.conv_transpose2d_s2_k3_nhwc(x, w)
    if bias is not None:
        y = ops.nhwc_as_nchw_view(y) + bias.view(1, -1, 1, 1)
        return y
    return ops.nhwc_as_nchw_view(y)
