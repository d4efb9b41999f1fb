"""``vtoonify_b200.op.conv2d_gradfix`` (a9: model/stylegan/op/conv2d_gradfix.py:22-75) against torch's own CPU convolutions,
including the ``groups = batch`` form ``ModulatedConv2d.forward`` uses (model/stylegan/model.py:273-304), and a restatement of that
forward built on the drop-in ops (what ``install_as_reference_ops()`` + the reference's unmodified module executes)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

TOL = 2e-4     # bf16x3 tensor-core path / fp32 FFMA path, relative to max|ref|


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def _close(y, ref, what, tol=TOL):
    assert tuple(y.shape) == tuple(ref.shape), f"{what}: shape {tuple(y.shape)} vs {tuple(ref.shape)}"
    err = (y.cpu().double() - ref.double()).abs().max().item()
    scale = max(1e-6, ref.abs().max().item())
    print(f"{what}: max|err| {err:.3e} (max|ref| {scale:.3f})")
    assert err <= tol * scale, f"{what}: {err:.3e} > {tol} * {scale:.3f}"


CONV_CASES = [
    # B, Cin, Cout, H, W, kh, kw, stride, padding, dilation, bias
    (2, 32, 64, 19, 23, 3, 3, 1, 1, 1, True),
    (1, 22, 32, 16, 24, 3, 3, 1, 1, 1, True),          # Cin not a multiple of 32 (the encoder's first conv)
    (2, 64, 48, 17, 20, 3, 3, 1, 1, 1, False),         # Cout padded to 64 and sliced
    (2, 64, 64, 18, 22, 3, 3, 2, 1, 1, True),          # stride 2
    (1, 32, 32, 20, 20, 3, 3, 1, 2, 2, False),         # dilation 2
    (2, 64, 32, 9, 11, 1, 1, 1, 0, 1, True),           # 1x1
    (1, 32, 32, 14, 15, 3, 3, 1, 0, 1, True),          # 'valid' padding: output smaller than input
    (2, 64, 3, 12, 13, 3, 3, 1, 1, 1, True),           # Cout <= 4, same geometry -> planar head
    (2, 64, 3, 12, 13, 3, 3, 1, 0, 1, True),           # Cout <= 4, padding 0 (round-1 bug: wrong shape) -> general path
    (1, 32, 3, 16, 16, 3, 3, 2, 1, 1, False),          # Cout <= 4, stride 2
    (1, 32, 8, 10, 12, 1, 3, 1, (0, 1), 1, True),      # rectangular kernel, per-axis padding
    (1, 32, 32, 16, 16, 5, 5, 1, 2, 1, False),         # 25 taps
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[f"c{i}" for i in range(len(CONV_CASES))])
def test_conv2d_vs_torch(case):
    from vtoonify_b200.op import conv2d_gradfix
    B, Cin, Cout, H, W, kh, kw, s, p, d, has_bias = case
    x = _rand((B, Cin, H, W), 1)
    w = _rand((Cout, Cin, kh, kw), 2, 1 / math.sqrt(Cin * kh * kw))
    b = _rand((Cout,), 3, 0.1) if has_bias else None
    ref = F.conv2d(x, w, b, stride=s, padding=p, dilation=d)
    y = conv2d_gradfix.conv2d(x.cuda(), w.cuda(), None if b is None else b.cuda(), stride=s, padding=p, dilation=d)
    _close(y, ref, f"conv2d {case}")


def test_conv2d_rejects_unsupported():
    from vtoonify_b200.op import conv2d_gradfix
    x, w = torch.zeros((2, 64, 8, 8)).cuda(), torch.zeros((64, 32, 3, 3)).cuda()
    with pytest.raises(NotImplementedError):
        conv2d_gradfix.conv2d(x, w, groups=2)                       # a true grouped conv over a batch of 2
    with pytest.raises(NotImplementedError):
        conv2d_gradfix.conv2d(x, torch.zeros((64, 64, 3, 3)).cuda(), stride=(1, 2))
    with pytest.raises(NotImplementedError):
        conv2d_gradfix.conv_transpose2d(x, torch.zeros((64, 64, 3, 3)).cuda(), stride=1)


@pytest.mark.parametrize("G,Cin,Cout,k,stride,pad", [(3, 32, 64, 3, 1, 1), (2, 64, 3, 1, 1, 0), (2, 32, 32, 3, 2, 0), (4, 64, 64, 3, 1, 1)])
def test_conv2d_groups_is_batch(G, Cin, Cout, k, stride, pad):
    """input [1, G*Cin, H, W], weight [G*Cout, Cin, k, k], groups=G  (model.py:291-301)"""
    from vtoonify_b200.op import conv2d_gradfix
    H, W = 13, 18
    x = _rand((1, G * Cin, H, W), 4)
    w = _rand((G * Cout, Cin, k, k), 5, 1 / math.sqrt(Cin * k * k))
    ref = F.conv2d(x, w, None, stride=stride, padding=pad, groups=G)
    y = conv2d_gradfix.conv2d(x.cuda(), w.cuda(), padding=pad, stride=stride, groups=G)
    assert y.is_contiguous()                                        # the reference .view()s the result to [B, Cout, H, W]
    _close(y, ref, f"conv2d groups={G} {Cin}->{Cout} k{k} s{stride}")


@pytest.mark.parametrize("G,Cin,Cout", [(1, 64, 32), (3, 32, 64), (2, 64, 48)])
def test_conv_transpose2d(G, Cin, Cout):
    """stride 2, padding 0, 3x3; groups=G with weight [G*Cin, Cout, 3, 3]  (model.py:273-283)"""
    from vtoonify_b200.op import conv2d_gradfix
    H, W = 9, 12
    x = _rand((1, G * Cin, H, W), 6)
    w = _rand((G * Cin, Cout, 3, 3), 7, 1 / math.sqrt(Cin * 9))
    ref = F.conv_transpose2d(x, w, None, stride=2, padding=0, groups=G)
    y = conv2d_gradfix.conv_transpose2d(x.cuda(), w.cuda(), padding=0, stride=2, groups=G)
    _close(y, ref, f"conv_transpose2d groups={G} {Cin}->{Cout}")


def _reference_modconv_forward(op, x, style, sd, prefix, scale, demodulate, upsample, downsample, blur_kernel, blur_pad, padding):
    """The fused branch of the reference's ModulatedConv2d.forward (model/stylegan/model.py:259-304), statement by
    statement, on top of an ``op`` package exporting ``upfirdn2d`` and ``conv2d_gradfix`` like model/stylegan/op."""
    batch, in_channel, height, width = x.shape
    W5 = sd[prefix + "weight"]                                       # [1, Cout, Cin, k, k]
    out_channel, k = W5.shape[1], W5.shape[3]
    s = F.linear(style, sd[prefix + "modulation.weight"] * (1 / math.sqrt(style.shape[1])), sd[prefix + "modulation.bias"])
    weight = scale * W5 * s.view(batch, 1, in_channel, 1, 1)
    if demodulate:
        demod = torch.rsqrt(weight.pow(2).sum([2, 3, 4]) + 1e-8)
        weight = weight * demod.view(batch, out_channel, 1, 1, 1)
    weight = weight.view(batch * out_channel, in_channel, k, k)
    if upsample:
        inp = x.view(1, batch * in_channel, height, width)
        weight = weight.view(batch, out_channel, in_channel, k, k).transpose(1, 2).reshape(batch * in_channel, out_channel, k, k)
        out = op.conv2d_gradfix.conv_transpose2d(inp, weight, padding=0, stride=2, groups=batch)
        _, _, h2, w2 = out.shape
        out = out.view(batch, out_channel, h2, w2)
        return op.upfirdn2d(out, blur_kernel, pad=blur_pad)
    if downsample:
        inp = op.upfirdn2d(x, blur_kernel, pad=blur_pad)
        _, _, h2, w2 = inp.shape
        inp = inp.reshape(1, batch * in_channel, h2, w2)
        out = op.conv2d_gradfix.conv2d(inp, weight, padding=0, stride=2, groups=batch)
    else:
        inp = x.view(1, batch * in_channel, height, width)
        out = op.conv2d_gradfix.conv2d(inp, weight, padding=padding, groups=batch)
    _, _, h2, w2 = out.shape
    return out.view(batch, out_channel, h2, w2)


@pytest.mark.parametrize("mode", ["plain", "up", "down", "torgb"])
def test_reference_modulated_conv_on_dropin_ops(mode):
    """install_as_reference_ops() level: the reference's own ModulatedConv2d.forward statements running
    on this package's ops == the oracle's modulated_conv2d."""
    import vtoonify_b200
    from oracle import vt_oracle as O
    from vtoonify_b200.stylegan import ModulatedConv2d
    from vtoonify_b200.weights import det_state_dict
    import sys
    saved = {k: sys.modules.get(k) for k in ("model.stylegan.op", "model.stylegan.op.conv2d_gradfix")}
    try:
        op = vtoonify_b200.install_as_reference_ops()
        assert sys.modules["model.stylegan.op"] is op and hasattr(op, "FusedLeakyReLU") and hasattr(op, "fused_leaky_relu")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    B, Cin, H, W = 3, 64, 12, 10
    Cout, k = (3, 1) if mode == "torgb" else (32, 3)
    m = ModulatedConv2d(Cin, Cout, k, 512, demodulate=mode != "torgb", upsample=mode == "up", downsample=mode == "down")
    sd = {"c." + kk: v for kk, v in det_state_dict(m, seed=11).items()}
    x, style = _rand((B, Cin, H, W), 8), _rand((B, 512), 9)
    ref = O.modulated_conv2d(x, style, sd, "c.", demodulate=mode != "torgb", upsample=mode == "up", downsample=mode == "down")
    dev = {kk: v.cuda() for kk, v in sd.items()}
    blur_k = dev.get("c.blur.kernel")
    y = _reference_modconv_forward(op, x.cuda(), style.cuda(), dev, "c.", m.scale, mode != "torgb", mode == "up", mode == "down",
                                   blur_k, tuple(m.blur.pad) if hasattr(m, "blur") else None, m.padding)
    _close(y, ref, f"reference ModulatedConv2d.forward [{mode}] on vtoonify_b200.op", tol=3e-4)
