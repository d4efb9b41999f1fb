"""GPU parity of the two reference custom ops and the small streaming kernels, through the C-ABI (ctypes)."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import vt_oracle as O

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def T(a):
    return torch.from_numpy(np.asarray(a))


def maxerr(a, b):
    assert tuple(a.shape) == tuple(b.shape), f"shape {tuple(a.shape)} vs {tuple(b.shape)}"
    return (a.double() - b.double()).abs().max().item()


def _cfg(g, i):
    cfg = json.loads(str(g[f"u{i}_cfg"]))
    up = tuple(cfg["up"]) if isinstance(cfg["up"], list) else cfg["up"]
    down = tuple(cfg["down"]) if isinstance(cfg["down"], list) else cfg["down"]
    return up, down, tuple(cfg["pad"])


def test_upfirdn2d_golden(golden):
    from vtoonify_b200.op import upfirdn2d
    g = golden("ops")
    for i in range(int(g["n_upfirdn"])):
        up, down, pad = _cfg(g, i)
        y = upfirdn2d(T(g[f"u{i}_x"]).cuda(), T(g[f"u{i}_k"]).cuda(), up=up, down=down, pad=pad).cpu()
        ref = T(g[f"u{i}_y"])
        assert maxerr(y, ref) <= 2e-6, f"case {i}: {maxerr(y, ref)}"


def test_upfirdn2d_random_vs_oracle():
    """index-exactness sweep: random shapes / up / down / signed pads / non-symmetric kernels vs the oracle."""
    from vtoonify_b200.op import upfirdn2d
    rng = np.random.RandomState(0)
    g = torch.Generator().manual_seed(0)
    n = 0
    while n < 60:
        H, W = int(rng.randint(1, 20)), int(rng.randint(1, 20))
        kh, kw = int(rng.randint(1, 7)), int(rng.randint(1, 7))
        up = (int(rng.randint(1, 4)), int(rng.randint(1, 4)))
        down = (int(rng.randint(1, 4)), int(rng.randint(1, 4)))
        pad = tuple(int(v) for v in rng.randint(-2, 5, size=4))
        oh = (H * up[1] + pad[2] + pad[3] - kh + down[1]) // down[1]
        ow = (W * up[0] + pad[0] + pad[1] - kw + down[0]) // down[0]
        if oh < 1 or ow < 1 or H * up[1] + min(pad[2], 0) + min(pad[3], 0) < 1 or W * up[0] + min(pad[0], 0) + min(pad[1], 0) < 1:
            continue
        x = torch.randn((2, 3, H, W), generator=g)
        k = torch.randn((kh, kw), generator=g)
        ref = O.upfirdn2d(x, k, up, down, pad)
        y = upfirdn2d(x.cuda(), k.cuda(), up=up, down=down, pad=pad).cpu()
        assert y.shape == ref.shape, (H, W, kh, kw, up, down, pad)
        assert maxerr(y, ref) <= 1e-5, (H, W, kh, kw, up, down, pad, maxerr(y, ref))
        n += 1


def test_upfirdn2d_hot_shapes_large():
    """The hot-path instances at realistic plane sizes, checked through closed forms + oracle on a crop."""
    from vtoonify_b200.op import upfirdn2d
    k = (O.make_kernel([1, 3, 3, 1]) * 4).cuda()
    g = torch.Generator().manual_seed(1)
    x = torch.randn((2, 4, 257, 513), generator=g)
    y = upfirdn2d(x.cuda(), k, pad=(1, 1)).cpu()
    assert y.shape == (2, 4, 256, 512)
    assert maxerr(y, O.upfirdn2d(x, k.cpu(), pad=(1, 1))) <= 2e-6
    s = torch.randn((2, 3, 128, 96), generator=g)
    u = upfirdn2d(s.cuda(), k, up=2, pad=(2, 1)).cpu()
    assert u.shape == (2, 3, 256, 192)
    assert maxerr(u, O.upfirdn2d(s, k.cpu(), up=2, pad=(2, 1))) <= 2e-6
    # linearity (size-independent property)
    a, b = torch.randn((1, 2, 65, 33), generator=g).cuda(), torch.randn((1, 2, 65, 33), generator=g).cuda()
    lhs = upfirdn2d(2 * a + b, k, pad=(1, 1))
    rhs = 2 * upfirdn2d(a, k, pad=(1, 1)) + upfirdn2d(b, k, pad=(1, 1))
    assert maxerr(lhs.cpu(), rhs.cpu()) <= 1e-5


def test_upfirdn2d_vs_c_oracle():
    from vtoonify_b200.op import upfirdn2d
    so = os.path.join(os.path.dirname(O.__file__), "_build", "libvt_oracle.so")
    if not os.path.exists(so):
        pytest.skip("C oracle not built")
    lib = ctypes.CDLL(so)
    g = torch.Generator().manual_seed(3)
    x = torch.randn((1, 2, 9, 7), generator=g)
    k = torch.randn((3, 4), generator=g)
    up, down, pad = (2, 3), (2, 1), (1, -1, 2, 0)
    oh = (9 * up[1] + pad[2] + pad[3] - 3 + down[1]) // down[1]
    ow = (7 * up[0] + pad[0] + pad[1] - 4 + down[0]) // down[0]
    out = torch.empty((1, 2, oh, ow))
    fp = ctypes.POINTER(ctypes.c_float)
    rc = lib.vo_upfirdn2d(ctypes.cast(x.data_ptr(), fp), ctypes.cast(k.data_ptr(), fp), ctypes.cast(out.data_ptr(), fp),
                          ctypes.c_long(2), 9, 7, 3, 4, up[0], up[1], down[0], down[1], *pad)
    assert rc == 0
    y = upfirdn2d(x.cuda(), k.cuda(), up=up, down=down, pad=pad).cpu()
    assert maxerr(y, out) <= 1e-5


def test_fused_leaky_relu_bit_exact(golden):
    from vtoonify_b200.op import FusedLeakyReLU, fused_leaky_relu
    g = golden("ops")
    for tag, args in (("f0", ()), ("f1", ()), ("f3", (0.1, 0.5))):
        y = fused_leaky_relu(T(g[tag + "_x"]).cuda(), T(g[tag + "_b"]).cuda(), *args).cpu()
        assert torch.equal(y, T(g[tag + "_y"])), tag
    y = fused_leaky_relu(T(g["f2_x"]).cuda(), None, 0.2, 1.0).cpu()
    assert torch.equal(y, T(g["f2_y"]))
    m = FusedLeakyReLU(5).cuda()
    m.bias.data.copy_(T(g["f0_b"]))
    assert torch.equal(m(T(g["f0_x"]).cuda()).cpu(), T(g["f0_y"]))
    # large + odd sizes (vector and scalar kernels), 64-bit offsets
    gen = torch.Generator().manual_seed(5)
    for shape in ((2, 7, 33, 31), (1, 32, 256, 256), (3, 5)):
        x = torch.randn(shape, generator=gen); b = torch.randn(shape[1], generator=gen)
        assert torch.equal(fused_leaky_relu(x.cuda(), b.cuda()).cpu(), O.fused_leaky_relu(x, b))


def test_cpu_tensor_raises():
    from vtoonify_b200 import _lib
    from vtoonify_b200.op import fused_leaky_relu, upfirdn2d
    with pytest.raises(_lib.VtError):
        upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))
    with pytest.raises(_lib.VtError):
        fused_leaky_relu(torch.zeros(1, 2, 4, 4), torch.zeros(2))


def test_layout_roundtrip_and_padding():
    from vtoonify_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn((2, 22, 17, 19), generator=g).cuda()
    n = ops.to_nhwc(x, 32, round_tf32=False)
    assert n.shape == (2, 17, 19, 32)
    assert torch.equal(n[..., :22], x.permute(0, 2, 3, 1))
    assert torch.count_nonzero(n[..., 22:]) == 0
    assert torch.equal(ops.to_nchw(n, 22), x)


def test_linear_pixelnorm(golden):
    from vtoonify_b200.stylegan import EqualLinear, PixelNorm
    from tests.shapes import layer_state_dict
    g = golden("layers")
    m = EqualLinear(512, 512, lr_mul=0.01, activation="fused_lrelu")
    m.load_state_dict(layer_state_dict("EqualLinear", "el")); m.cuda()
    assert maxerr(m(T(g["el_x"]).cuda()).cpu(), T(g["el_y"])) <= 2e-5
    assert maxerr(PixelNorm()(T(g["el_x"]).cuda()).cpu(), T(g["pn_y"])) <= 1e-6


def test_frame_transforms_bit_exact():
    from vtoonify_b200 import ops
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (2, 37, 53, 3), generator=g, dtype=torch.uint8)
    assert torch.equal(ops.frames_u8_to_f32(u8.cuda()).cpu(), O.frame_u8_to_f32(u8))
    img = torch.randn((2, 3, 37, 53), generator=g) * 1.5
    assert torch.equal(ops.f32_to_frames_u8(img.cuda(), swap_rb=True).cpu(), O.tensor2frame_u8(img, True))
    # every representable u8 round-trips: f32(u8) -> u8
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 1, 256, 1).repeat(1, 1, 1, 3)
    back = ops.f32_to_frames_u8(ops.frames_u8_to_f32(ramp.cuda()), swap_rb=False).cpu()
    assert (back.int() - ramp.int()).abs().max() <= 1   # truncation may lose one level; reference behaves identically
    assert torch.equal(back, O.tensor2frame_u8(O.frame_u8_to_f32(ramp), False))


def test_fir_nhwc_vs_oracle():
    from vtoonify_b200 import ops
    ops.set_precision("fp32")
    g = torch.Generator().manual_seed(4)
    x = torch.randn((2, 8, 13, 11), generator=g)
    k = O.make_kernel([1, 3, 3, 1]) * 4
    bias = torch.randn(8, generator=g); noise = torch.randn((2, 1, 12, 10), generator=g); nw = torch.tensor([0.3])
    ref = O.fused_leaky_relu(O.upfirdn2d(x, k, pad=(1, 1)) + nw * noise, bias)
    y = ops.fir_nhwc(ops.to_nhwc(x.cuda()), k.cuda(), (1, 1), bias=bias.cuda(), noise=noise.cuda(), noise_w=nw.cuda(), act=True)
    assert maxerr(ops.to_nchw(y).cpu(), ref) <= 2e-6
    k2 = torch.randn((3, 2), generator=g)
    y2 = ops.fir_nhwc(ops.to_nhwc(x.cuda()), k2.cuda(), (2, 1))
    assert maxerr(ops.to_nchw(y2).cpu(), O.upfirdn2d(x, k2, pad=(2, 1))) <= 1e-5
    ops.set_precision(ops.DEFAULT_PRECISION)


def test_adain_vs_oracle():
    from vtoonify_b200 import ops
    import torch.nn.functional as F
    ops.set_precision("fp32")
    g = torch.Generator().manual_seed(6)
    x = torch.randn((2, 64, 23, 17), generator=g) * 2 + 0.7
    x2 = torch.randn((2, 64, 23, 17), generator=g)
    gb = torch.randn((2, 128), generator=g)
    st = ops.instnorm_stats(ops.to_nhwc(x.cuda()))
    y = ops.to_nchw(ops.adain_apply(ops.to_nhwc(x.cuda()), st, gb.cuda())).cpu()
    ref = gb[:, :64, None, None] * F.instance_norm(x, eps=1e-5) + gb[:, 64:, None, None]
    assert maxerr(y, ref) <= 2e-5
    cat = torch.cat([x, (x - x2).abs()], 1)
    gb2 = torch.randn((2, 256), generator=g)
    st2 = ops.instnorm_stats(ops.to_nhwc(x.cuda()), ops.to_nhwc(x2.cuda()))
    y2 = ops.to_nchw(ops.adain_apply(ops.to_nhwc(x.cuda()), st2, gb2.cuda(), ops.to_nhwc(x2.cuda()))).cpu()
    ref2 = gb2[:, :128, None, None] * F.instance_norm(cat, eps=1e-5) + gb2[:, 128:, None, None]
    assert maxerr(y2, ref2) <= 2e-5
    ops.set_precision(ops.DEFAULT_PRECISION)


def test_upfirdn2d_tiled_and_generic_kernels_agree():
    """The shared-memory tiled kernel and the generic kernel implement the same index semantics (random configs incl.
    larger planes, negative pads, up/down up to 3) and both match the oracle."""
    from vtoonify_b200 import _lib
    from vtoonify_b200.op import upfirdn2d
    lib = _lib.load()
    rng = np.random.RandomState(7)
    g = torch.Generator().manual_seed(7)
    n = 0
    while n < 40:
        H, W = int(rng.randint(1, 90)), int(rng.randint(1, 300))
        kh, kw = int(rng.randint(1, 7)), int(rng.randint(1, 7))
        up = (int(rng.randint(1, 4)), int(rng.randint(1, 4)))
        down = (int(rng.randint(1, 4)), int(rng.randint(1, 4)))
        pad = tuple(int(v) for v in rng.randint(-2, 5, size=4))
        oh = (H * up[1] + pad[2] + pad[3] - kh + down[1]) // down[1]
        ow = (W * up[0] + pad[0] + pad[1] - kw + down[0]) // down[0]
        if oh < 1 or ow < 1 or H * up[1] + min(pad[2], 0) + min(pad[3], 0) < 1 or W * up[0] + min(pad[0], 0) + min(pad[1], 0) < 1:
            continue
        x = torch.randn((2, 3, H, W), generator=g)
        k = torch.randn((kh, kw), generator=g)
        ref = O.upfirdn2d(x, k, up, down, pad)
        outs = []
        for tiled in (2, 1, 0):
            old = lib.vt_set_option(b"upfirdn_tiled", tiled)
            try:
                outs.append(upfirdn2d(x.cuda(), k.cuda(), up=up, down=down, pad=pad).cpu())
            finally:
                lib.vt_set_option(b"upfirdn_tiled", old)
        for y in outs:
            assert y.shape == ref.shape, (H, W, kh, kw, up, down, pad)
            assert maxerr(y, ref) <= 2e-5 * max(1.0, ref.abs().max().item()), (H, W, kh, kw, up, down, pad, maxerr(y, ref))
        n += 1


@pytest.mark.parametrize("up,down", [(1, 1), (2, 1), (1, 2)])
def test_upfirdn2d_streaming_kernel(up, down):
    """The streaming 4x4 kernel (bulk-copy ring + register column filters; default for Blur / Upsample / Downsample) against the
    oracle and the generic kernel: strips (> 256 / 512 output columns), row chunks, ring laps, rows of 4k+1..4k+3 floats (every
    copy lead), pads 0..5 and crops, separable and full-rank taps, 1-pixel planes, a view that starts 4 bytes into its storage."""
    from vtoonify_b200 import _lib
    from vtoonify_b200.op import upfirdn2d
    lib = _lib.load()
    assert lib.vt_set_option(b"upfirdn_tiled", 2) == 2       # it is the default
    g = torch.Generator().manual_seed(100 * up + down)
    k1 = torch.tensor([1., 3., 3., 1.])
    ksep = k1[:, None] * k1[None, :] / 64 * (up * up)
    cases = [(1, 3, 70, 300), (2, 2, 33, 1025), (1, 1, 1, 1), (1, 2, 5, 3), (3, 1, 129, 515), (1, 1, 200, 64), (2, 1, 16, 2050)]
    pads = [(1, 1), (2, 1), (2, 2), (0, 0), (3, 0), (5, 4), (-1, 2), (1, -1)]
    n = 0
    for ci, (B, C, H, W) in enumerate(cases):
        for pi, pad in enumerate(pads):
            if (ci + pi) % 3 and ci > 1:
                continue                                   # every case with a third of the pads (all pads on the first two)
            oh = (H * up + pad[0] + pad[1] - 4 + down) // down
            ow = (W * up + pad[0] + pad[1] - 4 + down) // down
            if oh < 1 or ow < 1 or H * up + min(pad[0], 0) + min(pad[1], 0) < 1 or W * up + min(pad[0], 0) + min(pad[1], 0) < 1:
                continue
            for sep in (True, False):
                k = ksep if sep else torch.randn((4, 4), generator=g)
                x = torch.randn((B, C, H, W), generator=g)
                ref = O.upfirdn2d(x, k, up, down, pad)
                y = upfirdn2d(x.cuda(), k.cuda(), up=up, down=down, pad=pad).cpu()
                old = lib.vt_set_option(b"upfirdn_tiled", 0)
                try:
                    y0 = upfirdn2d(x.cuda(), k.cuda(), up=up, down=down, pad=pad).cpu()
                finally:
                    lib.vt_set_option(b"upfirdn_tiled", old)
                tol = 2e-6 * max(1.0, ref.abs().max().item())
                assert y.shape == ref.shape
                assert maxerr(y, ref) <= tol and maxerr(y, y0) <= tol, (B, C, H, W, pad, sep, maxerr(y, ref), maxerr(y, y0))
                n += 1
    # misaligned storage: the tensor starts 4 / 8 / 12 bytes into its allocation (first / last rows take the in-bounds path)
    for off in (1, 2, 3):
        buf = torch.randn(2 * 3 * 37 * 131 + off, generator=g).cuda()
        x = buf[off:].view(2, 3, 37, 131)
        assert x.data_ptr() % 16 == 4 * off
        ref = O.upfirdn2d(x.cpu(), ksep, up, down, (2, 1))
        assert maxerr(upfirdn2d(x, ksep.cuda(), up=up, down=down, pad=(2, 1)).cpu(), ref) <= 2e-6 * max(1.0, ref.abs().max().item())
    assert n >= 40


@pytest.mark.parametrize("shape", [(2, 32, 33, 65), (1, 64, 17, 130), (1, 8, 9, 7)])
def test_fir4_specialised_kernel_vs_generic_and_oracle(shape):
    """4x4 pad (1,1) Blur on NHWC: register-tiled kernel vs the generic kernel vs upfirdn2d of the oracle (odd sizes = the
    (2H+1)x(2W+1) output of a stride-2 transposed conv), with the fused noise + bias + leaky-relu tail."""
    from vtoonify_b200 import _lib, ops
    ops.set_precision("fp32")
    B, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn((B, C, H, W), generator=g)
    k = torch.randn((4, 4), generator=g)          # not separable on purpose: the kernel is used as given
    bias = torch.randn(C, generator=g); noise = torch.randn((B, 1, H - 1, W - 1), generator=g); nw = torch.tensor([0.3])
    ref = O.fused_leaky_relu(O.upfirdn2d(x, k, pad=(1, 1)) + nw * noise, bias)
    lib = _lib.load()
    outs = []
    for mode in (0, 1):
        old = lib.vt_set_option(b"fir4", mode)
        try:
            outs.append(ops.to_nchw(ops.fir_nhwc(ops.to_nhwc(x.cuda()), k.cuda(), (1, 1), bias=bias.cuda(), noise=noise.cuda(),
                                                 noise_w=nw.cuda(), act=True)).cpu())
        finally:
            lib.vt_set_option(b"fir4", old)
    ops.set_precision(ops.DEFAULT_PRECISION)
    assert maxerr(outs[0], ref) <= 1e-5 and maxerr(outs[1], ref) <= 1e-5, (maxerr(outs[0], ref), maxerr(outs[1], ref))
