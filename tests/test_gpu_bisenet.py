"""GPU parity of the BiSeNet parsing path (style_transfer.py:171-172, model/bisenet/model.py:230-254) against
the reference outputs in tests/golden/bisenet.npz and against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def net():
    from vtoonify_b200.bisenet import BiSeNet
    from vtoonify_b200.weights import det_state_dict
    m = BiSeNet(19).eval()
    m.load_state_dict(det_state_dict(m, seed=21), strict=True)
    return m.cuda()


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-4), ("bf16x3", 1e-3)])
def test_parsing_maps_golden(golden, net, prec, tol):
    from vtoonify_b200 import ops
    g = golden("bisenet")
    x, ref = T(g["x"]).float().cuda(), T(g["x_p"])
    ops.set_precision(prec)
    try:
        y = net.parsing_for_frames(x)
    finally:
        ops.set_precision(ops.DEFAULT_PRECISION)
    assert tuple(y.shape) == tuple(ref.shape)
    err = (y.cpu().double() - ref.double()).abs().max().item()
    rms = ref.pow(2).mean().sqrt().item()
    print(f"BiSeNet parsing maps [{prec}]: max|err| {err:.3e}, ref rms {rms:.3f}")
    assert err <= tol * max(1.0, rms)


def test_forward_matches_fused_frame_path(net):
    """forward(2 * up2(frames))[0] sampled at even pixels == parsing_for_frames(frames) (the fused read-out)"""
    import torch.nn.functional as F
    from oracle import vt_oracle as O
    from vtoonify_b200.weights import det_state_dict
    g = torch.Generator().manual_seed(5)
    frames = torch.rand((1, 3, 40, 56), generator=g) * 2 - 1
    x2 = 2 * F.interpolate(frames, scale_factor=2, mode="bilinear", align_corners=False)
    out, out16, out32 = net(x2.cuda())
    assert tuple(out.shape) == (1, 19, 80, 112) and tuple(out16.shape) == tuple(out32.shape) == tuple(out.shape)
    fused = net.parsing_for_frames(frames.cuda())
    assert (out[:, :, ::2, ::2] - fused).abs().max().item() <= 2e-3 * max(1.0, fused.abs().max().item())
    ref = O.parsing_for_vtoonify(det_state_dict(net, seed=21), frames)
    assert (fused.cpu() - ref).abs().max().item() <= 1e-3 * max(1.0, ref.pow(2).mean().sqrt().item())


def test_resamplers_vs_torch():
    import torch.nn.functional as F
    from vtoonify_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = torch.randn((2, 64, 13, 17), generator=g)
    xn = ops.to_nhwc(x.cuda())
    assert torch.equal(ops.to_nchw(ops.maxpool3x3s2(xn)).cpu(), F.max_pool2d(x, 3, 2, 1))
    assert torch.equal(ops.to_nchw(ops.resize_nearest(xn, 26, 33)).cpu(), F.interpolate(x, (26, 33), mode="nearest"))
    fr = torch.rand((2, 3, 9, 11), generator=g) * 2 - 1
    for up in (False, True):
        X = 2 * F.interpolate(fr, scale_factor=2, mode="bilinear", align_corners=False) if up else fr
        H, W = X.shape[2:]
        Ho, Wo = (H + 1) // 2, (W + 1) // 2
        Xp = F.pad(X, (0, 2 * Wo - W, 0, 2 * Ho - H))
        ref = torch.stack([Xp[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=1).reshape(2, 12, Ho, Wo)
        z = ops.frame_s2d(fr.cuda(), upsample2=up)
        assert tuple(z.shape) == (2, Ho, Wo, 32)
        assert (z[..., :12].permute(0, 3, 1, 2).cpu() - ref).abs().max().item() <= 1e-6 and z[..., 12:].abs().max().item() == 0
    lg = torch.randn((2, 19, 6, 7), generator=g)
    lgn = torch.zeros((2, 6, 7, 32)); lgn[..., :19] = lg.permute(0, 2, 3, 1)
    full = F.interpolate(lg, (24, 28), mode="bilinear", align_corners=True)
    assert (ops.logits_readout(lgn.cuda(), 19, 24, 28).cpu() - full).abs().max().item() <= 1e-5
    assert (ops.logits_readout(lgn.cuda(), 19, 24, 28, step=2, scale=0.5).cpu() - 0.5 * full[:, :, ::2, ::2]).abs().max().item() <= 1e-5
