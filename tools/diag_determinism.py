"""Repeat small forwards / ops many times and report any run that differs from the first (races, uninitialised reads)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import ops, _lib
from vtoonify_b200.op import upfirdn2d
from vtoonify_b200.vtoonify import VToonify
from vtoonify_b200.weights import det_inputs, det_state_dict
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
k1 = torch.tensor([1., 3., 3., 1.]); K = (k1[:, None] * k1[None, :] / 64).to(dev)
g = torch.Generator().manual_seed(1)
for name, shape, kw in (("up", (1, 3, 16, 16), dict(up=2, pad=(2, 1))), ("up", (1, 3, 64, 64), dict(up=2, pad=(2, 1))), ("blur", (2, 8, 33, 33), dict(pad=(1, 1))),
                        ("blur22", (2, 8, 33, 33), dict(pad=(2, 2))), ("down", (2, 8, 33, 33), dict(down=2, pad=(1, 1))), ("up", (4, 3, 288, 512), dict(up=2, pad=(2, 1)))):
    x = torch.randn(shape, generator=g).to(dev)
    kk = K * 4 if "up" in kw else K
    first = upfirdn2d(x, kk, **kw).clone()
    bad = 0
    for i in range(200):
        torch.empty(1 << 20, device=dev).normal_()        # perturb timing / smem contents of later blocks
        y = upfirdn2d(x, kk, **kw)
        if not torch.equal(y, first):
            bad += 1
    lib = _lib.load(); old = lib.vt_set_option(b"upfirdn_tiled", 0)
    ref = upfirdn2d(x, kk, **kw); lib.vt_set_option(b"upfirdn_tiled", old)
    print(f"upfirdn2d {name} {shape}: {bad}/200 runs differ from the first; max|stream - generic| {(first - ref).abs().max().item():.2e}")
for backbone, (B, H, W) in (("toonify", (1, 32, 32)), ("dualstylegan", (1, 32, 32)), ("dualstylegan", (2, 64, 96))):
    m = VToonify(backbone=backbone).eval(); m.load_state_dict(det_state_dict(m, seed=0), strict=True); m.to(dev)
    x, s = det_inputs(B, H, W, seed=100); x, s = x.to(dev), s.to(dev)
    first = m(x, s, d_s=0.5).clone()
    bad = 0
    for i in range(50):
        x2, _ = det_inputs(B, H, W, seed=200 + i)
        m(x2.to(dev), s, d_s=0.5)                          # a different frame in between
        y = m(x, s, d_s=0.5)
        if not torch.equal(y, first):
            bad += 1
            if bad == 1:
                print("   first mismatch: max diff", (y - first).abs().max().item(), "count", (y != first).sum().item())
    print(f"VToonify {backbone} {B}x{H}x{W}: {bad}/50 runs differ from the first")
