"""stride-2 3x3 layers of the encoder: one box per tap (single CTA) vs halo staging + CTA pairs (option tc_s2_halo)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import ops, _lib
lib = _lib.load(); dev = "cuda"
g = torch.Generator().manual_seed(0)
def run(x, w, b, H, W, prec):
    ops.set_precision(prec)
    Ho, Wo = ops.conv_out_size(H, 3, 2, 1, 1), ops.conv_out_size(W, 3, 2, 1, 1)
    y = ops.conv2d_nhwc([x], w, ops.conv_taps(3, 1), 2, Ho, Wo, bias=b, act=1)
    ops.set_precision(ops.DEFAULT_PRECISION)
    return y
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (B, Cin, Cout, H, W) in [(1, 64, 256, 33, 47), (2, 128, 256, 64, 96), (4, 128, 256, 576, 1024), (4, 256, 512, 288, 512), (4, 512, 512, 144, 256)]:
    x = torch.randn((B, H, W, Cin), generator=g).to(dev)
    w = ops.prep_weights((torch.randn((Cout, Cin, 3, 3), generator=g) / (3 * Cin ** 0.5)).to(dev), cin_pad=Cin, round_tf32=False)
    b = torch.randn(Cout, generator=g).to(dev)
    ref = run(x, w, b, H, W, "fp32") if H <= 64 else None
    res = []
    for mode in (0, 1):
        lib.vt_set_option(b"tc_s2_halo", mode)
        y = run(x, w, b, H, W, "bf16x3")
        ms = t(lambda: run(x, w, b, H, W, "bf16x3"))
        res.append((y, ms))
    lib.vt_set_option(b"tc_s2_halo", 0)
    err = (res[1][0] - res[0][0]).abs().max().item()
    e2 = (res[1][0] - ref).abs().max().item() if ref is not None else float("nan")
    print(f"{Cin}->{Cout} s2 {H}x{W} B{B}: per-tap {res[0][1]:.3f} ms, halo+pairs {res[1][1]:.3f} ms; max|halo - per-tap| {err:.2e}, |halo - fp32| {e2:.2e}")
