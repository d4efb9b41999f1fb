"""256 -> 128 up-convolution @288x512 (B=4): polyphase conv_transpose + NHWC FIR (the current route) against the row-strip up-conv
kernel with Cin = 256 (4 passes of 32 output channels), correctness vs the fp32 route and interleaved timing (tuning aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops

dev = torch.device("cuda:0")
K4 = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64 * 4).to(dev)


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    B, Cin, Cout, H, W = 4, 256, 128, 288, 512
    g = torch.Generator().manual_seed(1)
    x = torch.randn((B, H, W, Cin), generator=g).to(dev)
    wt = (torch.randn((Cout, Cin, 3, 3), generator=g) / (3 * Cin ** 0.5)).to(dev)
    w9 = ops.prep_weights(wt, cin_pad=Cin, round_tf32=False)
    wp = ops.prep_weights(wt, cin_pad=Cin)
    bias = (torch.randn(Cout, generator=g) * 0.2).to(dev)
    poly = lambda prec=None: ops.fir_nhwc(ops.conv_transpose2d_s2_k3_nhwc(x, wp, precision=prec), K4, (1, 1), bias=bias, act=True, slope=0.2, gain=2 ** 0.5)
    rsu = lambda: ops.conv_up2_rs_nhwc(x, w9, K4, bias=bias, act=1, slope=0.2, gain=2 ** 0.5)
    ref = poly("fp32")
    y = rsu()
    scale = ref.abs().max().item()
    print(f"rsu Cin=256 vs fp32 route: max|err| {(y - ref).abs().max().item():.3e} (max|ref| {scale:.2f}); "
          f"polyphase bf16x3 vs fp32: {(poly() - ref).abs().max().item():.3e}")
    del ref, y
    for rnd in range(3):
        print(f"round {rnd}: polyphase + FIR {timed(poly):.3f} ms   row-strip up {timed(rsu):.3f} ms", flush=True)
