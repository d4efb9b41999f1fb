"""Up-sampling ModulatedConv2d (Blur o conv_transpose2d): folded single-launch form vs polyphase conv_transpose + FIR.
   python tools/upconv_bench.py [precision]      (on the GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops

dev = torch.device("cuda:0")
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else ops.DEFAULT_PRECISION)
K4 = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64 * 4).to(dev)


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    for (Cin, Cout, H, W) in [(512, 512, 72, 128), (512, 256, 144, 256), (256, 128, 288, 512), (128, 64, 576, 1024), (64, 32, 1152, 2048)]:
        B = 4
        x = torch.randn((B, H, W, Cin), device=dev)
        w9 = ops.prep_weights(torch.randn((Cout, Cin, 3, 3), device=dev) / (3 * Cin ** 0.5), cin_pad=Cin)
        bias = torch.zeros(Cout, device=dev); noise = torch.randn((B, 1, 2 * H, 2 * W), device=dev); nw = torch.tensor([0.1], device=dev)
        wf = ops.fold_upconv_weights(w9, K4)
        folded = lambda: ops.conv_up2_folded_nhwc(x, wf, bias=bias, noise=noise, noise_w=nw, act=1, gain=1.41)
        def two_step():
            t = ops.conv_transpose2d_s2_k3_nhwc(x, w9)
            return ops.fir_nhwc(t, K4, (1, 1), bias=bias, noise=noise, noise_w=nw, act=True, gain=1.41)
        convt = lambda: ops.conv_transpose2d_s2_k3_nhwc(x, w9)
        a, b_, c = timeit(folded), timeit(two_step), timeit(convt)
        err = (folded() - two_step()).abs().max().item()
        print(f"{Cin:4d}->{Cout:4d} {H}x{W}: folded {a:7.3f} ms | conv_transpose (4 phase launches) {c:7.3f} + FIR {b_ - c:7.3f} = {b_:7.3f} ms   (max diff {err:.2e})")
