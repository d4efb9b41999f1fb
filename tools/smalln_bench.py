"""Small-N (Cout <= 4) 3x3 convolution kernels at the Fusion-module sizes of a 576x1024 frame batch: input-stationary vs gather.
   python tools/smalln_bench.py      (on the GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda:0")
ops.set_precision("fp32")


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    for (C, H, W) in [(128, 576, 1024), (256, 288, 512), (512, 144, 256), (512, 72, 128)]:
        B = 4
        x = torch.randn((B, H, W, C), device=dev); x2 = torch.randn((B, H, W, C), device=dev)
        fe = torch.randn((B, H, W, C), device=dev)
        w1 = ops.prep_weights(torch.randn((1, 2 * C, 3, 3), device=dev) / 50, cin_pad=2 * C)
        kc = torch.randn((1, 9, 1), device=dev)
        b1 = torch.zeros(1, device=dev)
        w3 = ops.prep_weights(torch.randn((3, C, 3, 3), device=dev) / 50, cin_pad=C)
        pl = torch.randn((B, 3, H, W), device=dev); wpl = torch.randn((9, 3, 3), device=dev)
        mask = lambda: ops.smalln_conv(x, w1, ops.conv_taps(3, 1), 1, B, H, W, bias=b1, act=_lib.ACT_RELU_TANH, mul_src=fe, src2=x2, tap_const=kc)
        skip = lambda: ops.smalln_conv(fe, w3, ops.conv_taps(3, 1), 3, B, H, W, planar=pl, planar_weight=wpl)
        mask_p = lambda: ops.smalln_conv(x, w1, ops.conv_taps(3, 1), 1, B, H, W, bias=b1, act=_lib.ACT_RELU_TANH, src2=x2, tap_const=kc)
        for name, fn, nbytes in (("mask  N=1 [x||x-x2|] (product)", mask_p, 4.0 * B * H * W * C * 2), ("mask  N=1 [x||x-x2|] + f_E*m", mask, 4.0 * B * H * W * C * 4), ("skip  N=3 + planar", skip, 4.0 * B * H * W * C)):
            res = []
            for mode in (0, 2):
                old = lib.vt_set_option(b"smalln_is", mode)
                res.append(timeit(fn))
                lib.vt_set_option(b"smalln_is", old)
            print(f"C={C:4d} {H}x{W}  {name:32s} gather {res[0]:7.3f} ms  input-stationary {res[1]:7.3f} ms  "
                  f"({nbytes / res[1] / 1e6:6.0f} GB/s algorithmic)")
