"""A/B of the M-tiles-per-work-item plan on the N = 128 layers (tuning aid).   python tools/exp_mt.py   (on the GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda:0")


def timed(fn, n=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def make(B, Cin, Cout, H, W, k=3, stride=1):
    x = torch.randn((B, H, W, Cin), device=dev)
    w = ops.prep_weights(torch.randn((Cout, Cin, k, k), device=dev) / (3 * Cin ** 0.5), cin_pad=Cin)
    bias = torch.zeros(Cout, device=dev)
    Ho, Wo = ops.conv_out_size(H, k, stride, k // 2, 1), ops.conv_out_size(W, k, stride, k // 2, 1)
    return (lambda: ops.conv2d_nhwc([x], w, ops.conv_taps(k, k // 2, 1), stride, Ho, Wo, bias=bias, act=1)), 2.0 * B * Ho * Wo * Cout * Cin * k * k


with torch.no_grad():
    cases = {"256->128 576x1024": make(4, 256, 128, 576, 1024), "128->128 576x1024": make(4, 128, 128, 576, 1024),
             "32->128 576x1024": make(4, 32, 128, 576, 1024), "128->256 s2 576x1024": make(4, 128, 256, 576, 1024, stride=2),
             "256->512 s2 288x512": make(4, 256, 512, 288, 512, stride=2)}
    for rnd in range(2):
        for key, vals in ((b"tc_mt", (0, 1, 2, 4)), (b"tc_s2_halo", (0, 1)), (b"tc_tgroup", (0, 1))):
            for v in vals:
                lib.vt_set_option(key, v)
                line = f"round {rnd} {key.decode()}={v}: "
                for name, (fn, fl) in cases.items():
                    try:
                        ms = timed(fn)
                        line += f"{name} {ms:6.3f} ms ({fl / ms / 1e9:4.0f} TF/s) | "
                    except Exception as e:
                        line += f"{name} ERR | "
                print(line, flush=True)
            lib.vt_set_option(key, 0)
