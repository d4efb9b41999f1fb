"""A/B of the row-strip up-convolution's epilogue forms (tuning aid).   python tools/exp_rsu.py   (on the GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda:0")
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
K4 = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64 * 4).to(dev)


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


def roles(fn):
    dbg.zero_()
    lib.vt_set_debug_buffer(dbg.data_ptr())
    fn(); torch.cuda.synchronize()
    lib.vt_set_debug_buffer(None)
    d = dbg.view(148, 16).double()
    tot = d[:, 0].mean().item()
    def pct(i): return 100 * d[:, i].mean().item() / max(tot, 1)
    return (f"prod A-empty {pct(1):4.1f} B-empty {pct(2):4.1f} | mma A-ready {pct(6):4.1f} B-full {pct(7):4.1f} row-empty {pct(8):4.1f} "
            f"| epi row-full {pct(11):4.1f} | xform tma-wait {pct(15):4.1f}")


def rsu_case(B, Cin, Cout, H, W):
    x = torch.randn((B, H, W, Cin), device=dev)
    w9 = ops.prep_weights(torch.randn((Cout, Cin, 3, 3), device=dev) / (3 * Cin ** 0.5), cin_pad=Cin, round_tf32=False)
    bias = torch.zeros(Cout, device=dev)
    noise = torch.randn((B, 1, 2 * H, 2 * W), device=dev)
    nw = torch.tensor([0.1], device=dev)
    fn = lambda: ops.conv_up2_rs_nhwc(x, w9, K4, bias=bias, noise=noise, noise_w=nw, act=1)
    ms = timed(fn)
    print(f"  up {Cin}->{Cout} {H}x{W} B{B}: {ms:7.3f} ms | {roles(fn)}", flush=True)


with torch.no_grad():
    for rnd in range(2):
        for epi, bst in ((0, 4), (1, 4), (1, 6), (0, 6)):
            lib.vt_set_option(b"rsu_epi", epi)
            lib.vt_set_option(b"rsu_bstages", bst)
            print(f"round {rnd}: rsu_epi {epi} rsu_bstages {bst}")
            rsu_case(4, 64, 32, 1152, 2048)
            rsu_case(4, 128, 64, 576, 1024)
    lib.vt_set_option(b"rsu_epi", 1)
    lib.vt_set_option(b"rsu_bstages", 6)
