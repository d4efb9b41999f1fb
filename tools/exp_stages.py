"""A/B of the conv_tc pipeline-stage policy on the dilated 512->512 layers and of the halo-staging threshold on the polyphase
up-convolution pieces (tuning aid).   python tools/exp_stages.py     (on the GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda:0")
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)


def timed(fn, n=8):
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
    return (f"prod A-empty {pct(1):4.1f} B-empty {pct(2):4.1f} | mma A-full {pct(6):4.1f} B-full {pct(7):4.1f} tmem-empty {pct(8):4.1f} "
            f"| epi tmem-full {pct(11):4.1f} | xform tma-wait {pct(15):4.1f}")


def conv_case(B, Cin, Cout, H, W, k=3, dil=1, affine=False):
    x = torch.randn((B, H, W, Cin), device=dev)
    w = ops.prep_weights(torch.randn((Cout, Cin, k, k), device=dev) / (3 * Cin ** 0.5), cin_pad=Cin)
    bias = torch.zeros(Cout, device=dev)
    res = torch.randn((B, H, W, Cout), device=dev)
    pad = dil * (k // 2)
    kw = {}
    if affine:
        kw["src_affine"] = [torch.rand((B, Cin, 2), device=dev) + 0.5]
    fn = lambda: ops.conv2d_nhwc([x], w, ops.conv_taps(k, pad, dil), 1, H, W, bias=bias, act=1, **kw)
    ms = timed(fn)
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    print(f"  {Cin}->{Cout} k{k} d{dil} {H}x{W} B{B}{' +affine' if affine else ''}: {ms:7.3f} ms {fl / ms / 1e9:5.0f} TF/s alg | {roles(fn)}", flush=True)


def poly_case(B, Cin, Cout, H, W):
    x = torch.randn((B, H, W, Cin), device=dev)
    w = ops.prep_weights(torch.randn((Cout, Cin, 3, 3), device=dev) / (3 * Cin ** 0.5), cin_pad=Cin)
    fn = lambda: ops.conv_transpose2d_s2_k3_nhwc(x, w)
    ms = timed(fn)
    fl = 2.0 * B * H * W * Cout * Cin * 9
    print(f"  convT {Cin}->{Cout} {H}x{W} B{B}: {ms:7.3f} ms {fl / ms / 1e9:5.0f} TF/s alg", flush=True)


import inspect
has_aff = "src_affine" in inspect.signature(ops.conv2d_nhwc).parameters
with torch.no_grad():
    for pol in (0, 1):
        lib.vt_set_option(b"tc_stage_policy", pol)
        print("stage policy", pol)
        for dil in (1, 2, 4):
            conv_case(4, 512, 512, 72, 128, dil=dil)
        if has_aff:
            conv_case(4, 512, 512, 72, 128, dil=4, affine=True)
    lib.vt_set_option(b"tc_stage_policy", 0)
    for pct in (50, 60, 100):
        lib.vt_set_option(b"tc_halo_pct", pct)
        print("halo pct", pct)
        poly_case(4, 512, 512, 72, 128)
        poly_case(4, 512, 256, 144, 256)
        poly_case(4, 256, 128, 288, 512)
        conv_case(4, 512, 512, 72, 128, k=1)
    lib.vt_set_option(b"tc_halo_pct", 50)
