"""Per-role wait-cycle breakdown of conv_tc_kernel for selected layer shapes (tuning aid).
   python tools/tc_role_timing.py            (on the GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda:0")
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
K4 = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64 * 4).to(dev)


def report(name, fn, flops):
    fn(); torch.cuda.synchronize()
    dbg.zero_()
    lib.vt_set_debug_buffer(dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.vt_set_debug_buffer(None)
    d = dbg.view(148, 16).double()
    ms = e0.elapsed_time(e1)
    tot = d[:, 0].mean().item()
    def pct(i): return 100 * d[:, i].mean().item() / max(tot, 1)
    print(f"{name:30s} {ms:7.3f} ms {flops / ms / 1e9:5.0f} TF/s cyc/CTA {tot:9.0f} | prod A-empty {pct(1):4.1f} B-empty {pct(2):4.1f} "
          f"| mma A-full {pct(6):4.1f} B-full {pct(7):4.1f} tmem-empty {pct(8):4.1f} "
          f"| epi tmem-full {pct(11):4.1f} tmem-ld {pct(12):4.1f} math {pct(13):4.1f} stage+store {pct(14):4.1f}"
          f" | xform tma-wait {pct(15):4.1f}")


def conv_case(B, Cin, Cout, H, W, k=3, dil=1, rgb=False):
    x = torch.randn((B, H, W, Cin), device=dev)
    w = ops.prep_weights(torch.randn((Cout, Cin, k, k), device=dev) / (3 * Cin ** 0.5), cin_pad=Cin)
    bias = torch.zeros(Cout, device=dev)
    pad = dil * (k // 2)
    r = None
    if rgb:   # fused ToRGB tail with the skip image (what conv2 of every generator level carries)
        r = {"w": torch.randn((B, 1, 3, Cout), device=dev) * 0.1, "bias": torch.zeros(3, device=dev),
             "skip": torch.randn((B, 3, H // 2, W // 2), device=dev), "kernel": K4}
    fn = lambda: ops.conv2d_nhwc([x], w, ops.conv_taps(k, pad, dil), 1, H, W, bias=bias, act=1, rgb=r)
    report(f"{Cin}->{Cout} k{k} d{dil} {H}x{W} B{B}{' +rgb' if rgb else ''}", fn, 2.0 * B * H * W * Cout * Cin * k * k)


def up_case(B, Cin, Cout, H, W):
    x = torch.randn((B, H, W, Cin), device=dev)
    w9 = ops.prep_weights(torch.randn((Cout, Cin, 3, 3), device=dev) / (3 * Cin ** 0.5), cin_pad=Cin, round_tf32=False)
    wf = ops.fold_upconv_weights(w9, K4)
    bias = torch.zeros(Cout, device=dev)
    fn = lambda: ops.conv_up2_folded_nhwc(x, wf, bias=bias, act=1)
    report(f"up {Cin}->{Cout} {H}x{W} B{B}", fn, 2.0 * B * H * W * Cout * Cin * 36)


def rsu_case(B, Cin, Cout, H, W):
    x = torch.randn((B, H, W, Cin), device=dev)
    w9 = ops.prep_weights(torch.randn((Cout, Cin, 3, 3), device=dev) / (3 * Cin ** 0.5), cin_pad=Cin, round_tf32=False)
    bias = torch.zeros(Cout, device=dev)
    fn = lambda: ops.conv_up2_rs_nhwc(x, w9, K4, bias=bias, act=1)
    report(f"up {Cin}->{Cout} {H}x{W} B{B} [row-strip up]", fn, 2.0 * B * H * W * Cout * Cin * 18)


ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else ops.DEFAULT_PRECISION)
print("precision", ops.get_precision())
with torch.no_grad():
    if len(sys.argv) > 2 and sys.argv[2] == "rsu":
        up_case(4, 64, 32, 1152, 2048)
        rsu_case(4, 64, 32, 1152, 2048)
        up_case(4, 128, 64, 576, 1024)
        rsu_case(4, 128, 64, 576, 1024)
        rsu_case(8, 64, 32, 512, 512)
        rsu_case(8, 128, 64, 256, 256)
        for rows in (8, 16, 32, 64):
            lib.vt_set_option(b"rsu_rows", rows)
            print("rsu rows_per_strip", rows)
            rsu_case(4, 64, 32, 1152, 2048)
        lib.vt_set_option(b"rsu_rows", 0)
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "rs":
        for rs in (False, True):
            ops.set_option("rs_conv", rs)
            print("row-strip kernel" if rs else "tap-by-tap kernel")
            for fmt in (("bf16", "f16") if rs else ("bf16",)):
                ops.set_option("rs_fmt", fmt)
                conv_case(4, 32, 32, 2304, 4096)
                conv_case(4, 32, 32, 2304, 4096, rgb=True)
                conv_case(4, 64, 64, 1152, 2048)
                conv_case(4, 64, 64, 1152, 2048, rgb=True)
                conv_case(8, 32, 32, 1024, 1024, rgb=True)
        ops.set_option("rs_fmt", "bf16")
        for strict in (1, 0):
            lib.vt_set_option(b"rs_strict", strict)
            print("strict cluster-scope release" if strict else "plain remote arrive")
            conv_case(4, 32, 32, 2304, 4096, rgb=True)
            conv_case(4, 64, 64, 1152, 2048, rgb=True)
        for cg in (1, 2):
            lib.vt_set_option(b"rs_cg", cg)
            print("rs_cg", cg)
            conv_case(4, 32, 32, 2304, 4096, rgb=True)
        lib.vt_set_option(b"rs_cg", 0)
        for rows in (48, 96):
            lib.vt_set_option(b"rs_rows", rows)
            print("rows_per_strip", rows)
            conv_case(4, 32, 32, 2304, 4096, rgb=True)
        lib.vt_set_option(b"rs_rows", 0)
        sys.exit(0)
    conv_case(4, 512, 512, 72, 128)
    conv_case(4, 256, 256, 288, 512)
    conv_case(4, 128, 128, 576, 1024)
    conv_case(4, 64, 64, 1152, 2048)
    conv_case(4, 32, 32, 2304, 4096)
    conv_case(4, 32, 32, 2304, 4096, rgb=True)
    conv_case(4, 64, 64, 1152, 2048, rgb=True)
    up_case(4, 64, 32, 1152, 2048)
    up_case(4, 128, 64, 576, 1024)
    up_case(4, 512, 256, 144, 256)
    conv_case(4, 256, 128, 576, 1024)
    conv_case(4, 128, 32, 576, 1024, k=1)
    conv_case(4, 128, 256, 288, 512, k=3)

out = torch.zeros(4, device=dev)
for mode, name in ((0, "ld x32 + wait"), (1, "4 x ld x32, one wait"), (2, "st x32 + wait"), (3, "4 x st x32, one wait")):
    _lib.check(lib.vt_selftest_tc_gemm(None, None, out.data_ptr(), 0, mode, 2000, 64, None))
    torch.cuda.synchronize()
    print(f"TMEM probe {name:22s}: {out[0].item():7.1f} cycles per 32-column access per warp (4 warps active) "
          f"= {128 * 32 * 4 / max(out[0].item(), 1e-9):7.1f} B/clk/SM")
