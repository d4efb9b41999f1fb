"""diagnostic: row-strip kernel vs tap-by-tap vs fp32 at the full-size layer shapes"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops
torch.set_grad_enabled(False)
lib = _lib.load()
dev = "cuda"
def run(B, C, H, W, rows=0, fmt="bf16"):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((B, H, W, C), generator=g).to(dev)
    w = ops.prep_weights((torch.randn((C, C, 3, 3), generator=g) / (3 * C ** 0.5)).to(dev), cin_pad=C, round_tf32=False)
    bias = torch.zeros(C, device=dev)
    taps = ops.conv_taps(3, 1)
    ops.set_option("rs_conv", False)
    ref = ops.conv2d_nhwc([x], w, taps, 1, H, W, bias=bias, act=1, precision="fp32")
    tap = ops.conv2d_nhwc([x], w, taps, 1, H, W, bias=bias, act=1)
    ops.set_option("rs_conv", True); ops.set_option("rs_fmt", fmt)
    lib.vt_set_option(b"rs_rows", rows)
    rs = ops.conv2d_nhwc([x], w, taps, 1, H, W, bias=bias, act=1)
    rs2 = ops.conv2d_nhwc([x], w, taps, 1, H, W, bias=bias, act=1)
    torch.cuda.synchronize()
    sc = float(ref.abs().max())
    d = (rs - ref).abs()
    idx = torch.nonzero(d == d.max())[0].tolist()
    print(f"B{B} C{C} {H}x{W} rows={rows} fmt={fmt}: tap-fp32 {float((tap-ref).abs().max())/sc:.2e}  rs-fp32 {float(d.max())/sc:.2e} at {idx}  "
          f"rs-rs {float((rs-rs2).abs().max())/sc:.2e}  rows with err>1e-4*sc: {int((d.amax(dim=(0,2,3)) > 1e-4*sc).sum())}")
    bad = (d.amax(dim=(0, 2, 3)) > 2e-5 * sc).nonzero().flatten().tolist()
    print("   bad rows (first 40):", bad[:40])
run(4, 32, 2304, 4096)
run(1, 32, 2304, 4096)
run(4, 64, 1152, 2048)
run(1, 32, 1024, 1024, rows=16)
run(4, 32, 2304, 4096, fmt="f16")
