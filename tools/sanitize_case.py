"""Tiny shapes through every hand-rolled mbarrier / TMEM pipeline, meant to run under compute-sanitizer (tools/run_sanitizer.sh):
conv_tc (per-tap, halo, CTA pair, folded up-conv, fused ToRGB), conv_rs (single CTA and pair, both epilogue groups, ring laps),
small-N heads, FIR, upfirdn2d, frame kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops
torch.set_grad_enabled(False)
lib = _lib.load()
dev = "cuda"
g = torch.Generator().manual_seed(0)
K4 = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64 * 4).to(dev)


def w_for(cout, cin, k=3):
    return ops.prep_weights((torch.randn((cout, cin, k, k), generator=g) / (k * cin ** 0.5)).to(dev), cin_pad=cin, round_tf32=False)


def conv(B, cin, cout, H, W, k=3, **kw):
    x = torch.randn((B, H, W, cin), generator=g).to(dev)
    return ops.conv2d_nhwc([x], w_for(cout, cin, k), ops.conv_taps(k, k // 2), 1, H, W, bias=torch.zeros(cout, device=dev), act=1, **kw)

ops.set_option("rs_min_width", 1)
for rs in (False, True):
    ops.set_option("rs_conv", rs)
    for cg in (1, 2):
        lib.vt_set_option(b"rs_cg", cg); lib.vt_set_option(b"rs_rows", 5)
        rgb = {"w": torch.randn((1, 1, 3, 32), generator=g).to(dev), "bias": torch.zeros(3, device=dev),
               "skip": torch.randn((1, 3, 9, 136 if cg == 2 else 72), generator=g).to(dev), "kernel": K4}
        conv(1, 32, 32, 18, 272 if cg == 2 else 144, rgb=rgb)
        conv(2, 64, 64, 17, 260 if cg == 2 else 130)
lib.vt_set_option(b"rs_cg", 0); lib.vt_set_option(b"rs_rows", 0)
ops.set_option("rs_conv", False)
conv(1, 64, 256, 24, 40)                       # N = 256 -> CTA pair, halo staging
conv(1, 128, 128, 20, 24)
conv(1, 64, 32, 16, 16, k=1)
lib.vt_set_option(b"tc_mode", 0); conv(1, 64, 64, 16, 24); lib.vt_set_option(b"tc_mode", 1)     # one TMA box per tap
x = torch.randn((1, 12, 16, 64), generator=g).to(dev)
w9 = w_for(32, 64)
ops.conv_up2_folded_nhwc(x, ops.fold_upconv_weights(w9, K4), bias=torch.zeros(32, device=dev), act=1)      # folded up-conv
t = ops.conv_transpose2d_s2_k3_nhwc(x, ops.prep_weights((torch.randn((32, 64, 3, 3), generator=g) / 24).to(dev), cin_pad=64))
ops.fir_nhwc(t, K4, (1, 1), bias=torch.zeros(32, device=dev), act=True)
xp = torch.randn((2, 3, 20, 24), generator=g).to(dev)
ops.upfirdn2d_planar(xp, K4, (2, 2), (1, 1), (2, 1, 2, 1)); ops.upfirdn2d_planar(xp, K4, (1, 1), (2, 2), (1, 1, 1, 1))
# streaming upfirdn2d (bulk-copy ring): rows of 4k+1 floats (every copy lead), several strips / chunks / ring laps, a tensor that
# starts 4 bytes into its allocation (with PYTORCH_NO_CUDA_MEMORY_CACHING=1 every tensor is its own cudaMalloc: exact bounds)
for shape in ((2, 3, 70, 1025), (1, 2, 9, 7), (1, 1, 130, 2053)):
    xs = torch.randn(shape, generator=g).to(dev)
    ops.upfirdn2d_planar(xs, K4 / 4, (1, 1), (1, 1), (1, 1, 1, 1)); ops.upfirdn2d_planar(xs, K4 / 4, (1, 1), (1, 1), (2, 2, 2, 2))
    ops.upfirdn2d_planar(xs, K4, (2, 2), (1, 1), (2, 1, 2, 1)); ops.upfirdn2d_planar(xs, K4 / 4, (1, 1), (2, 2), (1, 1, 1, 1))
    ops.upfirdn2d_planar(xs, torch.randn((4, 4), generator=g).to(dev), (1, 1), (1, 1), (1, 1, 1, 1))
buf = torch.randn(2 * 37 * 131 + 1, generator=g).to(dev)
ops.upfirdn2d_planar(buf[1:].view(1, 2, 37, 131), K4 / 4, (1, 1), (1, 1), (1, 1, 1, 1))
# image-only last layer (row-strip kernel, out = NULL), row-strip up-conv, mask head (input-stationary small-N kernel)
ops.set_option("rs_conv", True)
rgb1 = {"w": torch.randn((1, 1, 3, 32), generator=g).to(dev), "bias": torch.zeros(3, device=dev), "skip": None, "kernel": None, "only": True}
conv(1, 32, 32, 18, 272, rgb=rgb1)
ops.set_option("rs_conv", False)
xu = torch.randn((1, 10, 256, 64), generator=g).to(dev)
ops.conv_up2_rs_nhwc(xu, w_for(32, 64), K4, bias=torch.zeros(32, device=dev), act=1)
xm = torch.randn((1, 40, 70, 64), generator=g).to(dev); xm2 = torch.randn((1, 40, 70, 64), generator=g).to(dev)
ops.smalln_conv(xm, w_for(1, 128), ops.conv_taps(3, 1), 1, 1, 40, 70, bias=torch.zeros(1, device=dev), act=_lib.ACT_RELU_TANH, src2=xm2)
ops.fused_bias_act(xp, torch.zeros(3, device=dev), 0.2, 1.4)
fr = torch.randint(0, 256, (2, 33, 47, 3), generator=g, dtype=torch.uint8).to(dev)
ops.frame_prefilter_resize(fr, 2, (20, 15), (1, 14, 2, 19))
ops.f32_to_frames_u8(ops.frames_u8_to_f32(fr))
st = ops.instnorm_stats(torch.randn((2, 12, 16, 64), generator=g).to(dev))
torch.cuda.synchronize()
print("sanitize_case: done")
