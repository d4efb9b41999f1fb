"""Tiny shapes through the code paths added in the second session of round 2, for compute-sanitizer (tools/run_sanitizer.sh <tag>
tools/sanitize_case_r02b.py): per-warp output stores (4-row TMA boxes) incl. partial tiles, fused output statistics, the 2 + 6 stage
plan of big halo boxes (dilation 4), halo staging of the 2-tap polyphase pieces, conv_rsu's one-pass epilogue with the 6-deep weight
ring (single CTA and CTA pair, ring laps), instance-norm statistics with large chunks."""
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


def conv(B, cin, cout, H, W, k=3, dil=1, **kw):
    x = torch.randn((B, H, W, cin), generator=g).to(dev)
    return ops.conv2d_nhwc([x], w_for(cout, cin, k), ops.conv_taps(k, dil * (k // 2), dil), 1, H, W, bias=torch.zeros(cout, device=dev),
                           act=1, **kw)


ops.set_option("rs_conv", False)
ops.set_option("rs_min_width", 1)
conv(2, 64, 512, 19, 27, want_stats=True)                # 2 N tiles, CTA pairs, partial tiles, statistics
conv(1, 64, 256, 24, 40, dil=4, want_stats=True)         # big halo boxes: 2 halo + 6 weight stages
conv(1, 64, 128, 9, 7, want_stats=True)                  # 2 M tiles per work item, smaller than a tile
conv(1, 128, 32, 16, 24, k=1, want_stats=True)           # 1x1, 4 M tiles per work item
for ws in (0, 1):                                        # CTA-wide vs per-warp stores
    lib.vt_set_option(b"tc_warp_store", ws)
    conv(1, 64, 64, 17, 23)
lib.vt_set_option(b"tc_warp_store", 1)
x = torch.randn((1, 12, 16, 256), generator=g).to(dev)
ops.conv_transpose2d_s2_k3_nhwc(x, w_for(256, 256))      # polyphase pieces: 1, 2 and 4 taps (2-tap ones now halo-staged as CTA pairs)
for epi in (0, 1):
    lib.vt_set_option(b"rsu_epi", epi)
    for cg, W in ((1, 140), (2, 300)):
        lib.vt_set_option(b"rsu_cg", cg); lib.vt_set_option(b"rsu_rows", 4)
        xu = torch.randn((1, 11, W, 64), generator=g).to(dev)
        ops.conv_up2_rs_nhwc(xu, w_for(32, 64), K4, bias=torch.zeros(32, device=dev), act=1,
                             noise=torch.randn((1, 1, 22, 2 * W), generator=g).to(dev), noise_w=torch.tensor([0.3], device=dev))
        xu = torch.randn((1, 6, W, 128), generator=g).to(dev)
        ops.conv_up2_rs_nhwc(xu, w_for(64, 128), K4, bias=torch.zeros(64, device=dev), act=1)
lib.vt_set_option(b"rsu_epi", 1); lib.vt_set_option(b"rsu_cg", 0); lib.vt_set_option(b"rsu_rows", 0)
xs = torch.randn((2, 150, 201, 128), generator=g).to(dev)
ops.instnorm_stats(xs); ops.instnorm_stats(xs, torch.randn((2, 150, 201, 128), generator=g).to(dev))
torch.cuda.synchronize()
print("sanitize_case: done")
